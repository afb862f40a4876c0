"""CPU: the reference arm of bench.py (the reference's own AVX2 path on the host cores, oracle/_ref) prints one
JSON line with the contract's keys; ranks other than 0 stay silent.  (The b200 arm needs a GPU.)"""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT
from oracle import refbind

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built")


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--loci-per-gpu-sample", "48"], capture_output=True, text=True, env=env,
                          timeout=300)


def test_reference_arm_line():
    r = _run({})
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"] == "locus_lnL_evals_per_sec_full_tree" and line["unit"] == "locus-lnL evals/s"
    assert line["value"] > 0 and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("config3") and line["config"]["reference_sample_loci"] == 48
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_uses_the_config5_shard_shape_on_several_gpus():
    r = _run({"RANK": "0", "WORLD_SIZE": "8", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["config"]["workload"].startswith("config5") and line["config"]["patterns"] == 2000
    assert line["n_gpus"] == 8


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
