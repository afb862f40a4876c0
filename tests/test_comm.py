"""The collective of the C-ABI (bppgpu_comm_*, bppgpu_allreduce_sum) driven from a C host.

tests/c/two_engines.c is the shape BPP itself would use: one process, one engine per GPU, one pthread per
engine over a contiguous range of loci (threads.c:234-263), partial sums added by NCCL
(threads.c:544-558, 583-590).  With one GPU the same program runs a communicator of one rank."""
import os
import subprocess

import pytest

from helpers import ROOT
from bpp_b200 import build

SRC = os.path.join(ROOT, "tests", "c", "two_engines.c")
EXE = os.path.join(ROOT, "tests", "c", "two_engines")


def compile_program():
    build.build_cuda()
    build.build_host()
    pkg = os.path.join(ROOT, "bpp_b200")
    cmd = ["gcc", "-O1", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
           "-o", EXE, SRC, "-L", pkg, "-lbpphost", "-lbppgpu", "-Wl,-rpath," + pkg, "-lpthread", "-lm"]
    subprocess.check_call(cmd)
    return EXE


def test_c_program_compiles_against_the_public_headers():
    assert os.path.exists(compile_program())


def _run(n):
    exe = compile_program()
    r = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.gpu
def test_one_rank_communicator_from_c():
    out = _run(1)
    assert "OK" in out and "nccl" in out


@pytest.mark.gpu
def test_two_engines_in_one_process():
    from bpp_b200 import engine
    if engine.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    assert "OK" in _run(2)


@pytest.mark.gpu
def test_python_comm_one_rank_and_batch_allreduce():
    import numpy as np
    from bpp_b200 import engine, synth
    w = synth.make_workload("comm", n_loci=16, tips=5, sites=64, states=4, rate_cats=4, model="GTR", seed=11)
    e = engine.Engine(0)
    loci, trees = engine.load_workload(e, w)
    b = engine.Batch(e, loci)
    step = trees.full_pass_step()
    lnl, total = b.full_pass(step)
    c = engine.Comm(e, 1, 0, engine.Comm.unique_id())
    b.stage(step)
    b.run()
    b.allreduce_lnl_sum(c)
    lnl2, total2 = b.collect()
    assert total2 == total and np.array_equal(lnl, lnl2)
    assert np.array_equal(c.allreduce_sum([1.5, -2.0, 3.0, 4.0]), [1.5, -2.0, 3.0, 4.0])
    assert c.calls == 2
    c.destroy()
    b.destroy()
    e.close()
