"""Partial-update throughput: one gene-tree age move per locus across ALL loci of a batch as one launch
(gtree.c:4585, 5437-5467: 2-3 P-matrices, the root path's partials, root lnL), the schedule SURVEY.md 8f rank 2
asks for, with the compiled reference doing the same moves locus by locus on the host cores beside it.

  python tools/partial_update_bench.py [T ...]        -> one JSON line per tip count (default 8 16 48)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bpp_b200 import engine, synth  # noqa: E402


def run(T, n_loci, sites=1000, rate_cats=4, model="GTR", rounds=30, ref_loci=2048):
    ref_loci = min(ref_loci, n_loci)
    w = synth.make_workload("age%d" % T, n_loci=n_loci, tips=T, sites=sites, states=4, rate_cats=rate_cats,
                            model=model, seed=synth.SEED + T)
    eng = engine.Engine(0)
    loci, trees = engine.load_workload(eng, w)
    batch = engine.Batch(eng, loci)
    lnl0, _ = batch.full_pass(trees.full_pass_step())
    rng = np.random.default_rng(T)
    moves = []
    for _ in range(rounds):
        nodes, ages = engine.propose_ages(trees, rng)
        moves.append((nodes, ages, engine.age_move_step(trees, nodes, ages)))
    # parity + CPU baseline on a sample (the reference applies the same moves, in the same order)
    cpu = None
    from oracle import refbind
    if refbind.available() and not os.environ.get("PU_NO_CPU"):
        from helpers import ref_set_from_workload
        ws = w.subset(ref_loci)
        rs = ref_set_from_workload(ws)
        threads = os.cpu_count() or 1
        rs.full_pass_all(0, ws.n_loci, threads, 1)
        secs, ref = 0.0, None
        for nodes, ages, _ in moves:
            s, ref = rs.age_move_all(0, ws.n_loci, nodes[:ws.n_loci], ages[:ws.n_loci], threads)
            secs += s
        rs.close()
        cpu = {"moves_per_sec": ws.n_loci * rounds / secs, "cores": threads, "loci": ws.n_loci,
               "us_per_move_per_core": 1e6 * secs * threads / (ws.n_loci * rounds)}
    # e2e: every array of the step from host memory, stage + run + collect (the lists change with every move)
    out = np.zeros(w.n_loci)
    for _, _, step in moves[:3]:
        batch.full_pass(step)
    eng.reset_profile()
    eng.set_profiling(True)
    t0 = time.perf_counter()
    for _, _, step in moves:
        batch.stage(step)
        batch.run()
        lnl, total = batch.collect(out)
    secs = time.perf_counter() - t0
    prof = eng.profile()
    eng.set_profiling(False)
    err = None
    if cpu is not None:
        err = float(np.max(np.abs(lnl[:ref.size][:cpu["loci"]] - ref[:cpu["loci"]]) / np.abs(ref[:cpu["loci"]])))
    node_updates = float(sum(int(m[2][3].sum()) for m in moves))
    dev_ms = sum(v["ms"] for v in prof.values()) / rounds
    rec = {"what": "one age move per locus across all loci as one batch (root-path partial update)",
           "tips": T, "loci": n_loci, "patterns": sites, "rate_cats": rate_cats, "model": model, "rounds": rounds,
           "moves_per_sec_e2e": n_loci * rounds / secs, "us_per_batch_e2e": 1e6 * secs / rounds,
           "ms_per_batch_device": dev_ms, "moves_per_sec_device": n_loci / (dev_ms / 1000.0),
           "node_updates_per_sec_e2e": node_updates / secs, "mean_path_length": node_updates / (rounds * n_loci),
           "per_kernel_ms": {k: v["ms"] / rounds for k, v in prof.items()},
           "kernel": batch.kernel_name, "plan_stats": batch.plan_stats(), "cpu_reference": cpu, "max_rel_err_lnl_vs_reference": err,
           "speedup_e2e_vs_reference_all_cores": (n_loci * rounds / secs) / cpu["moves_per_sec"] if cpu else None}
    batch.destroy()
    eng.close()
    return rec


if __name__ == "__main__":
    tips = [int(a) for a in sys.argv[1:]] or [8, 16, 48]
    for T in tips:
        n = 10000 if T <= 16 else (4000 if T <= 48 else 1000)
        print(json.dumps(run(T, n)), flush=True)
