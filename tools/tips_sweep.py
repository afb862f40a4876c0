"""Device-resident full-pass time vs. tips per locus (one-chunk fast path up to 16 tips, chunk-by-chunk fast path up to
128 tips, general walker beyond).
Usage: tips_sweep.py [rate_cats] [model] [tips,tips,...]"""
import sys

sys.path.insert(0, ".")
from bpp_b200 import engine, synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = sys.argv[2] if len(sys.argv) > 2 else "GTR"
eng = engine.Engine(0)
TIPS = tuple(int(t) for t in sys.argv[3].split(",")) if len(sys.argv) > 3 else (8, 16, 17, 24, 32, 48, 64, 96, 128)
for tips in TIPS:
    n = max(200, 32000 // tips)
    rates = None if R in (1, 4) else [0.2 + 1.6 * k / (R - 1) for k in range(R)]
    w = synth.make_workload("sweep", n_loci=n, tips=tips, sites=1000, states=4, rate_cats=R, model=model, seed=3, rates=rates)
    loci, trees = engine.load_workload(eng, w)
    batch = engine.Batch(eng, loci)
    batch.set_waves(1)
    batch.stage(trees.full_pass_step())
    for _ in range(3):
        batch.run()
    K = 20
    eng.reset_profile(); eng.set_profiling(True)
    batch.timer_start()
    for _ in range(K):
        batch.run()
    ms = batch.timer_stop_ms() / K
    prof = {k: round(v["ms"] / max(1, v["launches"]), 3) for k, v in eng.profile().items() if v["launches"]}
    eng.set_profiling(False)
    node_updates = n * (tips - 1)
    st = batch.plan_stats() or {}
    print("R=%d %s tips=%2d loci=%5d: %.3f ms/pass  %.1f M node-updates/s  (%.2f GB/s of CLV writes)  %s %s walker=%s chunks=%s slots=%s smem=%s" % (
        R, model, tips, n, ms, node_updates / ms / 1e3, node_updates * 1000 * R * 32 / ms / 1e6, batch.kernel_name, prof,
        st.get("walker"), st.get("max_chunks"), st.get("slots"), st.get("smem_bytes")), flush=True)
    batch.destroy()
    for l in loci:
        l.destroy()
