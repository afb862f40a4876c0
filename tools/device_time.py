"""Device-resident step time per kernel for one config.  Usage: device_time.py config [n_loci] [scaling]"""
import os
import sys

sys.path.insert(0, ".")
from bpp_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else synth.CONFIGS[cfg]["n_loci"]
scaling = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
lg = None
if synth.CONFIGS[cfg].get("states", 4) == 20:
    import bench
    lg = bench.lg_tables()
w = synth.make_config(cfg, n_loci=n, scaling=scaling, lg=lg)
eng = engine.Engine(0, math=os.environ.get("BPPGPU_MATH", "exact"))
loci, trees = engine.load_workload(eng, w)
batch = engine.Batch(eng, loci)
batch.set_waves(1)
batch.stage(trees.full_pass_step())
for _ in range(5):
    batch.run()
batch.synchronize()
K = 100
eng.reset_profile(); eng.set_profiling(True)
batch.timer_start()
for _ in range(K):
    batch.run()
ms = batch.timer_stop_ms() / K
prof = eng.profile()
out, tot = batch.collect()
print("%s n=%d scaling=%d: %.4f ms/step  %.3f M evals/s  %s  sum %.6f  %s" % (
    cfg, n, scaling, ms, n / ms / 1e3, {k: round(v["ms"] / max(1, v["launches"]), 4) for k, v in prof.items()}, tot,
    batch.kernel_name), flush=True)
