"""A few device-resident full passes of one config, for ncu.  Usage: profile_step.py config [n_loci] [runs] [scaling]"""
import sys

sys.path.insert(0, ".")
from bpp_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else synth.CONFIGS[cfg]["n_loci"]
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scaling = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
lg = None
if synth.CONFIGS[cfg].get("states", 4) == 20:
    import bench
    lg = bench.lg_tables()
w = synth.make_config(cfg, n_loci=n, scaling=scaling, lg=lg)
eng = engine.Engine(0)
loci, trees = engine.load_workload(eng, w)
batch = engine.Batch(eng, loci)
batch.set_waves(1)
batch.stage(trees.full_pass_step())
for _ in range(runs):
    batch.run()
out, tot = batch.collect()
print("ok", tot)
