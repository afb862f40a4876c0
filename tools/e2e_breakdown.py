"""Wall-clock breakdown of the e2e step (stage / run / collect) for a few wave counts.  GPU box only."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bpp_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else synth.CONFIGS[cfg]["n_loci"]
w = synth.make_config(cfg, n_loci=n)
eng = engine.Engine(0)
loci, trees = engine.load_workload(eng, w)
batch = engine.Batch(eng, loci)
step = trees.full_pass_step()
pstep, holders = engine.pin_step(step)
prep = batch.prepare(pstep)
out = np.zeros(w.n_loci)
K = 200
for waves in (1, 2, 3):
    batch.set_waves(waves)
    for _ in range(5):
        batch.stage(prep); batch.run(); batch.collect(out)
    ts = tr = tc = 0.0
    t00 = time.perf_counter()
    for _ in range(K):
        t0 = time.perf_counter(); batch.stage(prep)
        t1 = time.perf_counter(); batch.run()
        t2 = time.perf_counter(); batch.collect(out)
        t3 = time.perf_counter()
        ts += t1 - t0; tr += t2 - t1; tc += t3 - t2
    tot = time.perf_counter() - t00
    # device span: event on the idle batch stream before stage -> event after the finish kernel
    span = 0.0
    for _ in range(50):
        batch.timer_start(); batch.stage(prep); batch.run(); span += batch.timer_stop_ms(); batch.collect(out)
    print("waves %d: step %.4f ms  (stage %.4f  run %.4f  collect %.4f)  %.2f M evals/s   device span %.4f ms" %
          (waves, 1e3 * tot / K, 1e3 * ts / K, 1e3 * tr / K, 1e3 * tc / K, w.n_loci * K / tot / 1e6, span / 50), flush=True)
# device-only reference
batch.set_waves(1)
batch.stage(prep)
for _ in range(5):
    batch.run()
batch.timer_start()
for _ in range(K):
    batch.run()
print("device only: %.4f ms/step" % (batch.timer_stop_ms() / K))
