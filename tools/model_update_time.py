"""Cost of a model-changing move over all loci (propose_alpha: new category rates; propose_qrates: new Q) followed by
a full pass, batched model sync + device eigen vs. the per-locus path.  Usage: model_update_time.py [config] [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bpp_b200 import engine, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
w = synth.make_config(cfg, n_loci=n)
eng = engine.Engine(0)
loci, trees = engine.load_workload(eng, w)
batch = engine.Batch(eng, loci)
step = trees.full_pass_step()
pstep, holders = engine.pin_step(step)
prep = batch.prepare(pstep)
out = np.zeros(w.n_loci)
batch.full_pass(pstep)
rng = np.random.default_rng(1)
for mode, thresh in (("per-locus copies, host eigen", "100000000"), ("one blob, device eigen", "8")):
    os.environ["BPPGPU_MODEL_BATCH_MIN"] = thresh
    for what in ("rates", "qrates"):
        ts = tp = 0.0
        K = 5
        for it in range(K):
            t0 = time.perf_counter()
            if what == "rates":
                r = np.sort(rng.uniform(0.1, 3.0, size=w.rate_cats))
                for l in loci:
                    l.set_category_rates(r)
            else:
                for i, l in enumerate(loci):
                    l.set_subst_params(rng.uniform(0.5, 1.5, size=6))
            t1 = time.perf_counter()
            batch.stage(prep); batch.run(); batch.collect(out)
            t2 = time.perf_counter()
            ts += t1 - t0; tp += t2 - t1
            print('   it %d: %.2f ms' % (it, 1e3 * (t2 - t1)))
        print("%-30s %-7s: host setters %.2f ms, sync + full pass %.2f ms (%d loci)" % (mode, what, 1e3 * ts / K, 1e3 * tp / K, n), flush=True)
