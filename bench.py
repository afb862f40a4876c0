#!/usr/bin/env python
"""bench.py -- locus-lnL evaluations/s of the full-tree likelihood pass (BASELINE.json metric).

A "step" is one full-tree pass over every locus of the rank: all 2T-2 P-matrices, all T-1 inner
CLVs (post-order) and the root log-likelihood of every locus -- the unit of work of the
reference's mixing move (prop_mixing.c:52-220) -- followed for N > 1 by one all-reduce of the
lnL sums (NCCL through torch.distributed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config config2|config3|config4|config5]
  python bench.py --impl reference ...     # the reference's own AVX2 path on the host cores

value  : inputs resident in HBM (staged once), CUDA-event timed on the launching stream.
e2e    : the same step through the public C-ABI call bppgpu_batch_full_pass with HOST buffers
         (pinned staging, one H2D, kernels, one D2H inside the timed region, wall clock).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bpp_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
METRIC = "locus_lnL_evals_per_sec_full_tree"
UNIT = "locus-lnL evals/s"


def lg_tables():
    d = np.load(os.path.join(GOLDEN, "lg_model.npz"))
    return d["rates"], d["freqs"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(name, n_loci, rank, scaling):
    lg = lg_tables() if name == "config4" else None
    return synth.make_config(name, n_loci=n_loci, scaling=scaling, seed=synth.SEED + 1000 * rank, lg=lg)


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU through NVML while the bench runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks", 0x100: "display_clocks",
               0x10: "sync_boost"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []           # (t, sm_mhz, power_w, reasons_mask, util)
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as ex:     # pragma: no cover
            self.err = repr(ex)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                ut = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.perf_counter(), sm, pw, rs, ut))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvml unavailable"}
        sel = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        mask = 0
        for s in sel:
            mask |= s[3]
        reasons = sorted(n for b, n in self.REASONS.items() if mask & b)
        return {"sm_mhz": float(np.median([s[1] for s in sel])), "sm_max_mhz": self.max_sm,
                "reasons": reasons, "power_w_max": max(s[2] for s in sel), "samples": len(sel)}


class DevScalar:
    """A device double exposed to torch through __cuda_array_interface__ (for the all-reduce)."""

    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def cpu_reference_rate(w, threads, target_seconds=12.0, max_loci=2048):
    """Time the reference's AVX2 path (oracle/_ref, unmodified bpp v4.8.7) on a bounded sample:
    full passes over the first `max_loci` loci with `threads` pthreads (static partition,
    threads.c:234-263).  Returns (loci/s, loci/s on one core, lnl of the sample, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import ref_set_from_workload
    ws = w.subset(max_loci)
    rs = ref_set_from_workload(ws)
    secs, lnl = rs.full_pass_all(0, ws.n_loci, threads, 1)          # warm-up + calibration
    passes = int(max(2, min(200, target_seconds / max(secs, 1e-6))))
    secs, lnl = rs.full_pass_all(0, ws.n_loci, threads, passes)
    rate = ws.n_loci * passes / secs
    n1 = max(1, ws.n_loci // max(1, threads))
    s1, _ = rs.full_pass_all(0, n1, 1, max(1, passes // 4))
    rate1 = n1 * max(1, passes // 4) / s1
    rs.close()
    return rate, rate1, lnl[:ws.n_loci].copy(), "%d loci x %d passes, %d pthreads" % (ws.n_loci, passes, threads)


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = build_workload(args.config, args.loci_per_gpu_sample, 0, args.scaling)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import ref_set_from_workload
    threads = os.cpu_count() or 1
    rs = ref_set_from_workload(w)
    for _ in range(args.warmup):
        rs.full_pass_all(0, w.n_loci, threads, 1)
    t = 0.0
    for _ in range(args.steps):
        s, _ = rs.full_pass_all(0, w.n_loci, threads, 1)
        t += s
    rs.close()
    value = w.n_loci * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(w, args, world, sample=w.n_loci),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "%d loci per step (bounded sample of the workload), %d pthreads, "
                                       "oracle/_ref = unmodified bpp v4.8.7, --arch avx2 path" % (w.n_loci, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(w, args, world, sample=None):
    c = {"workload": "%s: %d loci/GPU x %d patterns, %d tips, %s, %d rate cats, %d states, scaling=%d"
                     % (args.config, w.n_loci, w.sites, w.tips, w.model, w.rate_cats, w.states, int(w.scaling)),
         "config_name": args.config, "loci_per_gpu": w.n_loci, "patterns": w.sites, "tips": w.tips,
         "model": w.model, "rate_cats": w.rate_cats, "states": w.states, "scaling": int(w.scaling),
         "sharding": "loci statically sharded across %d GPU(s), one all-reduce of the lnL sum per step" % world,
         "l2_policy": "inputs larger than L2: every pass writes %.2f GB of CLVs (L2 = 126 MB), no flush needed"
                      % (w.n_loci * w.b_min() / 1e9),
         "seed": synth.SEED}
    if sample is not None:
        c["reference_sample_loci"] = sample
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="config2", choices=list(synth.CONFIGS))
    ap.add_argument("--loci", type=int, default=None, help="loci per GPU (default: the config's count; "
                    "config5 is sharded over the GPUs)")
    ap.add_argument("--scaling", type=int, default=0)
    ap.add_argument("--math", default="exact", choices=["exact", "fma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--loci-per-gpu-sample", type=int, default=2048,
                    help="loci per step of the --impl reference arm (bounded sample)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bpp_b200 import engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_loci = args.loci
    if n_loci is None:
        n_loci = synth.CONFIGS[args.config]["n_loci"]
        if args.config == "config5":
            n_loci = n_loci // world
    scaling_kind = "strong" if (args.config == "config5" and args.loci is None) else "weak"

    t_setup = time.perf_counter()
    w = build_workload(args.config, n_loci, rank, bool(args.scaling))
    eng = engine.Engine(local_rank, math=args.math)
    loci, trees = engine.load_workload(eng, w)
    batch = engine.Batch(eng, loci)
    step = trees.full_pass_step()
    t_setup = time.perf_counter() - t_setup

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- parity gate: GPU lnL vs the reference on a sample, before any timing
    lnl, total = batch.full_pass(step)
    parity = None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT))
        from oracle import refbind
        if refbind.available():
            threads = os.cpu_count() or 1
            rate, rate1, ref_lnl, desc = cpu_reference_rate(w, threads)
            err = float(np.max(np.abs(lnl[:ref_lnl.size] - ref_lnl) / np.abs(ref_lnl)))
            parity = {"max_rel_err_lnl_vs_reference": err, "loci_checked": int(ref_lnl.size), "bar": 1e-10,
                      "pass": bool(err <= 1e-10)}
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": desc + "; oracle/_ref = unmodified bpp v4.8.7 AVX2 path on the GPU box's host cores",
                   "value_1core": rate1}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    stream = torch.cuda.ExternalStream(batch.stream, device=local_rank)
    sum_t = torch.as_tensor(DevScalar(batch.lnl_sum_dev), device="cuda:%d" % local_rank) if world > 1 else None

    def allreduce():
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(sum_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    batch.stage(step)
    for _ in range(args.warmup):
        batch.run()
        allreduce()
    barrier()
    eng.reset_profile()
    eng.set_profiling(True)
    l0 = eng.launch_count
    t_region0 = time.perf_counter()
    batch.timer_start()
    for _ in range(args.steps):
        batch.run()
        allreduce()
    ms_total = batch.timer_stop_ms()
    barrier()
    t_region1 = time.perf_counter()
    launches = eng.launch_count - l0
    prof = eng.profile()
    eng.set_profiling(False)
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    ms_step = ms_total / args.steps
    value = world * w.n_loci / (ms_step / 1000.0) if scaling_kind == "weak" else world * w.n_loci / (ms_step / 1000.0)

    # ---- e2e: public C-ABI call with host buffers, H2D + kernels + D2H per step, wall clock
    # the step's host inputs live in pinned host memory, as a caller that wants throughput would keep them
    pstep, holders = engine.pin_step(step)
    for _ in range(3):
        batch.full_pass(pstep)
    barrier()
    t0 = time.perf_counter()
    prep = batch.prepare(pstep)
    out_buf = np.zeros(w.n_loci)
    for _ in range(args.steps):
        batch.stage(prep)
        batch.run()
        allreduce()
        out_lnl, out_sum = batch.collect(out_buf)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_region2 = time.perf_counter()
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    e2e_value = world * w.n_loci * args.steps / e2e_s
    n = w.n_loci
    n_mat, n_op = int(step[0].sum()), int(step[3].sum())
    h2d = 2 * (n + 1) * 4 + n_mat * 12 + n_op * 32 + n * 8
    d2h = (n + 1) * 8

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary(t_region0, t_region2)

    # ---- roofline of the dominant kernel (tree kernel), live CUDA-event timing
    peak, peak_src = measured_peak()
    tree_ms = prof["tree"]["ms"] / max(1, prof["tree"]["launches"])
    b_pass, b_min = w.b_pass(), w.b_min()
    achieved = b_pass * w.n_loci / (tree_ms / 1000.0) / 1e9 if tree_ms > 0 else None
    achieved_min = b_min * w.n_loci / (tree_ms / 1000.0) / 1e9 if tree_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "tree_kernel_s4" if w.states == 4 else
                ("tree_kernel_s20c" if w.states == 20 and not args.scaling else
                 ("tree_kernel_s20" if w.states == 20 else "tree_kernel_generic")),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_locus": b_pass,
                "note": "achieved = canonical node-streaming bytes B_pass (SURVEY.md 8d) x loci per launch / "
                        "tree-kernel CUDA-event time; the tree-fused kernel moves only the compulsory bytes, so "
                        "frac can exceed 1 -- achieved_compulsory is the bytes it really has to move",
                "achieved_compulsory": achieved_min, "frac_compulsory": achieved_min / peak if achieved_min else None,
                "compulsory_bytes_per_locus": b_min,
                "kernel_ms": tree_ms, "kernel_share_of_step": tree_ms / ms_step if ms_step else None,
                "per_kernel_ms": {k: (v["ms"] / max(1, v["launches"])) for k, v in prof.items()},
                "traffic": NCU_TRAFFIC.get(args.config) if not args.scaling and args.loci is None else None,
                "traffic_source": NCU_TRAFFIC_SOURCE.get(args.config)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": scaling_kind, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(w, args, world), "math": args.math,
                "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1000.0 * e2e_s / args.steps,
                        "what": "bppgpu_batch_stage+run+collect with host arrays in pinned memory: branch lengths, "
                                "P-matrix indices, pruning ops and root indices go H2D every step, n+1 doubles come "
                                "back; tip states are device-resident like the reference's tip CLVs (set once at "
                                "locus creation)"},
                "gpu_launches": int(launches), "clocks": clocks,
                "dataset_passes_per_sec": 1000.0 / ms_step, "setup_seconds": t_setup,
                "lnl_sum_check": float(out_sum), "hbm_bytes_allocated": eng.bytes_allocated}
        print(json.dumps(line), flush=True)

    batch.destroy()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per tree-kernel launch from the committed ncu capture
# (profiles/), filled in after each profiling pass; None = not captured for that config yet.
# Bytes per launch at the config's full size, scaling off.
NCU_TRAFFIC = {
    "config2": 2.180698e9 + 0.131435e9,      # tree_kernel_s4<1,exact,2>
    "config3": 19.141006e9 + 0.325830e9,     # tree_kernel_s4<4,exact,4>
    "config4": 4.476720e9 + 0.430410e9,      # tree_kernel_s20c<4>
}
NCU_TRAFFIC_SOURCE = {
    "config2": "profiles/r1_tree_v10_config2_ncu_summary.txt",
    "config3": "profiles/r1_tree_final_config3_ncu_summary.txt",
    "config4": "profiles/r1_s20c_v1_config4_ncu_summary.txt",
}


if __name__ == "__main__":
    main()
