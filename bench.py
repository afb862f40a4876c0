#!/usr/bin/env python
"""bench.py -- locus-lnL evaluations/s of the full-tree likelihood pass (BASELINE.json metric).

A "step" is one full-tree pass over every locus of the rank: all 2T-2 P-matrices, all T-1 inner
CLVs (post-order) and the root log-likelihood of every locus -- the unit of work of the
reference's mixing move (prop_mixing.c:52-220) -- followed for N > 1 by one NCCL all-reduce of the
lnL sums through the C-ABI (bppgpu_batch_allreduce_lnl_sum; threads.c:583-590).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config config2|config3|config4|config5]
  python bench.py --impl reference ...     # the reference's own AVX2 path on the host cores

Default workload (BASELINE.json configs): N = 1 -> config 3 (10 000 loci x 1 000 patterns, GTR+G4, 16 tips: the
largest single-GPU configuration), with configs 2 and 4 and the config-5 shard shape as sub-records of the same
JSON line; N > 1 -> the config-5 shard shape, 6 250 loci x 2 000 patterns per GPU, weak scaling, so that N = 8 IS
config 5 (50 000 loci).

value  : inputs resident in HBM (staged once), CUDA-event timed on the launching stream.
e2e    : the same step through the public C-ABI calls bppgpu_batch_stage / run / collect with HOST buffers
         (pinned, H2D + kernels + D2H inside the timed region, wall clock).
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
CONFIG5_GPUS = 8           # config 5 is defined on 8 GPUs: 50 000 loci / 8 = 6 250 per GPU


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


def ncu_traffic(config, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu
    capture of the SAME kernel instantiation at the config's full size (profiles/ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p)).get(config)
    except Exception:
        rec = None
    if not rec or rec.get("kernel") != kernel:
        return None, None
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("source")


def default_loci(config, world):
    n = synth.CONFIGS[config]["n_loci"]
    return n // CONFIG5_GPUS if config == "config5" else n


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


def ref_set(w):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import ref_set_from_workload
    return ref_set_from_workload(w)


def cpu_reference_rate(w, threads, target_seconds, max_loci, want_1core=True):
    """Time the reference's AVX2 path (oracle/_ref, unmodified bpp v4.8.7) on a bounded sample:
    full passes over the first `max_loci` loci with `threads` pthreads (static partition,
    threads.c:234-263).  Returns (loci/s, loci/s on one core, lnl of the sample, description)."""
    ws = w.subset(max_loci)
    rs = ref_set(ws)
    secs, lnl = rs.full_pass_all(0, ws.n_loci, threads, 1)          # warm-up + calibration
    passes = int(max(2, min(200, target_seconds / max(secs, 1e-6))))
    secs, lnl = rs.full_pass_all(0, ws.n_loci, threads, passes)
    rate = ws.n_loci * passes / secs
    rate1 = None
    if want_1core:
        n1 = max(1, ws.n_loci // max(1, threads))
        s1, _ = rs.full_pass_all(0, n1, 1, max(1, passes // 4))
        rate1 = n1 * max(1, passes // 4) / s1
    rs.close()
    return rate, rate1, lnl[:ws.n_loci].copy(), "%d loci x %d passes, %d pthreads" % (ws.n_loci, passes, threads)


def workload_config(w, name, world, sample=None):
    c = {"workload": "%s: %d loci/GPU x %d patterns, %d tips, %s, %d rate cats, %d states, scaling=%d"
                     % (name, w.n_loci, w.sites, w.tips, w.model, w.rate_cats, w.states, int(w.scaling)),
         "config_name": name, "loci_per_gpu": w.n_loci, "patterns": w.sites, "tips": w.tips,
         "model": w.model, "rate_cats": w.rate_cats, "states": w.states, "scaling": int(w.scaling),
         "sharding": "loci statically sharded across %d GPU(s), one NCCL all-reduce of the lnL sum per step" % world,
         "l2_policy": "inputs larger than L2: every pass writes %.2f GB of CLVs (L2 = 126 MB), no flush needed"
                      % (w.n_loci * w.b_min() / 1e9),
         "seed": synth.SEED}
    if name == "config5":
        c["config5_note"] = ("config 5 = 50 000 loci sharded over %d GPUs = %d loci per GPU; weak scaling in N, "
                             "N = %d is config 5 itself" % (CONFIG5_GPUS, w.n_loci, CONFIG5_GPUS))
    if sample is not None:
        c["reference_sample_loci"] = sample
    return c


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.config or ("config3" if world == 1 else "config5")
    w = build_workload(name, args.loci_per_gpu_sample, 0, args.scaling)
    threads = os.cpu_count() or 1
    rs = ref_set(w)
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
            "data": "synthetic", "config": workload_config(w, name, world, sample=w.n_loci),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "%d loci per step (bounded sample of the workload), %d pthreads, "
                                       "oracle/_ref = unmodified bpp v4.8.7, --arch avx2 path" % (w.n_loci, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Dist:
    """torch.distributed for the plumbing (rendezvous, barrier, max over ranks); the data-path sum goes through
    the C-ABI communicator."""

    def __init__(self, world, local_rank):
        self.world, self.local_rank = world, local_rank
        import torch
        self.torch = torch
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                    "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def share_bytes(self, payload):
        """rank 0's bytes to every rank"""
        if not self.dist:
            return payload
        t = self.torch.zeros(len(payload), dtype=self.torch.uint8, device="cuda")
        if self.dist.get_rank() == 0:
            t.copy_(self.torch.tensor(list(payload), dtype=self.torch.uint8))
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def measure(name, n_loci, args, rank, world, local_rank, D, steps, warmup, cpu_seconds, parity_loci, want_cpu,
            sampler):
    """One config on this rank's GPU: parity gate, `value`, `e2e`, roofline.  Returns the record (all ranks)."""
    from bpp_b200 import engine

    t_setup = time.perf_counter()
    w = build_workload(name, n_loci, rank, bool(args.scaling))
    eng = engine.Engine(local_rank, math=args.math)
    loci, trees = engine.load_workload(eng, w)
    batch = engine.Batch(eng, loci)
    step = trees.full_pass_step()
    t_setup = time.perf_counter() - t_setup
    comm = None
    if world > 1:
        uid = D.share_bytes(engine.Comm.unique_id() if rank == 0 else bytes(128))
        comm = engine.Comm(eng, world, rank, uid)

    # ---- parity gate on EVERY rank: GPU lnL vs the compiled reference on a sample of this
    #      rank's own loci, before any timing
    lnl, total = batch.full_pass(step)
    parity, cpu = None, None
    from oracle import refbind
    threads_all = os.cpu_count() or 1
    if not args.no_cpu_baseline and refbind.available():
        if want_cpu and rank == 0 and world == 1:
            rate, rate1, ref_lnl, desc = cpu_reference_rate(w, threads_all, cpu_seconds, parity_loci)
            cpu = {"value": rate, "unit": UNIT, "cores": threads_all, "kind": "reference",
                   "sample": desc + "; oracle/_ref = unmodified bpp v4.8.7 AVX2 path on the GPU box's host cores",
                   "value_1core": rate1}
        else:
            ws = w.subset(max(16, parity_loci // max(1, world)))
            rs = ref_set(ws)
            _, ref_lnl = rs.full_pass_all(0, ws.n_loci, max(1, threads_all // world), 1)
            ref_lnl = ref_lnl[:ws.n_loci].copy()
            rs.close()
        err = float(np.max(np.abs(lnl[:ref_lnl.size] - ref_lnl) / np.abs(ref_lnl)))
        worst = D.reduce(err, "max")
        parity = {"max_rel_err_lnl_vs_reference": worst, "loci_checked_per_rank": int(ref_lnl.size),
                  "ranks_checked": world, "bar": 1e-10, "pass": bool(worst <= 1e-10)}
    elif not args.no_cpu_baseline:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    def allreduce():
        if comm is not None:
            batch.allreduce_lnl_sum(comm)

    # ---- N > 1: this rank's own device-timed step without the collective (what N = 1 of the same shape does)
    solo_ms = None
    batch.stage(step)
    if world > 1:
        for _ in range(warmup):
            batch.run()
        D.barrier()
        batch.timer_start()
        for _ in range(steps):
            batch.run()
        solo_ms = D.reduce(batch.timer_stop_ms() / steps, "max")

    # ---- value: inputs resident in HBM
    for _ in range(warmup):
        batch.run()
        allreduce()
    D.barrier()
    eng.reset_profile()
    eng.set_profiling(True)
    l0 = eng.launch_count
    t_region0 = time.perf_counter()
    batch.timer_start()
    for _ in range(steps):
        batch.run()
        allreduce()
    ms_total = batch.timer_stop_ms()
    kname = batch.kernel_name                 # the instantiation the timed steps launched (the later stages replan)
    D.barrier()
    launches = eng.launch_count - l0
    prof = eng.profile()
    eng.set_profiling(False)
    ms_step = D.reduce(ms_total, "max") / steps
    value = world * w.n_loci / (ms_step / 1000.0)

    # ---- e2e: public C-ABI calls with host buffers, H2D + kernels + D2H per step, wall clock.
    #      A step is a whole-tree proposal (the mixing move): the lists are resident, the host flips the indices on
    #      the device (bppgpu_batch_flip_indices) and sends what changed -- the branch lengths, from pinned memory
    pstep, holders = engine.pin_step(step)
    n = w.n_loci
    n_mat, n_op = int(step[0].sum()), int(step[3].sum())
    out_buf = np.zeros(w.n_loci)
    out_sum = 0.0
    batch.stage(pstep)
    batch.run()
    for _ in range(4):                      # both index parities planned and cached
        batch.flip_indices()
        batch.set_branch_lengths(pstep[2])
        batch.run()
    batch.collect(out_buf)
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        batch.flip_indices()
        batch.set_branch_lengths(pstep[2])
        batch.run()
        allreduce()
        out_lnl, out_sum = batch.collect(out_buf)
    D.barrier()
    e2e_s = D.reduce(time.perf_counter() - t0, "max")
    e2e_value = world * w.n_loci * steps / e2e_s
    h2d = n_mat * 8
    d2h = (n + 1) * 8
    if steps % 2 == 1:
        batch.flip_indices()
    # ---- the same with every array of the step staged again (new traversals every step: nothing cached)
    for _ in range(2):
        batch.full_pass(pstep)
    D.barrier()
    t0 = time.perf_counter()
    prep = batch.prepare(pstep)
    for _ in range(steps):
        batch.stage(prep)
        batch.run()
        allreduce()
        out_lnl, out_sum = batch.collect(out_buf)
    D.barrier()
    e2e_restage_s = D.reduce(time.perf_counter() - t0, "max")
    t_region1 = time.perf_counter()
    for h in holders:
        h.free()

    # ---- roofline of the dominant kernel (tree kernel), live CUDA-event timing on the launching stream
    peak, peak_src = measured_peak()
    tree_ms = prof["tree"]["ms"] / max(1, steps)          # per step (one launch per step on this path)
    b_pass, b_min = w.b_pass(), w.b_min()
    achieved_min = b_min * w.n_loci / (tree_ms / 1000.0) / 1e9 if tree_ms > 0 else None
    achieved_can = b_pass * w.n_loci / (tree_ms / 1000.0) / 1e9 if tree_ms > 0 else None
    traffic, traffic_src = (None, None)
    if not args.scaling and n_loci == default_loci(name, world):
        traffic, traffic_src = ncu_traffic(name, kname)
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved_min, "peak": peak, "unit": "GB/s",
                "frac": achieved_min / peak if achieved_min else None, "peak_source": peak_src,
                "algorithmic_bytes_per_locus": b_min,
                "note": "achieved = compulsory bytes of the tree-fused pass (SURVEY.md 8d B_min: every inner CLV "
                        "written once, packed tips, weights, P-matrices) x loci per launch / tree-kernel CUDA-event "
                        "time; *_canonical uses SURVEY.md 8d's node-streaming B_pass (what the reference's per-node "
                        "kernels move), which a fused kernel exceeds",
                "achieved_canonical": achieved_can, "frac_canonical": achieved_can / peak if achieved_can else None,
                "canonical_bytes_per_locus": b_pass,
                "kernel_ms": tree_ms, "kernel_share_of_step": tree_ms / ms_step if ms_step else None,
                "per_kernel_ms": {k: v["ms"] / max(1, steps) for k, v in prof.items()},
                "traffic": traffic, "traffic_source": traffic_src,
                "frac_ncu": (traffic / (tree_ms / 1000.0) / 1e9 / peak) if (traffic and tree_ms > 0) else None}

    rec = {"value": value, "ms_per_step": ms_step, "config": workload_config(w, name, world),
           "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1000.0 * e2e_s / steps,
                   "what": "one mixing-move step through the public calls: bppgpu_batch_flip_indices (device-side "
                           "SWAP_* of every node) + bppgpu_batch_set_branch_lengths (pinned host array, H2D) + "
                           "bppgpu_batch_run + all-reduce + bppgpu_batch_collect (n+1 doubles D2H); op lists and "
                           "tip states are device-resident",
                   "restage": {"value": world * w.n_loci * steps / e2e_restage_s,
                               "ms_per_step": 1000.0 * e2e_restage_s / steps,
                               "h2d_bytes_per_step": n_mat * 12 + n_op * 32 + n * 8,
                               "what": "bppgpu_batch_stage+run+collect: every array of the step (branch lengths, "
                                       "P-matrix indices, pruning ops, roots) uploaded and planned again each step"}},
           "gpu_launches": int(launches), "dataset_passes_per_sec": 1000.0 / ms_step, "setup_seconds": t_setup,
           "lnl_sum_check": float(out_sum), "hbm_bytes_allocated": eng.bytes_allocated,
           "clocks": sampler.summary(t_region0, t_region1)}
    if solo_ms is not None:
        rec["n1_same_shape"] = {"value": w.n_loci / (solo_ms / 1000.0), "ms_per_step": solo_ms,
                                "what": "slowest rank's device-timed step of the same per-GPU workload without the "
                                        "all-reduce, measured in this run: the N = 1 point of THIS shape "
                                        "(the N = 1 default of bench.py is config 3)"}
        rec["nccl"] = {"allreduces": comm.calls, "via": "bppgpu_batch_allreduce_lnl_sum (C-ABI, libnccl via dlopen)"}
    if comm is not None:
        comm.destroy()
    batch.destroy()
    eng.close()           # frees every locus and the arena
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=list(synth.CONFIGS),
                    help="default: config3 on one GPU, the config-5 shard shape on several")
    ap.add_argument("--loci", type=int, default=None, help="loci per GPU (default: the config's count; "
                    "config5: 50 000 / 8 per GPU)")
    ap.add_argument("--scaling", type=int, default=0)
    ap.add_argument("--math", default="exact", choices=["exact", "fma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (configs 2, 4, 5-shard at N = 1)")
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

    D = Dist(world, local_rank)
    name = args.config or ("config3" if world == 1 else "config5")
    n_loci = args.loci if args.loci is not None else default_loci(name, world)

    sampler = ClockSampler(local_rank)
    sampler.start()
    main_rec = measure(name, n_loci, args, rank, world, local_rank, D, args.steps, args.warmup,
                       cpu_seconds=12.0, parity_loci=2048, want_cpu=True, sampler=sampler)
    subs = {}
    if world == 1 and args.config is None and args.loci is None and not args.no_sub:
        for sub in ("config2", "config4", "config5"):
            try:
                subs[sub if sub != "config5" else "config5_shard"] = measure(
                    sub, default_loci(sub, 1), args, rank, world, local_rank, D, args.steps, args.warmup,
                    cpu_seconds=5.0, parity_loci=1024, want_cpu=True, sampler=sampler)
            except Exception as ex:      # a sub-record must not cost the headline line
                subs[sub] = {"error": repr(ex)}
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank == 0:
        line = {"metric": METRIC, "value": main_rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "math": args.math}
        for k, v in main_rec.items():
            if k not in line:
                line[k] = v
        if subs:
            line["sub"] = subs
        print(json.dumps(line), flush=True)
    D.close()


if __name__ == "__main__":
    main()
