"""CPU, world_size 2, gloo: the N>1 host logic -- contiguous loci partition (threads.c:234-263) and
the all-reduce of the per-rank lnL sums (threads.c:583-590).  Per-locus values come from the oracle
(the GPU kernels are covered by the -m gpu tests; there is no collective on the data path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import F, ROOT, char_map, synth
from bpp_b200 import shard


def test_locus_range_matches_load_balance_none():
    for n, world in [(10, 3), (5, 5), (10000, 8), (50000, 8), (7, 2), (3, 4)]:
        got = [shard.locus_range(n, world, r) for r in range(world)]
        per, rem = divmod(n, world)
        start = 0
        for r, (first, count) in enumerate(got):
            assert first == start and count == per + (1 if r < rem else 0)
            start += count
        assert start == n


def _worker(rank, world, port, n_loci, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload("shard", n_loci=n_loci, tips=5, sites=23, states=4, rate_cats=4, model="GTR", seed=3)
    first, count = shard.locus_range(n_loci, world, rank)
    cm = char_map(4)
    local = sum(F.locus_from_workload(w, i, cm).full_pass() for i in range(first, first + count))
    total = shard.allreduce_sum(torch.tensor([local, float(count)], dtype=torch.float64))
    out[rank] = (local, float(total[0]), float(total[1]))
    dist.destroy_process_group()


def test_two_rank_lnl_sum_allreduce():
    n_loci, world = 7, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_loci, out), nprocs=world, join=True)
    w = synth.make_workload("shard", n_loci=n_loci, tips=5, sites=23, states=4, rate_cats=4, model="GTR", seed=3)
    cm = char_map(4)
    ref = [F.locus_from_workload(w, i, cm).full_pass() for i in range(n_loci)]
    assert out[0][2] == out[1][2] == n_loci                      # every locus owned exactly once
    assert out[0][1] == out[1][1]                                # same reduced value on every rank
    assert abs(out[0][1] - sum(ref)) <= 1e-12 * abs(sum(ref))
    assert abs(out[0][0] + out[1][0] - out[0][1]) <= 1e-12 * abs(out[0][1])


def test_zigzag_assignment_matches_reference_deal_and_balances():
    """load_balance_zigzag, threads.c:265-353: ascending loads dealt 0,1,2,2,1,0,0,1,2,..."""
    from bpp_b200 import shard
    loads = [50, 10, 40, 20, 30, 60, 70]            # ascending order of indices: 1,3,4,2,0,5,6
    got = shard.zigzag_assignment(loads, 3)
    assert got == [[1, 5, 6], [3, 0], [4, 2]]
    import random
    rng = random.Random(5)
    loads = [rng.randint(4, 40) * rng.randint(100, 3000) for _ in range(1000)]
    parts = shard.zigzag_assignment(loads, 8)
    assert sorted(i for p in parts for i in p) == list(range(1000))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    work = [sum(loads[i] for i in p) for p in parts]
    assert max(work) / min(work) < 1.02
    assert shard.zigzag_assignment([], 4) == [[], [], [], []]
    assert shard.zigzag_assignment([7, 3], 1) == [[1, 0]]
