"""Multi-GPU sharding of the likelihood path: loci are independent, so each rank owns a contiguous
range of loci and the only exchange is one all-reduce of the lnL sums per step.

Mirrors the reference's thread partition load_balance_none (threads.c:234-263: loci/threads each,
the remainder handed out one by one to the first threads) and the master's post-barrier sum of the
per-thread lnacceptance (threads.c:583-590)."""


def locus_range(n_loci, world, rank):
    """(first, count) of the loci owned by `rank` out of `world` ranks."""
    per, rem = divmod(n_loci, world)
    first = rank * per + min(rank, rem)
    return first, per + (1 if rank < rem else 0)


def zigzag_assignment(loads, world):
    """Ragged data: the reference's load_balance_zigzag (threads.c:265-353).  Loci are sorted by load
    (tips x sites, ascending) and dealt to the ranks in a zig-zag -- 0, 1, .., W-1, W-1, .., 1, 0, 0, 1, .. -- so
    that every rank gets the same number of loci (+-1) and nearly the same work.  Returns a list of `world` lists
    of locus indices (the order inside a rank follows the deal, like the reference's reordered msa_list)."""
    order = sorted(range(len(loads)), key=lambda i: (loads[i], i))
    out = [[] for _ in range(world)]
    core, inc = 0, 1
    for i in order:
        out[core].append(i)
        core += inc
        if core == world:
            inc, core = -1, world - 1
        elif core == -1:
            inc, core = 1, 0
    return out


def allreduce_sum(values, group=None):
    """Sum a small float64 vector over the ranks (NCCL on GPUs, gloo in the CPU tests); identity
    when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    t = values if isinstance(values, torch.Tensor) else torch.tensor(values, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
