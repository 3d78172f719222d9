"""Multi-GPU plumbing (one process per GPU, `torch.distributed` for rendezvous only).

The reference parallelises only across chains, with one whole `sample` call per thread/process
(`JuliaBUGS/docs/src/inference/parallel.md:5-63`, AbstractMCMC `MCMCThreads` / `MCMCDistributed`).
Here:

* chain sharding   -- rank r owns chains [c0, c1): no data-path collective at all;
* observation sharding -- rank r owns groups [i0, i1) of a plate (and the theta / gradient rows of that
  plate's local parameters); the per-chain `[logp, global adjoints]` are summed by ONE `ncclAllReduce`
  issued by libjbugs_cuda on the evaluation stream.

torch.distributed (gloo or nccl) is used only to hand the 128-byte ncclUniqueId from rank 0 to the others.
"""
from __future__ import annotations

from typing import Tuple

ALIGN = 32  # shard boundaries fall on staged-chunk boundaries (csrc: STAGE_G)


def chain_bounds(C: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous split of C chains; the first (C mod world) ranks get one more."""
    base, rem = divmod(C, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def shard_bounds(N: int, world: int, rank: int, align: int = ALIGN) -> Tuple[int, int]:
    """contiguous split of N plate groups into `world` shards whose boundaries are multiples of `align`
    (the last shard takes the remainder).  Every group belongs to exactly one shard."""
    blocks = (N + align - 1) // align
    b0, b1 = chain_bounds(blocks, world, rank)
    return min(b0 * align, N), min(b1 * align, N)


def broadcast_unique_id(make_id, rank: int, world: int, group=None) -> bytes:
    """rank 0 calls `make_id()` (-> 128 bytes); everyone gets the same bytes back."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return make_id()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    if rank == 0:
        raw = make_id()
        t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


def shard_observations(model, rank: int, world: int, group=None):
    """Give `model`'s device plan this rank's shard of every plate and attach the NCCL communicator."""
    from .cuda import CudaPlan

    cp = model.cuda
    for k, pl in enumerate(model.plan.plates):
        if not pl.has_obs:
            continue  # prior-only plate (coefficients of a dense predictor): replicated, counted once
        i0, i1 = shard_bounds(pl.N, world, rank)
        cp.set_shard(k, i0, i1)
    for dn in model.plan.dense:
        cp.set_dense_shard(*shard_bounds(dn.N, world, rank))  # rows of X, y; X^T R joins the all-reduce
    if world > 1:
        uid = broadcast_unique_id(CudaPlan.comm_unique_id, rank, world, group)
        cp.comm_init(world, rank, uid)
    return cp


def sample_chains_sharded(model, sampler, n_samples: int, n_adapts: int, n_chains: int, seed: int = 1234, monitor=None,
                          group=None):
    """`AbstractMCMC.sample(model, sampler, MCMCDistributed(), n, nchains)` of the reference
    (`JuliaBUGS/docs/src/inference/parallel.md:36-63`): the chains are split over the ranks (one GPU each), every rank
    runs the on-device sampler on its own chains with its own seed, no data-path collective; the draws are gathered
    on every rank with `all_gather_object`.  Returns a `Chains` holding all `n_chains` chains."""
    import numpy as np
    import torch.distributed as dist

    from . import hmc

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    c0, c1 = chain_bounds(n_chains, world, rank)
    local = hmc.sample(model, sampler, n_samples, n_adapts=n_adapts, n_chains=c1 - c0, seed=seed + 7919 * rank, monitor=monitor)
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, (local.samples, local.accept_rate, local.step_size, local.n_divergent), group=group)
    samples = np.concatenate([p[0] for p in parts], axis=2)
    w = np.array([p[0].shape[2] for p in parts], dtype=float)
    return hmc.Chains(local.names, samples, float(np.average([p[1] for p in parts], weights=w)),
                      float(np.mean([p[2] for p in parts])), model, n_divergent=int(sum(p[3] for p in parts)),
                      tree_depth=local.tree_depth)
