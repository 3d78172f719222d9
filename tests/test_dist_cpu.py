"""N > 1 host logic on CPU: world_size-2 gloo processes.  Each rank evaluates its observation shard with the C
oracle, the `[logp, gradient]` pieces are all-reduced, and the result must equal the unsharded evaluation --
the same decomposition libjbugs_cuda performs with ncclAllReduce (priors of the globals counted once)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "juliabugs.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import jbugs
    from c_oracle import COracle
    from helpers import random_thetas, synth_seeds
    from jbugs.dist import broadcast_unique_id, chain_bounds, shard_bounds

    data, b = synth_seeds(1000, seed=11)
    plan, _ = jbugs.lower(jbugs.examples.SEEDS, data)
    th0 = np.concatenate([[np.log(1 / 0.09), -0.84, 1.36, 0.09, -0.55], b])
    theta = random_thetas(th0, 6, seed=4, scale=0.05)
    co = COracle(plan)
    full_lp, full_g, _ = co.eval_batch(theta)
    # priors of the globals alone (empty shard) -- counted once after the reduction
    co.L.jbo_plan_set_shard(co.h, 0, 0, 0)
    prior_lp, prior_g, _ = co.eval_batch(theta)
    i0, i1 = shard_bounds(plan.plates[0].N, world, rank)
    co.L.jbo_plan_set_shard(co.h, 0, i0, i1)
    lp, g, _ = co.eval_batch(theta)
    part = torch.from_numpy(np.concatenate([(lp - prior_lp)[None, :], g - prior_g]))
    dist.all_reduce(part)
    tot = part.numpy()
    lp_sum = tot[0] + prior_lp
    g_sum = tot[1:] + prior_g
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank, world)
    ok = (np.allclose(lp_sum, full_lp, rtol=1e-13, atol=0)
          and np.allclose(g_sum, full_g, rtol=1e-10, atol=1e-10)
          and uid == bytes(range(128))
          and chain_bounds(7, world, rank) == ((0, 4) if rank == 0 else (4, 7)))
    out[rank] = (bool(ok), (i0, i1))
    dist.destroy_process_group()


def test_observation_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    (a0, a1), (b0, b1) = out[0][1], out[1][1]
    assert a0 == 0 and a1 == b0 and b1 == 1000 and a1 % 32 == 0


def test_shard_bounds_partition():
    from jbugs.dist import shard_bounds
    for N in (0, 1, 31, 32, 33, 1000, 10_000_019):
        for world in (1, 2, 4, 8):
            edges = [shard_bounds(N, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == N
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(e[0] % 32 == 0 or e[0] == N for e in edges)
