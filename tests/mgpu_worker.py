"""Launched by tests/test_gpu_multi.py under torch.distributed.run (one rank per GPU): observation sharding with
the NCCL all-reduce inside libjbugs_cuda must reproduce the single-GPU evaluation; chain sharding must reproduce
the corresponding columns."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "juliabugs.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import jbugs
    from c_oracle import COracle
    from helpers import random_thetas, synth_seeds
    from jbugs.dist import chain_bounds, shard_bounds, shard_observations

    N, C = 5000, 70
    data, b = synth_seeds(N, seed=21)
    th0 = np.concatenate([[np.log(1 / 0.09), -0.84, 1.36, 0.09, -0.55], b])
    theta = random_thetas(th0, C, seed=8, scale=0.05)
    ref_lp, ref_g, _ = COracle(jbugs.lower(jbugs.examples.SEEDS, data)[0]).eval_batch(theta)

    # --- observation sharding + all-reduce
    m = jbugs.set_evaluation_mode(jbugs.compile(jbugs.examples.SEEDS, data), jbugs.UseCUDABatch(device=local))
    cp = shard_observations(m, rank, world)
    lp, g, st = cp.eval_host(theta, layout="chain_fast")
    i0, i1 = shard_bounds(N, world, rank)
    assert np.all(st == 0)
    assert np.max(np.abs(lp - ref_lp) / np.abs(ref_lp)) < 1e-10, "logp after all-reduce"
    gs = np.maximum(np.abs(ref_g), 1e-4 * np.abs(ref_g).max(axis=0, keepdims=True))
    assert np.max(np.abs(g[:5] - ref_g[:5]) / gs[:5]) < 1e-10, "global gradients after all-reduce"
    assert np.max(np.abs(g[5 + i0:5 + i1] - ref_g[5 + i0:5 + i1]) / gs[5 + i0:5 + i1]) < 1e-10, "local gradients of the shard"
    lp_t = torch.from_numpy(lp).cuda()
    lp_all = [torch.empty_like(lp_t) for _ in range(world)]
    dist.all_gather(lp_all, lp_t)
    assert all(torch.equal(lp_all[0], x) for x in lp_all), "every rank must hold the identical logp"

    # --- chain sharding: no collective
    m2 = jbugs.set_evaluation_mode(jbugs.compile(jbugs.examples.SEEDS, data), jbugs.UseCUDABatch(device=local))
    c0, c1 = chain_bounds(C, world, rank)
    lp2, g2, _ = m2.cuda.eval_host(theta[:, c0:c1], layout="chain_fast")
    assert np.max(np.abs(lp2 - ref_lp[c0:c1]) / np.abs(ref_lp[c0:c1])) < 1e-10
    assert np.max(np.abs(g2 - ref_g[:, c0:c1]) / gs[:, c0:c1]) < 1e-10
    # --- dense predictor: rows of X sharded, P x C block of X^T R in the same all-reduce, prior plate replicated
    rng = np.random.default_rng(77)
    Nd, Pd, Cd = 4096, 40, 130
    X = np.asfortranarray(rng.standard_normal((Nd, Pd)))
    bt = rng.normal(0.0, 0.3, Pd)
    y = (rng.random(Nd) < 1.0 / (1.0 + np.exp(-X @ bt))).astype(float)
    ddata = {"N": Nd, "P": Pd, "X": X, "y": y}
    thd = random_thetas(np.concatenate([[0.5, 0.1], bt]), Cd, seed=4, scale=0.1)
    md = jbugs.set_evaluation_mode(jbugs.compile(jbugs.examples.DENSE_LOGISTIC, ddata), jbugs.UseCUDABatch(device=local))
    ref_lp, ref_g, _ = COracle(md.plan).eval_batch(thd)
    cpd = shard_observations(md, rank, world)
    lp3, g3, st3 = cpd.eval_host(thd, layout="chain_fast")
    assert np.all(st3 == 0)
    assert np.max(np.abs(lp3 - ref_lp) / np.abs(ref_lp)) < 1e-10, "dense: logp after all-reduce"
    gs3 = np.maximum(np.abs(ref_g), 1e-4 * np.abs(ref_g).max(axis=0, keepdims=True))
    assert np.max(np.abs(g3 - ref_g) / gs3) < 1e-10, "dense: every gradient row is complete on every rank"
    # --- the sampler with chains sharded over the ranks: every rank ends up with all chains, no collective on the path
    from jbugs import hmc
    from jbugs.dist import sample_chains_sharded
    rng = np.random.default_rng(5)
    yy = rng.normal(1.7, 1.0, 400)
    mm = jbugs.set_evaluation_mode(
        jbugs.compile("model { for (i in 1:N) { y[i] ~ dnorm(mu, 1.0) }\n mu ~ dnorm(0.0, 1.0E-6) }", {"N": 400, "y": yy}, {"mu": 0.0}),
        jbugs.UseCUDABatch(device=local))
    ch = sample_chains_sharded(mm, hmc.NUTS(max_depth=6, step_size=0.05), n_samples=200, n_adapts=200, n_chains=96 * world + 1)
    assert ch["mu"].shape == (200, 96 * world + 1)
    assert abs(ch["mu"].mean() - yy.mean()) < 8e-3 and abs(ch["mu"].std() - 0.05) < 5e-3
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
