"""World-size-2 `gloo` test of the N > 1 host logic (set sharding + gather) on CPU.  The per-set
values are produced by the CPU oracle here (no GPU in this tier); on the GPU box the same
`shard_range` / `gather_sets` wrap the CUDA path (bench.py)."""

import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from jaxoplanet_b200.distributed import shard_range

    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_sets, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.distributed import gather_sets, shard_range
    from oracle import ref_numpy as ref

    p, u, t, noise, sigma = W.c3(n_sets, 400, seed=5)
    y = 1 + noise
    lo, hi = shard_range(n_sets, rank, world)
    ll = []
    for s in range(lo, hi):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        m = 1 + ref.system_light_curve([bd], u[s], t).sum(-1)
        ll.append(ref.gaussian_loglike(m, y, np.full_like(t, sigma)))
    full = gather_sets(torch.tensor(ll, dtype=torch.float64), n_sets)
    np.save(os.path.join(out_dir, f"ll_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_two_rank_gather_matches_single(tmp_path):
    from jaxoplanet_b200 import workloads as W
    from oracle import ref_numpy as ref

    n_sets = 7  # ragged: 4 + 3
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n_sets, str(tmp_path)), nprocs=2, join=True)
    p, u, t, noise, sigma = W.c3(n_sets, 400, seed=5)
    y = 1 + noise
    want = []
    for s in range(n_sets):
        bd = ref.orbital_body(**{k: v[s] for k, v in p.items()})
        m = 1 + ref.system_light_curve([bd], u[s], t).sum(-1)
        want.append(ref.gaussian_loglike(m, y, np.full_like(t, sigma)))
    for r in range(2):
        got = np.load(tmp_path / f"ll_{r}.npy")
        np.testing.assert_array_equal(got, np.array(want))


def _oracle_eval(params, u, t, y, sigma):
    from oracle import ref_numpy as ref

    out = []
    n = np.shape(params["period"])[0]
    for s in range(n):
        bd = ref.orbital_body(**{k: v[s] for k, v in params.items()})
        us = u[s] if np.ndim(u) == 2 else u
        m = 1 + ref.system_light_curve([bd], us, t).sum(-1)
        out.append(ref.gaussian_loglike(m, y, np.broadcast_to(sigma, t.shape)))
    return torch.tensor(out, dtype=torch.float64)


def _worker_dist(rank, world, port, n_sets, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from jaxoplanet_b200 import workloads as W
    from jaxoplanet_b200.distributed import log_likelihood_distributed

    p, u, t, noise, sigma = W.c3(n_sets, 600, seed=9)
    y = 1 + noise
    sig = np.full_like(t, sigma) * (1 + 0.1 * np.sin(t))  # per-sample sigma
    full = log_likelihood_distributed(p, u, t, y, sig, evaluate=_oracle_eval)
    np.save(os.path.join(out_dir, f"lld_{n_sets}_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


def test_two_rank_set_and_time_sharding_match_single(tmp_path):
    """n_sets >= world: set-sharded + all-gather; n_sets < world (1 set, 2 ranks): time-sharded +
    all-reduce(sum) of the partial log-likelihoods (SURVEY 8e fallback)."""
    from jaxoplanet_b200 import workloads as W

    for n_sets in (5, 1):
        port = 31500 + (os.getpid() + n_sets) % 2000
        mp.spawn(_worker_dist, args=(2, port, n_sets, str(tmp_path)), nprocs=2, join=True)
        p, u, t, noise, sigma = W.c3(n_sets, 600, seed=9)
        sig = np.full_like(t, sigma) * (1 + 0.1 * np.sin(t))
        want = _oracle_eval(p, u, t, 1 + noise, sig).numpy()
        for r in range(2):
            got = np.load(tmp_path / f"lld_{n_sets}_{r}.npy")
            if n_sets >= 2:
                np.testing.assert_array_equal(got, want)
            else:  # the sum over time is split in two: rounding of the last bits
                np.testing.assert_allclose(got, want, rtol=1e-13)
