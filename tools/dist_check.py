"""2-rank NCCL check of jaxoplanet_b200.distributed.log_likelihood_distributed (run under torchrun)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxoplanet_b200 import workloads as W
from jaxoplanet_b200.distributed import log_likelihood_distributed, _cuda_evaluate

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for n_sets in (513, 1):
    p, u, t, noise, sigma = W.c3(n_sets, 50_000, seed=21)
    y = 1 + noise
    td, yd = torch.as_tensor(t).cuda(), torch.as_tensor(y).cuda()
    full = log_likelihood_distributed(p, u, td, yd, sigma)
    want = _cuda_evaluate(p, u, td, yd, sigma)
    err = float(torch.max(torch.abs(full - want) / torch.abs(want)))
    exact = bool(torch.equal(full, want))
    print(f"rank {rank} n_sets {n_sets}: shape {tuple(full.shape)} max rel err {err:.2e} bit-identical {exact}", flush=True)
    ok = ok and full.shape == (n_sets,) and err <= 1e-13 and (exact or n_sets == 1)
# the peer-store all-gather on its own: ragged blocks, many epochs back to back (a rank that runs ahead must not
# overwrite values a slower rank has not read), against NCCL's all-gather of the same blocks
from jaxoplanet_b200.distributed import gather_sets, peer_gather_for, shard_range
world = dist.get_world_size()
for n_total in (4096, 1001, world):
    pg = peer_gather_for(n_total, torch.device("cuda", local))
    print(f"rank {rank} n_total {n_total}: peer gather {'set up' if pg is not None else 'NOT available (NCCL fallback)'}", flush=True)
    if pg is None:
        ok = False
        continue
    lo, hi = shard_range(n_total, rank, world)
    bad = 0
    for e in range(200):
        mine = torch.arange(lo, hi, dtype=torch.float64, device="cuda") * (e + 1) + 0.25 * rank
        got = pg.gather(mine, lo)
        if rank == e % world:
            torch.cuda._sleep(200_000)  # (this rank lags: the others run ahead into the next epoch)
        want = torch.cat([torch.arange(*shard_range(n_total, r, world), dtype=torch.float64, device="cuda") * (e + 1) + 0.25 * r
                          for r in range(world)])
        bad += int(not torch.equal(got, want))
    st = pg.status()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mine = torch.ones(hi - lo, dtype=torch.float64, device="cuda")
    out = torch.empty(n_total, dtype=torch.float64, device="cuda")
    tp = tn = None
    if n_total % world == 0:
        for which in ("peer", "nccl"):
            dist.barrier(); torch.cuda.synchronize()
            ev0.record()
            for _ in range(200):
                if which == "peer":
                    pg.gather(mine, lo)
                else:
                    dist.all_gather_into_tensor(out, mine)
            ev1.record(); torch.cuda.synchronize()
            if which == "peer":
                tp = ev0.elapsed_time(ev1) / 200 * 1e3
            else:
                tn = ev0.elapsed_time(ev1) / 200 * 1e3
    print(f"rank {rank} n_total {n_total}: wrong epochs {bad} of 200, status {st}, us per gather: peer {tp}, nccl {tn}", flush=True)
    ok = ok and bad == 0 and st == 0
    g = gather_sets(torch.arange(lo, hi, dtype=torch.float64, device="cuda"), n_total)
    ok = ok and bool(torch.equal(g, torch.arange(n_total, dtype=torch.float64, device="cuda")))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
