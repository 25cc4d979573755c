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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
