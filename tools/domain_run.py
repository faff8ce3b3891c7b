"""Domain-decomposed season over the GPUs of one box (torchrun; NCCL halo exchange) checked against the same season
on one GPU.  usage: python -m torch.distributed.run --nproc-per-node N tools/domain_run.py [n] [days]
n = grid side (357 = 25 km, 1785 = 5 km), days = season length."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nesosim_b200 import domain, synthetic as S
from nesosim_b200.engine import SnowBudgetEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 357
T = int(sys.argv[2]) if len(sys.argv) > 2 else 11
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dx = {90: 100000, 357: 25000, 1785: 5000}.get(n, 50000)
mask = S.region_mask(dx=dx) if n in (90, 357) else S.region_mask(shape=(n, n), kind="disc")
forcing = S.make_season(mask, T, seed=41)
ic = S.make_ic(mask, seed=41)
params = [5.8e-7, 5., 1.45e-7, 2.2e-8]

def make(local_mask, num_days, dx_, forcing_local, params_row, ic_local):
    return domain.GpuStripStepper(local_mask, num_days, dx_, forcing_local, params_row, ic_local, device=local, atmlossInc=1)

dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
lo, hi, part = domain.run_decomposed_season(mask, T, dx, forcing, params, ic, rank, world, make)
torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
parts = [None] * world
dist.all_gather_object(parts, (lo, hi, {k: v for k, v in part.items() if k in ("snowDepths", "density", "snowAdv")}))
if rank == 0:
    eng = SnowBudgetEngine(mask, T, dx, n_members=1, device=local, atmlossInc=1)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    ref = {k: v[0].cpu().numpy() for k, v in eng.run_season([params], ic).items()}
    ok = True
    for lo_, hi_, p in parts:
        for k, v in p.items():
            ok &= bool(np.array_equal(v, ref[k][..., lo_:hi_, :], equal_nan=True))
    cells = n * n * (T - 1)
    print("domain decomposition %dx%d, %d days, %d ranks: value-identical to one GPU = %s; %.3f s incl. staging (%.2e cell-days/s)"
          % (n, n, T, world, ok, t1 - t0, cells / (t1 - t0)), flush=True)
    assert ok
dist.destroy_process_group()
