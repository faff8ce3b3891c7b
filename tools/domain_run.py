"""Domain-decomposed season over the GPUs of one box, checked against the same season on one GPU and timed.

usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           tools/domain_run.py [n] [days] [reps]
n = grid side (357 = 25 km, 1785 = 5 km), days = season length, reps = timed repetitions.

Three arms on the same synthetic season (8 generated days repeated to `days`; repetition changes neither the work nor
the parity check):
  one-gpu : rank 0 alone, the whole grid, general day kernel (one launch per day)
  nccl    : row strips, one engine call per day + torch.distributed batched isend/irecv of the ghost rows
  peer    : row strips, exchange fused into the day kernel (mailboxes in peer memory over CUDA IPC / NVLink)
Device time of a season = CUDA events around the enqueued season on every rank, max over ranks (peer, one-gpu) or wall
clock between synchronised barriers (nccl: the host drives every day).  Prints one JSON line per arm from rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nesosim_b200 import domain, synthetic as S
from nesosim_b200.engine import SnowBudgetEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 357
T = int(sys.argv[2]) if len(sys.argv) > 2 else 11
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dx = {90: 100000, 357: 25000, 1785: 5000}.get(n, 50000)
mask = S.region_mask(dx=dx) if n in (90, 357) else S.region_mask(shape=(n, n), kind="disc")
gen = S.make_season(mask, min(T, 8), seed=41)
idx = np.arange(T) % min(T, 8)
forcing = {k: (v if v is None else v[idx]) for k, v in gen.items()}
ic = S.make_ic(mask, seed=41)
params = [5.8e-7, 5., 1.45e-7, 2.2e-8]
cells = n * n * (T - 1)
KEYS = ("snowDepths", "density", "snowAdv", "snowLead")


def report(arm, ms, ok, extra=None):
    if rank == 0:
        line = {"arm": arm, "grid": [n, n], "days": T, "n_gpus": 1 if arm == "one-gpu" else world, "ms_per_season": ms,
                "us_per_day": 1e3 * ms / (T - 1), "cell_days_per_s": cells / (ms * 1e-3), "identical_to_one_gpu": ok}
        line.update(extra or {})
        print(json.dumps(line), flush=True)


def max_over_ranks(v):
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- one GPU: reference values + timing (rank 0 only; the others wait)
ref = None
if rank == 0:
    eng = SnowBudgetEngine(mask, T, dx, n_members=1, device=local, atmlossInc=1)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = eng.alloc_outputs()
    best = 1e30
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.run_season([params], ic, out)
        e1.record(); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    ref = {k: out[k][0].cpu().numpy() for k in KEYS}
    del out
    eng.close()
    torch.cuda.empty_cache()
    report("one-gpu", best, True)
dist.barrier()

# ---- peer-memory strips
eng = None
best = 1e30
for r in range(reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if eng is None:
        lo, hi, part, eng = domain.run_decomposed_season_peer(mask, T, dx, forcing, params, ic, rank, world, device=local,
                                                              atmlossInc=1)
        outs = eng._keep[1]
        continue
    ic_local = np.ascontiguousarray(ic[eng._strip_rows[2]:eng._strip_rows[3]])
    ic_dev = eng._dev(ic_local)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    eng.run_season([params], ic_dev, outs)
    e1.record(); torch.cuda.synchronize()
    assert not eng.strip_timed_out()
    best = min(best, max_over_ranks(e0.elapsed_time(e1)))
    dist.barrier()
parts = [None] * world
dist.all_gather_object(parts, (lo, hi, {k: part[k].cpu().numpy() for k in KEYS}))
ok = None
if rank == 0:
    ok = all(bool(np.array_equal(p[k], ref[k][..., lo_:hi_, :], equal_nan=True)) for lo_, hi_, p in parts for k in KEYS)
report("peer", best, ok, {"launches_per_day": 1, "exchange": "fused into the day kernel (peer-memory mailboxes + flags)"})
del outs, part
eng.close(); eng = None
torch.cuda.empty_cache()
dist.barrier()

if "nccl" not in os.environ.get("ARMS", "one-gpu,peer,nccl"):
    dist.destroy_process_group()
    sys.exit(0)

# ---- NCCL strips (host-driven exchange every day)
def make(local_mask, num_days, dx_, forcing_local, params_row, ic_local):
    return domain.GpuStripStepper(local_mask, num_days, dx_, forcing_local, params_row, ic_local, device=local, atmlossInc=1)

best = 1e30
clock = {}
def start_clock():
    torch.cuda.synchronize(); dist.barrier(); clock["t0"] = time.perf_counter()
def stop_clock():
    torch.cuda.synchronize(); dist.barrier(); clock["t1"] = time.perf_counter()
for r in range(2):
    lo, hi, part = domain.run_decomposed_season(mask, T, dx, forcing, params, ic, rank, world, make, on_ready=start_clock,
                                                on_done=stop_clock)
    if r: best = (clock["t1"] - clock["t0"]) * 1e3
parts = [None] * world
dist.all_gather_object(parts, (lo, hi, {k: part[k] for k in KEYS}))
if rank == 0:
    ok = all(bool(np.array_equal(p[k], ref[k][..., lo_:hi_, :], equal_nan=True)) for lo_, hi_, p in parts for k in KEYS)
report("nccl", best, ok, {"launches_per_day": 1, "exchange": "torch.distributed batched isend/irecv per day (NCCL), host-driven; wall clock of the day loop"})
dist.barrier()
dist.destroy_process_group()
