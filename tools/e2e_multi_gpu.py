"""The end-to-end host-buffer season (nesosim_run_season_host: what bench.py's `e2e` times) on every GPU of a box at
once, one process per GPU, under several settings of the drain -- plain or compacted, host threads per rank -- inside
ONE set of processes (the 26 GB of pinned arrays per rank are allocated once).
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_multi_gpu.py
       (env VARIANTS="K=V,K=V;..." overrides the list; @T in a value is replaced by cores // world, @T2 by twice that)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from nesosim_b200 import synthetic as S, _lib
from nesosim_b200.engine import SnowBudgetEngine

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
M, T, DX = 128, 260, 100000
mask = S.region_mask(dx=DX)
ny, nx = mask.shape
F = S.make_season(mask, T, seed=2024)
ic = S.make_ic(mask, seed=2024)
params = S.ensemble_params(M, seed=2024 + rank)
host_out = {n: torch.empty((M, T, 2, ny, nx) if n == "snowDepths" else (M, T, ny, nx), dtype=torch.float64, pin_memory=True)
            for n in _lib.OUTPUT_NAMES}
hf = {k: torch.from_numpy(np.ascontiguousarray(F[k])).pin_memory() for k in ("precip", "conc", "wind", "drift")}
ic_h = torch.from_numpy(np.ascontiguousarray(ic)).pin_memory()
cores = os.cpu_count() or 1


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# the link with every rank copying at once: plain pinned D2H, 4 x 1 GiB
d = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
h = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
h.copy_(d, non_blocking=True)
barrier()
t0 = time.perf_counter()
for _ in range(4):
    h.copy_(d, non_blocking=True)
barrier()
link = 4 * (1 << 30) / max_over_ranks(time.perf_counter() - t0) / 1e9
del d, h

variants = os.environ.get("VARIANTS", "NESOSIM_HOST_COMPACT=0,NESOSIM_HOST_THREADS=@T;NESOSIM_HOST_COMPACT=1,NESOSIM_HOST_THREADS=@T;"
                          "NESOSIM_HOST_COMPACT=1,NESOSIM_HOST_THREADS=@T2").split(";")
KEYS = ("NESOSIM_HOST_NO_SHARE", "NESOSIM_HOST_THREADS", "NESOSIM_HOST_BATCH_GB", "NESOSIM_HOST_COMPACT", "NESOSIM_HOST_HYBRID", "NESOSIM_HOST_CHUNK_MB", "NESOSIM_HOST_RING")
eng = SnowBudgetEngine(mask, T, DX, n_members=M, atmlossInc=1, device=local)
ref = None
for v in variants:
    for k in KEYS:
        os.environ.pop(k, None)
    v = v.replace("@T2", str(max(1, 2 * cores // world))).replace("@T", str(max(1, cores // world)))
    for kv in filter(None, v.split(",")):
        k, val = kv.split("=", 1)
        os.environ[k] = val
    eng.run_season_host(hf, params, ic_h, host_out)
    ts = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        _, up, down = eng.run_season_host(hf, params, ic_h, host_out)
        barrier()
        ts.append(max_over_ranks(time.perf_counter() - t0))
    info = eng.host_drain_info()
    blocks = eng.host_drain_blocks()
    # same arrays whatever the drain: a digest of this rank's result against the first variant's
    dig = [float(host_out[n][[0, M // 2, M - 1]].nan_to_num(nan=-7.0).sum()) for n in ("snowDepths", "density", "snowLead", "snowAcc")]
    if ref is None:
        ref = dig
    same = max_over_ranks(0.0 if dig == ref else 1.0) == 0.0
    if rank == 0:
        ms = 1e3 * min(ts)
        print(json.dumps({"gpus": world, "host_cores": cores, "variant": v, "drain": "compacted" if info[0] else "full", "blocks_packed": blocks[0], "blocks_plain": blocks[1],
                          "ms_all": [round(1e3 * t, 1) for t in ts], "ms_per_season": ms, "d2h_GB_per_gpu": down / 1e9,
                          "link_GBs_per_gpu": link, "same_digest_as_first": same,
                          "cell_days_per_s": world * M * ny * nx * (T - 1) / (ms * 1e-3)}), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
