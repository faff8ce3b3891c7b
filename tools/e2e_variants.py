"""End-to-end host-buffer season (nesosim_run_season_host, the call bench.py's `e2e` times) under a few settings of the
host path: member-independent arrays shared or not, number of replication threads, staging batch size.
usage: python tools/e2e_variants.py   (env VARIANTS="K=V,K=V;..." overrides the list)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nesosim_b200 import synthetic as S, _lib
from nesosim_b200.engine import SnowBudgetEngine
M, T, DX = 128, 260, 100000
mask = S.region_mask(dx=DX)
ny, nx = mask.shape
F = S.make_season(mask, T, seed=2024)
ic = S.make_ic(mask, seed=2024)
params = S.ensemble_params(M, seed=2024)
host_out = {n: torch.empty((M, T, 2, ny, nx) if n == "snowDepths" else (M, T, ny, nx), dtype=torch.float64, pin_memory=True)
            for n in _lib.OUTPUT_NAMES}
hf = {k: torch.from_numpy(np.ascontiguousarray(F[k])).pin_memory() for k in ("precip", "conc", "wind", "drift")}
ic_h = torch.from_numpy(np.ascontiguousarray(ic)).pin_memory()
# the link: plain pinned D2H of 4 GiB in 1 GiB pieces
d = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
h = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
link = 4 * (1 << 30) / (time.perf_counter() - t0) / 1e9
del d, h
variants = os.environ.get("VARIANTS", ";NESOSIM_HOST_NO_SHARE=1;NESOSIM_HOST_THREADS=2;NESOSIM_HOST_THREADS=4;NESOSIM_HOST_THREADS=16;"
                          "NESOSIM_HOST_BATCH_GB=2;NESOSIM_HOST_BATCH_GB=16;NESOSIM_HOST_BATCH_GB=2,NESOSIM_HOST_THREADS=4").split(";")
KEYS = ("NESOSIM_HOST_NO_SHARE", "NESOSIM_HOST_THREADS", "NESOSIM_HOST_BATCH_GB", "NESOSIM_HOST_COMPACT", "NESOSIM_HOST_HYBRID", "NESOSIM_HOST_CHUNK_MB")
for v in variants:
    for k in KEYS:
        os.environ.pop(k, None)
    for kv in filter(None, v.split(",")):
        k, val = kv.split("=", 1)
        os.environ[k] = val
    eng = SnowBudgetEngine(mask, T, DX, n_members=M, atmlossInc=1)
    eng.run_season_host(hf, params, ic_h, host_out)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, up, down = eng.run_season_host(hf, params, ic_h, host_out)
        ts.append(time.perf_counter() - t0)
    info = eng.host_drain_info()
    blocks = eng.host_drain_blocks()
    eng.close()
    ms = 1e3 * min(ts)
    print(json.dumps({"variant": v or "default", "drain": "compacted" if info[0] else "full", "blocks_packed": blocks[0], "blocks_plain": blocks[1], "chunks_copied_in_full": info[1],
                      "ms_all": [round(1e3 * t, 1) for t in ts], "host_cores": os.cpu_count(), "ms_per_season": ms, "d2h_GB": down / 1e9, "d2h_GBs": down / 1e9 / (ms * 1e-3),
                      "link_GBs": link, "fraction_of_link": down / 1e9 / (ms * 1e-3) / link,
                      "cell_days_per_s": M * ny * nx * (T - 1) / (ms * 1e-3)}), flush=True)
