"""Throughput of the general per-day path (day_step_kernel) on the larger grids: single season, M members.
usage: python tools/general_timing.py n days [members]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nesosim_b200 import synthetic as S
from nesosim_b200.engine import SnowBudgetEngine
n = int(sys.argv[1]); T = int(sys.argv[2]); M = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dx = {90: 100000, 357: 25000, 1785: 5000}[n]
mask = S.region_mask(dx=dx) if n in (90, 357) else S.region_mask(shape=(n, n), kind="disc")
F = S.make_season(mask, T, seed=1)
ic = S.make_ic(mask, seed=1)
params = S.ensemble_params(M, seed=1)
eng = SnowBudgetEngine(mask, T, dx, n_members=M, atmlossInc=1)
eng.set_path("general")
eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
out = eng.alloc_outputs()
ts = []
for rep in range(4):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_season(params, ic, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts[1:])
cells = M * n * n * (T - 1)
print("general path %dx%d, %d days, M=%d: %.3f ms/season, %.1f us/day, %.3e cell-days/s, %.0f GB/s algorithmic (%.1f%% of 6551)"
      % (n, n, T, M, ms, 1e3 * ms / (T - 1), cells / ms * 1e3, cells * (96 + 41.0 / M) / ms / 1e6, cells * (96 + 41.0 / M) / ms / 1e6 / 65.51))
