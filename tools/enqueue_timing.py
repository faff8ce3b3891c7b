"""Is the general path's day loop bound by the host enqueueing launches or by the device?  Wall time of the
nesosim_run_season call (asynchronous: returns when every day is enqueued) against the device time of the season.
usage: python tools/enqueue_timing.py n days"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nesosim_b200 import synthetic as S
from nesosim_b200.engine import SnowBudgetEngine
n = int(sys.argv[1]); T = int(sys.argv[2])
dx = {90: 100000, 357: 25000, 1785: 5000}[n]
mask = S.region_mask(dx=dx) if n in (90, 357) else S.region_mask(shape=(n, n), kind="disc")
gen = S.make_season(mask, min(T, 8), seed=1)
idx = np.arange(T) % min(T, 8)
F = {k: (v if v is None else torch.from_numpy(v[idx]).cuda()) for k, v in gen.items()}
ic = torch.from_numpy(S.make_ic(mask, seed=1)).cuda()
eng = SnowBudgetEngine(mask, T, dx, n_members=1, atmlossInc=1)
eng.set_path("general")
eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
out = eng.alloc_outputs()
p = [[5.8e-7, 5., 1.45e-7, 2.2e-8]]
for rep in range(4):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); eng.run_season(p, ic, out); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%dx%d %d days: host enqueue %.3f ms (%.2f us/day), device %.3f ms (%.2f us/day), wall to completion %.3f ms"
          % (n, n, T, (t1 - t0) * 1e3, (t1 - t0) * 1e6 / (T - 1), e0.elapsed_time(e1), e0.elapsed_time(e1) * 1e3 / (T - 1), (t2 - t0) * 1e3))
