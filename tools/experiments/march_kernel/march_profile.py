"""Short run of the marching day kernel on the 5 km grid for ncu (tools: ncu -k regex:day_march ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NESOSIM_DAY_KERNEL", "march")
import numpy as np, torch
from nesosim_b200 import synthetic as S
from nesosim_b200.engine import SnowBudgetEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1785
dx = {357: 25000, 1785: 5000}[n]
T = 7
mask = S.region_mask(dx=dx)
gen = S.make_season(mask, 2, seed=1)
idx = np.arange(T) % 2
F = {k: torch.from_numpy(v[idx]).cuda() for k, v in gen.items() if k != "temp"}
eng = SnowBudgetEngine(mask, T, dx, n_members=1, atmlossInc=1)
eng.set_path("general")
eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
out = eng.alloc_outputs()
eng.run_season([[5.8e-7, 5., 1.45e-7, 2.2e-8]], S.make_ic(mask, seed=1), out)
torch.cuda.synchronize()
print("ok", eng.last_day_kernel())
