"""Short single-wave run of the season-resident kernel for ncu (source-level stall attribution).
usage: python tools/ens_profile.py [members] [days] [variant] [cluster]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = int(sys.argv[1]) if len(sys.argv) > 1 else 22
T = int(sys.argv[2]) if len(sys.argv) > 2 else 41
if len(sys.argv) > 3 and sys.argv[3] not in ("", "-"):
    os.environ["NESOSIM_ENS_VARIANT"] = sys.argv[3]
if len(sys.argv) > 4 and sys.argv[4] not in ("", "-"):
    os.environ["NESOSIM_ENS_CLUSTER"] = sys.argv[4]
from nesosim_b200 import synthetic as S
from nesosim_b200.engine import SnowBudgetEngine

mask = S.region_mask(dx=100000)
F = S.make_season(mask, T, seed=1)
ic = S.make_ic(mask, seed=1)
params = S.ensemble_params(M, seed=1)
eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
eng.set_path("ensemble")
eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
out = eng.alloc_outputs()
for rep in range(2):
    eng.run_season(params, ic, out)
torch.cuda.synchronize()
print("ok")
