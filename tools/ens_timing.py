"""Debug aid: per-phase cycle totals of the season-resident kernel (NESOSIM_ENS_TIMING) and wall time per season.
usage: python tools/ens_timing.py [members] [variant] [cluster]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
if len(sys.argv) > 2 and sys.argv[2] not in ("", "-"):
    os.environ["NESOSIM_ENS_VARIANT"] = sys.argv[2]
if len(sys.argv) > 3 and sys.argv[3] not in ("", "-"):
    os.environ["NESOSIM_ENS_CLUSTER"] = sys.argv[3]
from nesosim_b200 import synthetic as S, _lib
from nesosim_b200.engine import SnowBudgetEngine

mask = S.region_mask(dx=100000)
T = 260
F = S.make_season(mask, T, seed=1)
ic = S.make_ic(mask, seed=1)
params = S.ensemble_params(M, seed=1)
eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
eng.set_path("ensemble")
eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
sets = {"all": _lib.OUTPUT_NAMES, "no_cum": tuple(n for n in _lib.OUTPUT_NAMES if n not in ("snowAcc", "snowOcean")), "density": ("density",)}
for label, names in sets.items():
    out = eng.alloc_outputs(names=names)
    os.environ.pop("NESOSIM_ENS_TIMING", None)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_season(params, ic, out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nbytes = M * mask.size * (T - 1) * 8 * {"all": 12, "no_cum": 10, "density": 1}[label]
    print("variant=%s cluster=%s M=%d outputs=%s: %.3f ms/season (best of %s) -> %.0f GB/s" % (
        os.environ.get("NESOSIM_ENS_VARIANT", "auto"), os.environ.get("NESOSIM_ENS_CLUSTER", "auto"), M, label,
        min(ts[1:]), ["%.2f" % t for t in ts], nbytes / min(ts[1:]) / 1e6), file=sys.stderr, flush=True)
    os.environ["NESOSIM_ENS_TIMING"] = "1"
    eng.run_season(params, ic, out)
    torch.cuda.synchronize()
    del out
    torch.cuda.empty_cache()
