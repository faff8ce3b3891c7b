"""Throughput of the general per-day path on the larger grids: single season, M members, for any set of build / launch
variants of the day kernels, checking that all variants agree.
usage: python tools/general_timing.py n days [members] [generated_days]
VARIANTS="ENV=val,ENV=val;ENV=val;..." -- one run per ';'-separated entry with those environment variables set, e.g.
  NESOSIM_DAY_THREADS=256|512, NESOSIM_LAND_SHORTCUT=0|1, NESOSIM_PDL=0|1
MASK=land|ocean replaces the mask by all land / all ocean (decomposition experiments); MASK=disc the synthetic disc."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nesosim_b200 import synthetic as S
from nesosim_b200.engine import SnowBudgetEngine
n = int(sys.argv[1]); T = int(sys.argv[2]); M = int(sys.argv[3]) if len(sys.argv) > 3 else 1
G = int(sys.argv[4]) if len(sys.argv) > 4 else min(T, 8)
dx = {90: 100000, 357: 25000, 1785: 5000}.get(n, 5000)
mask = S.region_mask(dx=dx) if n in (90, 357, 1785) and os.environ.get("MASK") != "disc" else S.region_mask(shape=(n, n), kind="disc")
gen = S.make_season(mask, G, seed=1)
if os.environ.get("MASK") == "land":
    mask = np.full_like(mask, 11)
elif os.environ.get("MASK") == "ocean":
    mask = np.full_like(mask, 8)
idx = np.arange(T) % G
F = {k: (v if v is None else torch.from_numpy(v[idx]).cuda()) for k, v in gen.items()}
ic = S.make_ic(mask, seed=1)
params = S.ensemble_params(M, seed=1)
cells = M * n * n * (T - 1)
land = float(np.mean((mask > 10) | (mask < 1)))
first = None
KEYS = ("NESOSIM_DAY_THREADS", "NESOSIM_LAND_SHORTCUT", "NESOSIM_PDL")
for variant in os.environ.get("VARIANTS", "NESOSIM_DAY_THREADS=256;NESOSIM_DAY_THREADS=512").split(";"):
    for k in KEYS:
        os.environ.pop(k, None)
    for kv in filter(None, variant.split(",")):
        k, v = kv.split("=", 1)
        os.environ[k] = v.replace("/", ",")
    eng = SnowBudgetEngine(mask, T, dx, n_members=M, atmlossInc=1)
    eng.set_path("general")
    eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    out = eng.alloc_outputs()
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run_season(params, ic, out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    digest = {k: (float(torch.nansum(v).item()), int(torch.isnan(v).sum().item())) for k, v in out.items()}
    if first is None:
        first = digest
    print(json.dumps({"grid": [n, n], "days": T, "members": M, "land_fraction": round(land, 3), "variant": variant,
                      "ms_per_season": ms, "us_per_day": 1e3 * ms / (T - 1),
                      "cell_days_per_s": cells / ms * 1e3, "algorithmic_GBs": cells * (96 + 41.0 / M) / ms / 1e6,
                      "frac_of_6551": cells * (96 + 41.0 / M) / ms / 1e6 / 6551, "same_digest_as_first": digest == first}),
          flush=True)
    del out
    eng.close()
    torch.cuda.empty_cache()
