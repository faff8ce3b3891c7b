"""Times the REFERENCE'S OWN calcBudget loop, executed verbatim from /root/reference (oracle/ref_loader.py, astropy
replaced by the restated convolve), on one core of the build container, next to the numpy port bench.py uses as its CPU
baseline on the GPU box (where /root/reference does not exist).  Writes profiles/r02_reference_verbatim.json.
usage: OMP_NUM_THREADS=1 python tools/time_reference_verbatim.py"""
import json, os, sys, time
os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from nesosim_b200 import synthetic as S
from oracle import nesosim_oracle as O, ref_loader

mask = S.region_mask(dx=100000)
T = 260
F = S.make_season(mask, T, seed=2024)
ic = S.make_ic(mask, seed=2024)
p = (5.8e-7, 5., 1.45e-7, 2.2e-8)
ref = ref_loader.load_reference()
ref_loader.set_globals(ref, *p)
cells = mask.size * (T - 1)
res = {}
for name, fn in (("reference_verbatim", lambda: ref_loader.run_reference_season(ref, F, ic, mask.astype(np.float64), 100000, dict(atmlossInc=1))),
                 ("numpy_port", lambda: O.run_season(F, ic, mask, 100000, O.Params(*p), O.Flags(atmlossInc=1)))):
    best = None
    with np.errstate(all="ignore"):
        for _ in range(3):
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    res[name] = {"seconds_per_member_season": best, "cell_days_per_s_per_core": cells / best}
    res[name + "_out"] = out
same = all(np.array_equal(res["reference_verbatim_out"][k], res["numpy_port_out"][k], equal_nan=True)
           for k in ("snowDepths", "density", "snowAdv", "snowLead"))
line = {"workload": "one member-season, 100 km grid (90x90), 260 days, run_multiseason parameters, 1 core (OMP_NUM_THREADS=1)",
        "where": "build container (no GPU); the GPU box has no /root/reference",
        "reference_verbatim": res["reference_verbatim"], "numpy_port": res["numpy_port"],
        "port_over_verbatim": res["numpy_port"]["cell_days_per_s_per_core"] / res["reference_verbatim"]["cell_days_per_s_per_core"],
        "identical_outputs": bool(same), "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")}
json.dump(line, open(os.path.join(ROOT, "profiles", "r02_reference_verbatim.json"), "w"), indent=1)
print(json.dumps(line))
