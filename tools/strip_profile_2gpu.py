"""Two row strips of the 5 km grid on TWO GPUs of one box, in ONE process, stepped day by day (strip A day x, strip B
day x, strip A day x+1, ...) -- the same day kernel, mailboxes and flags as the one-process-per-GPU production run, but
profilable: ncu serialises the kernels of a process in launch order, and in this order no kernel ever waits for a
flag that an EARLIER launch has not already raised.  Prints whether the owned rows equal a one-GPU run.

  ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum \
      --clock-control none -k regex:day_step_strip --csv --log-file out.csv python tools/strip_profile_2gpu.py
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nesosim_b200 import domain, synthetic as S
from nesosim_b200.engine import SnowBudgetEngine

assert torch.cuda.device_count() >= 2, "needs two GPUs"
rt = ctypes.CDLL("libcudart.so.12")
for a, b in ((0, 1), (1, 0)):
    torch.cuda.set_device(a)
    torch.zeros(1, device="cuda:%d" % a)
    rc = rt.cudaDeviceEnablePeerAccess(b, 0)
    assert rc in (0, 704), rc          # 704 = already enabled
dx, T = 5000, 6
mask = S.region_mask(dx=dx)
gen = S.make_season(mask, 2, seed=7)
idx = np.arange(T) % 2
forcing = {k: v[idx] for k, v in gen.items() if k != "temp"}
ic = S.make_ic(mask, seed=7)
params = [5.8e-7, 5., 1.45e-7, 2.2e-8]
strips = []
for r in range(2):
    torch.cuda.set_device(r)
    eng, lo, hi, elo, ehi = domain.make_strip_engine(mask, T, dx, forcing, r, 2, device=r, atmlossInc=1, timeout_s=20.0)
    strips.append([eng, lo, hi, elo, ehi, eng.alloc_outputs(), np.ascontiguousarray(ic[elo:ehi])])
blocks = [s[0].strip_block() for s in strips]
strips[0][0].strip_connect_local(None, blocks[1])
strips[1][0].strip_connect_local(blocks[0], None)
for x in range(T - 1):
    for r, s in enumerate(strips):
        torch.cuda.set_device(r)
        s[0].run_season([params], s[6], s[5], first_step=x, num_steps=1)
for r in range(2):
    torch.cuda.synchronize(r)
assert not any(s[0].strip_timed_out() for s in strips)
# one GPU, whole grid
torch.cuda.set_device(0)
one = SnowBudgetEngine(mask, T, dx, n_members=1, device=0, atmlossInc=1)
one.set_path("general")
one.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
ref = one.run_season([params], ic)
ok = True
for eng, lo, hi, elo, ehi, out, _ in strips:
    for k, v in out.items():
        mine = v[0][..., lo - elo:lo - elo + (hi - lo), :].to("cuda:0")
        ok = ok and bool(torch.equal(torch.nan_to_num(mine, nan=-7.0), torch.nan_to_num(ref[k][0][..., lo:hi, :], nan=-7.0)))
print("two GPUs, one process, day by day: identical to one GPU:", ok)
print("algorithmic bytes per day and direction: %d (2 layers x 2 rows x %d columns x 8 B) + 8 B flag" % (2 * 2 * mask.shape[1] * 8, mask.shape[1]))
