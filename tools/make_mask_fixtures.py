"""Regrid the NSIDC region mask onto the 100 km and 25 km model grids exactly as NESOSIM.main does
(NESOSIM.py:535-536) and store the result as package data for the synthetic benchmark / tests.

Run in the build container (needs /root/reference/anc_data):  python tools/make_mask_fixtures.py
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesosim_b200 import grid

ANC = '/root/reference/anc_data/'
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'nesosim_b200', 'data')
os.makedirs(out, exist_ok=True)
for dx in (100000, 25000):
    x, y, lat, lon, p = grid.create_grid(dxRes=dx)
    m = grid.region_mask_on_grid(ANC, x, y, p)
    m8 = m.astype(np.uint8)
    assert np.array_equal(m8, m)
    np.save(os.path.join(out, 'region_mask_%dkm.npy' % (dx // 1000)), m8)
    n = m8.size
    print(dx, m8.shape, 'ocean %.3f land %.3f lake %.3f' % (((m8 >= 1) & (m8 <= 10)).sum() / n, (m8 > 10).sum() / n, (m8 < 1).sum() / n),
          'corner', float(x[0, 0]), float(y[0, 0]))
