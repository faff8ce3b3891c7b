"""Helpers shared by the golden-vector tests (CPU oracle and GPU)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT_NAMES = ("snowDepths", "density", "snowAcc", "snowOcean", "snowAdv", "snowDiv", "snowLead", "snowAtm",
             "snowWindPackLoss", "snowWindPackGain", "snowWindPack")
SMALL_CASES = {
    "oneseason": dict(atmlossInc=0),
    "multiseason": dict(atmlossInc=1),
    "nodyn": dict(atmlossInc=1, dynamicsInc=0),
    "clim": dict(atmlossInc=1, densityType="clim"),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def canon_sha(a):
    b = np.array(a, dtype=np.float64, copy=True)
    b[np.isnan(b)] = np.nan
    b = b + 0.0
    return hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest()


def assert_identical(got, ref, name, rtol=1e-10):
    """Parity bar: identical NaN masks, every finite value within 1e-10 relative -- and in fact equal."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, name
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN mask differs: " + name
    fin = ~np.isnan(ref)
    assert np.allclose(got[fin], ref[fin], rtol=rtol, atol=1e-300), "outside 1e-10 relative: " + name
    assert np.array_equal(got, ref, equal_nan=True), "not value-identical: " + name
