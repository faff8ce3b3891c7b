"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol the header declares,
validates arguments, and refuses to compute without a device (no CPU fallback).  Also the exact
constant-division routine (host replica) against true division."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "nesosim_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nesosim_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from nesosim_b200 import _lib
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes table and header disagree"
    assert lib.nesosim_abi_version() == 1


def test_struct_layout_matches_header(lib):
    from nesosim_b200 import _lib
    # sizes implied by the header: 4 int32 + 6 double + 9 double + 1 double + 6 int32
    assert C.sizeof(_lib.Config) == 16 + 8 * 16 + 24
    assert C.sizeof(_lib.MemberParams) == 32
    assert C.sizeof(_lib.Outputs) == 11 * 8 + 16


def _cfg(**kw):
    from nesosim_b200 import _lib
    from nesosim_b200.engine import conv_constants
    c = _lib.Config()
    c.ny, c.nx, c.num_days, c.n_members = 4, 5, 3, 1
    c.dx, c.deltaT = 100000., 86400.
    c.snowDensityFresh, c.snowDensityOld, c.minSnowD, c.minConc = 200., 350., 0.02, 0.15
    w, d = conv_constants()
    c.conv_weights = (C.c_double * 9)(*w.tolist())
    c.conv_divisor = d
    c.dynamicsInc = c.leadlossInc = c.windpackInc = 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_argument_validation_and_no_cpu_fallback(lib):
    from nesosim_b200 import _lib
    mask = np.full((4, 5), 8, dtype=np.uint8)
    h = C.c_void_p()
    mp = mask.ctypes.data_as(C.c_void_p)
    assert lib.nesosim_create(None, mp, C.byref(h)) == _lib.ERR_ARG
    assert lib.nesosim_create(C.byref(_cfg(ny=1)), mp, C.byref(h)) == _lib.ERR_ARG
    assert b"np.gradient" in lib.nesosim_last_error()
    assert lib.nesosim_create(C.byref(_cfg(num_days=1)), mp, C.byref(h)) == _lib.ERR_ARG
    assert lib.nesosim_create(C.byref(_cfg(n_members=0)), mp, C.byref(h)) == _lib.ERR_ARG
    assert lib.nesosim_create(C.byref(_cfg(dx=0.0)), mp, C.byref(h)) == _lib.ERR_ARG
    if lib.nesosim_device_count() == 0:
        # the product path must fail loudly without a GPU
        assert lib.nesosim_create(C.byref(_cfg()), mp, C.byref(h)) == _lib.ERR_CUDA
        assert b"no CPU path" in lib.nesosim_last_error()
        with pytest.raises(_lib.NesosimError):
            from nesosim_b200.engine import SnowBudgetEngine
            SnowBudgetEngine(mask, 3, 100000)
    assert lib.nesosim_destroy(None) == 0
    assert lib.nesosim_smooth(None, None, 3, 3, None, 1.0, None) == _lib.ERR_ARG
    # strips of a decomposed grid: nothing to set up, export or query without a context
    assert lib.nesosim_strip_setup(None, 1, 0) == _lib.ERR_ARG
    assert lib.nesosim_strip_export(None, C.create_string_buffer(64)) == _lib.ERR_ARG
    assert lib.nesosim_strip_status(None, C.byref(C.c_int(0))) == _lib.ERR_ARG
    assert lib.nesosim_strip_set_timeout(None, 1.0) == _lib.ERR_ARG
    assert lib.nesosim_strip_block(None, None, None) == _lib.ERR_STATE
    assert lib.nesosim_strip_connect(None, None, None) == _lib.ERR_STATE
    assert lib.nesosim_strip_connect_local(None, None, None) == _lib.ERR_STATE
    # asynchronous and calibration modes
    assert lib.nesosim_set_async(None, 1) == _lib.ERR_ARG
    assert lib.nesosim_sync(None, None) == _lib.ERR_ARG
    assert lib.nesosim_set_observations(None, 0, None, None, None, None) == _lib.ERR_ARG
    assert lib.nesosim_run_season_misfit(None, None, None, 0, None, None, None) == _lib.ERR_ARG


def test_missing_library_is_an_import_error(monkeypatch):
    from nesosim_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnesosim_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_constant_division_is_exact(lib):
    from nesosim_b200.engine import conv_constants
    ksum = conv_constants()[1]
    rng = np.random.default_rng(0)
    divisors = [100000., 200000., 50000., 25000., 10000., 5000., 200., ksum, 1.0, 298.94, 3.0, 1e5 / 3]
    for c in divisors:
        fast = lib.nesosim_const_div_is_fast(c)
        xs = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.integers(-12, 12, 20000),
                             rng.random(5000), [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, -1e250, 5e-324, c, -c,
                                                3 * c, c / 3, np.nextafter(c, 0), np.nextafter(c, 2 * c)]])
        with np.errstate(all="ignore"):
            want = xs / c
        got = np.array([lib.nesosim_const_div_eval_host(float(x), c) for x in xs])
        assert np.array_equal(got, want, equal_nan=True), (c, fast)
        assert np.array_equal(np.signbit(got[~np.isnan(got)]), np.signbit(want[~np.isnan(want)]))
    # grid spacings, the fresh-snow density and the Gaussian kernel sum must all take the 3-operation path
    for c in (100000., 200000., 25000., 50000., 5000., 10000., 200., ksum):
        assert lib.nesosim_const_div_is_fast(c) == 1, c


@pytest.mark.parametrize("pps,T,offset", [(1, 7, 0), (2, 5, 0), (1, 2, 0), (2, 9, 1)])
def test_unpack_member_array_layout(lib, pps, T, offset):
    """The host half of the compacted drain (drain_kernels.cuh layout): ocean cells of every plane + land cells of the
    first three slots -> full planes, later slots repeating the third slot's land cells; NaNs keep their bits; a
    destination that is only 8-byte aligned (odd offset) takes the unaligned path of the streaming stores."""
    rng = np.random.default_rng(pps * 100 + T)
    ny, nx = 9, 11
    mask = rng.choice(np.array([0, 3, 8, 8, 11, 12], dtype=np.uint8), size=(ny, nx))
    plane = ny * nx
    land = ((mask > 10) | (mask < 1)).ravel()
    full = rng.normal(size=(T, pps, plane))
    full[rng.random(full.shape) < 0.1] = np.nan
    head = min(3, T)
    for s in range(head, T):                      # the contract of the packed form: land constant after the head slots
        full[s][:, land] = full[head - 1][:, land]
    planes = full.reshape(T * pps, plane)
    packed = np.concatenate([planes[:, ~land].ravel(), planes[:head * pps][:, land].ravel()])
    buf = np.full(T * pps * plane + 2, -7.0)
    dst = buf[offset:offset + T * pps * plane]
    rc = lib.nesosim_unpack_member_array(mask.ctypes.data_as(C.c_void_p), plane, pps, T,
                                         packed.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(dst.view(np.uint64), planes.ravel().view(np.uint64))
    assert buf[offset + T * pps * plane:].tolist() == [-7.0] * (2 - offset) and (offset == 0 or buf[0] == -7.0)
    assert lib.nesosim_unpack_member_array(None, plane, pps, T, packed.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p)) != 0
