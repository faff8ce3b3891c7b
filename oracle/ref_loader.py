"""Stub-import the *unmodified* reference ``NESOSIM.py`` so its hot-path functions run verbatim.

TEST INFRASTRUCTURE ONLY, and only usable where ``/root/reference`` exists (the build container) -- it is
what pins ``oracle/nesosim_oracle.py`` and what ``tests/golden/make_golden.py`` uses to generate fixtures.
Nothing that runs on the GPU box imports this module.

``import NESOSIM`` needs xarray, cartopy, astropy, netCDF4, pyproj and matplotlib (NESOSIM.py:37-48,
utils.py:21-34); none is installed.  None is touched by ``calcBudget`` and its callees except the two
astropy symbols, so the I/O libraries are replaced by empty stub modules and ``astropy.convolution`` by the
restatement in ``oracle/astropy_restated.py``.  No reference source is copied.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_SOURCE = "/root/reference/source"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SOURCE, "NESOSIM.py"))


class _Any:
    """Placeholder for any attribute of a stubbed I/O library (default-argument values such as
    ``plt.cm.viridis`` are evaluated when the reference's ``utils.py`` is imported)."""

    def __getattr__(self, name):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()


def _stub(name, **attrs):
    m = _StubModule(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference(conv_variant="post_divide"):
    """Return the reference's ``NESOSIM`` module object (imported from /root/reference, not copied)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_SOURCE)
    from . import astropy_restated as ar

    class _Gaussian2DKernel:
        def __init__(self, x_stddev, y_stddev=None, theta=0.0, **kw):
            self.array = ar.gaussian2d_kernel(x_stddev=x_stddev, y_stddev=y_stddev, theta=theta,
                                              x_size=kw["x_size"], y_size=kw["y_size"])

    def _convolve(array, kernel):
        return ar.convolve_fill0(array, kernel.array, variant=conv_variant)

    for name in ("xarray", "netCDF4", "pyproj", "matplotlib", "matplotlib.pyplot", "matplotlib.colorbar",
                 "matplotlib.cm", "cartopy", "astropy"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["matplotlib"].colorbar = sys.modules["matplotlib.colorbar"]
    _stub("cartopy.crs", Projection=object, NorthPolarStereo=lambda *a, **k: None)
    sys.modules["cartopy"].crs = sys.modules["cartopy.crs"]
    _stub("astropy.convolution", convolve=_convolve, Gaussian2DKernel=_Gaussian2DKernel)
    sys.modules["astropy"].convolution = sys.modules["astropy.convolution"]

    if REFERENCE_SOURCE not in sys.path:
        sys.path.insert(0, REFERENCE_SOURCE)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mod in ("utils", "NESOSIM"):
            sys.modules.pop(mod, None)
        ref = importlib.import_module("NESOSIM")
    # the reference rebinds these names on import; make the stubbed pair explicit
    ref.convolve = _convolve
    ref.Gaussian2DKernel = _Gaussian2DKernel
    return ref


def set_globals(ref, windPackFactor, windPackThresh, leadLossFactor, atmLossFactor, ancDataPath=None):
    """What ``main`` does at NESOSIM.py:527-541 (the hot path reads these module globals)."""
    ref.snowDensityFresh = 200.
    ref.snowDensityOld = 350.
    ref.minSnowD = 0.02
    ref.minConc = 0.15
    ref.deltaT = 60. * 60. * 24.
    ref.leadLossFactor = leadLossFactor
    ref.windPackThresh = windPackThresh
    ref.windPackFactor = windPackFactor
    ref.atmLossFactor = atmLossFactor
    if ancDataPath is not None:
        ref.ancDataPath = ancDataPath


def run_reference_season(ref, forcing, ic, mask, dx, flags, num_steps=None, day_of_year=None):
    """Drive the reference's own ``genEmptyArrays`` + ``calcBudget`` exactly as ``main``'s loop does
    (NESOSIM.py:586-649) on pre-loaded forcing; returns the 15 arrays in a dict keyed by the reference's names."""
    import numpy as np
    T, ny, nx = forcing["precip"].shape
    names = ("precipDays", "iceConcDays", "windDays", "tempDays", "snowDepths", "density", "snowDiv",
             "snowAdv", "snowAcc", "snowOcean", "snowWindPack", "snowWindPackLoss", "snowWindPackGain",
             "snowLead", "snowAtm")
    s = dict(zip(names, ref.genEmptyArrays(T, ny, nx)))
    temp = forcing.get("temp")
    if temp is None:
        temp = np.full((T, ny, nx), np.nan)
    if ic is not None:
        icd = np.array(ic, dtype=float)
        icd[np.where(forcing["conc"][0] < ref.minConc)] = 0
        s["snowDepths"][0, 0] = icd * 0.5
        s["snowDepths"][0, 1] = icd * 0.5
    steps = T - 1 if num_steps is None else num_steps
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for x in range(steps):
            dayT = 1 + x if day_of_year is None else day_of_year[x]
            ref.calcBudget(None, None, s["snowDepths"], forcing["conc"][x], forcing["precip"][x],
                           forcing["drift"][x], forcing["wind"][x], temp[x], s["density"], s["precipDays"],
                           s["iceConcDays"], s["windDays"], s["tempDays"], s["snowAcc"], s["snowOcean"],
                           s["snowAdv"], s["snowDiv"], s["snowLead"], s["snowAtm"], s["snowWindPackLoss"],
                           s["snowWindPackGain"], s["snowWindPack"], mask, dx, x, dayT,
                           densityType=flags.get("densityType", "variable"),
                           dynamicsInc=flags.get("dynamicsInc", 1), leadlossInc=flags.get("leadlossInc", 1),
                           windpackInc=flags.get("windpackInc", 1), atmlossInc=flags.get("atmlossInc", 0))
    if steps == T - 1:
        s["precipDays"][T - 1] = forcing["precip"][T - 1]
        s["iceConcDays"][T - 1] = forcing["conc"][T - 1]
        s["windDays"][T - 1] = forcing["wind"][T - 1]
        s["tempDays"][T - 1] = temp[T - 1]
    return s
