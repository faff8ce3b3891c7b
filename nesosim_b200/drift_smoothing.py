"""Drop-ins for the reference's offline drift regridders ``utils.int_smooth_drifts_v2`` / ``int_smooth_drifts_v3``
(/root/reference/source/utils.py:259-297 and 299-338), same names, arguments and return value.

What they do in the reference: drop the source points flagged by its pole-hole "bodge" mask (NaN drift AND latitude
above 90 degrees -- never true, so every point stays and NaNs flow on; kept as it is), interpolate both drift
components linearly from the scattered source points onto the model grid (scipy ``griddata`` in v2, a
``LinearNDInterpolator`` over a Delaunay triangulation the caller built once in v3), smooth each gridded component
with astropy's ``convolve(Gaussian2DKernel(sigma_factor, x_size=3))`` -- the one place where NaNs reach the
convolution, i.e. its NaN-interpolating branch -- and mask the result where the gridded input was NaN.

Here the interpolation stays on the CPU (scipy, as in the reference) and the smoothing tail runs on the GPU through
``nesosim_smooth`` (``engine.smooth_gridded_drift``).  ``smoother`` replaces that tail (tests on a machine without a GPU
pass the CPU restatement of the convolution); there is no CPU fallback: without it the native library is required."""
import numpy as np


def _kept_points(xptsF, yptsF, latsF, driftFmon):
    """Source coordinates and drift components of the points the reference keeps (utils.py:268-276, 310-318)."""
    flagged = np.isnan(np.asarray(driftFmon[1], dtype=np.float64)) & (np.asarray(latsF) > 90)
    keep = ~flagged
    return (np.asarray(xptsF)[keep], np.asarray(yptsF)[keep],
            np.asarray(driftFmon[0])[keep], np.asarray(driftFmon[1])[keep])


def _smooth_and_mask(driftFGx, driftFGy, sigma_factor, x_size_val, smoother):
    if smoother is None:
        from .engine import smooth_gridded_drift as smoother
    return smoother(driftFGx, driftFGy, sigma_factor=sigma_factor, x_size_val=x_size_val)


def int_smooth_drifts_v2(xptsG, yptsG, xptsF, yptsF, latsF, driftFmon, sigma_factor=1, x_size_val=3, truncate=1,
                         smoother=None):
    """``utils.int_smooth_drifts_v2`` (utils.py:259-297).  ``truncate`` is accepted and unused, as in the reference.
    Returns the masked array ``(2, nx, ny)`` of smoothed drift components on the model grid."""
    from scipy.interpolate import griddata
    x, y, u, v = _kept_points(xptsF, yptsF, latsF, driftFmon)
    driftFGx = griddata((x, y), u, (xptsG, yptsG), method='linear')
    driftFGy = griddata((x, y), v, (xptsG, yptsG), method='linear')
    return _smooth_and_mask(driftFGx, driftFGy, sigma_factor, x_size_val, smoother)


def int_smooth_drifts_v3(tri, xptsG, yptsG, xptsF, yptsF, latsF, driftFmon, sigma_factor=1, x_size_val=3, truncate=1,
                         smoother=None):
    """``utils.int_smooth_drifts_v3`` (utils.py:299-338): as v2, interpolating over the caller's Delaunay
    triangulation ``tri`` of the kept source points."""
    from scipy.interpolate import LinearNDInterpolator
    _, _, u, v = _kept_points(xptsF, yptsF, latsF, driftFmon)
    driftFGx = LinearNDInterpolator(tri, u.flatten())((xptsG, yptsG))
    driftFGy = LinearNDInterpolator(tri, v.flatten())((xptsG, yptsG))
    return _smooth_and_mask(driftFGx, driftFGy, sigma_factor, x_size_val, smoother)
