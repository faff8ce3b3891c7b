"""CPU restatement of the two astropy.convolution symbols NESOSIM's hot path calls.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

Third-party dependency restated here: **astropy** (module ``astropy.convolution``), *unpinned* in the
reference (``/root/reference/environment.yml:18`` is the bare string ``astropy``; with ``python=3.7.*`` at
``environment.yml:9`` the newest installable release is astropy 4.3.1).  astropy is not installed in this image
and there is no network, so the algorithm is restated from its published behaviour and anchored on the
reference's own call site, ``smooth_snow`` (``/root/reference/source/NESOSIM.py:170-187``):

    kernel = Gaussian2DKernel(x_stddev=1, x_size=3, y_size=3)
    arr    = convolve(arr, kernel)          # all defaults

PARITY UNPINNED at this boundary: the reference ships no golden vectors or tests, so the
restatement cannot be checked against astropy output.  Two published normalisation orders are provided
(they differ by <= a few ulp, far below the 1e-10 parity tolerance):

``post_divide`` (default; astropy 3.1 - 4.x ``convolve.py``)
    plain branch   : out = (sum_k pad*g_k) / g.sum()     (C loop with the raw kernel, ``result /= kernel_sum`` after)
    NaN branch     : out = top/bot, ``bot==0 -> arr[i,j]``  (normalisation inside the C loop, raw kernel)
``pre_normalised`` (astropy >= 5: the kernel is divided by its sum before the C loop)
    plain branch   : out = sum_k pad*(g_k/g.sum())
    NaN branch     : out = top/bot with the normalised weights
"""
import numpy as np

__all__ = ["gaussian2d_kernel", "convolve_fill0", "nan_interpolate_flag"]


def gaussian2d_kernel(x_stddev=1, x_size=3, y_size=3, y_stddev=None, theta=0.0):
    """``Gaussian2DKernel(x_stddev, x_size=, y_size=).array`` (mode='center').

    Follows astropy.modeling ``Gaussian2D.evaluate`` with ``amplitude = 1/(2*pi*sx*sy)`` sampled at integer
    offsets ``arange(-(size//2), size//2+1)`` on a ``meshgrid``; the call site is ``NESOSIM.py:184`` (note the
    reference passes ``y_size=x_size_val``).
    """
    if y_stddev is None:
        y_stddev = x_stddev
    x_stddev = float(x_stddev)
    y_stddev = float(y_stddev)
    amplitude = 1. / (2 * np.pi * x_stddev * y_stddev)
    xs = np.arange(-(int(x_size) // 2), int(x_size) // 2 + 1)
    ys = np.arange(-(int(y_size) // 2), int(y_size) // 2 + 1)
    x, y = np.meshgrid(xs, ys)
    x = x.astype(float)
    y = y.astype(float)
    cost2 = np.cos(theta) ** 2
    sint2 = np.sin(theta) ** 2
    sin2t = np.sin(2. * theta)
    xstd2 = x_stddev ** 2
    ystd2 = y_stddev ** 2
    xdiff = x - 0.0
    ydiff = y - 0.0
    a = 0.5 * ((cost2 / xstd2) + (sint2 / ystd2))
    b = 0.5 * ((sin2t / xstd2) - (sin2t / ystd2))
    c = 0.5 * ((sint2 / xstd2) + (cost2 / ystd2))
    return amplitude * np.exp(-((a * xdiff ** 2) + (b * xdiff * ydiff) + (c * ydiff ** 2)))


def nan_interpolate_flag(arr):
    """astropy: ``nan_interpolate = (nan_treatment == 'interpolate') and np.isnan(array.sum())``."""
    with np.errstate(all="ignore"):
        return bool(np.isnan(np.asarray(arr, dtype=float).sum()))


def convolve_fill0(arr, kernel, variant="post_divide"):
    """``astropy.convolution.convolve(arr, kernel)`` with all defaults, for an odd-sized 2-D kernel.

    Defaults: boundary='fill', fill_value=0.0, nan_treatment='interpolate', normalize_kernel=True,
    mask=None, preserve_nan=False.  Per output cell the C loop runs ii (rows) outer, jj (columns) inner over
    the zero-padded input, multiplies by the *flipped* kernel ``g[nky-1-ii, nkx-1-jj]`` and accumulates into a
    double that starts at 0.0 -- reproduced here with whole-array shifted slices in the same tap order, so
    every output element sees the same sequence of IEEE operations.
    """
    f = np.array(arr, dtype=float, order="C")
    g = np.array(kernel, dtype=float, order="C")
    nky, nkx = g.shape
    assert nky % 2 == 1 and nkx % 2 == 1, "convolve requires odd kernel axes"
    wy, wx = nky // 2, nkx // 2
    ny, nx = f.shape
    ksum = g.sum()
    if variant == "pre_normalised":
        g = g / ksum
    elif variant != "post_divide":
        raise ValueError(variant)
    interp = nan_interpolate_flag(f)

    pad = np.zeros((ny + 2 * wy, nx + 2 * wx), dtype=float)
    pad[wy:wy + ny, wx:wx + nx] = f

    top = np.zeros((ny, nx), dtype=float)
    with np.errstate(all="ignore"):
        if not interp:
            for ii in range(nky):
                for jj in range(nkx):
                    top += pad[ii:ii + ny, jj:jj + nx] * g[nky - 1 - ii, nkx - 1 - jj]
            if variant == "post_divide":
                top /= ksum
            return top
        bot = np.zeros((ny, nx), dtype=float)
        for ii in range(nky):
            for jj in range(nkx):
                val = pad[ii:ii + ny, jj:jj + nx]
                ker = g[nky - 1 - ii, nkx - 1 - jj]
                ok = ~np.isnan(val)
                # masked in-place adds: cells whose tap is NaN skip both accumulations
                np.add(top, val * ker, out=top, where=ok)
                np.add(bot, ker, out=bot, where=ok)
        out = np.where(bot == 0, f, top / bot)
    return out


def convolve_fill0_scalar(arr, kernel, variant="post_divide"):
    """Literal per-cell loop form of :func:`convolve_fill0` (slow; used by tests to pin the vectorised form)."""
    f = np.array(arr, dtype=float, order="C")
    g = np.array(kernel, dtype=float, order="C")
    nky, nkx = g.shape
    wy, wx = nky // 2, nkx // 2
    ny, nx = f.shape
    ksum = g.sum()
    if variant == "pre_normalised":
        g = g / ksum
    interp = nan_interpolate_flag(f)
    out = np.zeros_like(f)
    for i in range(ny):
        for j in range(nx):
            top = 0.0
            bot = 0.0
            for ii in range(nky):
                for jj in range(nkx):
                    r, c = i + ii - wy, j + jj - wx
                    val = f[r, c] if (0 <= r < ny and 0 <= c < nx) else 0.0
                    ker = g[nky - 1 - ii, nkx - 1 - jj]
                    if interp:
                        if not np.isnan(val):
                            top += val * ker
                            bot += ker
                    else:
                        top += val * ker
            if interp:
                out[i, j] = f[i, j] if bot == 0 else top / bot
            else:
                out[i, j] = top / ksum if variant == "post_divide" else top
    return out
