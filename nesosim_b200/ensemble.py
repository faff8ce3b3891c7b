"""Calibration-ensemble driver (the reference has no counterpart -- it only runs one parameter set
per ``main`` call): M parameter sets over one forcing, sharded over the ranks of a box, with the outputs selected and
reduced ON THE DEVICE so that only what a calibration needs crosses PCIe (the full contract is 96 B per member-cell-day,
192 GB for 1024 members).

    misfit = run_ensemble(mask, forcing, ic, params, dx, obs=(day_idx, row_idx, col_idx, depth_obs))

Each rank runs its block of members through ``nesosim_run_season`` with only ``snowDepths`` requested, evaluates the
modelled snow depth over ice ``(h0+h1)/iceConc`` (what ``main`` writes as ``snow_depth``, NESOSIM.py:654) at the
observation points, and returns the per-member sum of squared differences; with ``world > 1`` the ranks' results are
all-gathered into member order (the only communication).
"""
import numpy as np

from . import sharding


def depth_misfit(snowDepths, conc, obs):
    """Per-member sum of squared (modelled - observed) snow depth over ice at the observation points.
    ``snowDepths``: CUDA tensor (M,T,2,ny,nx); ``conc``: CUDA tensor (T,ny,nx); ``obs`` = (day, row, col, depth) arrays.
    NaN model values (land, missing forcing) are skipped, as a calibration against IceBridge-style data would."""
    import torch
    day, row, col, depth = (torch.as_tensor(np.asarray(a), device=snowDepths.device) for a in obs)
    day, row, col = day.long(), row.long(), col.long()
    h = snowDepths[:, day, 0, row, col] + snowDepths[:, day, 1, row, col]          # (M, n_obs)
    model = h / conc[day, row, col].unsqueeze(0)
    diff = model - depth.to(torch.float64).unsqueeze(0)
    ok = torch.isfinite(diff)
    return torch.where(ok, diff * diff, torch.zeros_like(diff)).sum(dim=1), ok.sum(dim=1)


def run_ensemble(mask, forcing, ic, params, dx, obs, rank=0, world=1, device=0, group=None, fused=None, **flags):
    """Returns (misfit[M], n_used[M]) as numpy arrays in member order (identical on every rank).  ``fused``: None =
    inside the season kernel where it applies, True = insist on it, False = depths to HBM + reduction there (the
    round-1 path, kept as the cross-check)."""
    import torch
    from .engine import SnowBudgetEngine
    params = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    M = len(params)
    lo, hi = sharding.member_range(M, rank, world)
    T = forcing["precip"].shape[0]
    eng = SnowBudgetEngine(mask, T, dx, n_members=hi - lo, device=device, **flags)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    from . import _lib
    try:
        # fused: the observation operator and the reduction run inside the season-resident kernel, nothing is stored
        mis, used = eng.run_season_misfit(params[lo:hi], ic, obs)
    except _lib.NesosimError as e:
        if fused is True or e.code != _lib.ERR_ARG:
            raise
        # grids the season-resident kernel does not take: depths to HBM, reduced there (still nothing crosses PCIe)
        out = eng.run_season(params[lo:hi], ic, eng.alloc_outputs(names=("snowDepths",)))
        mis, used = depth_misfit(out["snowDepths"], eng._forcing[1], obs)
    if fused is False:
        out = eng.run_season(params[lo:hi], ic, eng.alloc_outputs(names=("snowDepths",)))
        mis, used = depth_misfit(out["snowDepths"], eng._forcing[1], obs)
    eng.close()
    if world > 1:
        mis = sharding.gather_member_results(mis, M, rank, world, group)
        used = sharding.gather_member_results(used, M, rank, world, group)
    return mis.cpu().numpy(), used.cpu().numpy()
