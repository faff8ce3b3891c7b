"""Domain decomposition of one season over several ranks, for grids too large to be worth a single GPU (the 5 km
pan-Arctic case of BASELINE.json; SURVEY.md §8e "Space").

The fused dependency radius of one budget step is two cells (np.gradient r=1 followed by the 3x3 Gaussian r=1, reference
``NESOSIM.py:204-213`` and ``:184-185``), and only the two depth layers cross cell boundaries.  So the grid is cut into
contiguous row strips, every rank runs the ordinary engine on its strip EXTENDED by two ghost rows towards each
neighbour, and after every day the ghost rows of ``snowDepths[x+1]`` are replaced by the neighbour's owned rows -- a
(2 layers x 2 rows x nx) message per neighbour per day (57 KB at nx = 1785).  Nothing else is ever exchanged: the
forcing ghost rows are staged once, accumulators never look sideways.  Owned rows come out value-identical to the
single-domain run: inside the extended strip the stencils of the owned rows only ever touch real data, and the
one-sided differences / zero padding the kernels apply at the edges of the LOCAL grid only reach the ghost rows --
except on the first and last strip, where the local edge IS the global edge.

Two drivers share the decomposition:

* ``run_decomposed_season_peer`` (production): the exchange is FUSED INTO THE DAY KERNEL.  Every strip's boundary CTAs
  store their new depths straight into the neighbour's mailbox in peer memory (NVLink between the GPUs of a box; CUDA
  IPC between the one-process-per-GPU ranks) and raise a flag there; the next day's boundary CTAs wait on their own
  flag.  One native call enqueues the whole season -- one launch per day, no collective, no host round trip.
  ``torch.distributed`` only carries the 64-byte IPC handles once and the barriers around the season.
* ``run_decomposed_season`` (baseline / CPU-testable): one engine call per day followed by a ``torch.distributed``
  batched isend/irecv of the ghost rows (NCCL on GPUs; gloo in the CPU tests).  The stepping is behind a small
  interface so the same driver runs the GPU engine and, in the CPU test of the exchange logic, the numpy oracle.
"""
import numpy as np

from . import sharding

GHOST = 2


def strip_rows(ny, rank, world):
    """Owned global rows [lo, hi) of `rank`, and the extended range [elo, ehi) including ghost rows."""
    lo, hi = sharding.member_range(ny, rank, world)
    return lo, hi, max(lo - GHOST, 0), min(hi + GHOST, ny)


def balanced_cuts(mask, world, ocean_weight=2.3):
    """Row boundaries ``[0, c1, ..., ny]`` of ``world`` strips with (nearly) equal modelled cost instead of equal height:
    a day costs roughly ``ocean_weight`` on an ocean cell for 1 on a land cell (all-land tiles take the closed-form
    shortcut), and the ocean is not spread evenly over the rows of the polar grid -- at eight strips of the 5 km grid
    the heaviest equal-height strip carries 11 % more than the mean, and the slowest strip sets the pace of all.
    Deterministic in the mask, so every rank computes the same cuts; every strip keeps at least 2*GHOST rows."""
    mask = np.asarray(mask)
    ny = mask.shape[0]
    land = (mask > 10) | (mask < 1)
    row_cost = mask.shape[1] + (ocean_weight - 1.0) * (~land).sum(axis=1)
    cum = np.concatenate([[0.0], np.cumsum(row_cost)])
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(cum, cum[-1] * r / world))
        c = max(c, cuts[-1] + 2 * GHOST)
        c = min(c, ny - 2 * GHOST * (world - r))
        cuts.append(c)
    cuts.append(ny)
    if any(b - a < GHOST for a, b in zip(cuts[:-1], cuts[1:])):
        raise ValueError("strips must own at least %d rows (ny=%d over %d ranks)" % (GHOST, ny, world))
    return cuts


def strip_rows_balanced(mask, rank, world):
    """Like ``strip_rows`` but with the cost-balanced cuts of ``balanced_cuts``."""
    cuts = balanced_cuts(mask, world)
    ny = np.asarray(mask).shape[0]
    lo, hi = cuts[rank], cuts[rank + 1]
    return lo, hi, max(lo - GHOST, 0), min(hi + GHOST, ny)


class GpuStripStepper:
    """One strip on one GPU through the C ABI (``nesosim_run_season`` one day at a time, general kernels)."""

    def __init__(self, local_mask, num_days, dx, forcing_local, params_row, ic_local, device=0, **flags):
        from .engine import SnowBudgetEngine
        self.eng = SnowBudgetEngine(local_mask, num_days, dx, n_members=1, device=device, **flags)
        self.eng.set_path("general")
        self.eng.set_forcing(forcing_local["precip"], forcing_local["conc"], forcing_local["wind"], forcing_local["drift"],
                             forcing_local.get("rho_clim"))
        self.params = [list(params_row)]
        self.ic = None if ic_local is None else self.eng._dev(ic_local)      # staged once, not once per day
        self.out = self.eng.alloc_outputs()

    def step(self, x):
        self.eng.run_season(self.params, self.ic, self.out, first_step=x, num_steps=1)

    def depths(self, slot):
        """Writable view (2, rows, nx) of snowDepths[slot] (a torch tensor on the strip's device)."""
        return self.out["snowDepths"][0, slot]

    def result(self, rows):
        import torch
        torch.cuda.synchronize()
        return {k: v[0][..., rows, :].cpu().numpy() for k, v in self.out.items()}


def exchange_ghost_rows(depths, rank, world, top_ghost, bottom_ghost, group=None):
    """Swap boundary rows of one (2, rows, nx) depth slot with the neighbouring ranks.

    ``top_ghost`` / ``bottom_ghost``: number of ghost rows this strip has above / below (0 at the global edges).
    Sends the first / last ``GHOST`` OWNED rows, receives into the ghost rows.  All four transfers are posted as one
    batch so NCCL runs them as a single grouped operation."""
    import torch
    import torch.distributed as dist
    ops, landing = [], []
    nrows = depths.shape[1]
    if rank > 0 and top_ghost:
        send = depths[:, top_ghost:top_ghost + GHOST].contiguous()
        recv = torch.empty_like(depths[:, :top_ghost].contiguous())
        ops += [dist.P2POp(dist.isend, send, rank - 1, group=group), dist.P2POp(dist.irecv, recv, rank - 1, group=group)]
        landing.append((slice(0, top_ghost), recv))
    if rank < world - 1 and bottom_ghost:
        send = depths[:, nrows - bottom_ghost - GHOST:nrows - bottom_ghost].contiguous()
        recv = torch.empty_like(depths[:, nrows - bottom_ghost:].contiguous())
        ops += [dist.P2POp(dist.isend, send, rank + 1, group=group), dist.P2POp(dist.irecv, recv, rank + 1, group=group)]
        landing.append((slice(nrows - bottom_ghost, nrows), recv))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for rows, buf in landing:
        depths[:, rows] = buf


def slice_rows(forcing, elo, ehi):
    out = {}
    for k, v in forcing.items():
        if v is None or k == "rho_clim":
            out[k] = v
        else:
            out[k] = np.ascontiguousarray(v[..., elo:ehi, :])
    return out


def run_decomposed_season(mask, num_days, dx, forcing, params_row, ic, rank, world, make_stepper, group=None,
                          on_ready=None, on_done=None):
    """This rank's part of one season: returns (lo, hi, {array: owned rows}).  ``forcing`` / ``ic`` / ``mask`` are the
    GLOBAL arrays (each rank slices its rows; only the slices go to the device).  ``make_stepper(local_mask, num_days,
    dx, forcing_local, params_row, ic_local)`` builds the strip stepper (``GpuStripStepper`` in production)."""
    ny = mask.shape[0]
    lo, hi, elo, ehi = strip_rows(ny, rank, world)
    if hi - lo < GHOST:
        raise ValueError("strips must own at least %d rows (ny=%d over %d ranks)" % (GHOST, ny, world))
    stepper = make_stepper(np.ascontiguousarray(mask[elo:ehi]), num_days, dx, slice_rows(forcing, elo, ehi), params_row,
                           None if ic is None else np.ascontiguousarray(ic[elo:ehi]))
    top, bottom = lo - elo, ehi - hi
    if on_ready is not None:
        on_ready()                         # the strip is staged (tools/domain_run.py starts its clock here)
    for x in range(num_days - 1):
        stepper.step(x)
        if world > 1:
            exchange_ghost_rows(stepper.depths(x + 1), rank, world, top, bottom, group=group)
    if on_done is not None:
        on_done()                          # ... and stops it here, before the strip is copied to the host
    return lo, hi, stepper.result(slice(top, top + (hi - lo)))


def run_decomposed_season_one_process(mask, num_days, dx, forcing, params_row, ic, n_strips, make_stepper):
    """All strips in ONE process (one GPU), stepped day by day with the ghost rows copied between them directly: the
    same decomposition and the same per-strip calls as the multi-rank driver, for boxes with fewer GPUs than strips
    and for the single-GPU parity test.  Returns the assembled global arrays."""
    ny = mask.shape[0]
    strips = []
    for r in range(n_strips):
        lo, hi, elo, ehi = strip_rows(ny, r, n_strips)
        st = make_stepper(np.ascontiguousarray(mask[elo:ehi]), num_days, dx, slice_rows(forcing, elo, ehi), params_row,
                          None if ic is None else np.ascontiguousarray(ic[elo:ehi]))
        strips.append((lo, hi, elo, ehi, st))
    for x in range(num_days - 1):
        for s in strips:
            s[4].step(x)
        for a, b in zip(strips[:-1], strips[1:]):           # a above b
            da, db = a[4].depths(x + 1), b[4].depths(x + 1)
            top_b, bot_a = b[0] - b[2], a[3] - a[1]
            na = da.shape[1]
            send_down = da[:, na - bot_a - GHOST:na - bot_a].clone()      # a's last owned rows -> b's top ghost rows
            send_up = db[:, top_b:top_b + GHOST].clone()                  # b's first owned rows -> a's bottom ghost rows
            db[:, :top_b] = send_down[:, GHOST - top_b:]
            da[:, na - bot_a:] = send_up[:, :bot_a]
    out = None
    for lo, hi, elo, ehi, st in strips:
        part = st.result(slice(lo - elo, lo - elo + (hi - lo)))
        if out is None:
            out = {k: np.empty(v.shape[:-2] + (ny, v.shape[-1]), dtype=v.dtype) for k, v in part.items()}
        for k, v in part.items():
            out[k][..., lo:hi, :] = v
    return out


# ------------------------------------------------------------------------------ fused peer-memory exchange

def make_strip_engine(mask, num_days, dx, forcing, rank, world, device=0, timeout_s=None, day_index=None, balance=False,
                      **flags):
    """Engine on this rank's extended strip with its forcing staged and ``nesosim_strip_setup`` done.
    Returns (engine, lo, hi, elo, ehi).  ``day_index`` (length ``num_days``): the forcing arrays hold a few generated
    days and day x of the season is ``forcing[...][day_index[x]]`` -- the repetition happens on the device after the
    rows have been sliced, so a benchmark never materialises the whole season of the whole grid on the host."""
    from .engine import SnowBudgetEngine
    ny = mask.shape[0]
    lo, hi, elo, ehi = strip_rows_balanced(mask, rank, world) if balance else strip_rows(ny, rank, world)
    if hi - lo < GHOST:
        raise ValueError("strips must own at least %d rows (ny=%d over %d ranks)" % (GHOST, ny, world))
    eng = SnowBudgetEngine(np.ascontiguousarray(mask[elo:ehi]), num_days, dx, n_members=1, device=device, **flags)
    eng.set_path("general")
    f = slice_rows(forcing, elo, ehi)
    if day_index is not None:
        import torch
        idx = torch.as_tensor(np.asarray(day_index), device="cuda:%d" % device, dtype=torch.long)
        f = {k: (v if (v is None or k == "rho_clim") else eng._dev(v)[idx].contiguous()) for k, v in f.items()}
    eng.set_forcing(f["precip"], f["conc"], f["wind"], f["drift"], f.get("rho_clim"))
    eng.strip_setup(rank > 0, rank < world - 1, timeout_s=timeout_s)
    return eng, lo, hi, elo, ehi


def check_strips(eng, rank, world, group=None):
    """After a season: did any strip give up waiting for its neighbour?  A strip that timed out carried on with stale
    ghost rows, so its results (and its neighbours') are wrong -- EVERY rank raises, together (the exchange of the
    flags doubles as the barrier that ends the season)."""
    import torch.distributed as dist
    flags = [None] * world
    dist.all_gather_object(flags, bool(eng.strip_timed_out()), group=group)
    if any(flags):
        raise RuntimeError("strip(s) %s did not receive their neighbours' ghost rows in time (nesosim_strip_status); "
                           "the season's results are invalid" % [r for r, f in enumerate(flags) if f])


def run_decomposed_season_peer(mask, num_days, dx, forcing, params_row, ic, rank, world, device=0, group=None,
                               outputs=None, engine=None, timeout_s=None, day_index=None, balance=False, **flags):
    """This rank's strip of one season with the ghost-row exchange fused into the day kernel (peer memory).
    ``torch.distributed`` must be initialised (any backend: it moves 64-byte handles and barriers only).
    Returns (lo, hi, {array: device tensor of the OWNED rows}, engine); pass ``engine`` back in to run further seasons
    on the same strip without re-staging."""
    import torch
    import torch.distributed as dist
    if engine is None:
        eng, lo, hi, elo, ehi = make_strip_engine(mask, num_days, dx, forcing, rank, world, device=device,
                                                  timeout_s=timeout_s, day_index=day_index, balance=balance, **flags)
        handles = [None] * world
        dist.all_gather_object(handles, eng.strip_export(), group=group)
        eng.strip_connect(handles[rank - 1] if rank > 0 else None, handles[rank + 1] if rank < world - 1 else None)
        eng._strip_rows = (lo, hi, elo, ehi)
    else:
        eng = engine
        lo, hi, elo, ehi = eng._strip_rows
    ic_local = None if ic is None else np.ascontiguousarray(ic[elo:ehi])
    torch.cuda.synchronize(device)
    dist.barrier(group=group)              # every strip's previous season is complete and every mailbox attached
    out = eng.run_season([list(params_row)], ic_local, outputs)
    torch.cuda.synchronize(device)
    check_strips(eng, rank, world, group)
    own = slice(lo - elo, lo - elo + (hi - lo))
    return lo, hi, {k: v[0][..., own, :] for k, v in out.items()}, eng


def run_decomposed_season_peer_one_process(mask, num_days, dx, forcing, params_row, ic, n_strips, device=0,
                                           whole_season_per_strip=False, balance=False, **flags):
    """All strips as separate contexts on ONE GPU, wired through device pointers instead of IPC handles: the same day
    kernel, mailboxes and flags as the multi-GPU run.  Default: the strips take turns day by day on one stream (no wait
    ever spins).  ``whole_season_per_strip``: each strip's whole season is enqueued on its own stream, so the strips
    really run concurrently and synchronise through their flags, as they do across GPUs."""
    import torch
    ny = mask.shape[0]
    strips = []
    for r in range(n_strips):
        eng, lo, hi, elo, ehi = make_strip_engine(mask, num_days, dx, forcing, r, n_strips, device=device, balance=balance, **flags)
        strips.append([eng, lo, hi, elo, ehi, eng.alloc_outputs(), None if ic is None else np.ascontiguousarray(ic[elo:ehi])])
    blocks = [s[0].strip_block() for s in strips]
    for r, s in enumerate(strips):
        s[0].strip_connect_local(blocks[r - 1] if r > 0 else None, blocks[r + 1] if r < n_strips - 1 else None)
    p = [list(params_row)]
    if whole_season_per_strip:
        streams = [torch.cuda.Stream(device) for _ in strips]
        torch.cuda.synchronize(device)
        for st, s in zip(streams, strips):
            with torch.cuda.stream(st):
                s[0].run_season(p, s[6], s[5])
    else:
        for x in range(num_days - 1):
            for s in strips:
                s[0].run_season(p, s[6], s[5], first_step=x, num_steps=1)
    torch.cuda.synchronize(device)
    out = None
    for eng, lo, hi, elo, ehi, o, _ in strips:
        if eng.strip_timed_out():
            raise RuntimeError("a strip timed out waiting for its neighbour")
        for k, v in o.items():
            part = v[0][..., lo - elo:lo - elo + (hi - lo), :].cpu().numpy()
            if out is None:
                out = {}
            if k not in out:
                out[k] = np.empty(part.shape[:-2] + (ny, part.shape[-1]), dtype=part.dtype)
            out[k][..., lo:hi, :] = part
    return out
