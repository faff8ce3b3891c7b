"""Row-strip domain decomposition (nesosim_b200/domain.py): the exchange logic on CPU with the numpy oracle as the
strip stepper (two gloo ranks, and several strips in one process), and the GPU engine as the stepper on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nesosim_b200 import domain
from nesosim_b200 import synthetic as S
from oracle import nesosim_oracle as O

PARAMS = [5.8e-7, 5., 1.45e-7, 2.2e-8]
NAMES = ("snowDepths", "density", "snowAcc", "snowOcean", "snowAdv", "snowDiv", "snowLead", "snowAtm",
         "snowWindPackLoss", "snowWindPackGain", "snowWindPack")


class OracleStripStepper:
    """The CPU oracle on one extended strip, one day at a time (test stand-in for GpuStripStepper)."""

    def __init__(self, local_mask, num_days, dx, forcing_local, params_row, ic_local):
        self.mask, self.dx, self.f = local_mask, dx, forcing_local
        self.p = O.Params(windPackFactor=params_row[0], windPackThresh=params_row[1], leadLossFactor=params_row[2],
                          atmLossFactor=params_row[3])
        self.fl = O.Flags(atmlossInc=1)
        T, ny, nx = forcing_local["precip"].shape
        self.s = O.gen_empty_arrays(T, ny, nx)
        if ic_local is not None:
            half = O.initial_depths(ic_local, forcing_local["conc"][0], self.p)
            self.s["snowDepths"][0, 0] = half
            self.s["snowDepths"][0, 1] = half
        self._t = {}

    def step(self, x):
        f = self.f
        O.calc_budget(self.s, f["conc"][x], f["precip"][x], f["drift"][x], f["wind"][x], np.full(f["conc"][x].shape, np.nan),
                      self.mask, self.dx, x, self.p, self.fl)

    def depths(self, slot):
        return torch.from_numpy(self.s["snowDepths"][slot])      # shares memory with the numpy state

    def result(self, rows):
        return {k: self.s[k][..., rows, :].copy() for k in NAMES}


def setup(ny=23, nx=17, T=7, seed=5):
    mask = S.region_mask(shape=(ny, nx), kind="disc")
    forcing = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed) * 2
    p = O.Params(windPackFactor=PARAMS[0], windPackThresh=PARAMS[1], leadLossFactor=PARAMS[2], atmLossFactor=PARAMS[3])
    ref = O.run_season(forcing, ic, mask, 50000, p, O.Flags(atmlossInc=1))
    return mask, forcing, ic, ref


def test_strip_rows_cover_the_grid_with_two_ghost_rows():
    for ny, world in ((23, 2), (90, 3), (1785, 8)):
        prev_hi = 0
        for r in range(world):
            lo, hi, elo, ehi = domain.strip_rows(ny, r, world)
            assert lo == prev_hi and elo == max(lo - 2, 0) and ehi == min(hi + 2, ny)
            prev_hi = hi
        assert prev_hi == ny


def test_cost_balanced_cuts_cover_the_grid():
    """domain.balanced_cuts: contiguous strips over all rows, no strip thinner than its ghost rows, every strip's
    modelled cost within a row's worth of the mean, and closer to it than equal-height strips are."""
    from nesosim_b200 import sharding
    mask = S.region_mask(dx=25000)
    land = (mask > 10) | (mask < 1)
    row_cost = mask.shape[1] + 1.3 * (~land).sum(axis=1)
    for world in (2, 3, 8):
        cuts = domain.balanced_cuts(mask, world)
        assert cuts[0] == 0 and cuts[-1] == mask.shape[0] and len(cuts) == world + 1
        assert all(b - a >= 2 * domain.GHOST for a, b in zip(cuts[:-1], cuts[1:]))
        cost = np.array([row_cost[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])])
        equal = np.array([row_cost[slice(*sharding.member_range(mask.shape[0], r, world))].sum() for r in range(world)])
        assert cost.max() - cost.mean() <= row_cost.max() + 1e-9
        assert cost.max() <= equal.max() + 1e-9
        for r in range(world):
            lo, hi, elo, ehi = domain.strip_rows_balanced(mask, r, world)
            assert (lo, hi) == (cuts[r], cuts[r + 1]) and elo == max(lo - 2, 0) and ehi == min(hi + 2, mask.shape[0])
    tiny = S.region_mask(shape=(13, 9), kind="disc")
    assert domain.balanced_cuts(tiny, 3)[-1] == 13
    with pytest.raises(ValueError):
        domain.balanced_cuts(tiny, 8)


@pytest.mark.parametrize("n_strips", [2, 3, 5])
def test_strips_in_one_process_reproduce_the_single_domain_run(n_strips):
    mask, forcing, ic, ref = setup()
    got = domain.run_decomposed_season_one_process(mask, 7, 50000, forcing, PARAMS, ic, n_strips, OracleStripStepper)
    for k in NAMES:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


def _rank(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mask, forcing, ic, _ = setup()
    lo, hi, part = domain.run_decomposed_season(mask, 7, 50000, forcing, PARAMS, ic, rank, world, OracleStripStepper)
    q.put((rank, lo, hi, part))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_exchange_ghost_rows():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, _, _, ref = setup()
    for rank, lo, hi, part in parts:
        for k in NAMES:
            assert np.array_equal(part[k], ref[k][..., lo:hi, :], equal_nan=True), (k, rank)


class _FakeStrip:
    def __init__(self, timed_out):
        self._t = timed_out

    def strip_timed_out(self):
        return self._t


def _check_rank(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for flags in ((False, False), (False, True)):
        try:
            domain.check_strips(_FakeStrip(flags[rank]), rank, world)
            out.append("ok")
        except RuntimeError as e:
            out.append(str(e))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_a_strip_time_out_is_raised_on_every_rank():
    """domain.check_strips: the ranks exchange their time-out marks; one mark fails the season everywhere."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_check_rank, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        assert got[rank][0] == "ok"
        assert "strip(s) [1] did not receive" in got[rank][1]


@pytest.mark.gpu
def test_gpu_strips_reproduce_the_single_domain_run(cuda):
    mask = S.region_mask(dx=100000)
    T = 6
    forcing = S.make_season(mask, T, seed=11)
    ic = S.make_ic(mask, seed=11)
    p = O.Params(windPackFactor=PARAMS[0], windPackThresh=PARAMS[1], leadLossFactor=PARAMS[2], atmLossFactor=PARAMS[3])
    ref = O.run_season(forcing, ic, mask, 100000, p, O.Flags(atmlossInc=1))

    def make(local_mask, num_days, dx, forcing_local, params_row, ic_local):
        return domain.GpuStripStepper(local_mask, num_days, dx, forcing_local, params_row, ic_local, atmlossInc=1)

    got = domain.run_decomposed_season_one_process(mask, T, 100000, forcing, PARAMS, ic, 4, make)
    for k in NAMES:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


# ------------------------------------------------------------------ exchange fused into the day kernel (peer memory)

def _gpu_reference(mask, T, dx, forcing, ic, **flags):
    from nesosim_b200.engine import SnowBudgetEngine
    eng = SnowBudgetEngine(mask, T, dx, n_members=1, **flags)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    return {k: v[0].cpu().numpy() for k, v in eng.run_season([PARAMS], ic).items()}


@pytest.mark.gpu
@pytest.mark.parametrize("n_strips", [2, 3, 5])
def test_peer_strips_on_a_small_grid_match_the_oracle(cuda, n_strips):
    """Strips of 4-12 rows: one tile row is the top AND the bottom boundary set of its strip."""
    mask, forcing, ic, ref = setup()
    got = domain.run_decomposed_season_peer_one_process(mask, 7, 50000, forcing, PARAMS, ic, n_strips, atmlossInc=1)
    for k in NAMES:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx,n_strips,concurrent", [(90, 90, 4, False), (357, 357, 3, False), (357, 357, 3, True),
                                                        (131, 70, 2, True), (200, 45, 7, True)])
def test_peer_strips_match_the_single_domain_run(cuda, ny, nx, n_strips, concurrent):
    """Same day kernel, mailboxes and flags as across GPUs, all strips on one device; `concurrent` = every strip's whole
    season on its own stream, synchronised only through the flags."""
    mask = S.region_mask(dx=100000) if (ny, nx) == (90, 90) else S.region_mask(shape=(ny, nx), kind="disc")
    T = 9
    forcing = S.make_season(mask, T, seed=23)
    ic = S.make_ic(mask, seed=23)
    ref = _gpu_reference(mask, T, 25000, forcing, ic, atmlossInc=1)
    got = domain.run_decomposed_season_peer_one_process(mask, T, 25000, forcing, PARAMS, ic, n_strips,
                                                        whole_season_per_strip=concurrent, atmlossInc=1)
    for k in NAMES:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


@pytest.mark.gpu
def test_cost_balanced_peer_strips_match_the_single_domain_run(cuda):
    """Strips of unequal height (rows cut for equal modelled cost) through the mailboxes, concurrent streams."""
    mask = S.region_mask(dx=25000)
    T = 7
    forcing = S.make_season(mask, T, seed=29)
    ic = S.make_ic(mask, seed=29)
    ref = _gpu_reference(mask, T, 25000, forcing, ic, atmlossInc=1)
    for n_strips in (3, 6):
        got = domain.run_decomposed_season_peer_one_process(mask, T, 25000, forcing, PARAMS, ic, n_strips,
                                                            whole_season_per_strip=True, balance=True, atmlossInc=1)
        for k in NAMES:
            assert np.array_equal(got[k], ref[k], equal_nan=True), (k, n_strips)


@pytest.mark.gpu
def test_peer_strips_run_two_seasons_and_without_dynamics(cuda):
    """The flag epoch: a second season on the same contexts; and dynamicsInc=0, where nothing is exchanged."""
    import torch
    mask = S.region_mask(shape=(64, 40), kind="disc")
    T = 6
    forcing = S.make_season(mask, T, seed=3)
    ic = S.make_ic(mask, seed=3)
    for flags in (dict(atmlossInc=1), dict(dynamicsInc=0)):
        ref = _gpu_reference(mask, T, 50000, forcing, ic, **flags)
        strips = [domain.make_strip_engine(mask, T, 50000, forcing, r, 3, **flags) for r in range(3)]
        blocks = [s[0].strip_block() for s in strips]
        for r, s in enumerate(strips):
            s[0].strip_connect_local(blocks[r - 1] if r > 0 else None, blocks[r + 1] if r < 2 else None)
        for season in range(2):
            outs = []
            streams = [torch.cuda.Stream() for _ in strips]
            torch.cuda.synchronize()
            for st, (eng, lo, hi, elo, ehi) in zip(streams, strips):
                with torch.cuda.stream(st):
                    outs.append(eng.run_season([PARAMS], np.ascontiguousarray(ic[elo:ehi])))
            torch.cuda.synchronize()
            for (eng, lo, hi, elo, ehi), o in zip(strips, outs):
                assert not eng.strip_timed_out()
                for k in NAMES:
                    assert np.array_equal(o[k][0][..., lo - elo:hi - elo, :].cpu().numpy(), ref[k][..., lo:hi, :],
                                          equal_nan=True), (k, season, flags)


@pytest.mark.gpu
def test_a_strip_without_its_neighbour_times_out_instead_of_hanging(cuda):
    mask = S.region_mask(shape=(40, 40), kind="disc")
    T = 4
    forcing = S.make_season(mask, T, seed=3)
    a = domain.make_strip_engine(mask, T, 50000, forcing, 0, 2, timeout_s=0.2)[0]
    b = domain.make_strip_engine(mask, T, 50000, forcing, 1, 2, timeout_s=0.2)[0]
    a.strip_connect_local(None, b.strip_block())
    a.run_season([PARAMS], None)          # b never runs: a's second day waits for ghost rows that never come
    cuda.cuda.synchronize()
    assert a.strip_timed_out()


def _peer_rank(rank, world, port, q, n_dev):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % n_dev
    torch.cuda.set_device(dev)
    mask = S.region_mask(shape=(120, 80), kind="disc")
    T = 6
    forcing = S.make_season(mask, T, seed=9)
    ic = S.make_ic(mask, seed=9)
    lo, hi, part, eng = domain.run_decomposed_season_peer(mask, T, 25000, forcing, PARAMS, ic, rank, world, device=dev,
                                                          atmlossInc=1, timeout_s=30.0)
    q.put((rank, lo, hi, {k: v.cpu().numpy() for k, v in part.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_strips_across_processes_over_cuda_ipc(cuda):
    """One process per strip, mailboxes attached through CUDA IPC handles: on a multi-GPU box one GPU per rank (peer
    stores over NVLink), on a one-GPU box both ranks share the device (time-sliced contexts)."""
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs: two processes time-slicing one GPU cannot wait on each other's kernels")
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_rank, args=(r, world, port, q, cuda.cuda.device_count())) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mask = S.region_mask(shape=(120, 80), kind="disc")
    forcing = S.make_season(mask, 6, seed=9)
    ref = _gpu_reference(mask, 6, 25000, forcing, S.make_ic(mask, seed=9), atmlossInc=1)
    for rank, lo, hi, part in parts:
        for k in NAMES:
            assert np.array_equal(part[k], ref[k][..., lo:hi, :], equal_nan=True), (k, rank)


def _peer_rank_with_a_silent_neighbour(rank, world, port, q, n_dev):
    """Rank 0 runs its season; rank 1 attaches its mailboxes but never launches: rank 0's boundary CTAs give up after the
    time-out, and the check that ends the season must fail on BOTH ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % n_dev
    torch.cuda.set_device(dev)
    mask = S.region_mask(shape=(120, 80), kind="disc")
    T = 4
    forcing = S.make_season(mask, T, seed=9)
    ic = S.make_ic(mask, seed=9)
    eng, lo, hi, elo, ehi = domain.make_strip_engine(mask, T, 25000, forcing, rank, world, device=dev, timeout_s=0.3, atmlossInc=1)
    handles = [None] * world
    dist.all_gather_object(handles, eng.strip_export())
    eng.strip_connect(handles[rank - 1] if rank > 0 else None, handles[rank + 1] if rank < world - 1 else None)
    dist.barrier()
    if rank == 0:
        eng.run_season([list(PARAMS)], np.ascontiguousarray(ic[elo:ehi]))
        torch.cuda.synchronize(dev)
    failed = False
    try:
        domain.check_strips(eng, rank, world)
    except RuntimeError as e:
        failed = "did not receive" in str(e)
    q.put((rank, failed, bool(eng.strip_timed_out())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_a_strip_time_out_fails_the_season_on_every_rank(cuda):
    """A neighbour that never delivers: the waiting strip's kernels give up after the time-out (no hung GPU), its own
    status says so, and ``domain.check_strips`` raises on every rank -- the results of such a season are invalid."""
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (one process per strip)")
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_rank_with_a_silent_neighbour, args=(r, world, port, q, cuda.cuda.device_count())) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == (0, True, True)          # rank 0 timed out itself ...
    assert got[1] == (1, True, False)         # ... and rank 1, which did not, fails with it
