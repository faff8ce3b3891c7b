"""Multi-GPU host logic on CPU: members/seasons are sharded with no data-path collective; two gloo ranks each run
their members (through the CPU oracle here -- the GPU engine is exercised by the -m gpu tests) and gather a small
per-member result; the union must equal the single-process run."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nesosim_b200 import sharding
from nesosim_b200 import synthetic as S
from oracle import nesosim_oracle as O


def test_member_ranges_partition_every_count():
    for world in (1, 2, 3, 4, 8):
        for M in (1, 7, 8, 128, 1024, 1027):
            spans = [sharding.member_range(M, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert sharding.member_range(1024, 3, 8) == (384, 512)


def test_season_assignment_round_robin():
    years = list(range(1980, 2021))
    got = [sharding.season_assignment(years, r, 8) for r in range(8)]
    assert sorted(sum(got, [])) == years and got[0][:2] == [1980, 1988] and len(got[0]) == 6 and len(got[7]) == 5


def _member_metric(forcing, ic, mask, row):
    p = O.Params(windPackFactor=row[0], windPackThresh=row[1], leadLossFactor=row[2], atmLossFactor=row[3])
    out = O.run_season(forcing, ic, mask, 100000, p, O.Flags(atmlossInc=1))
    return float(np.nansum(out["snowDepths"][-1]))


def _worker(rank, world, port, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mask = S.region_mask(shape=(14, 12), kind="disc")
    forcing = S.make_season(mask, 5, seed=3)
    ic = S.make_ic(mask, seed=3)
    params = S.ensemble_params(M, seed=3)
    mine = sharding.shard_params(params, rank, world)
    local = torch.tensor([_member_metric(forcing, ic, mask, row) for row in mine], dtype=torch.float64)
    allv = sharding.gather_member_results(local, M, rank, world)
    if rank == 0:
        q.put(allv.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_single_process_ensemble():
    M, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mask = S.region_mask(shape=(14, 12), kind="disc")
    forcing = S.make_season(mask, 5, seed=3)
    ic = S.make_ic(mask, seed=3)
    params = S.ensemble_params(M, seed=3)
    exp = np.array([_member_metric(forcing, ic, mask, row) for row in params])
    assert np.array_equal(got, exp)


def test_multiseason_batch_of_the_bench_matches_baseline_step_count():
    """BASELINE.json configs[3]: 41 seasons, Sep 1 - Apr 30, 9891 steps in all; dealt round-robin."""
    import bench
    assert len(bench.MULTI_YEARS) == 41 and bench.MULTI_YEARS[0] == 1980 and bench.MULTI_YEARS[-1] == 2020
    days = [bench.season_days(y) for y in bench.MULTI_YEARS]
    assert set(days) == {242, 243} and sum(d - 1 for d in days) == 9891
    assert bench.season_days(1983) == 243 and bench.season_days(1999) == 243 and bench.season_days(2018) == 242
    for world in (1, 2, 4, 8):
        parts = [sharding.season_assignment(bench.MULTI_YEARS, r, world) for r in range(world)]
        assert sorted(y for p in parts for y in p) == bench.MULTI_YEARS
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_bench_end_to_end_leg_runs_against_a_stand_in_engine(monkeypatch):
    """bench.run_e2e (the host-buffer leg of the JSON line) end to end on the CPU with a stand-in engine: the record's
    keys, the byte counts it reports, the drain mode it asks the library for at one and at several ranks, and the check
    of the host arrays against the device-resident result."""
    import types
    import torch
    import bench

    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"}))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    for k in ("NESOSIM_HOST_THREADS", "NESOSIM_HOST_COMPACT"):
        monkeypatch.delenv(k, raising=False)

    class Engine:
        M, T, ny, nx = 3, 4, 5, 6
        calls = 0

        def run_season_host(self, forcing, params, ic, outputs):
            Engine.calls += 1
            for i, n in enumerate(sorted(outputs)):
                outputs[n].fill_(float(i))
                outputs[n][0, 0].fill_(float("nan"))
            return outputs, 111, 222

        def host_drain_info(self):
            return os.environ.get("NESOSIM_HOST_COMPACT") == "1", 0

        def host_drain_blocks(self):
            return (20, 7) if os.environ.get("NESOSIM_HOST_COMPACT") == "1" else (0, 0)

    eng = Engine()
    forcing = {k: np.zeros((4, 5, 6)) for k in ("precip", "conc", "wind")}
    forcing["drift"] = np.zeros((4, 2, 5, 6))
    args = types.SimpleNamespace(steps=2, e2e_steps=3)
    monkeypatch.setattr(os, "cpu_count", lambda: 16)
    rec = bench.run_e2e(args, eng, forcing, None, np.zeros((5, 6)), 0, 1, 1000.0, lambda: None)
    assert Engine.calls == 3 and rec["steps"] == 2 and rec["value"] > 0 and rec["unit"] == bench.UNIT
    assert rec["h2d_bytes_per_step"] == 111 and rec["d2h_bytes_per_step"] == 222
    assert rec["host_array_bytes_per_step"] == 12 * 3 * 4 * 5 * 6 * 8
    assert rec["drain"].startswith("compacted") and rec["host_threads"] == 16
    assert rec["blocks_packed"] == 20 and rec["blocks_copied_whole_by_the_link"] == 7
    assert "host_arrays_identical_to_device_result" not in rec
    # against a "device" result: equal (NaNs in the same places), then different
    from nesosim_b200 import _lib
    dev = {}
    for i, n in enumerate(sorted(_lib.OUTPUT_NAMES)):
        dev[n] = torch.full((3, 4, 2, 5, 6) if n == "snowDepths" else (3, 4, 5, 6), float(i), dtype=torch.float64)
        dev[n][0, 0] = float("nan")
    rec = bench.run_e2e(args, eng, forcing, None, np.zeros((5, 6)), 0, 1, 1000.0, lambda: None, dev)
    assert rec["host_arrays_identical_to_device_result"] is True
    dev["density"][2, 3, 4, 5] += 1.0
    rec = bench.run_e2e(args, eng, forcing, None, np.zeros((5, 6)), 0, 1, 1000.0, lambda: None, dev)
    assert rec["host_arrays_identical_to_device_result"] is False
    # several ranks on a box: the plain drain, and this rank's share of the cores
    assert bench.host_drain_settings(1, 16) == (16, True) and bench.host_drain_settings(1, 64) == (16, True)
    assert bench.host_drain_settings(2, 24) == (12, False) and bench.host_drain_settings(8, 32) == (4, False)
    assert bench.host_drain_settings(1, 2) == (2, False) and bench.host_drain_settings(8, 4) == (1, False)
