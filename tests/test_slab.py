"""Slab decomposition (babelbrain_b200/slab.py): host logic of the multi-GPU path.  The exchange
protocol the CUDA library follows (halo_exchange_plan) is executed here by two gloo ranks on CPU with
a stand-in update that has the solver's reach along i (two planes one way, one the other, alternating
like the stress / particle half-steps) and compared with the undecomposed result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from babelbrain_b200.slab import (HALO, SlabPlan, assemble_maps, assemble_sensors, expand_runs, halo_exchange_plan, merge_sensor_runs,
                                  merge_sensor_tables, sensor_rows_of_slab)


def test_plan_covers_the_grid():
    for n1, nr in ((240, 1), (240, 2), (1080, 8), (601, 4), (33, 8)):
        p = SlabPlan(n1, nr)
        assert p.owned(0)[0] == 0 and p.owned(nr - 1)[1] == n1
        sizes = [p.owned(r)[1] - p.owned(r)[0] for r in range(nr)]
        assert sum(sizes) == n1 and max(sizes) - min(sizes) <= 1
        for r in range(nr):
            lo, hi = p.with_halo(r)
            assert lo == max(p.owned(r)[0] - HALO, 0) and hi == min(p.owned(r)[1] + HALO, n1)
            assert p.neighbours(r) == (r - 1 if r else None, r + 1 if r < nr - 1 else None)
            for i in range(*p.owned(r)):
                assert p.rank_of_plane(i) == r
    with pytest.raises(ValueError):
        SlabPlan(12, 8)
    assert SlabPlan(1080, 8).halo_bytes_per_half_step(1080, 1088) == 3 * 2 * 2 * 1080 * 1088 * 4


def test_sensor_rows_and_assembly():
    rng = np.random.default_rng(0)
    shape = (20, 6, 7)
    sm = rng.random(shape) < 0.3
    index = (np.flatnonzero(sm.reshape(-1, order='F')) + 1).astype(np.uint32)     # IndexSensorMap convention
    plan = SlabPlan(shape[0], 3)
    rows = [sensor_rows_of_slab(index, shape, *plan.owned(r)) for r in range(3)]
    assert sorted(np.concatenate(rows).tolist()) == list(range(index.size))
    data = rng.random((index.size, 4)).astype(np.float32)
    out = assemble_sensors(index.size, 4, rows, [data[r] for r in rows])
    assert np.array_equal(out, data)
    vol = rng.random(shape).astype(np.float32)
    assert np.array_equal(assemble_maps(shape, plan, [vol[slice(*plan.owned(r))] for r in range(3)]), vol)


@pytest.mark.parametrize('shape,nranks,density', [((23, 9, 11), 3, 0.4), ((64, 5, 7), 8, 0.05), ((16, 4, 4), 2, 1.0), ((16, 4, 4), 4, 0.0)])
def test_merge_of_slab_sensor_tables_is_the_global_table(shape, nranks, density):
    """Each slab's device-built table (ascending Fortran-order indices of its planes) merged without sorting
    equals the table of the whole SensorMap, ragged slabs and empty lines included."""
    rng = np.random.default_rng(3)
    sen = rng.random(shape) < density
    expect = (np.flatnonzero(sen.reshape(-1, order='F')) + 1).astype(np.uint32)
    plan = SlabPlan(shape[0], nranks, 2)
    local = []
    for r in range(nranks):
        i0, i1 = plan.owned(r)
        m = np.zeros(shape, bool)
        m[i0:i1] = sen[i0:i1]
        local.append((np.flatnonzero(m.reshape(-1, order='F')) + 1).astype(np.uint32))
    index, rows = merge_sensor_tables(local, shape[0], shape[1] * shape[2])
    assert index.dtype == np.uint32 and np.array_equal(index, expect)
    for r in range(nranks):
        assert np.array_equal(index[rows[r]], local[r])
        assert np.array_equal(rows[r], sensor_rows_of_slab(expect, shape, *plan.owned(r)))
    # the same merge as runs of consecutive rows (what the multi-GPU gather uses), placed by the library's host helper
    import ctypes
    from babelbrain_b200 import _capi, build
    build.build()
    L = _capi.lib()
    ntot, runs = merge_sensor_runs(local, shape[0], shape[1] * shape[2])
    assert ntot == expect.size
    index2 = np.zeros(ntot, np.uint32)
    traces = [np.arange(t.size * 3, dtype=np.float32).reshape(-1, 3) + 1000 * r for r, t in enumerate(local)]
    full = np.zeros((ntot, 3), np.float32)
    for r in range(nranks):
        dst, src, cnt = runs[r]
        assert np.array_equal(expand_runs(runs[r]), rows[r]) and int(cnt.sum()) == local[r].size and np.all(cnt > 0)
        for out, part, nb in ((index2, local[r], 4), (full, traces[r], 12)):
            assert L.bb_host_scatter_runs(_capi.ptr(out), _capi.ptr(part), _capi.ptr(dst), _capi.ptr(src), _capi.ptr(cnt), dst.size, nb) == 0
    assert np.array_equal(index2, expect)
    for r in range(nranks):
        assert np.array_equal(full[rows[r]], traces[r])


def _step(a, forward):
    """stand-in half-step along axis 0 with the 4-point staggered reach (zero outside)"""
    p = np.pad(a, ((2, 2), (0, 0)))
    if forward:   # f(i+1) - f(i) and f(i+2) - f(i-1)
        return a + 0.1 * (1.125 * (p[3:-1] - p[2:-2]) - (p[4:] - p[1:-3]) / 24)
    return a + 0.1 * (1.125 * (p[2:-2] - p[1:-3]) - (p[3:-1] - p[:-4]) / 24)


def _worker(rank, world, port, n1, width, nsteps, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    plan = SlabPlan(n1, world)
    i0, i1 = plan.owned(rank)
    nown = i1 - i0
    full = np.random.default_rng(7).random((n1, width))
    loc = np.zeros((nown + 2 * HALO, width))
    lo, hi = plan.with_halo(rank)
    loc[HALO - (i0 - lo):HALO + nown + (hi - i1)] = full[lo:hi]              # owned planes + the halos that exist
    for n in range(nsteps):
        new = _step(loc, forward=bool(n & 1))
        loc[HALO:HALO + nown] = new[HALO:HALO + nown]                        # only owned planes are updated
        reqs, bufs = [], []
        for op, peer, first, cnt in halo_exchange_plan(rank, world, nown):
            if op == 'send':
                reqs.append(dist.isend(torch.from_numpy(loc[first:first + cnt].copy()), peer))
            else:
                t = torch.empty((cnt, width), dtype=torch.float64)
                bufs.append((first, cnt, t))
                reqs.append(dist.irecv(t, peer))
        for r in reqs:
            r.wait()
        for first, cnt, t in bufs:
            loc[first:first + cnt] = t.numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, loc[HALO:HALO + nown])
    if rank == 0:
        q.put(assemble_maps((n1, width), plan, gathered).astype(np.float64))
    dist.destroy_process_group()


def test_two_rank_halo_protocol_matches_single_domain():
    n1, width, nsteps, world = 23, 5, 12, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n1, width, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.random.default_rng(7).random((n1, width))
    for n in range(nsteps):
        ref = _step(ref, forward=bool(n & 1))
    assert np.allclose(got, ref, rtol=0, atol=1e-6)   # assemble_maps returns float32 volumes
