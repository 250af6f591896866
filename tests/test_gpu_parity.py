"""GPU parity tests proper: the CUDA path (through the reference-shaped Python API -> C ABI) against
the C oracle on the same seeded inputs.  Tolerance (north_star): float32 pressure-amplitude map
relative L2 <= 1e-4 with an identical peak voxel; index maps bit-exact."""
import numpy as np
import pytest

import oracle
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab, collect_results, PropagationModel

pytestmark = pytest.mark.gpu
TOL = 1e-4
DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')


def rl2(a, b):
    wide = np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64
    a = np.asarray(a, wide)
    b = np.asarray(b, wide)
    nb = float(np.linalg.norm(b))
    d = float(np.linalg.norm(a - b))
    return d / nb if nb > 0 else d


def same_peak(a, b, tie=1e-5):
    """Identical peak voxel; a mirror-symmetric workload has twin maxima that differ by rounding
    only, so a voxel where the oracle itself is within `tie` (relative) of its maximum also counts."""
    ia, ib = int(np.argmax(a)), int(np.argmax(b))
    if ia == ib:
        return True
    fa, fb = a.reshape(-1), b.reshape(-1)
    return abs(float(fb[ia]) - float(fb[ib])) <= tie * float(fb[ib]) and abs(float(fa[ia]) - float(fa[ib])) <= tie * float(fa[ia])


def run_cuda(w, variant=0, **over):
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    kw.update(over)
    s = FdtdSlab(*w['args'], kernel_variant=variant, **kw)
    s.run()
    out = collect_results(s)
    last = {k: s.get_map(2, k) for k in ('Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmaxy', 'Pressure')}
    s.close()
    return out, last


def run_oracle(w, **over):
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    kw.update(over)
    return oracle.run_c(*w['args'], want_last=True, **kw)


CASES = [('single_water', (40, 44, 56), 5, 8), ('ctx500_skull', (56, 48, 72), 6, 8),
         ('ctx500_skull', (40, 70, 45), 4, 6), ('h317_skull', (48, 48, 48), 3, 6),
         ('dome_stress', (64, 64, 48), 4, 8), ('hires_1mhz', (44, 40, 70), 3, 12)]


@pytest.mark.parametrize('variant', [0, 1, 2, 3])
@pytest.mark.parametrize('name,shape,periods,pml', CASES)
def test_parity_small(name, shape, periods, pml, variant):
    w = workloads.make_workload(name, shape=shape, periods=periods, pml=pml)
    (Sensor, RMS, Peak, IP), last = run_cuda(w, variant)
    ref = run_oracle(w)
    assert np.array_equal(IP['IndexSensorMap'], ref['IndexSensorMap'])
    assert np.allclose(Sensor['time'], ref['Sensor']['time'], rtol=1e-12)
    for k in ref['RMS']:
        assert rl2(RMS[k], ref['RMS'][k]) <= TOL, (k, rl2(RMS[k], ref['RMS'][k]))
    for k in ref['Peak']:
        assert rl2(Peak[k], ref['Peak'][k]) <= TOL, k
    main = RMS if 'Pressure' in RMS else Peak
    refm = ref['RMS'] if 'Pressure' in ref['RMS'] else ref['Peak']
    assert same_peak(main['Pressure'], refm['Pressure'])
    assert rl2(Sensor['Pressure'], ref['Sensor']['Pressure']) <= TOL
    for k, v in last.items():
        assert rl2(v, ref['LastMap'][k]) <= 5 * TOL, (k, rl2(v, ref['LastMap'][k]))


def test_all_maps_and_peak():
    w = workloads.make_workload('ctx500_skull', shape=(48, 44, 52), periods=4, pml=6)
    maps = ['ALLV', 'Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Sigmaxy', 'Sigmaxz', 'Sigmayz', 'Pressure']
    over = dict(SelMapsRMSPeakList=maps, SelMapsSensorsList=maps, SelRMSorPeak=3)
    (Sensor, RMS, Peak, IP), _ = run_cuda(w, 0, **over)
    ref = run_oracle(w, **over)
    for k in maps:
        assert rl2(RMS[k], ref['RMS'][k]) <= TOL, ('rms', k, rl2(RMS[k], ref['RMS'][k]))
        assert rl2(Peak[k], ref['Peak'][k]) <= TOL, ('peak', k, rl2(Peak[k], ref['Peak'][k]))
        assert rl2(Sensor[k], ref['Sensor'][k]) <= TOL, ('sensor', k)


def test_reflector_and_hard_source():
    w = workloads.make_workload('ctx500_skull', shape=(44, 44, 56), periods=4, pml=6)
    refl = np.zeros(w['args'][0].shape, np.uint32)
    refl[18:24, 10:30, 30:34] = 1
    for ts in (0, 1):
        over = dict(ReflectorMask=refl, TypeSource=ts)
        (Sensor, RMS, Peak, IP), _ = run_cuda(w, 0, **over)
        ref = run_oracle(w, **over)
        assert rl2(RMS['Pressure'], ref['RMS']['Pressure']) <= TOL
        assert np.all(RMS['Pressure'][18:24, 10:30, 30:34] == 0)


def test_many_materials_uint16_labels():
    """CT-style maps: > 127 materials switch the device labels to uint16 (BabelIntegrationBASE.py:1268)."""
    w = workloads.make_workload('ctx500_skull', shape=(40, 40, 48), periods=3, pml=6)
    MM, ML = w['args'][0], w['args'][1]
    rng = np.random.default_rng(3)
    nb = 300
    bone = np.tile(ML[2], (nb, 1)) * (1 + 0.1 * rng.random((nb, 5)))
    ML2 = np.vstack([ML, bone])
    MM2 = MM.copy()
    sel = MM == 2
    MM2[sel] = 5 + rng.integers(0, nb, sel.sum()).astype(np.uint32)
    args = (MM2, ML2) + w['args'][2:]
    w2 = dict(args=args, kwargs=dict(w['kwargs'], QCorrection=np.concatenate([w['kwargs']['QCorrection'], np.full(nb, 3.0)])), meta=w['meta'])
    (Sensor, RMS, Peak, IP), _ = run_cuda(w2, 0)
    ref = run_oracle(w2)
    assert rl2(RMS['Pressure'], ref['RMS']['Pressure']) <= TOL


def test_public_api_tuple_and_errors():
    w = workloads.make_workload('single_water', shape=(40, 40, 48), periods=3, pml=8)
    PM = PropagationModel()
    r = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    assert len(r) == 4
    Sensor, LastMap, RMS, IP = r
    assert RMS['Pressure'].dtype == np.float32 and RMS['Pressure'].flags.writeable
    RMS['Pressure'] *= 2.0
    assert Sensor['Pressure'].shape == (IP['IndexSensorMap'].size, Sensor['time'].size)
    assert Sensor['time'].size % (w['meta']['ppp'] // w['meta']['sub']) == 0
    assert LastMap['Vz'].shape == w['args'][0].shape
    r5 = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], SelRMSorPeak=3))
    assert len(r5) == 5
    bad = list(w['args'])
    bad[0] = bad[0].astype(np.int32)
    with pytest.raises(TypeError):
        PM.StaggeredFDTD_3D_with_relaxation(*bad, **w['kwargs'])
    bad = list(w['args'])
    bad[0] = bad[0] + 7
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*bad, **w['kwargs'])
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], DT=w['kwargs']['DT'] * 10))


def test_full_size_config1_against_oracle():
    """BASELINE config 1 (250 kHz water, 120x120x160, 720 steps) at full size."""
    w = workloads.make_workload('single_water')
    (Sensor, RMS, Peak, IP), _ = run_cuda(w, 0)
    ref = run_oracle(w)
    assert rl2(RMS['Pressure'], ref['RMS']['Pressure']) <= TOL
    assert same_peak(RMS['Pressure'], ref['RMS']['Pressure'])
    assert rl2(Sensor['Pressure'], ref['Sensor']['Pressure']) <= TOL


def test_properties_linearity_and_variants_large():
    """Size-independent properties on a larger domain: linear in source amplitude; the tiled and the
    direct kernels agree; a homogeneous lossless medium has p == -Sigmaxx."""
    w = workloads.make_workload('ctx500_skull', shape=(120, 112, 160), periods=10)
    (S1, R1, _, _), l1 = run_cuda(w, 0)
    (S2, R2, _, _), l2 = run_cuda(w, 1)
    assert rl2(R1['Pressure'], R2['Pressure']) <= 1e-5
    args = list(w['args'])
    args[4] = args[4] * 3.0
    (S3, R3, _, _), _ = run_cuda(dict(args=tuple(args), kwargs=w['kwargs']), 0)
    assert rl2(R3['Pressure'], 3.0 * R1['Pressure'].astype(np.float64)) <= 1e-5
    ww = workloads.make_workload('single_water', shape=(96, 96, 128), periods=8)
    (S4, R4, _, _), l4 = run_cuda(ww, 0, SelMapsRMSPeakList=['Pressure', 'Sigmaxx'])
    assert rl2(R4['Pressure'], R4['Sigmaxx']) <= 1e-5


@pytest.mark.parametrize('name', ['water_focus', 'skull_plane', 'skull_allmaps', 'dome_stress'])
def test_against_committed_golden_vectors(name):
    """The CUDA path against tests/golden/*.npz (float64 oracle outputs, see make_golden.py)."""
    import os
    from tests.golden import make_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'))
    w = make_golden.build_case(name)
    assert str(g['digest']) == make_golden.inputs_digest(w)
    (Sensor, RMS, Peak, IP), last = run_cuda(w, 0)
    assert np.array_equal(IP['IndexSensorMap'], g['IndexSensorMap'])       # bit-exact index map
    assert IP['IndexSensorMap'].dtype == g['IndexSensorMap'].dtype
    for key in g.files:
        kind, _, mapname = key.partition('_')
        got = {'RMS': RMS, 'Peak': Peak, 'Sensor': Sensor}.get(kind)
        if got is None or mapname not in got:
            continue
        assert rl2(got[mapname], g[key]) <= TOL, (key, rl2(got[mapname], g[key]))
    main = RMS if 'Pressure' in RMS else Peak
    assert same_peak(main['Pressure'], g['RMS_Pressure'] if 'RMS_Pressure' in g.files else g['Peak_Pressure'])


def test_device_sensor_table_matches_numpy_order():
    """bb_fdtd_set_sensor_map: irregular sensor masks give the IndexSensorMap the reference would
    (ascending 1-based Fortran-order index), bit-exact, and rows in that order."""
    w = workloads.make_workload('single_water', shape=(40, 36, 44), periods=2, pml=6)
    rng = np.random.default_rng(11)
    sen = (rng.random(w['args'][0].shape) < 0.07).astype(np.uint32)
    sen[:6] = 0; sen[-6:] = 0; sen[:, :6] = 0; sen[:, -6:] = 0; sen[:, :, :7] = 0; sen[:, :, -6:] = 0
    args = w['args'][:7] + (sen,)
    w2 = dict(args=args, kwargs=w['kwargs'])
    (Sensor, RMS, Peak, IP), _ = run_cuda(w2, 0)
    ref = run_oracle(w2)
    expect = (np.flatnonzero(sen.reshape(-1, order='F')) + 1).astype(np.uint32)
    assert np.array_equal(IP['IndexSensorMap'], expect) and np.array_equal(ref['IndexSensorMap'], expect)
    assert rl2(Sensor['Pressure'], ref['Sensor']['Pressure']) <= TOL


@pytest.mark.parametrize('halo', ['peer', 'nccl'])
def test_two_gpu_slabs_match_single_gpu(halo):
    """Slab decomposition (two handles on two devices, one thread each) against the single-GPU run of the same
    inputs, with the NVLink halo push (boundary CTAs store into the neighbour's halo planes) and with the NCCL
    send/recv exchange.  Skipped on a one-GPU box."""
    import threading
    from babelbrain_b200 import _capi
    from babelbrain_b200.slab import SlabPlan, assemble_maps
    if _capi.device_count() < 2:
        pytest.skip('needs two GPUs')
    w = workloads.make_workload('ctx500_skull', shape=(64, 56, 72), periods=5, pml=8)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    (S1, R1, _, IP1), _ = run_cuda(w, 0)
    uid = FdtdSlab.nccl_unique_id() if halo == 'nccl' else None
    out, err, exports = [None, None], [], [None, None]
    gate = threading.Barrier(2, timeout=120)

    def rank_main(r):
        try:
            s = FdtdSlab(*w['args'], device=r, rank=r, nranks=2, **kw)
            if halo == 'nccl':
                s.comm_init(uid)
            else:
                exports[r] = s.peer_export()
                gate.wait()
                s.peer_attach(exports[r - 1] if r > 0 else None, exports[r + 1] if r < 1 else None)
                gate.wait()
            s.run()
            out[r] = (s.get_map(0, 'Pressure'), s.get_sensors('Pressure'), s.sensor_rows, s.IndexSensorMap)
            gate.wait()          # nobody frees memory a neighbour may still be writing to
            s.close()
        except Exception as e:   # noqa: BLE001
            err.append(e)
            gate.abort()
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not err, err
    plan = SlabPlan(64, 2, 8)
    full = assemble_maps((64, 56, 72), plan, [out[0][0], out[1][0]])
    assert rl2(full, R1['Pressure']) <= 1e-6                                # same kernels, same order of operations
    assert np.array_equal(out[0][3], IP1['IndexSensorMap'])
    sens = np.zeros_like(S1['Pressure'])
    for r in range(2):
        sens[out[r][2]] = out[r][1]
    assert rl2(sens, S1['Pressure']) <= 1e-6


def test_sources_streamed_during_the_run_match_the_upfront_upload():
    """bb_fdtd_set_source_functions_streamed (what the public call uses): the table goes up in chunks of time samples while
    the time loop runs; results must be those of the upload-at-once path bit for bit, for a float64 table and for a float32
    view with a row stride, over several chunks (512 samples each for a small source count) and a partial run + rest."""
    w = workloads.make_workload('ctx500_skull', shape=(30, 26, 36), periods=30, pml=4)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    args = list(w['args'])
    SF = np.asarray(args[4])
    assert SF.shape[1] > 1100                                   # at least three chunks
    wide = np.zeros((SF.shape[0], SF.shape[1] + 7), np.float32)
    wide[:, :SF.shape[1]] = SF
    for table in (SF, wide[:, :SF.shape[1]]):
        args[4] = table
        out = []
        for streamed in (False, True):
            s = FdtdSlab(*args, stream_sources=streamed, **kw)
            if streamed:
                s.run(700)                                      # stop inside the second chunk, then finish
            s.run()
            out.append((s.get_map(0, 'Pressure').copy(), s.get_sensors('Pressure').copy()))
            s.close()
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        assert out[0][0].max() > 0


def test_sensor_rows_placed_by_runs_on_the_device():
    """bb_fdtd_get_sensors_runs (the gather of a multi-GPU run through the public call): a slab's rows, cut into runs,
    land in their places of a larger page-locked table; everything else in the table stays untouched."""
    import ctypes
    from babelbrain_b200 import _capi
    w = workloads.make_workload('ctx500_skull', shape=(30, 26, 36), periods=3, pml=4)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    s = FdtdSlab(*w['args'], **kw)
    s.run()
    ref = s.get_sensors('Pressure')
    n, nsam = ref.shape
    cuts = np.array([0, n // 3, n // 2, n], np.int64)
    src, cnt = np.ascontiguousarray(cuts[:-1]), np.ascontiguousarray(np.diff(cuts))
    dst = np.array([cnt[2] + 3 + cnt[1] + 2, cnt[2] + 3, 0], np.int64)           # the three runs in reverse order, with gaps
    rows = int(n + 7)
    table = _capi.pinned.empty((rows, nsam), np.float32)
    table[:] = -7.0
    done = ctypes.c_int(0)
    call = lambda t, d, sr, c: _capi.lib().bb_fdtd_get_sensors_runs(s._h, _capi.MAP_ID['Pressure'], _capi.ptr(t), t.shape[0], _capi.ptr(d),
                                                                    _capi.ptr(sr), _capi.ptr(c), d.size, ctypes.byref(done))
    _capi.check(call(table, dst, src, cnt))
    assert done.value == 1
    expect = np.full((rows, nsam), -7.0, np.float32)
    for d, a, c in zip(dst, src, cnt):
        expect[d:d + c] = ref[a:a + c]
    assert np.array_equal(table, expect)
    pageable = np.zeros((rows, nsam), np.float32)                                  # not page-locked: nothing written, the caller falls back
    _capi.check(call(pageable, dst, src, cnt))
    assert done.value == 0 and not pageable.any()
    bad = src.copy(); bad[1] += 1                                                  # runs that do not tile the slab's rows
    assert call(table, dst, bad, cnt) != 0
    too_far = dst.copy(); too_far[0] += 10                                         # a run that leaves the table
    assert call(table, too_far, src, cnt) != 0
    s.close()


def test_public_call_on_two_gpus_returns_the_single_gpu_arrays(monkeypatch):
    """NumberGPUs=2 (or BABELB200_NGPUS=2 for an unmodified caller) through the reference-shaped call: same
    tuple, whole-grid arrays, IndexSensorMap bit-exact, maps and sensor traces equal to the one-GPU run."""
    from babelbrain_b200 import _capi
    if _capi.device_count() < 2:
        pytest.skip('needs two GPUs')
    w = workloads.make_workload('h317_skull', shape=(66, 52, 60), periods=4, pml=8)
    PM = PropagationModel()
    S1, L1, R1, P1, IP1 = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    vz1 = np.array(L1['Vz'])
    S2, L2, R2, P2, IP2 = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], NumberGPUs=2, **w['kwargs'])
    assert np.array_equal(IP2['IndexSensorMap'], IP1['IndexSensorMap'])
    assert S2['Pressure'].shape == S1['Pressure'].shape and rl2(S2['Pressure'], S1['Pressure']) <= 1e-6
    for k in R1:
        assert R2[k].shape == w['args'][0].shape and R2[k].flags.writeable
        assert rl2(R2[k], R1[k]) <= 1e-6 and rl2(P2[k], P1[k]) <= 1e-6
    assert rl2(L2['Vz'], vz1) <= 1e-6
    monkeypatch.setenv('BABELB200_NGPUS', '2')
    S3, _, R3, P3, IP3 = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    assert len(PM.last_timing['devices']) == 2 and np.array_equal(R3['Pressure'], R2['Pressure'])
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*w['args'], NumberGPUs=64, **w['kwargs'])


@pytest.mark.parametrize('ratio,tissue_in_shell', [(0.0, False), (0.1, True), (0.05, False)])
def test_long_run_stays_bounded_and_matches_the_oracle_for_both_layers(ratio, tissue_in_shell):
    """100 periods (4800 steps).  MPMLRatio 0 -- the classical split-field layer, the default -- on the caller's kind
    of map (water inside the shell, BabelIntegrationBASE.py:2154-2159); the multi-axial layer on the map the classical
    one cannot survive (tissue inside the shell, tests/test_oracle.py).  Both stay at the steady state and agree with
    the oracle run at the same ratio."""
    w = workloads.make_workload('ctx500_skull', shape=(40, 36, 56), periods=100, pml=6, tissue_in_shell=tissue_in_shell)
    (Sensor, RMS, _, _), _ = run_cuda(w, MPMLRatio=ratio)
    ref = run_oracle(w, MPMLRatio=ratio)
    assert float(RMS['Pressure'].max()) < 3e6
    assert rl2(RMS['Pressure'], ref['RMS']['Pressure']) <= TOL and same_peak(RMS['Pressure'], ref['RMS']['Pressure'])
    assert rl2(Sensor['Pressure'], ref['Sensor']['Pressure']) <= TOL


def test_mpml_ratio_through_the_public_call():
    """MPMLRatio is a keyword of the public call (default 0 = classical layer); both layers against the oracle on a
    short run of a map with tissue inside the shell, where they differ at the 1e-2 level."""
    w = workloads.make_workload('ctx500_skull', shape=(48, 40, 64), periods=8, pml=8, tissue_in_shell=True)
    PM = PropagationModel()
    maps = {}
    for ratio in (None, 0.0, 0.1):
        kw = dict(w['kwargs']) if ratio is None else dict(w['kwargs'], MPMLRatio=ratio)
        _, _, RMS, _ = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **kw)
        ref = run_oracle(w, MPMLRatio=0.0 if ratio is None else ratio)
        assert rl2(RMS['Pressure'], ref['RMS']['Pressure']) <= TOL, ratio
        maps[ratio] = np.array(RMS['Pressure'])
    assert np.array_equal(maps[None], maps[0.0])
    assert rl2(maps[0.1], maps[0.0]) > 1e-3
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], MPMLRatio=-0.5))


def test_full_size_ctx500_properties():
    """BASELINE configs[1] at full size (240x240x320, 2544 steps), where the oracle is too slow to be the
    checker: size-independent properties instead.  Linear in the source amplitude, zero RMS inside the PML
    shell, sensor table = every voxel of the sensor box in IndexSensorMap order, p == -Sigmaxx wherever the
    wave has only crossed lossless water, finite everywhere."""
    w = workloads.make_workload('ctx500_skull')
    over = dict(SelMapsRMSPeakList=['Pressure', 'Sigmaxx'])
    (S1, R1, _, IP), _ = run_cuda(w, 0, **over)
    p = R1['Pressure']
    assert p.shape == (240, 240, 320) and np.isfinite(p).all() and np.isfinite(S1['Pressure']).all()
    pml = w['meta']['pml']
    assert not p[:pml].any() and not p[-pml:].any() and not p[:, :pml].any() and not p[:, :, -pml:].any()
    assert p[pml:-pml, pml:-pml, pml + 1:-pml].min() > 0
    n1, n2, n3 = p.shape
    assert IP['IndexSensorMap'].size == (n1 - 2 * pml) * (n2 - 2 * pml) * (n3 - 2 * pml - 1)
    assert S1['Pressure'].shape == (IP['IndexSensorMap'].size, 2 * w['meta']['ppp'] // w['meta']['sub'])
    assert np.all(np.diff(IP['IndexSensorMap'].astype(np.int64)) > 0)
    # in front of the skull the medium is lossless water: pressure and -Sigmaxx coincide there
    water = (w['args'][0][pml:-pml, pml:-pml, pml + 1:pml + 24] == 0).all()
    assert water
    assert rl2(R1['Pressure'][pml:-pml, pml:-pml, pml + 1:pml + 20], R1['Sigmaxx'][pml:-pml, pml:-pml, pml + 1:pml + 20]) <= 1e-4
    args = list(w['args'])
    args[4] = args[4] * 0.5
    (S2, R2, _, _), _ = run_cuda(dict(args=tuple(args), kwargs=w['kwargs']), 0, **over)
    assert rl2(R2['Pressure'], 0.5 * p.astype(np.float64)) <= 1e-5
    assert rl2(S2['Pressure'], 0.5 * S1['Pressure'].astype(np.float64)) <= 1e-5


def test_rayleigh_variants_against_oracle():
    """ForwardSimple as the transducer files call it: whole-grid fields, single points (phase programming,
    BabelIntegrationANNULAR_ARRAY.py:383), attenuating wavenumber, MaxDistance, per-point amplitudes (u0step)."""
    from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple, InitCuda
    InitCuda('B200')
    rng = np.random.default_rng(21)
    nsrc = 700
    center = (rng.random((nsrc, 3)).astype(np.float32) - 0.5) * 0.05
    center[:, 2] = -0.04 - 0.01 * rng.random(nsrc).astype(np.float32)
    ds = np.full((nsrc, 1), 3e-6, np.float32)
    u0 = (rng.random(nsrc) + 1j * rng.random(nsrc)).astype(np.complex64)
    k0 = 2 * np.pi * 7e5 / 1500
    for npts in (1, 128, 5000):
        rf = (rng.random((npts, 3)).astype(np.float32) - 0.5) * 0.06
        rf[:, 2] = rng.random(npts).astype(np.float32) * 0.08
        for k in (k0 + 0j, k0 - 4.0j):
            got = ForwardSimple(np.array(k).astype(np.complex64), center, ds, u0, rf)
            ref = oracle.rayleigh_numpy(np.complex64(k), center, ds, u0, rf)
            assert got.dtype == np.complex64 and got.shape == (npts,)
            # a single point is an ill-conditioned sum (700 random phases cancel): float32 itself is at 3e-5 there
            assert rl2(got, ref) <= (TOL if npts > 1 else 2 * TOL), (npts, k, rl2(got, ref))
        got = ForwardSimple(np.array(k0 + 0j).astype(np.complex64), center, ds, u0, rf, MaxDistance=0.07)
        ref = oracle.rayleigh_numpy(np.complex64(k0), center, ds, u0, rf, MaxDistance=0.07)
        assert rl2(got, ref) <= TOL
    # per-point source amplitudes
    npts = 64
    rf = (rng.random((npts, 3)).astype(np.float32) - 0.5) * 0.06
    u0pp = (rng.random((npts, nsrc)) + 1j * rng.random((npts, nsrc))).astype(np.complex64)
    got = ForwardSimple(np.array(k0 + 0j).astype(np.complex64), center, ds, u0pp.reshape(-1), rf, u0step=nsrc)
    ref = np.array([oracle.rayleigh_numpy(np.complex64(k0), center, ds, u0pp[n], rf[n:n + 1])[0] for n in range(npts)])
    assert rl2(got, ref) <= TOL
    with pytest.raises(ValueError):
        ForwardSimple(np.array(k0 + 0j).astype(np.complex64), center, ds[:-1], u0, rf)


def test_edge_cases_empty_sources_sensors_peak_only_short_pulse():
    """Empty and ragged inputs: no source voxel, no sensor voxel, peak-only maps, a pulse table shorter than
    the run (the source stops, BabelIntegrationSingle.py:315-316), DT=None (stable step), odd sizes that are
    not multiples of the 8 x 64 tile, CheckOnlyParams."""
    w = workloads.make_workload('ctx500_skull', shape=(37, 43, 67), periods=4, pml=5)
    MM, ML, f, SM, SF, h, T, SEN = w['args']
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    # no sources: everything stays exactly zero
    s = FdtdSlab(MM, ML, f, np.zeros_like(SM), SF, h, T, SEN, **kw)
    s.run()
    Sensor, RMS, Peak, IP = collect_results(s)
    s.close()
    assert not RMS['Pressure'].any() and not Sensor['Pressure'].any()
    # no sensors: empty table with the right number of columns
    (S0, R0, _, IP0), _ = run_cuda(dict(args=(MM, ML, f, SM, SF, h, T, np.zeros_like(SEN)), kwargs=w['kwargs']), 0)
    assert S0['Pressure'].shape == (0, S0['time'].size) and IP0['IndexSensorMap'].size == 0
    ref = run_oracle(w)
    assert rl2(R0['Pressure'], ref['RMS']['Pressure']) <= TOL
    # peak only
    (S2, _, P2, _), _ = run_cuda(w, 0, SelRMSorPeak=2)
    refp = run_oracle(w, SelRMSorPeak=2)
    assert rl2(P2['Pressure'], refp['Peak']['Pressure']) <= TOL
    r4 = PropagationModel().StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], SelRMSorPeak=2))
    assert len(r4) == 4 and rl2(r4[2]['Pressure'], refp['Peak']['Pressure']) <= TOL
    # pulse shorter than the run, and the solver's own stable step
    short = SF[:, :SF.shape[1] // 2]
    w3 = dict(args=(MM, ML, f, SM, short, h, T, SEN), kwargs=dict(w['kwargs'], DT=None, AlphaCFL=0.8, SensorStart=4))   # 0.8: inside the O(2,4) stability limit 6/7
    (S3, R3, _, _), _ = run_cuda(w3, 0)
    ref3 = run_oracle(w3)
    assert S3['time'].size == ref3['Sensor']['time'].size
    assert rl2(R3['Pressure'], ref3['RMS']['Pressure']) <= TOL and rl2(S3['Pressure'], ref3['Sensor']['Pressure']) <= TOL
    assert PropagationModel().StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], CheckOnlyParams=True)) is None


def test_full_size_ctx500_against_the_committed_oracle_reduction():
    """BASELINE configs[1] (the benchmark workload) at FULL size, 240x240x320 and all 2544 steps, against the float64
    C oracle's run of the same inputs (tests/golden/make_ctx500_full_golden.py, ~17 min of CPU, committed as a
    reduction): identical peak voxel, the three RMS planes and the line through it, the norm of the whole map, every
    97th sensor trace and index, all within the north-star tolerance."""
    import os
    from tests.golden import make_ctx500_full_golden as G
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ctx500_full.npz'))
    w = G.build_case()
    assert str(g['digest']) == G.digest(w), 'the seeded workload generator drifted: regenerate tests/golden/ctx500_full.npz'
    (Sensor, RMS, _, IP), _ = run_cuda(w, 0)
    p = RMS['Pressure']
    pk = tuple(int(x) for x in g['peak_voxel'])
    assert tuple(int(x) for x in np.unravel_index(int(np.argmax(p)), p.shape)) == pk
    assert abs(float(p[pk]) / float(g['peak_value']) - 1) <= TOL
    assert rl2(p[pk[0]], g['plane_i']) <= TOL and rl2(p[:, pk[1]], g['plane_j']) <= TOL and rl2(p[:, :, pk[2]], g['plane_k']) <= TOL
    assert rl2(p[pk[0], pk[1]], g['line_k']) <= TOL
    assert abs(float(np.linalg.norm(p.astype(np.float64))) / float(g['rms_norm']) - 1) <= TOL
    assert IP['IndexSensorMap'].size == int(g['nsensors'])
    assert np.array_equal(IP['IndexSensorMap'][::G.ROW_STRIDE], g['index_rows'])
    assert rl2(Sensor['Pressure'][::G.ROW_STRIDE], g['sensor_rows']) <= TOL
    assert abs(float(np.linalg.norm(Sensor['Pressure'].astype(np.float64))) / float(g['sensor_norm']) - 1) <= TOL
    assert np.allclose(Sensor['time'], g['time'], rtol=1e-12) and int(g['steps']) == w['meta']['steps']
