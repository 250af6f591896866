"""Host-side preparation (babelbrain_b200/hostprep.py, the float64 part of the reference's
CalculateMatricesForPropagation / StaggeredFDTD_3D_with_relaxation) against the independent
restatement inside the oracle, plus the bookkeeping rules the caller relies on
(TranscranialModeling/BabelIntegrationBASE.py:1799-1829, 2082-2109, 2490-2511)."""
import numpy as np
import pytest

from oracle import fdtd_numpy
from babelbrain_b200 import hostprep, workloads


@pytest.mark.parametrize('f', [250e3, 500e3, 700e3, 1e6])
@pytest.mark.parametrize('qfc', [True, False])
def test_material_table_matches_oracle(f, qfc):
    rows = workloads.material_rows(f)
    ML = np.array(list(rows.values()))
    qc = np.array([1.0, 1.0, 3.0, 3.0, 1.0])
    h = workloads.SHEAR_FLOOR_SOS / f / 6
    T, A = hostprep.material_table(ML, f, qfc, h, qc)
    R = fdtd_numpy.material_tables(ML, f, qfc, h, qc)
    for col, key in enumerate(('M', 'G', 'L', 'B', 'tauL', 'tauS', 'ots', 'K')):
        assert np.allclose(T[:, col], R[key], rtol=1e-13, atol=0), key
    assert np.allclose(A['QL'], R['QL']) and np.allclose(A['QS'], R['QS'])
    # fluids carry no shear modulus / shear relaxation; lossless water has no memory variable at all
    assert T[0, 1] == 0 and T[0, 4] == 0 and T[0, 5] == 0 and T[0, 6] == 0
    assert np.all(T[[1, 4], 1] == 0) and np.all(T[[1, 4], 4] > 0)
    assert hostprep.stable_dt(ML, h, 0.5) == pytest.approx(fdtd_numpy.ideal_dt(ML, h, 0.5), rel=1e-15)
    # linear in AlphaCFL up to the stability limit of the O(2,4) scheme, 6/7, and capped there (the caller's water-only
    # normalisation step asks for AlphaCFL = 1.0, BabelIntegrationBASE.py:1801)
    assert hostprep.stable_dt(ML, h, 0.5) == pytest.approx(0.5 * np.sqrt(3) / 3 * h / ML[:, 1].max(), rel=1e-15)
    assert hostprep.stable_dt(ML, h, 1.0) == pytest.approx(6 / 7 * np.sqrt(3) / 3 * h / ML[:, 1].max(), rel=1e-15)
    assert hostprep.stable_dt(ML, h, 1.0) == fdtd_numpy.ideal_dt(ML, h, 0.99) == hostprep.hard_limit_dt(ML, h)


def test_relaxation_fit_meets_q_exactly():
    f = 500e3
    QL = np.array([0.0, 30.0, 12.0, 80.0])
    QS = np.array([0.0, 0.0, 8.0, 40.0])
    tauL, tauS, ots = hostprep.relaxation_fit(f, QL, QS)
    w = 2 * np.pi * f
    for m in range(4):
        if ots[m] == 0:
            assert tauL[m] == 0 and tauS[m] == 0
            continue
        x = w / ots[m]
        for tau, Q in ((tauL[m], QL[m]), (tauS[m], QS[m])):
            if Q > 0:
                y = x * (1 + tau)
                assert (1 + x * y) / (y - x) == pytest.approx(Q, rel=1e-12)   # Q(omega) of a standard linear solid


def test_pml_table_and_steps():
    P, h, dt = 12, 3.675e-4, 4.1667e-8
    t = hostprep.pml_table(P, h, dt, 2476.0, 1e-5)
    o = np.stack(fdtd_numpy.pml_damping(P, h, 2476.0, 1e-5))
    assert t.shape == (2, P + 1) and np.allclose(t, o, rtol=1e-14)
    assert t[0, 0] == 0.0 and t[0, P] == pytest.approx(np.log(1e5) * 3 * 2476.0 / (2 * P * h))   # no damping at depth 0, d0 at depth P
    assert np.all(np.diff(t[0]) > 0) and np.all(np.diff(t[1]) > 0) and np.all(t[1, :-1] > t[0, :-1]) and np.all(t[1, :-1] < t[0, 1:])
    # the classical coefficients follow from the damping (oracle/fdtd_numpy.py: pml_tables)
    inv, dx, invh, dxh = fdtd_numpy.pml_tables(P, h, dt, 2476.0, 1e-5)
    assert np.allclose(inv, 1 / (1 / dt + t[0] / 2)) and np.allclose(dxh, 1 / dt - t[1] / 2)
    assert hostprep.MPML_RATIO == fdtd_numpy.MPML_RATIO == 0.0      # classical split-field layer by default
    # TimeSimulation = dt*steps must give back `steps` despite floating point (BabelIntegrationBASE.py:2089)
    for steps in (720, 2544, 5616, 11250):
        assert hostprep.number_of_steps(dt * steps, dt) == steps
    assert hostprep.number_of_steps(dt * 100.4, dt) == 101


def test_sampling_rules_of_the_caller():
    """SensorSubSampling / SensorStart as UpdateConditions derives them: the number of samples is a
    whole number of periods (BabelIntegrationBASE.py:2490-2496 slices by PPP/SensorSubSampling)."""
    for name in ('single_water', 'ctx500_skull', 'h317_skull', 'hires_1mhz'):
        cfg = workloads.CONFIGS[name]
        rows = np.array(list(workloads.material_rows(cfg['frequency']).values()))
        S = workloads.sizing(cfg['frequency'], cfg['ppw'], rows, cfg['shape'])
        n = hostprep.sample_steps(S['steps'], S['sub'], S['sensor_start'])
        per_period = S['ppp'] // S['sub']
        assert S['ppp'] % S['sub'] == 0 and per_period >= 4
        assert n.size == 2 * per_period and n.size % per_period == 0
        assert n[0] == S['sensor_start'] * S['sub'] and np.all(np.diff(n) == S['sub'])
        assert S['dt'] <= S['dt_ideal'] * (1 + 1e-12)


def test_sizing_reproduces_the_survey_table():
    """PPP and step counts of the BASELINE configs."""
    expect = {'single_water': (30, 720), 'ctx500_skull': (48, 2544), 'h317_skull': (72, 5616), 'hires_1mhz': (75, 11250)}
    for name, (ppp, steps) in expect.items():
        cfg = workloads.CONFIGS[name]
        rows = np.array(list(workloads.material_rows(cfg['frequency']).values()))
        if name == 'single_water':
            rows = rows[:1]
        S = workloads.sizing(cfg['frequency'], cfg['ppw'], rows, cfg['shape'])
        assert (S['ppp'], S['steps']) == (ppp, steps), (name, S['ppp'], S['steps'])


def test_maps_mask_and_errors():
    assert hostprep.maps_mask(['Pressure']) == 1 << 10
    assert hostprep.maps_mask(['ALLV', 'Vz']) == 0b1001
    with pytest.raises(ValueError):
        hostprep.maps_mask(['Pressur'])
    with pytest.raises(ValueError):
        hostprep.material_table(np.ones((2, 4)), 5e5, True, 1e-3)


def test_calculate_matrices_tuple():
    from babelbrain_b200.propagation import PropagationModel
    ML = np.array(list(workloads.material_rows(5e5).values()))
    h = 3.675e-4
    r = PropagationModel().CalculateMatricesForPropagation(np.zeros((10, 10, 5), np.uint32), ML, 5e5, True, h, 0.5)
    assert len(r) == 10
    assert r[0] == pytest.approx(0.5 * np.sqrt(3) / 3 * h / ML[:, 1].max())
    o = fdtd_numpy.calculate_matrices_for_propagation(None, ML, 5e5, True, h, 0.5)
    for a, b in zip(r[1:], o[1:]):
        assert np.allclose(a, b, rtol=1e-12)
