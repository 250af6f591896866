"""Bio-heat solver: the NumPy oracle against closed forms on CPU; the CUDA path
(BabelViscoFDTD.tools.RayleighAndBHTE.BHTE / BHTEMultiplePressureFields -> bb_bhte_run) against the oracle on the GPU.
Tolerance (float32 CUDA vs float64 oracle over hundreds of explicit steps): temperature within 1e-3 K (1e-4 of the
temperature rise), CEM43 dose -- an exponential of the temperature -- within 1e-3 relative L2."""
import numpy as np
import pytest

from oracle import bhte_numpy
from babelbrain_b200 import thermal

ML = {'Density': np.array([1000.0, 1116.0, 1896.5, 1738.0, 1041.0]), 'SoS': np.array([1500.0, 1537.0, 2476.0, 2220.0, 1562.0]),
      'Attenuation': np.array([0.0, 2.3, 81.0, 81.0, 3.45]), 'SpecificHeat': np.array([4178.0, 3391.0, 1313.0, 2274.0, 3630.0]),
      'Conductivity': np.array([0.6, 0.37, 0.32, 0.31, 0.51]), 'Perfusion': np.array([0.0, 106.0, 10.0, 30.0, 559.0]),
      'Absorption': np.array([0.0, 0.85, 0.16, 0.15, 0.85]), 'InitTemperature': np.full(5, 37.0)}    # CalculateTemperatureEffects.py:780-792


def case(shape=(28, 24, 36), seed=3):
    rng = np.random.default_rng(seed)
    MM = np.zeros(shape, np.uint32)
    MM[:, :, 8:10] = 1
    MM[:, :, 10:16] = 2
    MM[6:20, 5:18, 12:14] = 3
    MM[:, :, 16:] = 4
    x, y, z = np.meshgrid(*(np.arange(n) for n in shape), indexing='ij')
    P = 5.5e6 * np.exp(-((x - 14) ** 2 + (y - 12) ** 2) / 18.0 - (z - 22) ** 2 / 60.0) * (1 + 0.05 * rng.standard_normal(shape))
    return MM, P.astype(np.float32)


def test_dose_rule_closed_forms():
    # constant temperature: dose = dt * R^(43 - T); at 43 C one second of exposure is one second of dose
    assert bhte_numpy.cem43_increment(43.0, 43.0, 0.1) == pytest.approx(0.1)
    assert bhte_numpy.cem43_increment(44.0, 44.0, 0.1) == pytest.approx(0.2)
    assert bhte_numpy.cem43_increment(41.0, 41.0, 0.1) == pytest.approx(0.1 / 16)
    # a ramp inside one regime equals the integral of R^(43-T(t)) dt
    t = np.linspace(0, 0.1, 20001)
    for a, b in ((38.0, 41.5), (44.0, 47.0), (41.0, 45.0), (46.0, 42.0)):
        T = a + (b - a) * t / 0.1
        ref = np.trapezoid(np.where(T >= 43, 0.5, 0.25) ** (43 - T), t)
        assert bhte_numpy.cem43_increment(a, b, 0.1) == pytest.approx(ref, rel=2e-4)


def test_oracle_energy_and_steady_state():
    MM, P = case()
    dx, dt = 0.5e-3, 0.01
    bh, perf, q = bhte_numpy.coefficients(ML, 5, dx, dt)
    assert np.all(bh < 1 / 6) and perf[0] == 0 and q[0] == 0            # water: no perfusion, no absorption
    Q = thermal.heat_source(P, MM.astype(np.int64), ML, dx, dt)[None]
    assert np.allclose(Q[0], (P.astype(float) ** 2 * q[MM]).astype(np.float32), rtol=1e-6)
    # no beam: the volume stays at the core temperature and the dose grows at dt * 0.25^6 per step
    T, D, _, _ = bhte_numpy.run(Q, MM, ML, dx, 40, np.full(40, -1), dt=dt)
    assert np.allclose(T, 37.0) and np.allclose(D[1:-1, 1:-1, 1:-1], 40 * dt * 0.25 ** 6)
    # beam on: heating where tissue absorbs, none in water; faces keep their value
    T, D, _, _ = bhte_numpy.run(Q, MM, ML, dx, 60, np.zeros(60, int), dt=dt)
    assert T[14, 12, 22] > 37.5 and np.allclose(T[:, :, 2:5], 37.0, atol=1e-3) and np.all(T[0] == 37.0) and np.all(T[:, :, -1] == 37.0)
    with pytest.raises(ValueError):
        thermal.getBHTECoefficient(0.6, 1000.0, 4178.0, 0.1e-3, 1.0, dt=0.05)   # unstable explicit step


def test_field_schedule_of_several_foci():
    s = thermal.field_schedule(np.array([[2, 1], [3, 2]]), 19)
    assert s.tolist() == [0, 0, -1, 1, 1, 1, -1, -1] * 2 + [0, 0, -1]


@pytest.mark.gpu
def test_bhte_against_the_oracle():
    from BabelViscoFDTD.tools.RayleighAndBHTE import BHTE, InitCuda
    InitCuda('')
    MM, P = case()
    dx, dt, steps, on = 0.5e-3, 0.01, 300, 200
    pts = np.zeros(MM.shape, np.uint32)
    pts[14, 12, 22] = 1
    pts[10, 9, 12] = 2
    T, D, Slice, Qarr, TP = BHTE(P, MM, ML, dx, steps, on, 12, nFactorMonitoring=3, dt=dt, DutyCycle=0.6, MonitoringPointsMap=pts, stableTemp=37.0)
    Q = thermal.heat_source(P, MM.astype(np.int64), ML, dx, dt, 0.6)[None]
    assert np.array_equal(Qarr, Q[0])
    sched = np.where(np.arange(steps) < on, 0, -1)
    rT, rD, rS, rP = bhte_numpy.run(Q, MM, ML, dx, steps, sched, dt=dt, LocationMonitoring=12, nFactorMonitoring=3, MonitoringPointsMap=pts)
    assert T.dtype == np.float32 and T.shape == MM.shape and Slice.shape == (MM.shape[0], MM.shape[2], steps // 3) and TP.shape == (2, steps)
    errs = dict(T=float(np.abs(T - rT).max()), rise=float(np.abs(rT - 37.0).max()), D=float(np.linalg.norm(D - rD) / np.linalg.norm(rD)),
                S=float(np.abs(Slice - rS).max()), P=float(np.abs(TP - rP).max()))
    print('BHTE vs oracle:', errs)
    assert errs['T'] < 1e-3 and errs['T'] / errs['rise'] < 1e-4, errs
    assert errs['D'] < 1e-3, errs
    assert errs['S'] < 1e-3 and errs['P'] < 1e-3, errs
    assert rT.max() > 39.0                                               # the case does heat
    # second segment continuing from the first (beam off), as RunBHTECycles chains them (CalculateTemperatureEffects.py:406-420)
    T2, D2, _, _ = BHTE(P * 0, MM, ML, dx, 100, 0, -1, dt=dt, initT0=T, initDose=D)
    rT2, rD2, _, _ = bhte_numpy.run(Q * 0, MM, ML, dx, 100, np.full(100, -1), dt=dt, initT0=rT, initDose=rD)
    assert np.abs(T2 - rT2).max() < 1e-3 and np.linalg.norm(D2 - rD2) / np.linalg.norm(rD2) < 1e-3 and T2.max() < T.max()


@pytest.mark.gpu
def test_bhte_multiple_pressure_fields_against_the_oracle():
    from BabelViscoFDTD.tools.RayleighAndBHTE import BHTEMultiplePressureFields
    MM, P = case()
    P2 = np.roll(P, 5, axis=0)
    dx, dt, steps = 0.5e-3, 0.01, 240
    onoff = np.array([[20, 10], [15, 5]], np.int32)
    T, D, Slice, QL = BHTEMultiplePressureFields(np.stack([P, P2]), MM, ML, dx, steps, onoff, -1, dt=dt)
    Q = np.stack([thermal.heat_source(x, MM.astype(np.int64), ML, dx, dt) for x in (P, P2)])
    assert np.array_equal(QL, Q) and not Slice.any()
    rT, rD, _, _ = bhte_numpy.run(Q, MM, ML, dx, steps, thermal.field_schedule(onoff, steps), dt=dt)
    assert np.abs(T - rT).max() < 1e-3 and np.linalg.norm(D - rD) / np.linalg.norm(rD) < 1e-3
