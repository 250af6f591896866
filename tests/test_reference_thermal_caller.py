"""The UNMODIFIED driver of BabelBrain's thermal step -- ThermalModeling/CalculateTemperatureEffects.py: RunBHTECycles
(:259-460): beam-on BHTE, cooling-off BHTE, pause between groups, initT0 / initDose chained from cycle to cycle,
MonitoringPointsMap traces stacked along time -- on this repository's BHTE.

CPU: the driver runs with BHTE answered by the oracle (oracle/bhte_numpy.py): pins the harness, the argument
conventions the driver uses (MaterialList entries as Python lists, nStepsOn, LocationMonitoring = the y index `cy`,
five return values when MonitoringPointsMap is given) and two closed forms of the scheme.
GPU: the same driver on BabelViscoFDTD.tools.RayleighAndBHTE.BHTE of the shim (-> bb_bhte_run) against that oracle run:
temperatures within 2e-3 K, dose within 2e-3 relative, traces within 2e-3 K over 2 x (30 on/off + 20 off) + 15 steps.
"""
import numpy as np
import pytest

from tests import refcaller
from oracle import bhte_numpy

pytestmark = pytest.mark.skipif(refcaller.reference_root() is None, reason='no reference tree (python tests/make_ref_install.py)')

# CalculateTemperatureEffects.py:780-792: Python lists, water / skin / cortical / trabecular / brain
ML = {'Density': np.array([1000.0, 1116.0, 1896.5, 1738.0, 1041.0]), 'SoS': np.array([1500.0, 1537.0, 2476.0, 2220.0, 1562.0]),
      'Attenuation': np.array([0.0, 2.3, 81.0, 81.0, 3.45]), 'SpecificHeat': [4178.0, 3391.0, 1313.0, 2274.0, 3630.0],
      'Conductivity': [0.6, 0.37, 0.32, 0.31, 0.51], 'Perfusion': [0.0, 106.0, 10.0, 30.0, 559.0],
      'Absorption': [0, 0.85, 0.16, 0.15, 0.85], 'InitTemperature': [37.0] * 5}
DX, DT = 0.5e-3, 0.01
CYCLE = dict(Repetitions=2, TotalIterations=2, TotalDurationBetweenGroups=15, TotalDurationStepsOff=20, TotalDurationSteps=30, nStepsOn=18,
             nFactorMonitoring=5, DutyCycle=0.3)


def case(shape=(28, 24, 36)):
    MM = np.zeros(shape, np.uint32)
    MM[:, :, 8:10] = 1
    MM[:, :, 10:16] = 2
    MM[6:20, 5:18, 12:14] = 3
    MM[:, :, 16:] = 4
    x, y, z = np.meshgrid(*(np.arange(n) for n in shape), indexing='ij')
    P = (9.0e6 * np.exp(-((x - 14) ** 2 + (y - 12) ** 2) / 18.0 - (z - 22) ** 2 / 60.0)).astype(np.float32)
    MP = np.zeros(shape, np.uint32)          # ids 1..3: focus, skull, skin (CalculateTemperatureEffects.py:1003-1022)
    MP[14, 12, 22], MP[14, 12, 12], MP[14, 12, 8] = 1, 2, 3
    return MM, P, MP


def oracle_bhte(Pressure, MaterialMap, MaterialList, dx, TotalDurationSteps, nStepsOn, LocationMonitoring, nFactorMonitoring=1, dt=0.1,
                blood_rho=1050, blood_ct=3617, stableTemp=37.0, DutyCycle=1.0, Backend='CUDA', MonitoringPointsMap=None, initT0=None,
                initDose=None):
    MM = np.asarray(MaterialMap).astype(np.int64)
    _, _, q = bhte_numpy.coefficients(MaterialList, int(MM.max()) + 1, dx, dt, blood_rho, blood_ct)
    Q = (np.asarray(Pressure, np.float64) ** 2 * q[MM] * DutyCycle).astype(np.float32)
    steps = int(TotalDurationSteps)
    T, D, Slice, pts = bhte_numpy.run(Q[None], MM, MaterialList, dx, steps, np.where(np.arange(steps) < int(nStepsOn), 0, -1), dt=dt,
                                      stableTemp=stableTemp, initT0=initT0, initDose=initDose, LocationMonitoring=LocationMonitoring,
                                      nFactorMonitoring=nFactorMonitoring, MonitoringPointsMap=MonitoringPointsMap)
    return (T, D, Slice, Q, pts) if MonitoringPointsMap is not None else (T, D, Slice, Q)


def oracle_bhte_multi(PressureFields, MaterialMap, MaterialList, dx, TotalDurationSteps, nStepsOnOffList, LocationMonitoring, nFactorMonitoring=1,
                      dt=0.1, blood_rho=1050, blood_ct=3617, stableTemp=37.0, Backend='CUDA', MonitoringPointsMap=None, initT0=None, initDose=None):
    MM = np.asarray(MaterialMap).astype(np.int64)
    _, _, q = bhte_numpy.coefficients(MaterialList, int(MM.max()) + 1, dx, dt, blood_rho, blood_ct)
    Q = (np.asarray(PressureFields, np.float64) ** 2 * q[MM][None]).astype(np.float32)
    steps = int(TotalDurationSteps)
    turn = []                                  # the foci take turns: on, off, next focus ... (CalculateTemperatureEffects.py:716-736)
    for m, (on, off) in enumerate(np.asarray(nStepsOnOffList).reshape(-1, 2)):
        turn += [m] * int(on) + [-1] * int(off)
    sched = np.array((turn * (steps // len(turn) + 1))[:steps])
    T, D, Slice, pts = bhte_numpy.run(Q, MM, MaterialList, dx, steps, sched, dt=dt, stableTemp=stableTemp, initT0=initT0, initDose=initDose,
                                      LocationMonitoring=LocationMonitoring, nFactorMonitoring=nFactorMonitoring,
                                      MonitoringPointsMap=MonitoringPointsMap)
    return (T, D, Slice, Q, pts) if MonitoringPointsMap is not None else (T, D, Slice, Q)


def drive_multi(mod, bhte=None, multi=None):
    """The electronic-steering branch of the driver: a list of input files, one pressure field each, BHTEMultiplePressureFields
    while the beam is on and plain BHTE with a zero field while it is off."""
    MM, P, MP = case()
    P2 = np.roll(P, 5, axis=0) * 0.8
    if bhte is not None:
        mod.BHTE, mod.BHTEMultiplePressureFields = bhte, multi
    return mod.RunBHTECycles(nCurrent=0, LimitBHTEIterationsPerProcess=100, InputPData=['a_DataForSim.h5', 'b_DataForSim.h5'],
                             PMaps=np.stack([P, P2]), MaterialMap=MM, MaterialList=ML, dx=DX, cy=12, dt=DT, Backend='CUDA',
                             MonitoringPointsMap=MP, stableTemp=37.0, TemperaturePoints=None, FinalTemp=None, FinalDose=None,
                             PreviousData=None, **dict(CYCLE, nStepsOn=np.array([[3, 2], [4, 1]], np.int32)))


def drive(mod, bhte=None):
    """RunBHTECycles of the unmodified driver; bhte = None keeps the BHTE it imported (the shim's)."""
    MM, P, MP = case()
    if bhte is not None:
        mod.BHTE = bhte
    return mod.RunBHTECycles(nCurrent=0, LimitBHTEIterationsPerProcess=100, InputPData='Single_DataForSim.h5', PMaps=P, MaterialMap=MM,
                             MaterialList=ML, dx=DX, cy=12, dt=DT, Backend='CUDA', MonitoringPointsMap=MP, stableTemp=37.0,
                             TemperaturePoints=None, FinalTemp=None, FinalDose=None, PreviousData=None, **CYCLE)


def test_unmodified_thermal_driver_on_the_oracle():
    mod = refcaller.load_thermal()
    TMax, Dose, FinalT, FinalD, Pts, nxt = drive(mod, oracle_bhte)
    MM, P, MP = case()
    total = 2 * (CYCLE['TotalDurationSteps'] + CYCLE['TotalDurationStepsOff']) + CYCLE['TotalDurationBetweenGroups']
    assert nxt == 2 and Pts.shape == (3, total) and TMax.shape == MM.shape
    focus = Pts[0]
    # heating while the beam is on, cooling afterwards, the second cycle starting from where the first ended
    on, off = CYCLE['nStepsOn'], CYCLE['TotalDurationSteps'] + CYCLE['TotalDurationStepsOff']
    assert np.all(np.diff(focus[:on]) > 0) and np.all(np.diff(focus[on:off]) < 0)
    assert focus[off] > focus[off - 1] and focus[off + on - 1] > focus[on - 1]
    # ResTempMax is the largest END-of-sonication-call temperature over the cycles (:399-402), not the peak inside a call
    n_call = CYCLE['TotalDurationSteps']
    assert TMax[14, 12, 22] == pytest.approx(max(focus[n_call - 1], focus[off + n_call - 1]), abs=1e-4)
    assert TMax[14, 12, 22] > FinalT[14, 12, 22] > 37.0
    # first step at the focus: the temperature rises by DutyCycle * p^2 * q(brain), nothing else has acted yet
    _, _, q = bhte_numpy.coefficients(ML, 5, DX, DT)
    assert focus[0] - 37.0 == pytest.approx(CYCLE['DutyCycle'] * float(P[14, 12, 22]) ** 2 * q[4], rel=1e-5)
    # water neither absorbs nor is perfused: far from the beam it stays at the core temperature
    assert np.allclose(FinalT[:, :, 1:4], 37.0, atol=1e-3)
    # dose in seconds, monotone, and dominated by the hot spot
    assert np.all(FinalD >= Dose - 1e-12) and np.unravel_index(np.argmax(FinalD), FinalD.shape)[2] >= 10
    # the steering branch of the driver runs as well, and two foci heat two places
    TMax2 = drive_multi(mod, oracle_bhte, oracle_bhte_multi)[0]
    assert TMax2[14, 12, 22] > 37.5 and TMax2[19, 12, 22] > 37.5


@pytest.mark.gpu
def test_unmodified_thermal_driver_through_the_cuda_path():
    mod = refcaller.load_thermal()
    from BabelViscoFDTD.tools.RayleighAndBHTE import InitCuda
    InitCuda('B200')
    shim_bhte = mod.BHTE                       # what the driver imported: babelbrain_b200.thermal.BHTE
    assert shim_bhte.__module__ == 'babelbrain_b200.thermal'
    got = drive(mod)
    ref = drive(mod, oracle_bhte)
    mod.BHTE = shim_bhte
    names = ('ResTempMax', 'ResDose', 'FinalTemp', 'FinalDose', 'TemperaturePoints')
    for name, a, b in zip(names, got, ref):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if 'Dose' in name:
            assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(b), name
        else:
            assert np.abs(a - b).max() <= 2e-3, (name, np.abs(a - b).max())
    assert got[5] == ref[5] == 2
    assert got[0].max() - 37.0 > 1.0           # the comparison is not between two cold volumes
    shim_multi = mod.BHTEMultiplePressureFields
    assert shim_multi.__module__ == 'babelbrain_b200.thermal'
    got = drive_multi(mod)
    ref = drive_multi(mod, oracle_bhte, oracle_bhte_multi)
    mod.BHTE, mod.BHTEMultiplePressureFields = shim_bhte, shim_multi
    for name, a, b in zip(names, got, ref):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if 'Dose' in name:
            assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(b), name
        else:
            assert np.abs(a - b).max() <= 2e-3, (name, np.abs(a - b).max())
    assert got[0].max() - 37.0 > 1.0
