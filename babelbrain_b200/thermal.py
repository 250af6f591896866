"""
Host-side mirror of ``BabelViscoFDTD.tools.RayleighAndBHTE.BHTE`` / ``BHTEMultiplePressureFields`` -- the bio-heat
(Pennes) solver of BabelBrain's thermal step (ThermalModeling/CalculateTemperatureEffects.py:14, :365-395, :406, :439,
:960-990).  Same arguments, same return tuples; the time loop runs in libbabelb200.so
(csrc/bhte.cu) through ``bb_bhte_run``.  No CPU fallback.

The arithmetic lives in the BabelViscoFDTD package, which is absent from the reference tree (PARITY UNPINNED, like the
FDTD solver): the coefficient formulas and the CEM43 dose rule below restate the published scheme; interface facts
(argument order, MaterialList keys, dose in seconds, MonitoringPointsMap ids, slice layout) are anchored on the caller.
"""
import ctypes
import numpy as np

from . import _capi


def getBHTECoefficient(kappa, rho, c_t, h, t_int, dt=0.1):
    """kappa dt / (rho c_t h^2); the explicit scheme is stable below 1/6."""
    coeff = kappa * dt / (rho * c_t * h ** 2)
    if coeff >= 1.0 / 6.0:
        best_nt = np.ceil(6 * kappa * t_int / (rho * c_t * h ** 2))
        raise ValueError('The time step %g s is too large for the explicit bio-heat scheme (coefficient %g >= 1/6); '
                         'use at least %d steps for %g s' % (dt, coeff, best_nt, t_int))
    return coeff


def getPerfusionCoefficient(w_b, c_t, blood_rho, blood_ct, dt=0.1):
    """w_b in ml/min/kg -> per-step relaxation towards the core temperature."""
    return w_b / 60.0 * 1.0e-6 * blood_rho * blood_ct * dt / c_t


def getQCoeff(rho, SoS, alpha, c_t, Absorption, h, dt):
    """Temperature rise per step per Pa^2 of pressure amplitude."""
    return dt / (2.0 * rho ** 2 * SoS * h * c_t) * Absorption * (1.0 - np.exp(-2.0 * h * alpha))


def _tables(MaterialMap, MaterialList, dx, TotalDurationSteps, dt, blood_rho, blood_ct):
    nmat = int(MaterialMap.max()) + 1
    for k in ('Density', 'SoS', 'Attenuation', 'SpecificHeat', 'Conductivity', 'Perfusion', 'Absorption', 'InitTemperature'):
        if len(MaterialList[k]) < nmat:
            raise ValueError('MaterialList[%r] has %d entries but MaterialMap holds label %d' % (k, len(MaterialList[k]), nmat - 1))
    bh = np.zeros(nmat, np.float32)
    perf = np.zeros(nmat, np.float32)
    for n in range(nmat):
        bh[n] = getBHTECoefficient(MaterialList['Conductivity'][n], MaterialList['Density'][n], MaterialList['SpecificHeat'][n],
                                   dx, TotalDurationSteps * dt, dt=dt)
        perf[n] = getPerfusionCoefficient(MaterialList['Perfusion'][n], MaterialList['SpecificHeat'][n], blood_rho, blood_ct, dt=dt)
    return nmat, bh, perf


def heat_source(Pressure, MaterialMap, MaterialList, dx, dt, DutyCycle=1.0):
    """Qarr: temperature added per step while the beam is on (float32 volume)."""
    nmat = int(MaterialMap.max()) + 1
    qc = np.array([getQCoeff(MaterialList['Density'][n], MaterialList['SoS'][n], MaterialList['Attenuation'][n],
                             MaterialList['SpecificHeat'][n], MaterialList['Absorption'][n], dx, dt) for n in range(nmat)])
    P = np.asarray(Pressure, dtype=np.float64)
    return np.ascontiguousarray(P * P * qc[MaterialMap] * DutyCycle, dtype=np.float32)


def _run(Q, MaterialMap, MaterialList, dx, TotalDurationSteps, schedule, LocationMonitoring, nFactorMonitoring, dt, blood_rho,
         blood_ct, stableTemp, MonitoringPointsMap, initT0, initDose):
    _capi.require_gpu()
    from .rayleigh import _state
    MaterialMap = np.ascontiguousarray(MaterialMap, dtype=np.uint32)
    if MaterialMap.ndim != 3:
        raise ValueError('MaterialMap must be a 3-D volume')
    N1, N2, N3 = MaterialMap.shape
    TotalDurationSteps = int(TotalDurationSteps)
    nmat, bh, perf = _tables(MaterialMap, MaterialList, dx, TotalDurationSteps, dt, blood_rho, blood_ct)
    if initT0 is None:
        T = np.asarray(MaterialList['InitTemperature'], dtype=np.float32)[MaterialMap]
    else:
        T = np.array(initT0, dtype=np.float32)
    D = np.zeros(MaterialMap.shape, np.float32) if initDose is None else np.array(initDose, dtype=np.float32)
    if T.shape != MaterialMap.shape or D.shape != MaterialMap.shape:
        raise ValueError('initT0 / initDose must have the shape of MaterialMap')
    T, D = np.ascontiguousarray(T), np.ascontiguousarray(D)
    nmon = max(int(nFactorMonitoring), 1)
    # the slice is returned with its full shape even when monitoring is disabled (LocationMonitoring < 0), as zeros
    MonitorSlice = np.zeros((N1, N3, TotalDurationSteps // nmon), np.float32)
    want_slice = 0 <= int(LocationMonitoring) < N2 and MonitorSlice.size > 0
    points = tpoints = None
    npoints = 0
    if MonitoringPointsMap is not None:
        points = np.ascontiguousarray(MonitoringPointsMap, dtype=np.uint32)
        if points.shape != MaterialMap.shape:
            raise ValueError('MonitoringPointsMap must have the shape of MaterialMap')
        npoints = int((points > 0).sum())
        if npoints and int(points.max()) > npoints:
            raise ValueError('MonitoringPointsMap ids must be 1..number of points')
        tpoints = np.zeros((max(npoints, 1), TotalDurationSteps), np.float32)
        if npoints == 0:
            points = None
    sched = np.ascontiguousarray(schedule, dtype=np.int16)
    ms = ctypes.c_double(0.0)
    if TotalDurationSteps > 0:
        _capi.check(_capi.lib().bb_bhte_run(N1, N2, N3, nmat, Q.shape[0], _capi.ptr(Q), _capi.ptr(MaterialMap), _capi.ptr(bh), _capi.ptr(perf),
                                            _capi.ptr(T), _capi.ptr(D), _capi.ptr(sched), TotalDurationSteps, float(dt), float(stableTemp),
                                            int(LocationMonitoring) if want_slice else -1, nmon, _capi.ptr(MonitorSlice) if want_slice else None,
                                            _capi.ptr(points), npoints, _capi.ptr(tpoints) if points is not None else None,
                                            _state['device'], ctypes.byref(ms)))
    _state['last_bhte_ms'] = ms.value
    if MonitoringPointsMap is not None:
        return T, D, MonitorSlice, (tpoints[:npoints] if npoints else np.zeros((0, TotalDurationSteps), np.float32))
    return T, D, MonitorSlice, None


def BHTE(Pressure, MaterialMap, MaterialList, dx, TotalDurationSteps, nStepsOn, LocationMonitoring, nFactorMonitoring=1, dt=0.1,
         blood_rho=1050, blood_ct=3617, stableTemp=37.0, DutyCycle=1.0, Backend='CUDA', MonitoringPointsMap=None, initT0=None,
         initDose=None):
    """One pressure field: the beam heats during the first nStepsOn of TotalDurationSteps steps.
    Returns (ResTemp, ResDose, MonitorSlice, Qarr) and, when MonitoringPointsMap is given, TemperaturePoints as a fifth
    value.  ResDose is the CEM43 dose in seconds (the caller divides by 60, CalculateTemperatureEffects.py:1134)."""
    MaterialMap = np.asarray(MaterialMap)
    Qarr = heat_source(Pressure, MaterialMap.astype(np.int64), MaterialList, dx, dt, DutyCycle)
    steps = int(TotalDurationSteps)
    schedule = np.where(np.arange(steps) < int(nStepsOn), 0, -1)
    T, D, Slice, Pts = _run(Qarr[None], MaterialMap, MaterialList, dx, steps, schedule, LocationMonitoring, nFactorMonitoring, dt,
                            blood_rho, blood_ct, stableTemp, MonitoringPointsMap, initT0, initDose)
    if MonitoringPointsMap is not None:
        return T, D, Slice, Qarr, Pts
    return T, D, Slice, Qarr


def field_schedule(nStepsOnOffList, TotalDurationSteps):
    """Which pressure field heats at each step: the fields take turns, field m on for nStepsOnOffList[m,0] steps and off
    for nStepsOnOffList[m,1], the cycle repeating until TotalDurationSteps (CalculateTemperatureEffects.py:733-736)."""
    onoff = np.asarray(nStepsOnOffList, dtype=np.int64).reshape(-1, 2)
    cycle = []
    for m, (on, off) in enumerate(onoff):
        cycle += [m] * int(on) + [-1] * int(off)
    if not cycle:
        raise ValueError('nStepsOnOffList describes an empty cycle')
    reps = -(-int(TotalDurationSteps) // len(cycle))
    return np.array((cycle * max(reps, 1))[:int(TotalDurationSteps)], dtype=np.int16)


def BHTEMultiplePressureFields(PressureFields, MaterialMap, MaterialList, dx, TotalDurationSteps, nStepsOnOffList, LocationMonitoring,
                               nFactorMonitoring=1, dt=0.1, blood_rho=1050, blood_ct=3617, stableTemp=37.0, Backend='CUDA',
                               MonitoringPointsMap=None, initT0=None, initDose=None):
    """Several pressure fields (electronic steering / multi-focus) taking turns.  PressureFields is (N, N1, N2, N3).
    Returns (ResTemp, ResDose, MonitorSlice, QArrList[, TemperaturePoints])."""
    P = np.asarray(PressureFields)
    MaterialMap = np.asarray(MaterialMap)
    if P.ndim != 4 or P.shape[1:] != MaterialMap.shape:
        raise ValueError('PressureFields must be (N,) + MaterialMap.shape')
    if np.asarray(nStepsOnOffList).reshape(-1, 2).shape[0] != P.shape[0]:
        raise ValueError('nStepsOnOffList needs one (on, off) pair per pressure field')
    Q = np.stack([heat_source(P[m], MaterialMap.astype(np.int64), MaterialList, dx, dt, 1.0) for m in range(P.shape[0])])
    schedule = field_schedule(nStepsOnOffList, TotalDurationSteps)
    T, D, Slice, Pts = _run(np.ascontiguousarray(Q), MaterialMap, MaterialList, dx, int(TotalDurationSteps), schedule, LocationMonitoring,
                            nFactorMonitoring, dt, blood_rho, blood_ct, stableTemp, MonitoringPointsMap, initT0, initDose)
    if MonitoringPointsMap is not None:
        return T, D, Slice, Q, Pts
    return T, D, Slice, Q
