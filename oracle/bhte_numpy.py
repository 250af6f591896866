"""ORACLE (test infrastructure only) -- NumPy restatement of the bio-heat solver behind
BabelViscoFDTD.tools.RayleighAndBHTE.BHTE / BHTEMultiplePressureFields as BabelBrain's thermal step calls it
(ThermalModeling/CalculateTemperatureEffects.py:365-395, :406, :439, :960-990).

PARITY UNPINNED: the package holding the reference arithmetic is absent from /root/reference and from this image; the
reference tree holds no test vector for the thermal solver.  The scheme below is the published one -- explicit Pennes
equation on the 7-point stencil with per-material conduction / perfusion coefficients, heat source from the pressure
amplitude, CEM43 dose integrated exactly along each step's linear temperature ramp -- with the interface facts (argument
order, MaterialList keys, dose in seconds: the caller divides by 60 at :1134, monitoring ids at :1003-1022) taken from
the caller.  Only tests/ import this module.
"""
import numpy as np


def coefficients(MaterialList, nmat, dx, dt, blood_rho=1050.0, blood_ct=3617.0):
    k, rho, ct = (np.asarray(MaterialList[x], float)[:nmat] for x in ('Conductivity', 'Density', 'SpecificHeat'))
    bh = k * dt / (rho * ct * dx ** 2)
    perf = np.asarray(MaterialList['Perfusion'], float)[:nmat] / 60.0 * 1e-6 * blood_rho * blood_ct * dt / ct
    sos, att, ab = (np.asarray(MaterialList[x], float)[:nmat] for x in ('SoS', 'Attenuation', 'Absorption'))
    q = dt / (2.0 * rho ** 2 * sos * dx * ct) * ab * (1.0 - np.exp(-2.0 * dx * att))
    return bh, perf, q


def cem43_increment(t0, t1, dt):
    """Integral of R^(43 - T) over one step with T linear from t0 to t1; R = 0.5 above 43 C and 0.25 below, the ramp
    split where it crosses 43 C; steps that change the temperature by < 1e-4 use the end value."""
    t0, t1 = np.asarray(t0, float), np.asarray(t1, float)
    r1 = np.where(t0 >= 43.0, 0.5, 0.25)
    r2 = np.where(t1 >= 43.0, 0.5, 0.25)
    flat = np.abs(t1 - t0) < 1e-4
    d = np.where(flat, 1.0, t1 - t0)
    out = dt * r2 ** (43.0 - t1)
    same = (r2 ** (43.0 - t1) - r1 ** (43.0 - t0)) / (-d / dt * np.log(r1))
    with np.errstate(divide='ignore', invalid='ignore'):
        dtp = dt * (43.0 - t0) / d
        cross = (1.0 - r1 ** (43.0 - t0)) / (-(43.0 - t0) / dtp * np.log(r1)) + (r2 ** (43.0 - t1) - 1.0) / ((43.0 - t1) / (dt - dtp) * np.log(r2))
    return np.where(flat, out, np.where(r1 == r2, same, cross))


def run(Q, MaterialMap, MaterialList, dx, TotalDurationSteps, schedule, dt=0.1, stableTemp=37.0, initT0=None, initDose=None,
        LocationMonitoring=-1, nFactorMonitoring=1, MonitoringPointsMap=None, dtype=np.float64):
    """Q: (nfields, N1, N2, N3) temperature added per step by each field; schedule[n] = field heating at step n or -1."""
    MM = np.asarray(MaterialMap).astype(np.int64)
    nmat = int(MM.max()) + 1
    bh, perf, _ = coefficients(MaterialList, nmat, dx, dt)
    bhv, pfv = bh[MM].astype(dtype), perf[MM].astype(dtype)
    T = (np.asarray(MaterialList['InitTemperature'], float)[MM] if initT0 is None else np.asarray(initT0)).astype(dtype)
    D = (np.zeros(MM.shape) if initDose is None else np.asarray(initDose)).astype(dtype)
    steps = int(TotalDurationSteps)
    N1, N2, N3 = MM.shape
    Slice = np.zeros((N1, N3, steps // nFactorMonitoring), np.float32)
    pts = None
    if MonitoringPointsMap is not None:
        ids = np.asarray(MonitoringPointsMap)
        where = [tuple(np.argwhere(ids == n + 1)[0]) for n in range(int((ids > 0).sum()))]
        pts = np.zeros((len(where), steps), np.float32)
    inner = (slice(1, -1),) * 3
    for n in range(steps):
        c = T[inner]
        lap = (T[2:, 1:-1, 1:-1] + T[:-2, 1:-1, 1:-1] + T[1:-1, 2:, 1:-1] + T[1:-1, :-2, 1:-1] + T[1:-1, 1:-1, 2:] + T[1:-1, 1:-1, :-2] - 6.0 * c)
        new = c + bhv[inner] * lap + pfv[inner] * (stableTemp - c)
        if schedule[n] >= 0:
            new = new + np.asarray(Q[schedule[n]], dtype)[inner]
        D[inner] += cem43_increment(c, new, dt).astype(dtype)
        T = T.copy()
        T[inner] = new
        if 0 <= LocationMonitoring < N2 and n % nFactorMonitoring == 0 and n // nFactorMonitoring < Slice.shape[2]:
            if 0 < LocationMonitoring < N2 - 1:
                Slice[1:-1, 1:-1, n // nFactorMonitoring] = T[1:-1, LocationMonitoring, 1:-1]
        if pts is not None:
            for m, w in enumerate(where):
                if all(0 < w[a] < MM.shape[a] - 1 for a in range(3)):
                    pts[m, n] = T[w]
    return T, D, Slice, pts
