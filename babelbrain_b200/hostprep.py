"""
Host-side (float64) preparation shared by every run: per-material coefficient tables, the stable
time step, the PML damping tables and the step/sample bookkeeping.  This is the part of
``PropagationModel.CalculateMatricesForPropagation`` / ``StaggeredFDTD_3D_with_relaxation`` that runs
before the time loop (callers: TranscranialModeling/BabelIntegrationBASE.py:1799,1801,2338).

The arithmetic of BabelViscoFDTD is not available in the reference tree (un-vendored pip package,
environment_linux.yml:44); the formulas below restate the published scheme and are cross-checked
against the independent restatement in oracle/fdtd_numpy.py by tests/test_hostprep.py.
"""
import numpy as np

MAP_NAMES = ['ALLV', 'Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Sigmaxy', 'Sigmaxz', 'Sigmayz', 'Pressure']


def maps_mask(names):
    m = 0
    for n in names:
        if n not in MAP_NAMES:
            raise ValueError('unknown map name %r (valid: %s)' % (n, MAP_NAMES))
        m |= 1 << MAP_NAMES.index(n)
    return m


def quality_factors(MaterialProperties, Frequency, QCorrection=1.0):
    """Low-loss Q = omega/(2 c alpha) per mode, times the per-material QCorrection
    (BabelIntegrationBASE.py:1249-1272 passes 1 for fluids, 3 for bone)."""
    MP = np.atleast_2d(np.asarray(MaterialProperties, dtype=np.float64))
    w = 2.0 * np.pi * float(Frequency)
    qc = np.broadcast_to(np.asarray(QCorrection, dtype=np.float64), (MP.shape[0],))
    with np.errstate(divide='ignore', invalid='ignore'):
        QL = np.where((MP[:, 3] > 0) & (MP[:, 1] > 0), w / (2.0 * MP[:, 1] * MP[:, 3]), 0.0) * qc
        QS = np.where((MP[:, 4] > 0) & (MP[:, 2] > 0), w / (2.0 * MP[:, 2] * MP[:, 4]), 0.0) * qc
    return QL, QS


def relaxation_fit(Frequency, QL, QS):
    """Single standard-linear-solid (tau-method): common tau_sigma per material from the L mode
    (S mode when only shear attenuates); per-mode tau so that Q(omega) is met exactly."""
    w = 2.0 * np.pi * float(Frequency)
    qref = np.where(QL > 0, QL, QS)
    on = qref > 0
    iq = np.where(on, 1.0 / np.where(on, qref, 1.0), 0.0)
    x = np.where(on, np.sqrt(1.0 + iq * iq) - iq, 1.0)   # omega * tau_sigma
    ots = np.where(on, w / x, 0.0)

    def tau_of(Q):
        has = on & (Q > 0)
        iqm = np.where(has, 1.0 / np.where(has, Q, 1.0), 0.0)
        y = (x + iqm) / (1.0 - x * iqm)                   # omega * tau_epsilon
        return np.where(has, y / x - 1.0, 0.0)
    return tau_of(QL), tau_of(QS), ots


def dispersion_factor(Frequency, tau, ots):
    """relaxed speed / phase speed at omega for the SLS (QfactorCorrection)."""
    w = 2.0 * np.pi * float(Frequency)
    on = ots > 0
    x = np.where(on, w / np.where(on, ots, 1.0), 0.0)
    y = x * (1.0 + tau)
    F = (1.0 + 1j * y) / (1.0 + 1j * x)
    return np.where(on, np.real(1.0 / np.sqrt(F)), 1.0)


def material_table(MaterialProperties, Frequency, QfactorCorrection, SpatialStep, QCorrection=1.0):
    """(nmat, 8) float64 rows [M, G, L, B, tauL, tauS, 1/tau_sigma, K], moduli divided by h."""
    MP = np.atleast_2d(np.asarray(MaterialProperties, dtype=np.float64))
    if MP.shape[1] != 5:
        raise ValueError('MaterialProperties must be (N,5): rho, cL, cS, attL, attS')
    h = float(SpatialStep)
    QL, QS = quality_factors(MP, Frequency, QCorrection)
    tauL, tauS, ots = relaxation_fit(Frequency, QL, QS)
    rho, cL, cS = MP[:, 0], MP[:, 1], MP[:, 2]
    cLr, cSr = cL, cS
    if QfactorCorrection:
        cLr = cL * dispersion_factor(Frequency, tauL, ots)
        cSr = cS * dispersion_factor(Frequency, tauS, ots)
    M = rho * cLr * cLr / h
    G = rho * cSr * cSr / h
    T = np.stack([M, G, M - 2.0 * G, 1.0 / (rho * h), tauL, tauS, ots, rho * cL * cL / h], axis=1)
    return T, dict(QL=QL, QS=QS, cLr=cLr, cSr=cSr)


# Courant number of the O(2,4) staggered scheme at its stability limit in 3-D, relative to h/(sqrt(3) c):
# 1/(9/8 + 1/24) = 6/7.
CFL_LIMIT_O24 = 6.0 / 7.0


def stable_dt(MaterialProperties, SpatialStep, AlphaCFL):
    """Stable time step: min(AlphaCFL, 6/7) * (sqrt(3)/3) * h / max c_L.

    The linear part is pinned by the caller: with AlphaCFL = 0.5 (BabelIntegrationBASE.py:934) it reproduces the awkward
    points-per-period the caller special-cases (47 for cortical bone at 500 kHz and 6 PPW, 71 at 9 PPW, :1811-1824).
    The limit at 6/7 -- the largest Courant fraction at which the fourth-order scheme is stable at all -- is pinned by
    the caller's second use of this function, the water-only step at AlphaCFL = 1.0 (:1801) that normalises its
    dispersion-correction polynomial (:1674, :2433-2440): only with (6/7) h/(sqrt(3) c) there does the corrected FDTD
    amplitude agree with the Rayleigh integral in water as the reference's own 309-case comparison records it
    (OfflineBatchExamples/CompareRayleightWithFDTD/SummaryAnalysis.xlsx: +0.1 ... +0.9 % at the peak; an uncapped
    step gives -14 %, tests/test_reference_caller.py).  [Recalled scheme corrected by in-tree evidence; DESIGN.md 2.]"""
    MP = np.atleast_2d(np.asarray(MaterialProperties, dtype=np.float64))
    return min(float(AlphaCFL), CFL_LIMIT_O24) * np.sqrt(3.0) / 3.0 * float(SpatialStep) / MP[:, 1].max()


def hard_limit_dt(MaterialProperties, SpatialStep):
    """The step above which the scheme is certainly unstable (the O(2,4) limit for the fastest material)."""
    return stable_dt(MaterialProperties, SpatialStep, CFL_LIMIT_O24)


# Damping of the split parts of the absorbing layer.  0 = the classical split-field layer (every part damped along its
# own axis only), which is what the restated scheme uses and the default.  BabelBrain always hands the solver a water
# shell (UpdateConditions writes the tissue mask only inside the PML offsets, BabelIntegrationBASE.py:1853-1862,
# :2154-2159), and there the classical layer is stable (9600 steps, profiles/r2_pml_stability.txt).  A label map with
# fluid-solid interfaces *inside* the shell makes it grow without bound (profiles/r1_pml_stability.txt); for such maps
# the public call takes MPMLRatio > 0 (0.05-0.1): the multi-axial layer of Meza-Fajardo & Papageorgiou (BSSA 2008),
# every part also damped by that fraction of the damping of the two other axes.
MPML_RATIO = 0.0


def pml_table(NDelta, SpatialStep, dt, Vmax, ReflectionLimit):
    """(2, NDelta+1) float64: the damping d at integer depth xi = 0..NDelta and at half depth xi + 0.5
    (quadratic profile, d0 = ln(1/R) 3 Vmax / (2 NDelta h)).  A part advances as
    f <- (f (1/dt - d/2) + C D) / (1/dt + d/2)."""
    P = int(NDelta)
    d0 = np.log(1.0 / float(ReflectionLimit)) * 3.0 * float(Vmax) / (2.0 * P * float(SpatialStep))
    xi = np.arange(P + 1, dtype=np.float64)
    return np.stack([d0 * (xi / P) ** 2, d0 * ((xi + 0.5) / P) ** 2])


def number_of_steps(DurationSimulation, dt):
    r = float(DurationSimulation) / float(dt)
    if abs(r - round(r)) < 1e-6 * max(1.0, r):
        return int(round(r))
    return int(np.ceil(r))


def sample_steps(steps, SensorSubSampling, SensorStart):
    n = np.arange(0, steps, int(SensorSubSampling))
    return n[n // int(SensorSubSampling) >= int(SensorStart)]
