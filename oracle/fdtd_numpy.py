"""
ORACLE (test infrastructure, NOT product code) -- float64 NumPy restatement of the
3-D staggered-grid isotropic viscoelastic FDTD solver BabelBrain calls as
``PModel.StaggeredFDTD_3D_with_relaxation`` (call sites
TranscranialModeling/BabelIntegrationBASE.py:2338-2365, 2374-2398, 2401-2428) and of
``PModel.CalculateMatricesForPropagation`` (BabelIntegrationBASE.py:1799,1801).

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
BabelViscoFDTD (pinned ==1.2.4 in environment_linux.yml:44, ==1.2.6 in
environment_win-314.yml:55) which is NOT vendored in /root/reference, is not installed in this
image and cannot be fetched (no network).  The reference tree holds no test, fixture or golden
vector for the solver (Tests/ is git-ignored, .gitignore:140).  This file therefore restates the
*published* scheme (Virieux 1986 staggered grid; Blanch/Robertsson/Symes 1995 tau-method with one
standard-linear-solid; Collino & Tsogka 2001 split-field PML; Pichardo et al. PMB 2017, cited at
BabelIntegrationBASE.py:72) and anchors on the reference's own call sites for every interface
fact (dtypes, shapes, index conventions, sampling windows).  Every scheme item that could differ
from upstream is an isolated, named function so it can be corrected in one place.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (babelbrain_b200) never does.

Conventions (firm, from the caller):
  * arrays are (N1,N2,N3) numpy C-order; linear sensor index is 1-based Fortran order
    i + j*N1 + k*N1*N2 + 1 (decoded at BabelIntegrationBASE.py:2503-2511)
  * MaterialList rows [rho, cL, cS, alphaL(Np/m), alphaS(Np/m)] (BabelIntegrationBASE.py:143,2261)
  * SourceMap 0 = none, else 1-based row of SourceFunctions (BabelIntegrationSingle.py:326-344)
  * SourceFunctions (Nsrc, Nt_src) (BabelIntegrationSingle.py:335)
  * RMS window: steps n >= SensorStart*SensorSubSampling (BabelIntegrationBASE.py:2108-2109)
"""
import numpy as np

CA = 9.0 / 8.0
CB = 1.0 / 24.0

MAP_BITS = {'ALLV': 0x1, 'Vx': 0x2, 'Vy': 0x4, 'Vz': 0x8, 'Sigmaxx': 0x10, 'Sigmayy': 0x20,
            'Sigmazz': 0x40, 'Sigmaxy': 0x80, 'Sigmaxz': 0x100, 'Sigmayz': 0x200, 'Pressure': 0x400}
MAP_ORDER = ['ALLV', 'Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Sigmaxy', 'Sigmaxz',
             'Sigmayz', 'Pressure']


# --------------------------------------------------------------------------------------------
# (a)/(b) per-material tables and relaxation fit
# --------------------------------------------------------------------------------------------
def q_from_attenuation(omega, c, alpha):
    """Low-loss quality factor Q = omega / (2 c alpha); 0 where no attenuation (alpha Np/m)."""
    c = np.asarray(c, float)
    alpha = np.asarray(alpha, float)
    Q = np.zeros_like(c)
    ok = (alpha > 0) & (c > 0)
    Q[ok] = omega / (2.0 * c[ok] * alpha[ok])
    return Q


def sls_fit(omega, QL, QS):
    """One standard-linear-solid per material, common tau_sigma for both modes (tau-method).
    x = omega*tau_sigma = sqrt(1+1/Q^2) - 1/Q taken from the L mode (from S if only S attenuates);
    y = omega*tau_eps   = (x + 1/Q)/(1 - x/Q) for each mode so that Q(omega) is met exactly;
    tau = y/x - 1.  Returns tauL, tauS, one_over_tau_sigma (0 where lossless)."""
    n = QL.shape[0]
    tauL = np.zeros(n)
    tauS = np.zeros(n)
    ots = np.zeros(n)
    for m in range(n):
        qref = QL[m] if QL[m] > 0 else QS[m]
        if qref <= 0:
            continue
        x = np.sqrt(1.0 + 1.0 / qref ** 2) - 1.0 / qref
        ots[m] = omega / x
        for Q, out in ((QL[m], tauL), (QS[m], tauS)):
            if Q > 0:
                y = (x + 1.0 / Q) / (1.0 - x / Q)
                out[m] = y / x - 1.0
    return tauL, tauS, ots


def phase_velocity_factor(omega, tau, ots):
    """c_relaxed / c_phase(omega) for the SLS: Re(1/sqrt((1+j y)/(1+j x))), x=omega*tau_sigma,
    y = x(1+tau).  Used by QfactorCorrection so that the phase speed at omega is the nominal c."""
    f = np.ones_like(tau)
    on = ots > 0
    x = np.zeros_like(tau)
    x[on] = omega / ots[on]
    y = x * (1.0 + tau)
    F = (1.0 + 1j * y[on]) / (1.0 + 1j * x[on])
    f[on] = np.real(1.0 / np.sqrt(F))
    return f


def material_tables(MaterialProperties, Frequency, QfactorCorrection, h, QCorrection=1.0):
    """Per-material coefficient tables (float64).  Keys:
    M = rho cL^2/h (relaxed), G = rho cS^2/h, L = M-2G, B = 1/(rho h), tauL, tauS, ots,
    K = rho cL_nominal^2 / h (pressure scaling)."""
    MP = np.asarray(MaterialProperties, float)
    rho, cL, cS, aL, aS = (MP[:, n].copy() for n in range(5))
    omega = 2.0 * np.pi * Frequency
    QC = np.ones(MP.shape[0]) * np.asarray(QCorrection, float)
    QL = q_from_attenuation(omega, cL, aL) * QC
    QS = q_from_attenuation(omega, cS, aS) * QC
    tauL, tauS, ots = sls_fit(omega, QL, QS)
    cLr = cL.copy()
    cSr = cS.copy()
    if QfactorCorrection:
        cLr = cL * phase_velocity_factor(omega, tauL, ots)
        cSr = cS * phase_velocity_factor(omega, tauS, ots)
    M = rho * cLr ** 2 / h
    G = rho * cSr ** 2 / h
    return dict(M=M, G=G, L=M - 2.0 * G, B=1.0 / (rho * h), tauL=tauL, tauS=tauS, ots=ots,
                K=rho * cL ** 2 / h, QL=QL, QS=QS, cLr=cLr, cSr=cSr)


def ideal_dt(MaterialProperties, h, AlphaCFL):
    """(c) dt_ideal = min(AlphaCFL, 6/7) * (sqrt(3)/3) * h / max cL.  The linear part reproduces the caller's special-cased
    points-per-period (BabelIntegrationBASE.py:1811-1824); the cap at the O(2,4) stability limit 6/7 = 1/(9/8 + 1/24) is
    what makes the caller's water-only normalisation step (AlphaCFL = 1.0, :1801) and its dispersion-correction
    polynomial (:1674) reproduce the FDTD-vs-Rayleigh agreement recorded in SummaryAnalysis.xlsx
    (tests/test_reference_caller.py)."""
    return min(float(AlphaCFL), 6.0 / 7.0) * np.sqrt(3.0) / 3.0 * h / np.max(np.asarray(MaterialProperties, float)[:, 1])


def calculate_matrices_for_propagation(MaterialMap, MaterialProperties, Frequency,
                                       QfactorCorrection, h, AlphaCFL, QCorrection=1.0):
    """10-tuple, element 0 = dt (the only element the caller uses, BabelIntegrationBASE.py:1799)."""
    T = material_tables(MaterialProperties, Frequency, QfactorCorrection, h, QCorrection)
    dt = ideal_dt(MaterialProperties, h, AlphaCFL)
    rho = np.asarray(MaterialProperties, float)[:, 0]
    return (dt, rho, T['G'] * h, T['M'] * h, T['L'] * h, T['tauL'], T['tauS'],
            np.where(T['ots'] > 0, 1.0 / np.where(T['ots'] > 0, T['ots'], 1.0), 0.0), T['QL'], T['QS'])


# --------------------------------------------------------------------------------------------
# (g) PML profiles
# --------------------------------------------------------------------------------------------
def pml_tables(P, h, dt, Vmax, ReflectionLimit):
    """Quadratic damping d(xi)=d0 (xi/P)^2, d0 = ln(1/R) 3 Vmax/(2 P h), sampled at integer depth
    xi=0..P and half depth xi+0.5.  Returns InvDXDT, DXDT, InvDXDThp, DXDThp (each P+1):
    f <- InvDXDT * (f*DXDT + C*Diff)  with InvDXDT = 1/(1/dt + d/2), DXDT = 1/dt - d/2."""
    d0 = np.log(1.0 / ReflectionLimit) * 3.0 * Vmax / (2.0 * P * h)
    xi = np.arange(P + 1, dtype=float)
    d = d0 * (xi / P) ** 2
    dhp = d0 * ((xi + 0.5) / P) ** 2
    return 1.0 / (1.0 / dt + d / 2), (1.0 / dt - d / 2), 1.0 / (1.0 / dt + dhp / 2), (1.0 / dt - dhp / 2)


# (g') multi-axial damping (optional, MPMLRatio > 0).  The classical split-field layer above (each part damped along
# its own axis only) is the default: BabelBrain's label maps are water inside the shell (BabelIntegrationBASE.py:2110,
# :2154-2159) and there it is stable.  Where a fluid-solid interface runs INTO the layer it grows without bound
# (e-folding ~130 steps; same map with the solid kept out of the layer, or damping off: bounded).
# Meza-Fajardo & Papageorgiou (BSSA 2008) cure that by adding a fraction of each axis' damping to the parts of the
# other two axes:
#     d_eff(part of axis a) = d_a(own staggering) + ratio * (d_b + d_c)     (d_b, d_c at integer nodes)
MPML_RATIO = 0.0


def pml_damping(P, h, Vmax, ReflectionLimit):
    """The damping values behind pml_tables: d at integer depth 0..P and at half depth xi+0.5."""
    d0 = np.log(1.0 / ReflectionLimit) * 3.0 * Vmax / (2.0 * P * h)
    xi = np.arange(P + 1, dtype=float)
    return d0 * (xi / P) ** 2, d0 * ((xi + 0.5) / P) ** 2


def pml_depth(N, P):
    """Depth tables for one axis.  Integer nodes n: depth P-n on the low side (n<P), n-(N-P-1) on
    the high side (n>=N-P), 0 inside.  Half nodes n+1/2: index into the half-point table:
    low side P-1-n (depth P-n-0.5), high side n-(N-P-1) (depth +0.5); inside -> -1 (no damping)."""
    n = np.arange(N)
    di = np.zeros(N, int)
    dh = -np.ones(N, int)
    lo = n < P
    hi = n >= N - P
    di[lo] = P - n[lo]
    di[hi] = n[hi] - (N - P - 1)
    dh[lo] = P - 1 - n[lo]
    dh[hi] = n[hi] - (N - P - 1)
    return di, dh, (lo | hi)


# --------------------------------------------------------------------------------------------
# differences with the edge rules
# --------------------------------------------------------------------------------------------
def _shift(f, axis, s):
    """g[n] = f[n+s] along axis, zero outside."""
    g = np.zeros_like(f)
    N = f.shape[axis]
    src = [slice(None)] * 3
    dst = [slice(None)] * 3
    if s >= 0:
        src[axis] = slice(s, N)
        dst[axis] = slice(0, N - s)
    else:
        src[axis] = slice(0, N + s)
        dst[axis] = slice(-s, N)
    g[tuple(dst)] = f[tuple(src)]
    return g


def _axis_index(shape, axis):
    sh = [1, 1, 1]
    sh[axis] = shape[axis]
    return np.arange(shape[axis]).reshape(sh)


def dbwd(f, axis):
    """Backward staggered difference landing on n: 4th order for 1<n<N-1, 2-point at n=1 and
    n=N-1, 0 at n=0."""
    N = f.shape[axis]
    n = _axis_index(f.shape, axis)
    d2 = f - _shift(f, axis, -1)
    d4 = CA * d2 - CB * (_shift(f, axis, 1) - _shift(f, axis, -2))
    return np.where((n > 1) & (n < N - 1), d4, np.where(n > 0, d2, 0.0))


def dfwd(f, axis):
    """Forward staggered difference landing on n+1/2: 4th order for 0<n<N-2, 2-point at n=0 and
    n=N-2, 0 at n=N-1."""
    N = f.shape[axis]
    n = _axis_index(f.shape, axis)
    d2 = _shift(f, axis, 1) - f
    d4 = CA * d2 - CB * (_shift(f, axis, 2) - _shift(f, axis, -1))
    return np.where((n > 0) & (n < N - 2), d4, np.where(n < N - 1, d2, 0.0))


def _nb(a, axis):
    """a at the +1 neighbour along axis (edge-clamped; the clamped cells are never updated)."""
    idx = np.minimum(np.arange(a.shape[axis]) + 1, a.shape[axis] - 1)
    return np.take(a, idx, axis=axis)


def harmonic4(g1, g2, g3, g4):
    """(d) edge rigidity: 4/(1/g1+..+1/g4), 0 if any gi is 0."""
    prod = g1 * g2 * g3 * g4
    ok = prod != 0
    out = np.zeros_like(g1)
    out[ok] = 4.0 / (1.0 / g1[ok] + 1.0 / g2[ok] + 1.0 / g3[ok] + 1.0 / g4[ok])
    return out


# --------------------------------------------------------------------------------------------
# the solver
# --------------------------------------------------------------------------------------------
def number_of_steps(DurationSimulation, dt):
    """Steps n = 0..steps-1 with n*dt < DurationSimulation (fuzz-tolerant to the caller's
    TimeSimulation = dt*ntSteps, BabelIntegrationBASE.py:2089)."""
    r = DurationSimulation / dt
    if abs(r - round(r)) < 1e-6 * max(1.0, r):
        return int(round(r))
    return int(np.ceil(r))


def run(MaterialMap, MaterialProperties, Frequency, SourceMap, SourceFunctions, SpatialStep,
        DurationSimulation, SensorMap, Ox=1.0, Oy=1.0, Oz=1.0, NDelta=12, DT=None,
        ReflectionLimit=1e-5, AlphaCFL=1.0, TypeSource=0, QfactorCorrection=True, QCorrection=1.0,
        SelRMSorPeak=1, SelMapsRMSPeakList=('Pressure',), SelMapsSensorsList=('Pressure',),
        SensorSubSampling=2, SensorStart=0, ReflectorMask=None, dtype=np.float64, MPMLRatio=None):
    """Whole simulation; returns dict(Sensor, LastMap, RMS, Peak, IndexSensorMap, steps, dt)."""
    MaterialMap = np.asarray(MaterialMap)
    N1, N2, N3 = MaterialMap.shape
    P = int(NDelta)
    h = float(SpatialStep)
    MP = np.asarray(MaterialProperties, float)
    T = material_tables(MP, Frequency, QfactorCorrection, h, QCorrection)
    dt_id = ideal_dt(MP, h, AlphaCFL)
    dt = dt_id if DT is None else float(DT)
    if dt > dt_id * (1 + 1e-9):
        raise ValueError('DT larger than the stable step')
    steps = number_of_steps(DurationSimulation, dt)
    dmp, dmphp = (a.astype(dtype) for a in pml_damping(P, h, MP[:, 1].max(), ReflectionLimit))
    dt = dtype(dt)
    ratio = dtype(MPML_RATIO if MPMLRatio is None else MPMLRatio)

    mm = MaterialMap.astype(np.int64)
    tab = {k: T[k].astype(dtype) for k in ('M', 'G', 'L', 'B', 'tauL', 'tauS', 'ots', 'K')}
    M, G, L, B, tauL, tauS, ots = (tab[k][mm] for k in ('M', 'G', 'L', 'B', 'tauL', 'tauS', 'ots'))

    # --- staggered material averages (d),(f)
    def edge(ax_a, ax_b):
        g1, g2, g3, g4 = G, _nb(G, ax_a), _nb(G, ax_b), _nb(_nb(G, ax_a), ax_b)
        rig = harmonic4(g1, g2, g3, g4)
        t = 0.25 * (tauS + _nb(tauS, ax_a) + _nb(tauS, ax_b) + _nb(_nb(tauS, ax_a), ax_b))
        return rig, np.where(rig != 0, t, tauS)
    Rig = {}
    TauE = {}
    for name, (a, b) in (('xy', (0, 1)), ('xz', (0, 2)), ('yz', (1, 2))):
        Rig[name], TauE[name] = edge(a, b)
    Bx = 0.5 * (B + _nb(B, 0))
    By = 0.5 * (B + _nb(B, 1))
    Bz = 0.5 * (B + _nb(B, 2))

    # --- region masks
    dep = [pml_depth(N, P) for N in (N1, N2, N3)]
    inI = dep[0][2].reshape(N1, 1, 1)
    inJ = dep[1][2].reshape(1, N2, 1)
    inK = dep[2][2].reshape(1, 1, N3)
    pml = inI | inJ | inK
    interior = ~pml
    ii, jj, kk = np.ogrid[:N1, :N2, :N3]
    upd = pml & (ii < N1 - 1) & (jj < N2 - 1) & (kk < N3 - 1)

    def damp(axis, half):
        di, dh, _ = dep[axis]
        sh = [1, 1, 1]
        sh[axis] = -1
        v = np.where(dh >= 0, dmphp[np.maximum(dh, 0)], dtype(0)) if half else dmp[di]
        return v.astype(dtype).reshape(sh)

    def coef(axis, half):
        """(InvDXDT, DXDT) of a split part of `axis` at every cell: own damping at the part's staggering plus
        MPML_RATIO times the integer-node damping of the two other axes."""
        o1, o2 = [a for a in (0, 1, 2) if a != axis]
        d = damp(axis, half) + ratio * (damp(o1, False) + damp(o2, False))
        return 1 / (1 / dt + d / 2), 1 / dt - d / 2
    cI, cJ, cK = coef(0, False), coef(1, False), coef(2, False)
    hI, hJ, hK = coef(0, True), coef(1, True), coef(2, True)

    z = lambda: np.zeros((N1, N2, N3), dtype)
    Vx, Vy, Vz = z(), z(), z()
    Sxx, Syy, Szz, Sxy, Sxz, Syz = z(), z(), z(), z(), z(), z()
    Rxx, Ryy, Rzz, Rxy, Rxz, Ryz = z(), z(), z(), z(), z(), z()
    Pr = z()
    sp = {n: z() for n in ('Vx_x', 'Vx_y', 'Vx_z', 'Vy_x', 'Vy_y', 'Vy_z', 'Vz_x', 'Vz_y', 'Vz_z',
                           'Sxx_x', 'Sxx_y', 'Sxx_z', 'Syy_x', 'Syy_y', 'Syy_z', 'Szz_x', 'Szz_y', 'Szz_z',
                           'Sxy_x', 'Sxy_y', 'Sxz_x', 'Sxz_z', 'Syz_y', 'Syz_z')}

    def split(name, c, C, D):
        a, b = c
        sp[name] = np.where(upd, a * (sp[name] * b + C * D), sp[name])
        return sp[name]

    # sources
    SourceMap = np.asarray(SourceMap)
    src_idx = np.nonzero(SourceMap)
    src_id = SourceMap[src_idx].astype(np.int64) - 1
    SF = np.asarray(SourceFunctions, float)
    nt_src = SF.shape[1]

    def bro(O):
        O = np.asarray(O, float)
        return (np.ones((N1, N2, N3)) * O.reshape(-1)[0])[src_idx] if O.size == 1 else O[src_idx]
    ox, oy, oz = (bro(O).astype(dtype) for O in (Ox, Oy, Oz))

    # sensors
    SensorMap = np.asarray(SensorMap)
    IndexSensorMap = (np.flatnonzero(SensorMap.flatten(order='F')) + 1).astype(np.uint32)
    si = (IndexSensorMap.astype(np.int64) - 1)
    s_i = si % N1
    s_j = (si // N1) % N2
    s_k = si // (N1 * N2)
    sub = int(SensorSubSampling)
    n0 = int(SensorStart) * sub
    sample_steps = [n for n in range(steps) if n % sub == 0 and n // sub >= SensorStart]
    Sensor = {k: np.zeros((len(si), len(sample_steps)), dtype) for k in SelMapsSensorsList}
    Sensor['time'] = np.array(sample_steps, float) * float(dt)

    sel = [k for k in MAP_ORDER if k in SelMapsRMSPeakList]
    acc = {k: z() for k in sel} if (SelRMSorPeak & 1) else {}
    peak = {k: z() for k in sel} if (SelRMSorPeak & 2) else {}
    refl = None if ReflectorMask is None else (np.asarray(ReflectorMask) != 0)

    def accumulate(fields):
        for k, v in fields.items():
            if k in acc:
                acc[k] += np.where(interior, v * v, 0)
            if k in peak:
                peak[k] = np.where(interior, np.maximum(peak[k], v), peak[k])

    half = dtype(0.5)
    for n in range(steps):
        # ---------------- stress half-step
        Dxx, Dyy, Dzz = dbwd(Vx, 0), dbwd(Vy, 1), dbwd(Vz, 2)
        th = Dxx + Dyy + Dzz
        att = (tauL != 0) | (tauS != 0)
        LM = M * (1 + tauL)
        Mi2 = 2 * G * (1 + tauS)
        LMC = dt * M * (tauL * ots)
        MC = dt * 2 * G * (tauS * ots)
        den = 1 + dt * half * ots
        num = 1 - dt * half * ots
        Pr = np.where(interior, Pr + dt * th, Pr)
        for S, R, oth, nm, cc in ((Sxx, Rxx, Dyy + Dzz, 'Sxx', Dxx), (Syy, Ryy, Dxx + Dzz, 'Syy', Dyy),
                                  (Szz, Rzz, Dxx + Dyy, 'Szz', Dzz)):
            NextR = np.where(att, (num * R - LMC * th + MC * oth) / den, R)
            Sint = S + dt * (LM * th - Mi2 * oth + np.where(att, half * (R + NextR), 0))
            # PML split parts: own-direction part uses M (lambda+2mu), the others lambda
            px = split(nm + '_x', cI, M if nm == 'Sxx' else L, Dxx)
            py = split(nm + '_y', cJ, M if nm == 'Syy' else L, Dyy)
            pz = split(nm + '_z', cK, M if nm == 'Szz' else L, Dzz)
            S[...] = np.where(interior, Sint, np.where(upd, px + py + pz, S))
            R[...] = np.where(interior, NextR, R)
        for S, R, nm, (fa, aa, ca, na), (fb, ab, cb, nb_) in (
                (Sxy, Rxy, 'xy', (Vy, 0, hI, 'Sxy_x'), (Vx, 1, hJ, 'Sxy_y')),
                (Sxz, Rxz, 'xz', (Vz, 0, hI, 'Sxz_x'), (Vx, 2, hK, 'Sxz_z')),
                (Syz, Ryz, 'yz', (Vz, 1, hJ, 'Syz_y'), (Vy, 2, hK, 'Syz_z'))):
            Da, Db = dfwd(fa, aa), dfwd(fb, ab)
            D = Da + Db
            rig, te = Rig[nm], TauE[nm]
            on = interior & (rig != 0)
            ta = on & (te != 0)
            NextR = np.where(ta, (num * R - dt * (rig * (te * ots)) * D) / den, R)
            Sint = S + dt * (rig * (1 + te) * D + np.where(ta, half * (R + NextR), 0))
            pa = split(na, ca, rig, Da)
            pb = split(nb_, cb, rig, Db)
            S[...] = np.where(on, Sint, np.where(upd, pa + pb, S))
            R[...] = np.where(ta, NextR, R)
        if refl is not None:
            for S in (Sxx, Syy, Szz, Sxy, Sxz, Syz, Pr):
                S[refl] = 0
        pscaled = -tab['K'][mm] * Pr
        if n >= n0:  # RMS/peak are taken inside the half-step, before this step's source is added
            accumulate({'Sigmaxx': Sxx, 'Sigmayy': Syy, 'Sigmazz': Szz, 'Sigmaxy': Sxy, 'Sigmaxz': Sxz,
                        'Sigmayz': Syz, 'Pressure': pscaled})
        if TypeSource >= 2 and n < nt_src and len(src_id):
            val = SF[src_id, n].astype(dtype) * ox
            for S in (Sxx, Syy, Szz):
                if TypeSource == 2:
                    S[src_idx] += val
                else:
                    S[src_idx] = val
        # ---------------- particle half-step
        for V, Bv, nm, (f1, k1, a1, c1), (f2, k2, a2, c2), (f3, k3, a3, c3) in (
                (Vx, Bx, 'Vx', (Sxx, 'f', 0, hI), (Sxy, 'b', 1, cJ), (Sxz, 'b', 2, cK)),
                (Vy, By, 'Vy', (Sxy, 'b', 0, cI), (Syy, 'f', 1, hJ), (Syz, 'b', 2, cK)),
                (Vz, Bz, 'Vz', (Sxz, 'b', 0, cI), (Syz, 'b', 1, cJ), (Szz, 'f', 2, hK))):
            D1 = dfwd(f1, a1) if k1 == 'f' else dbwd(f1, a1)
            D2 = dfwd(f2, a2) if k2 == 'f' else dbwd(f2, a2)
            D3 = dfwd(f3, a3) if k3 == 'f' else dbwd(f3, a3)
            Vint = V + dt * Bv * (D1 + D2 + D3)
            p1 = split(nm + '_x', c1, Bv, D1)
            p2 = split(nm + '_y', c2, Bv, D2)
            p3 = split(nm + '_z', c3, Bv, D3)
            V[...] = np.where(interior, Vint, np.where(upd, p1 + p2 + p3, V))
        if refl is not None:
            for V in (Vx, Vy, Vz):
                V[refl] = 0
        if n >= n0:
            accumulate({'Vx': Vx, 'Vy': Vy, 'Vz': Vz})
            if 'ALLV' in acc:
                acc['ALLV'] += np.where(interior, Vx * Vx + Vy * Vy + Vz * Vz, 0)
            if 'ALLV' in peak:
                peak['ALLV'] = np.where(interior, np.maximum(peak['ALLV'], Vx * Vx + Vy * Vy + Vz * Vz), peak['ALLV'])
        if TypeSource < 2 and n < nt_src and len(src_id):
            val = SF[src_id, n].astype(dtype)
            for V, o in ((Vx, ox), (Vy, oy), (Vz, oz)):
                if TypeSource == 0:
                    V[src_idx] += val * o
                else:
                    V[src_idx] = val * o
        # ---------------- sensors
        if n in sample_steps:
            q = sample_steps.index(n)
            fields = {'Vx': Vx, 'Vy': Vy, 'Vz': Vz, 'Sigmaxx': Sxx, 'Sigmayy': Syy, 'Sigmazz': Szz,
                      'Sigmaxy': Sxy, 'Sigmaxz': Sxz, 'Sigmayz': Syz, 'Pressure': pscaled}
            for k in SelMapsSensorsList:
                if k == 'ALLV':
                    Sensor[k][:, q] = np.sqrt(Vx ** 2 + Vy ** 2 + Vz ** 2)[s_i, s_j, s_k]
                else:
                    Sensor[k][:, q] = fields[k][s_i, s_j, s_k]

    nacc = max(steps - n0, 1)
    RMS = {k: np.sqrt(v / nacc) for k, v in acc.items()}
    Peak = {k: (np.sqrt(v) if k == 'ALLV' else v) for k, v in peak.items()}
    LastMap = {'Vx': Vx, 'Vy': Vy, 'Vz': Vz, 'Sigmaxx': Sxx, 'Sigmayy': Syy, 'Sigmazz': Szz,
               'Sigmaxy': Sxy, 'Sigmaxz': Sxz, 'Sigmayz': Syz, 'Pressure': -tab['K'][mm] * Pr}
    return dict(Sensor=Sensor, LastMap=LastMap, RMS=RMS, Peak=Peak, IndexSensorMap=IndexSensorMap,
                steps=steps, dt=float(dt))
