"""CPU tests of the oracle itself (test infrastructure): the C/OpenMP restatement against the float64
NumPy restatement, both against the committed golden vectors, and checks that do not need the
reference: symmetry, p == -Sigmaxx in a lossless fluid, PML absorption, linearity, Rayleigh."""
import os

import numpy as np
import pytest

import oracle
from oracle import fdtd_numpy
from babelbrain_b200 import workloads
from tests.golden import make_golden

DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rl2(a, b):
    wide = np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64
    a, b = np.asarray(a, wide), np.asarray(b, wide)
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / nb if nb > 0 else float(np.linalg.norm(a - b))


def kwargs_of(w):
    return {k: v for k, v in w['kwargs'].items() if k not in DROP}


@pytest.fixture(scope='module', autouse=True)
def _build():
    oracle.build()


@pytest.mark.parametrize('name', list(make_golden.CASES))
def test_oracles_reproduce_golden(name):
    """float64 C oracle == golden (float64 NumPy) to rounding; float32 C oracle within the 1e-4 budget."""
    g = np.load(os.path.join(GOLD, name + '.npz'))
    w = make_golden.build_case(name)
    assert str(g['digest']) == make_golden.inputs_digest(w), 'the seeded workload generator drifted'
    r64 = oracle.run_c(*w['args'], dtype=np.float64, want_last=True, **kwargs_of(w))
    r32 = oracle.run_c(*w['args'], dtype=np.float32, want_last=True, **kwargs_of(w))
    assert np.array_equal(r64['IndexSensorMap'], g['IndexSensorMap'])
    assert int(g['steps']) == r64['steps'] == r32['steps']
    assert np.allclose(r64['Sensor']['time'], g['time'], rtol=1e-12)
    for key in g.files:
        kind, _, mapname = key.partition('_')
        if kind not in ('RMS', 'Peak', 'Sensor', 'Last'):
            continue
        grp = {'RMS': 'RMS', 'Peak': 'Peak', 'Sensor': 'Sensor', 'Last': 'LastMap'}[kind]
        assert rl2(r64[grp][mapname], g[key]) < 2e-6, (key, rl2(r64[grp][mapname], g[key]))   # golden is stored as float32
        assert rl2(r32[grp][mapname], g[key]) < 1e-4, (key, rl2(r32[grp][mapname], g[key]))


def test_numpy_and_c_agree_with_reflector_and_hard_sources():
    w = workloads.make_workload('ctx500_skull', shape=(20, 22, 26), periods=2, pml=4)
    refl = np.zeros(w['args'][0].shape, np.uint32)
    refl[8:11, 6:14, 14:17] = 1
    for ts in (0, 1):
        kw = dict(kwargs_of(w), ReflectorMask=refl, TypeSource=ts)
        a = fdtd_numpy.run(*w['args'], dtype=np.float64, **kw)
        b = oracle.run_c(*w['args'], dtype=np.float64, **kw)
        assert rl2(b['RMS']['Pressure'], a['RMS']['Pressure']) < 1e-10
        assert rl2(b['Sensor']['Pressure'], a['Sensor']['Pressure']) < 1e-10
        assert np.all(a['RMS']['Pressure'][8:11, 6:14, 14:17] == 0)


def test_lossless_fluid_pressure_equals_minus_sigma():
    w = workloads.make_workload('single_water', shape=(26, 26, 34), periods=3, pml=5)
    kw = dict(kwargs_of(w), SelMapsRMSPeakList=['Pressure', 'Sigmaxx', 'Sigmayy', 'Sigmaxy'])
    r = oracle.run_c(*w['args'], dtype=np.float64, **kw)
    assert rl2(r['RMS']['Pressure'], r['RMS']['Sigmaxx']) < 1e-10
    assert rl2(r['RMS']['Sigmayy'], r['RMS']['Sigmaxx']) < 1e-10
    assert np.all(r['RMS']['Sigmaxy'] == 0)


def test_symmetry_and_linearity():
    w = workloads.make_workload('single_water', shape=(28, 28, 36), periods=4, pml=5)
    r = oracle.run_c(*w['args'], dtype=np.float64, **kwargs_of(w))
    p = r['RMS']['Pressure']
    # x and y are treated identically: exact swap symmetry; the staggered grid is not mirror symmetric
    # (half-cell offsets, the last PML cell is never updated), so mirroring only holds approximately
    assert rl2(np.swapaxes(p, 0, 1), p) < 1e-9
    assert rl2(p[::-1], p) < 5e-3 and rl2(p[:, ::-1], p) < 5e-3
    args = list(w['args'])
    args[4] = args[4] * 2.5
    r2 = oracle.run_c(*args, dtype=np.float64, **kwargs_of(w))
    assert rl2(r2['RMS']['Pressure'], 2.5 * p) < 1e-12
    assert np.all(p[:5] == 0) and np.all(p[:, :, -5:] == 0)     # the RMS map excludes the PML shell


def test_pml_absorbs():
    """A short burst must leave the domain: after many transit times what remains is a small
    fraction of the peak field energy (a rigid wall would keep all of it; the 10-cell layer sits a
    few wavelengths from an abruptly switched source, so grazing and low-frequency content sets the floor)."""
    w = workloads.make_workload('single_water', shape=(36, 36, 40), periods=14, pml=10)
    MM, ML, f, SM, SF, h, T, SEN = w['args']
    SF = SF.copy()
    ppp = w['meta']['ppp']
    SF[:, 2 * ppp:] = 0.0                                      # two periods of drive, then silence
    kw = dict(kwargs_of(w), SensorStart=0, SelMapsSensorsList=['Pressure'])
    r = oracle.run_c(MM, ML, f, SM, SF, h, T, SEN, dtype=np.float64, **kw)
    e = (r['Sensor']['Pressure'] ** 2).sum(0)                 # energy proxy per sample
    assert e[-1] < 5e-3 * e.max(), (e[-1], e.max())
    assert e[-4:].mean() < e[-12:-8].mean()                   # and it keeps draining


def test_attenuating_medium_decays_at_the_requested_rate():
    """Beam along k in a homogeneous fluid, once lossless and once attenuating: diffraction and the
    lateral PML act on both alike to first order, so the amplitude ratio decays as exp(-alpha z).  The
    fitted alpha must be within 15% of the MaterialList value at the drive frequency: tight enough to
    catch a unit or factor-of-two slip in the relaxation fit (Np vs dB, Q vs 2Q), loose enough for the
    near-field ripple of a 4-wavelength aperture."""
    f = 500e3
    alpha = 40.0                                               # Np/m
    shape = (64, 64, 150)
    amps = []
    for a in (0.0, alpha):
        ML = np.array([[1000.0, 1500.0, 0.0, a, 0.0]])
        S = workloads.sizing(f, 9, np.array([[1000.0, 1500.0, 0.0, alpha, 0.0]]), shape, pml=6, periods=40)
        h, dt, steps = S['h'], S['dt'], S['steps']
        MM = np.zeros(shape, np.uint32)
        SM = np.zeros(shape, np.uint32)
        SM[6:-6, 6:-6, 6] = 1
        SF = workloads.cw_sources(np.array([1e5]), np.array([0.0]), f, dt, steps)
        SEN = np.zeros(shape, np.uint32)
        SEN[32, 32, 8:-8] = 1
        r = oracle.run_c(MM, ML, f, SM, SF, h, dt * steps, SEN, dtype=np.float64, NDelta=6, DT=dt, Ox=np.array([0.0]), Oy=np.array([0.0]),
                         Oz=np.array([1.0 / 1.5e6]), SelMapsRMSPeakList=['Pressure'], SelMapsSensorsList=['Pressure'], SelRMSorPeak=1,
                         SensorSubSampling=S['sub'], SensorStart=S['sensor_start'], QfactorCorrection=True)
        amps.append(r['RMS']['Pressure'][32, 32, 20:120])
    z = np.arange(20, 120) * h
    slope = np.polyfit(z, np.log(amps[1] / amps[0]), 1)[0]
    assert abs(-slope - alpha) < 0.15 * alpha, (-slope, alpha)


def test_rayleigh_c_against_numpy():
    rng = np.random.default_rng(5)
    center = (rng.random((200, 3)).astype(np.float32) - 0.5) * 0.04
    center[:, 2] -= 0.06
    ds = np.full((200, 1), 2e-6, np.float32)
    u0 = (rng.random(200) + 1j * rng.random(200)).astype(np.complex64)
    rf = (rng.random((500, 3)).astype(np.float32) - 0.5) * 0.04
    for k in (2 * np.pi * 5e5 / 1500 + 0j, 2 * np.pi * 5e5 / 1500 - 3.0j):
        a = oracle.rayleigh_c(np.complex64(k), center, ds, u0, rf, dtype=np.float64)
        b = oracle.rayleigh_numpy(np.complex64(k), center, ds, u0, rf)
        assert rl2(a, b) < 1e-10
        a32 = oracle.rayleigh_c(np.complex64(k), center, ds, u0, rf, dtype=np.float32)
        assert rl2(a32, b) < 1e-4


def test_classical_layer_is_stable_on_a_water_shell_and_the_multiaxial_one_where_tissue_enters_it():
    """The caller's label maps are water inside the absorbing shell (BabelIntegrationBASE.py:2110,:2154-2159): there the
    classical split-field layer (MPMLRatio 0, the default) settles to a steady state.  A map whose fluid-solid interfaces
    run INTO the layer (tissue_in_shell) makes the classical layer grow by many orders of magnitude over 100 periods; the
    multi-axial layer (MPMLRatio 0.1) settles (profiles/r1_pml_stability.txt, profiles/r2_pml_stability.txt)."""
    def rms_max(periods, ratio, tissue_in_shell):
        w = workloads.make_workload('ctx500_skull', shape=(40, 36, 56), periods=periods, pml=6, tissue_in_shell=tissue_in_shell)
        return float(oracle.run_c(*w['args'], MPMLRatio=ratio, **kwargs_of(w))['RMS']['Pressure'].max())
    assert fdtd_numpy.MPML_RATIO == 0.0
    assert abs(rms_max(100, None, False) / rms_max(50, None, False) - 1) < 0.02
    steady = rms_max(50, 0.1, True)
    late = rms_max(100, 0.1, True)
    assert abs(late / steady - 1) < 0.02
    assert rms_max(100, 0.0, True) > 1e4 * late


def test_field_of_a_focusing_source_plane_agrees_with_the_rayleigh_integral():
    """The two halves of the hot path against each other ("independent checks"): a water domain driven
    by a focusing source plane, steady-state amplitude from the single-bin DFT of the sensor traces, against the
    Rayleigh integral of the same source distribution (pistons of area h^2, u0 = A exp(j phi)).  A soft velocity
    source of amplitude A/(rho c) added every step radiates a plane wave of pressure A h / (2 c dt), which fixes the
    scale between the two; at 6 points per wavelength the maps agree to a few per cent (numerical dispersion)."""
    from oracle import phase_data
    from babelbrain_b200.sources import CWSourceFunctions
    pml = 10
    w = workloads.make_workload('single_water', shape=(72, 72, 96), pml=pml)
    m = w['meta']
    MM, ML, f, SM, _, h, T, SEN = w['args']
    cw0 = m['cw_sources']
    cw = CWSourceFunctions(cw0.amplitude, -cw0.phase, cw0.Frequency, cw0.TemporalStep, T)      # outer pixels lead: focusing
    r = oracle.run_c(MM, ML, f, SM, cw.dense(), h, T, SEN, **kwargs_of(w))
    _, fo, _ = phase_data.calculate_phase_data(r['Sensor']['time'], r['Sensor']['Pressure'], r['IndexSensorMap'], MM.shape, f,
                                               m['ppp'], m['sub'])
    n1, n2, n3 = MM.shape
    x = (np.arange(n1) - n1 / 2 + 0.5) * h
    y = (np.arange(n2) - n2 / 2 + 0.5) * h
    z = np.arange(n3) * h
    ii, jj = np.nonzero(SM[:, :, pml])
    rows = SM[ii, jj, pml] - 1
    center = np.stack([x[ii], y[jj], np.full(ii.size, (pml + 0.5) * h)], 1)       # Vz sits half a cell above the source plane
    u0 = cw.amplitude[rows] * np.exp(1j * cw.phase[rows])
    box = (slice(pml, -pml), slice(pml, -pml), slice(pml + 8, -pml))
    X, Y, Z = np.meshgrid(x[box[0]], y[box[1]], z[box[2]], indexing='ij')
    pr = np.abs(oracle.rayleigh_c(2 * np.pi * f / 1500.0 + 0j, center, np.full(ii.size, h * h), u0,
                                  np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1), dtype=np.float64)).reshape(X.shape)
    a = np.abs(fo)[box]
    scale = float((a * pr).sum() / (pr * pr).sum())
    assert abs(scale / (h / (2 * 1500.0 * m['dt'])) - 1) < 0.08
    assert np.linalg.norm(a - scale * pr) / np.linalg.norm(a) < 0.12
    pa, pb = np.unravel_index(a.argmax(), a.shape), np.unravel_index(pr.argmax(), pr.shape)
    assert max(abs(int(u) - int(v)) for u, v in zip(pa, pb)) <= 1 and pa[2] > 10     # a focus inside the volume, same voxel +-1
