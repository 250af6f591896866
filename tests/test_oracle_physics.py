"""Closed-form checks of the restated viscoelastic scheme that need no reference implementation (the solver package is not in
the reference tree): what a tone burst does at a fluid-solid interface and how fast shear and
compressional plane waves travel in a solid.  They pin the parts of the oracle the reference's own FDTD-vs-Rayleigh
acceptance data (water only) cannot reach -- shear stresses, the rigidity averaging at interfaces, the staggering of the
stress components -- to textbook acoustics, and with the oracle the CUDA path, which reproduces it to ~1e-7
(tests/test_gpu_parity.py).  C/OpenMP oracle in float64, small grids, a few seconds each."""
import numpy as np
import pytest

import oracle

F0 = 500e3
WATER = [1000.0, 1500.0, 0.0, 0.0, 0.0]
BONE = [1850.0, 2800.0, 1500.0, 0.0, 0.0]             # lossless: the closed forms below are the lossless ones


@pytest.fixture(scope='module', autouse=True)
def _build():
    oracle.build()


def burst(dt, steps, cycles=5):
    t = np.arange(steps) * dt
    T = cycles / F0
    return np.where(t < T, np.sin(2 * np.pi * F0 * t) * np.sin(np.pi * t / T) ** 2, 0.0)[None, :]


def run(MM, ML, h, dt, steps, src_k, O, maps, sensors_k, mpml=0.1, n12=40):
    shape = MM.shape
    SM = np.zeros(shape, np.uint32)
    SM[:, :, src_k] = 1                                # the whole cross-section, layer included: a plane wave on the axis
    SEN = np.zeros(shape, np.uint32)
    SEN[n12 // 2, n12 // 2, sensors_k] = 1
    ox, oy, oz = O
    r = oracle.run_c(MM, np.array(ML), F0, SM, burst(dt, steps), h, dt * steps, SEN, dtype=np.float64, NDelta=8, DT=dt,
                     Ox=np.array([ox]), Oy=np.array([oy]), Oz=np.array([oz]), SelMapsRMSPeakList=['Pressure'], SelMapsSensorsList=list(maps),
                     SelRMSorPeak=1, SensorSubSampling=1, SensorStart=0, QfactorCorrection=False, MPMLRatio=mpml)
    order = np.argsort(np.asarray(sensors_k))          # IndexSensorMap order = increasing k here (same i, j)
    idx = np.asarray(r['IndexSensorMap']).astype(np.int64) - 1
    k_of_row = idx // (shape[0] * shape[1])
    assert np.array_equal(np.sort(k_of_row), np.sort(np.asarray(sensors_k)))
    rows = {int(k): n for n, k in enumerate(k_of_row)}
    return {m: {int(k): r['Sensor'][m][rows[int(k)]] for k in sensors_k} for m in maps}, order


def test_normal_incidence_on_a_solid_half_space():
    """Water over bone, burst travelling along k: the transmitted normal stress and particle velocity and the reflected
    pressure follow the impedance ratio, T_sigma = 2 Z2/(Z1+Z2) = 1.551, T_v = 2 Z1/(Z1+Z2) = 0.449, R = (Z2-Z1)/(Z2+Z1) =
    0.551.  Measured on this 40-cell-wide grid at 12 points per wavelength: 1.511, 0.445, 0.542 (64 cells wide: 1.523, 0.452,
    0.535) -- the few-percent remainder is the absorbing side walls two wavelengths from the axis and the half-cell offset
    between the stress and the velocity nodes at the interface; a wrong averaging rule or a missing factor is tens of
    percent."""
    n12, n3, ppw = 40, 150, 12
    h = 1500.0 / F0 / ppw
    dt = 0.4 * h / 2800.0 / np.sqrt(3.0)
    # interface, source plane, sensors just in front of and just behind the interface: the side walls of this narrow grid are
    # absorbing layers, so the "plane" wave loses amplitude along k -- alike in both runs up to the interface, not behind it
    k_if, k_src, k_w, k_s = 90, 12, 86, 93
    steps = int(1.15 * ((k_if - k_src) * h / 1500.0 + (k_if - k_w) * h / 1500.0 + 5 / F0) / dt)
    out = {}
    for name, ML in (('water', [WATER, WATER]), ('bone', [WATER, BONE])):
        MM = np.zeros((n12, n12, n3), np.uint32)
        MM[:, :, k_if:] = 1
        out[name], _ = run(MM, ML, h, dt, steps, k_src, (0.0, 0.0, 1.0 / 1.5e6), ('Pressure', 'Sigmazz', 'Vz'), [k_w, k_s], n12=n12)
    Z1, Z2 = 1000.0 * 1500.0, 1850.0 * 2800.0
    env = lambda x: np.abs(x).max()
    # transmitted: first arrival at k_s, against the undisturbed burst at the same place of the all-water run
    t_sigma = env(out['bone']['Sigmazz'][k_s]) / env(out['water']['Sigmazz'][k_s])
    t_v = env(out['bone']['Vz'][k_s]) / env(out['water']['Vz'][k_s])
    assert t_sigma == pytest.approx(2 * Z2 / (Z1 + Z2), rel=0.04), t_sigma
    assert t_v == pytest.approx(2 * Z1 / (Z1 + Z2), rel=0.03), t_v
    # reflected: what the bone run adds at k_w to the all-water run
    refl = out['bone']['Pressure'][k_w] - out['water']['Pressure'][k_w]
    r = env(refl) / env(out['water']['Pressure'][k_w])
    assert r == pytest.approx((Z2 - Z1) / (Z2 + Z1), rel=0.04), r
    # in the water the pressure is minus every normal stress (no shear there)
    assert np.allclose(out['water']['Pressure'][k_w], -out['water']['Sigmazz'][k_w], atol=1e-9 * env(out['water']['Pressure'][k_w]))


@pytest.mark.parametrize('polarisation,speed', [('shear', 1500.0), ('compressional', 2800.0)])
def test_plane_wave_speeds_in_a_solid(polarisation, speed):
    """Homogeneous bone, a source plane moving the particles along i (shear wave along k) or along k (compressional wave
    along k).  The grid is a narrow channel between absorbing side walls, i.e. a waveguide: the carrier travels faster than
    the bulk speed and the envelope slower, and for a guided mode (phase speed) x (group speed) = c^2.  Both are measured
    from the burst's arrival at two depths; their geometric mean must be c_S or c_L of the MaterialList row within 1 %
    (measured: 1547 x 1465 -> 1505 for 1500, 2877 x 2764 -> 2820 for 2800; water gives 1507 for 1500 the same way)."""
    from scipy.signal import hilbert
    n12, n3, ppw = 40, 200, 12
    h = speed / F0 / ppw
    dt = 0.4 * h / 2800.0 / np.sqrt(3.0)
    k_src, k1, k2 = 12, 50, 150
    steps = int(1.1 * ((k2 - k_src) * h / speed + 5 / F0) / dt)
    MM = np.zeros((n12, n12, n3), np.uint32)
    O, comp = ((1.0, 0.0, 0.0), 'Vx') if polarisation == 'shear' else ((0.0, 0.0, 1.0), 'Vz')
    tr, _ = run(MM, [BONE], h, dt, steps, k_src, O, (comp,), [k1, k2], n12=n12)
    a, b = tr[comp][k1], tr[comp][k2]

    def delay(x, y):                                   # lag of y behind x in steps, refined with a parabola through the peak
        c = np.correlate(y, x, 'full')
        p = int(np.argmax(c))
        frac = 0.5 * (c[p - 1] - c[p + 1]) / (c[p - 1] - 2 * c[p] + c[p + 1])
        return p - (len(x) - 1) + frac
    dist = (k2 - k1) * h
    v_phase = dist / (delay(a, b) * dt)
    v_group = dist / (delay(np.abs(hilbert(a)), np.abs(hilbert(b))) * dt)
    assert v_phase > speed > v_group, (v_phase, v_group)
    assert np.sqrt(v_phase * v_group) == pytest.approx(speed, rel=0.01), (v_phase, v_group, speed)


@pytest.mark.parametrize('polarisation,speed,column', [('shear', 1500.0, 4), ('compressional', 2800.0, 3)])
def test_attenuation_of_shear_and_compressional_waves_in_a_solid(polarisation, speed, column):
    """The same channel once with lossless and once with attenuating bone (alpha_L = 60 Np/m, alpha_S = 120 Np/m at the drive
    frequency): walls and diffraction act on both bursts alike, so the amplitude ratio between two depths differs by
    exp(-alpha dz).  The fitted alpha must be within 15 % of the MaterialList value of that wave type -- the relaxation fit of
    the shear modulus is a separate code path from the compressional one (memory variables of the shear stresses)."""
    n12, n3, ppw = 40, 160, 12
    h = speed / F0 / ppw
    dt = 0.4 * h / 2800.0 / np.sqrt(3.0)
    k_src, k1, k2 = 12, 40, 120
    steps = int(1.1 * ((k2 - k_src) * h / speed + 5 / F0) / dt)
    MM = np.zeros((n12, n12, n3), np.uint32)
    O, comp = ((1.0, 0.0, 0.0), 'Vx') if polarisation == 'shear' else ((0.0, 0.0, 1.0), 'Vz')
    lossy = [1850.0, 2800.0, 1500.0, 60.0, 120.0]
    drop = []
    for row in (BONE, lossy):
        tr, _ = run(MM, [row], h, dt, steps, k_src, O, (comp,), [k1, k2], n12=n12)
        drop.append(np.log(np.abs(tr[comp][k2]).max() / np.abs(tr[comp][k1]).max()))
    alpha = -(drop[1] - drop[0]) / ((k2 - k1) * h)
    assert alpha == pytest.approx(lossy[column], rel=0.15), (alpha, lossy[column])
