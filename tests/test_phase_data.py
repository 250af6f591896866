"""Phase / amplitude extraction.  The oracle (oracle/phase_data.py) is pinned bit for bit
against outputs of the reference's own CalculatePhaseData (tests/golden/phase_data_ref.npz, produced by
tests/golden/make_phase_golden.py from /root/reference); the CUDA path is checked against the oracle on the sensor
traces of real simulations."""
import numpy as np
import pytest

from oracle import phase_data
from babelbrain_b200 import workloads

GOLD = __import__('os').path.join(__import__('os').path.dirname(__file__), 'golden', 'phase_data_ref.npz')


def test_oracle_reproduces_the_reference_method_bit_for_bit():
    g = np.load(GOLD)
    ph, fo, pk = phase_data.calculate_phase_data(g['time'], g['pressure'], g['index'], tuple(g['shape']), float(g['frequency']),
                                                 int(g['ppp']), int(g['sub']))
    assert fo.dtype == np.complex64 and ph.dtype == np.float32 and pk.dtype == np.float32
    assert np.array_equal(fo, g['PressMapFourier']) and np.array_equal(ph, g['PhaseMap']) and np.array_equal(pk, g['PressMapPeak'])
    assert np.count_nonzero(fo) == g['index'].size                  # zeros exactly where there is no sensor


def test_oracle_recovers_amplitude_and_phase_of_a_tone():
    f, ppp, sub = 500e3, 48, 12
    t = (100 + np.arange(8)) * sub / (f * ppp)
    amp, ph0 = 3.0e4, 0.7
    p = (amp * np.sin(2 * np.pi * f * t + ph0)).astype(np.float32)[None, :]
    idx = np.array([1 + 2 + 3 * 5 + 4 * 5 * 6], np.uint32)           # voxel (2, 3, 4) of a (5, 6, 7) grid
    ph, fo, pk = phase_data.calculate_phase_data(t, p, idx, (5, 6, 7), f, ppp, sub)
    assert abs(abs(fo[2, 3, 4]) - amp) < 1e-3 * amp
    expect = np.angle(np.exp(1j * (2 * np.pi * f * t[0] + ph0 - np.pi / 2)))
    assert abs(np.angle(np.exp(1j * (ph[2, 3, 4] - expect)))) < 1e-3
    with pytest.raises(ValueError):
        phase_data.calculate_phase_data(t[:7], p[:, :7], idx, (5, 6, 7), f, ppp, sub)


def _compare(res, Sensor, IP, shape, meta):
    ph, fo, pk = phase_data.calculate_phase_data(Sensor['time'], Sensor['Pressure'], IP['IndexSensorMap'], shape, meta['frequency'],
                                                 meta['ppp'], meta['sub'])
    assert res['IndSpectrum'] == phase_data.spectrum_index(Sensor['time'], meta['frequency'])
    scale = np.abs(fo).max()
    assert scale > 0
    assert np.abs(res['PressMapFourier'] - fo).max() <= 2e-6 * scale          # float32 DFT of <= 10 samples
    assert np.array_equal(res['PressMapPeak'], pk)                            # a maximum of the same float32 samples: bit-exact
    assert np.array_equal(res['PressMapFourier'] == 0, fo == 0)               # same support (sensor voxels only)
    strong = np.abs(fo) > 1e-3 * scale
    dphi = np.angle(np.exp(1j * (res['PhaseMap'].astype(np.float64) - ph)))
    assert np.abs(dphi[strong]).max() <= 1e-3
    assert int(np.argmax(np.abs(res['PressMapFourier']))) == int(np.argmax(np.abs(fo)))


@pytest.mark.gpu
@pytest.mark.parametrize('name,shape,pml', [('ctx500_skull', (56, 48, 72), 8), ('single_water', (40, 44, 56), 8)])
def test_device_phase_data_matches_the_oracle(name, shape, pml):
    from babelbrain_b200.propagation import PropagationModel
    w = workloads.make_workload(name, shape=shape, periods=6, pml=pml)
    PM = PropagationModel()
    Sensor, _, RMS, IP = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    res = PM.CalculatePhaseDataOnDevice(w['meta']['frequency'])
    _compare(res, Sensor, IP, shape, w['meta'])
    # amplitude from the spectrum ~ sqrt(2) * RMS for a steady tone (the caller's own cross-check, BabelIntegrationBASE.py:2440)
    focus = np.unravel_index(np.argmax(RMS['Pressure']), shape)
    assert abs(abs(res['PressMapFourier'][focus]) / (np.sqrt(2) * RMS['Pressure'][focus]) - 1) < 0.05
    with pytest.raises(ValueError):
        PM.CalculatePhaseDataOnDevice(w['meta']['frequency'], MapName='Vx')


@pytest.mark.gpu
def test_device_phase_data_on_two_gpus():
    from babelbrain_b200 import _capi
    from babelbrain_b200.propagation import PropagationModel
    if _capi.device_count() < 2:
        pytest.skip('needs two GPUs')
    shape = (66, 52, 60)
    w = workloads.make_workload('ctx500_skull', shape=shape, periods=5, pml=8)
    PM = PropagationModel()
    Sensor, _, RMS, IP = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], NumberGPUs=2, **w['kwargs'])
    _compare(PM.CalculatePhaseDataOnDevice(w['meta']['frequency']), Sensor, IP, shape, w['meta'])
