"""
ORACLE (test infrastructure only) -- NumPy restatement of the phase / amplitude extraction BabelBrain runs on the
host after a simulation: SimulationConditionsBASE.CalculatePhaseData, forward branch (bRefocused=False),
/root/reference/TranscranialModeling/BabelIntegrationBASE.py:2489-2518.

PINNED: tests/golden/phase_data_ref.npz holds outputs of the reference's own method, executed from
/root/reference by tests/golden/make_phase_golden.py on seeded inputs; tests/test_phase_data.py checks this
restatement against them bit for bit.  The CUDA path (bb_fdtd_get_phase_data) is then checked against this file.
"""
import numpy as np


def spectrum_index(sensor_time, frequency):
    """:2489 and :2498-2499 -- the FFT bin closest to the driving frequency."""
    time_step = np.diff(sensor_time).mean()
    freqs = np.fft.fftfreq(sensor_time.size, time_step)
    return int(np.argmin(np.abs(freqs - frequency)))


def calculate_phase_data(sensor_time, sensor_pressure, index_sensor_map, shape, frequency, ppp, sensor_subsampling):
    """Returns (PhaseMap float32, PressMapFourier complex64, PressMapPeak float32), each of `shape` = (N1,N2,N3).
    sensor_pressure is (Nsensors, Nsamples) float32 in IndexSensorMap row order; index_sensor_map is the 1-based
    Fortran-order linear index i + j*N1 + k*N1*N2 + 1 (:2503, :2508-2511)."""
    n1, n2, n3 = shape
    phase = np.zeros(shape, np.float32)                    # :2474-2476
    fourier = np.zeros(shape, np.complex64)
    peak = np.zeros(shape, np.float32)
    # :2491-2497 -- the caller truncates when the sample count is not a whole number of periods (and slices the wrong
    # axis of Sensor['Pressure'] doing so); the solver always returns whole periods, so this is an assertion here
    if sensor_time.shape[0] % (ppp / sensor_subsampling) != 0:
        raise ValueError('sample count %d is not a multiple of PPP/SensorSubSampling' % sensor_time.shape[0])
    ind = spectrum_index(sensor_time, frequency)
    pressure = np.ascontiguousarray(sensor_pressure)       # :2501
    index = index_sensor_map - 1                           # :2503
    nstep = 100000                                         # :2504
    for n in range(0, pressure.shape[0], nstep):
        top = min(n + nstep, pressure.shape[0])
        fsignal = np.fft.fft(pressure[n:top, :], axis=1)   # :2507 (numpy.fft when mkl_fft is absent, :34-37)
        k = index[n:top] // (n1 * n2)
        j = index[n:top] % (n1 * n2)
        i = j % n1
        j = j // n1
        fsignal = fsignal[:, ind]
        phase[i, j, k] = np.angle(fsignal)
        fourier[i, j, k] = fsignal
        peak[i, j, k] = pressure[n:top, :].max(axis=1)
    fourier *= 2 / sensor_time.size                        # :2518
    return phase, fourier, peak
