"""
ORACLE package -- test infrastructure only (see the headers of fdtd_numpy.py / fdtd_oracle.c).
PARITY UNPINNED: BabelViscoFDTD (the package holding the reference arithmetic) is absent from
/root/reference and from this image; there are no reference golden vectors for this path.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes
import os
import subprocess
import numpy as np

from . import fdtd_numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile liboracle_f32.so / liboracle_f64.so with the committed Makefile (gcc + OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ('fdtd_oracle.c', 'rayleigh_oracle.c', 'Makefile')]
    newest = max(os.path.getmtime(s) for s in srcs)
    for name in ('liboracle_f32.so', 'liboracle_f64.so'):
        p = os.path.join(_HERE, name)
        if force or not os.path.exists(p) or os.path.getmtime(p) < newest:
            subprocess.run(['make', '-C', _HERE, '-B'], check=True, capture_output=True)
            break


class _Params(ctypes.Structure):
    _fields_ = [('n1', ctypes.c_int32), ('n2', ctypes.c_int32), ('n3', ctypes.c_int32),
                ('pml', ctypes.c_int32), ('nmat', ctypes.c_int32), ('nsrc', ctypes.c_int32),
                ('nt_src', ctypes.c_int32), ('steps', ctypes.c_int32), ('type_source', ctypes.c_int32),
                ('sel_rms_peak', ctypes.c_int32), ('sel_maps_rms', ctypes.c_uint32),
                ('sel_maps_sensor', ctypes.c_uint32), ('sensor_subsampling', ctypes.c_int32),
                ('sensor_start', ctypes.c_int32), ('nsrc_cells', ctypes.c_int64),
                ('nsensors', ctypes.c_int64), ('dt', ctypes.c_double), ('mpml', ctypes.c_double)]


def load(dtype=np.float32):
    key = np.dtype(dtype).name
    if key not in _LIBS:
        name = 'liboracle_f32.so' if key == 'float32' else 'liboracle_f64.so'
        path = os.path.join(_HERE, name)
        try:
            if not os.path.exists(path):
                build()
            lib = ctypes.CDLL(path)
        except OSError:  # built on another host CPU: rebuild here
            build(force=True)
            lib = ctypes.CDLL(path)
        lib.oracle_fdtd_run.restype = ctypes.c_int
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIBS[key] = lib
    return _LIBS[key]


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def mask_of(names):
    m = 0
    for n in names:
        m |= fdtd_numpy.MAP_BITS[n]
    return m


def run_c(MaterialMap, MaterialProperties, Frequency, SourceMap, SourceFunctions, SpatialStep,
          DurationSimulation, SensorMap, Ox=1.0, Oy=1.0, Oz=1.0, NDelta=12, DT=None,
          ReflectionLimit=1e-5, AlphaCFL=1.0, TypeSource=0, QfactorCorrection=True, QCorrection=1.0,
          SelRMSorPeak=1, SelMapsRMSPeakList=('Pressure',), SelMapsSensorsList=('Pressure',),
          SensorSubSampling=2, SensorStart=0, ReflectorMask=None, dtype=np.float32, want_last=False,
          steps_override=None, MPMLRatio=None):
    """Same arguments / result dict as fdtd_numpy.run, computed by the C/OpenMP oracle."""
    F = fdtd_numpy
    lib = load(dtype)
    MaterialMap = np.ascontiguousarray(MaterialMap, dtype=np.uint32)
    N1, N2, N3 = MaterialMap.shape
    N = N1 * N2 * N3
    MP = np.asarray(MaterialProperties, float)
    h = float(SpatialStep)
    T = F.material_tables(MP, Frequency, QfactorCorrection, h, QCorrection)
    dt_id = F.ideal_dt(MP, h, AlphaCFL)
    dt = dt_id if DT is None else float(DT)
    steps = F.number_of_steps(DurationSimulation, dt) if steps_override is None else int(steps_override)
    tables = np.ascontiguousarray(np.stack([T[k] for k in ('M', 'G', 'L', 'B', 'tauL', 'tauS', 'ots', 'K')]), dtype=dtype)
    pmltab = np.ascontiguousarray(np.stack(F.pml_damping(int(NDelta), h, MP[:, 1].max(), ReflectionLimit)), dtype=dtype)
    SourceMap = np.asarray(SourceMap)
    src_cell = np.flatnonzero(SourceMap.reshape(-1)).astype(np.int64)
    src_id = (SourceMap.reshape(-1)[src_cell].astype(np.int64) - 1).astype(np.int32)
    SF = np.asarray(SourceFunctions)
    srcfun = np.ascontiguousarray(SF.T, dtype=dtype)

    def bro(O):
        O = np.asarray(O, float)
        return np.full(src_cell.shape, O.reshape(-1)[0], dtype) if O.size == 1 else np.ascontiguousarray(O.reshape(-1)[src_cell], dtype=dtype)
    ox, oy, oz = bro(Ox), bro(Oy), bro(Oz)
    SensorMap = np.asarray(SensorMap)
    IndexSensorMap = (np.flatnonzero(SensorMap.flatten(order='F')) + 1).astype(np.uint32)
    si = IndexSensorMap.astype(np.int64) - 1
    s_i, s_j, s_k = si % N1, (si // N1) % N2, si // (N1 * N2)
    sensor_cell = np.ascontiguousarray((s_i * N2 + s_j) * N3 + s_k, dtype=np.int64)
    sub = int(SensorSubSampling)
    nsamples = len([n for n in range(steps) if n % sub == 0 and n // sub >= SensorStart])
    sel = [k for k in F.MAP_ORDER if k in SelMapsRMSPeakList]
    sels = [k for k in F.MAP_ORDER if k in SelMapsSensorsList]
    prm = _Params(N1, N2, N3, int(NDelta), MP.shape[0], SF.shape[0], SF.shape[1], steps, int(TypeSource),
                  int(SelRMSorPeak), mask_of(sel), mask_of(sels), sub, int(SensorStart), len(src_cell),
                  len(sensor_cell), dt, F.MPML_RATIO if MPMLRatio is None else float(MPMLRatio))
    out_rms = np.zeros((len(sel), N1, N2, N3), dtype) if SelRMSorPeak & 1 else None
    out_peak = np.zeros((len(sel), N1, N2, N3), dtype) if SelRMSorPeak & 2 else None
    out_sensor = np.zeros((len(sels), len(sensor_cell), nsamples), dtype)
    out_last = np.zeros((10, N1, N2, N3), dtype) if want_last else None
    refl = None if ReflectorMask is None else np.ascontiguousarray(ReflectorMask, dtype=np.uint32)
    r = lib.oracle_fdtd_run(ctypes.byref(prm), _p(MaterialMap), _p(tables), _p(pmltab), _p(src_cell), _p(src_id),
                            _p(ox), _p(oy), _p(oz), _p(srcfun), _p(sensor_cell), _p(refl), _p(out_rms),
                            _p(out_peak), _p(out_sensor), _p(out_last))
    assert r == nsamples, (r, nsamples)
    Sensor = {k: out_sensor[n] for n, k in enumerate(sels)}
    Sensor['time'] = np.array([n for n in range(steps) if n % sub == 0 and n // sub >= SensorStart], float) * dt
    res = dict(Sensor=Sensor, RMS={k: out_rms[n] for n, k in enumerate(sel)} if out_rms is not None else {},
               Peak={k: out_peak[n] for n, k in enumerate(sel)} if out_peak is not None else {},
               IndexSensorMap=IndexSensorMap, steps=steps, dt=dt, threads=lib.oracle_num_threads())
    if want_last:
        res['LastMap'] = {k: out_last[n] for n, k in enumerate(F.MAP_ORDER[1:])}
    return res


def rayleigh_c(cwvnb, center, ds, u0, rf, MaxDistance=-1.0, dtype=np.float32):
    """C oracle of ForwardSimple; returns complex array (Npts,)."""
    lib = load(dtype)
    cdt = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    center = np.ascontiguousarray(center, dtype=dtype).reshape(-1, 3)
    ds = np.ascontiguousarray(np.asarray(ds).reshape(-1), dtype=dtype)
    u0 = np.ascontiguousarray(np.asarray(u0).reshape(-1), dtype=cdt)
    rf = np.ascontiguousarray(rf, dtype=dtype).reshape(-1, 3)
    out = np.zeros(rf.shape[0], cdt)
    k = complex(np.asarray(cwvnb).reshape(-1)[0])
    R = ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double
    lib.oracle_rayleigh_forward(R(k.real), R(k.imag), ctypes.c_int64(center.shape[0]), _p(center), _p(ds), _p(u0),
                                ctypes.c_int64(rf.shape[0]), _p(rf), _p(out), R(MaxDistance))
    return out


def rayleigh_numpy(cwvnb, center, ds, u0, rf, MaxDistance=-1.0):
    """float64 NumPy restatement of ForwardSimple (small sizes only)."""
    k = complex(np.asarray(cwvnb).reshape(-1)[0])
    center = np.asarray(center, float).reshape(-1, 3)
    ds = np.asarray(ds, float).reshape(-1)
    u0 = np.asarray(u0).reshape(-1).astype(complex)
    rf = np.asarray(rf, float).reshape(-1, 3)
    out = np.zeros(rf.shape[0], complex)
    for p0 in range(0, rf.shape[0], 4096):
        d = rf[p0:p0 + 4096, None, :] - center[None, :, :]
        R = np.sqrt((d * d).sum(-1))
        w = ds[None, :] * np.exp(k.imag * R) / R * u0[None, :] * np.exp(-1j * k.real * R)
        if MaxDistance > 0:
            w = np.where(R > MaxDistance, 0, w)
        out[p0:p0 + 4096] = 1j * k * w.sum(1) / (2 * np.pi)
    return out
