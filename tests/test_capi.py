"""The C-ABI library: it loads, exports every symbol include/babelb200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from babelbrain_b200 import _capi, build, workloads
from babelbrain_b200.propagation import PropagationModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _capi.lib()


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'babelb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bb_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'include/babelb200.h declares %s but libbabelb200.so does not export it' % n
    assert sorted(_capi.SYMBOLS) == names, 'babelbrain_b200/_capi.py SYMBOLS is out of sync with the header'


def test_struct_layouts_match_the_header(lib):
    # bb_fdtd_desc: 21 int32/uint32 fields, then a double (8-byte aligned)
    assert ctypes.sizeof(_capi.FdtdDesc) == 96
    assert _capi.FdtdDesc.dt.offset == 88
    assert ctypes.sizeof(_capi.FdtdStats) == 5 * 8 + 8 * 8
    assert lib.bb_version().decode().startswith('babelb200')


def _no_gpu(lib):
    return lib.bb_device_count() <= 0


def test_no_cpu_fallback_without_a_device(lib):
    if not _no_gpu(lib):
        pytest.skip('a CUDA device is present')
    d = _capi.FdtdDesc(n1=32, n2=32, n3=32, i0=0, i1=32, pml=4, nmat=1, nsrc=1, nt_src=1, steps=1, sel_rms_peak=1, sel_maps_rms=1 << 10,
                       sel_maps_sensor=1 << 10, sensor_subsampling=1, device=0, rank=0, nranks=1, dt=1e-8)
    h = ctypes.c_void_p()
    rc = lib.bb_fdtd_create(ctypes.byref(d), ctypes.byref(h))
    assert rc == 2 and lib.bb_last_error()                       # BB_ERR_CUDA with a message
    out = np.zeros(2, np.float32)
    one = np.ones(3, np.float32)
    rc = lib.bb_rayleigh_forward(1.0, 0.0, 1, _capi.ptr(one), _capi.ptr(one), _capi.ptr(one), 1, _capi.ptr(one), _capi.ptr(out), -1.0, 0, 0, None)
    assert rc == 2
    w = workloads.make_workload('single_water', shape=(24, 24, 28), periods=1, pml=4)
    with pytest.raises(_capi.BabelB200Error):
        PropagationModel().StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])


def test_argument_errors_come_before_the_device(lib):
    """dtype / shape errors are the caller's and are raised as TypeError / ValueError like the reference does."""
    w = workloads.make_workload('single_water', shape=(24, 24, 28), periods=1, pml=4)
    PM = PropagationModel()
    bad = list(w['args'])
    bad[3] = bad[3].astype(np.int64)
    with pytest.raises(TypeError):
        PM.StaggeredFDTD_3D_with_relaxation(*bad, **w['kwargs'])
    bad = list(w['args'])
    bad[7] = bad[7][:-1]
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*bad, **w['kwargs'])
    bad = list(w['args'])
    bad[0] = bad[0] + 3
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*bad, **w['kwargs'])
    with pytest.raises(ValueError):
        PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], DT=1.0))
    with pytest.raises(NotImplementedError):
        PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **dict(w['kwargs'], IntervalSnapshots=10))


def test_drop_in_import_surface():
    """The names BabelBrain imports from BabelViscoFDTD resolve to this package."""
    from BabelViscoFDTD.PropagationModel import PropagationModel as PM2
    from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple, InitCuda, InitOpenCL, InitMetal, SpeedofSoundWater  # noqa: F401
    from BabelViscoFDTD.H5pySimple import ReadFromH5py, SaveToH5py  # noqa: F401
    import BabelViscoFDTD.StaggeredFDTD_3D_With_Relaxation_CUDA as C
    assert PM2 is PropagationModel and callable(C.ListDevices)
    assert 1480 < SpeedofSoundWater(20.0) < 1485


def test_shim_keeps_the_solvers_and_delegates_file_io_to_a_genuine_install(tmp_path):
    """With a genuine BabelViscoFDTD further down sys.path, the shim keeps everything that computes (ForwardSimple,
    PropagationModel, BHTE / BHTEMultiplePressureFields) and hands out the genuine H5pySimple (same on-disk format)."""
    import subprocess
    import sys
    up = tmp_path / 'site' / 'BabelViscoFDTD'
    (up / 'tools').mkdir(parents=True)
    (up / '__init__.py').write_text('__version__ = "1.2.4"\n')
    (up / 'tools' / '__init__.py').write_text('')
    (up / 'tools' / 'RayleighAndBHTE.py').write_text('def BHTE(*a, **k):\n    return "genuine BHTE"\n'
                                                    'def ForwardSimple(*a, **k):\n    return "genuine FS"\n')
    (up / 'H5pySimple.py').write_text('def ReadFromH5py(f):\n    return "genuine read"\ndef SaveToH5py(d, f):\n    return "genuine save"\n')
    code = ('import sys; sys.path.insert(0, %r); sys.path.append(%r)\n'
            'from BabelViscoFDTD.tools.RayleighAndBHTE import BHTE, BHTEMultiplePressureFields, ForwardSimple\n'
            'from BabelViscoFDTD.H5pySimple import ReadFromH5py\n'
            'from BabelViscoFDTD.PropagationModel import PropagationModel\n'
            'print(BHTE.__module__, BHTEMultiplePressureFields.__module__, ForwardSimple.__module__, ReadFromH5py(None), PropagationModel.__module__)\n'
            % (ROOT, str(tmp_path / 'site')))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout.split('\n')[0]
    assert out == 'babelbrain_b200.thermal babelbrain_b200.thermal babelbrain_b200.rayleigh genuine read babelbrain_b200.propagation'


def test_bhte_has_no_cpu_fallback(lib):
    from BabelViscoFDTD.tools.RayleighAndBHTE import BHTE
    if not _no_gpu(lib):
        pytest.skip('a CUDA device is present')
    ML = {k: np.ones(1) for k in ('Density', 'SoS', 'Attenuation', 'SpecificHeat', 'Conductivity', 'Perfusion', 'Absorption', 'InitTemperature')}
    with pytest.raises(_capi.BabelB200Error):
        BHTE(np.zeros((4, 4, 4), np.float32), np.zeros((4, 4, 4), np.uint32), ML, 1e-3, 2, 1, -1, dt=0.01)


def test_host_nonzero_matches_numpy(lib):
    """bb_host_nonzero_u32 (source cells of the caller's SourceMap, host threads only): indices and values of np.flatnonzero,
    for empty, sparse, dense, unaligned and multi-threaded (> 1 M cells per thread) volumes."""
    from babelbrain_b200.propagation import _nonzero_u32
    rng = np.random.default_rng(5)
    for shape in ((3, 4, 5), (50, 40, 30), (129, 65, 33), (160, 160, 130)):
        for density in (0.0, 0.002, 0.4, 1.0):
            a = ((rng.random(shape) < density) * rng.integers(1, 2 ** 31, shape)).astype(np.uint32)
            for view in (a, a[1:]):                              # the second starts off a 32-byte boundary for odd plane sizes
                idx, val = _nonzero_u32(view)
                ref = np.flatnonzero(view.reshape(-1))
                assert idx.dtype == np.int64 and val.dtype == np.uint32
                assert np.array_equal(idx, ref) and np.array_equal(val, view.reshape(-1)[ref])


def test_host_scatter_rows_places_slab_rows(lib):
    """The gather helper of the multi-GPU path is plain host code (no device): out[rows[r]] = data[r]."""
    import ctypes
    lib.bb_host_scatter_rows.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 2
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(500, 120, replace=False)).astype(np.int64)
    data = rng.random((120, 10)).astype(np.float32)
    out = np.zeros((500, 10), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    assert lib.bb_host_scatter_rows(p(out), p(rows), p(data), 120, 40) == 0
    ref = np.zeros_like(out)
    ref[rows] = data
    assert np.array_equal(out, ref)
    assert lib.bb_host_scatter_rows(p(out), p(rows), p(data), 0, 40) == 0
    assert lib.bb_host_scatter_rows(None, p(rows), p(data), 3, 40) != 0


def test_pinned_pool_is_thread_safe():
    """The per-GPU threads of a slab-decomposed run allocate result buffers concurrently and finalizers give blocks back
    from any thread: no block may ever back two live arrays.  Ordinary memory stands in for page-locked
    memory, so this runs without a device."""
    import gc
    import threading

    class Block:
        def __init__(self, nbytes):
            self.buf = (ctypes.c_char * nbytes)()
            self.ptr, self.nbytes = ctypes.addressof(self.buf), nbytes

        def release(self):
            self.ptr = None

    pool = _capi.PinnedPool(max_idle_bytes=64 << 20, block_type=Block)
    errors = []

    def hammer(seed):
        rng = np.random.default_rng(seed)
        try:
            for it in range(300):
                sizes = rng.integers(1, 5000, 4)
                arrs = [pool.empty((int(n),), np.int64) for n in sizes]
                for t, a in enumerate(arrs):
                    a[:] = seed * 1000003 + it * 7 + t
                for t, a in enumerate(arrs):
                    if not (a == seed * 1000003 + it * 7 + t).all():
                        errors.append('two live arrays share a block')
                del arrs
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=hammer, args=(s,)) for s in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    gc.collect()
    assert not errors, errors[:3]
    assert pool.hits > 0 and len(set(id(b) for b in pool.idle)) == len(pool.idle)
    assert pool.idle_bytes == sum(b.nbytes for b in pool.idle)
