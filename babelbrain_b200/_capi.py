"""ctypes binding of include/babelb200.h.  There is no CPU fallback: if libbabelb200.so is missing
or a call fails, an exception is raised."""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BB_LIB', os.path.join(_HERE, 'libbabelb200.so'))

MAP_NAMES = ['ALLV', 'Vx', 'Vy', 'Vz', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Sigmaxy', 'Sigmaxz', 'Sigmayz', 'Pressure']
MAP_ID = {n: i for i, n in enumerate(MAP_NAMES)}
NCOEF = 8


class BabelB200Error(RuntimeError):
    pass


class FdtdDesc(ctypes.Structure):
    _fields_ = [('n1', ctypes.c_int32), ('n2', ctypes.c_int32), ('n3', ctypes.c_int32),
                ('i0', ctypes.c_int32), ('i1', ctypes.c_int32), ('pml', ctypes.c_int32),
                ('nmat', ctypes.c_int32), ('nsrc', ctypes.c_int32), ('nt_src', ctypes.c_int32),
                ('steps', ctypes.c_int32), ('type_source', ctypes.c_int32), ('sel_rms_peak', ctypes.c_int32),
                ('sel_maps_rms', ctypes.c_uint32), ('sel_maps_sensor', ctypes.c_uint32),
                ('sensor_subsampling', ctypes.c_int32), ('sensor_start', ctypes.c_int32),
                ('device', ctypes.c_int32), ('rank', ctypes.c_int32), ('nranks', ctypes.c_int32),
                ('kernel_variant', ctypes.c_int32), ('mpml_ratio', ctypes.c_float), ('dt', ctypes.c_double)]


class PeerInfo(ctypes.Structure):
    _fields_ = [('pid', ctypes.c_int64), ('device', ctypes.c_int32), ('nown', ctypes.c_int32), ('n2', ctypes.c_int32),
                ('pitch', ctypes.c_int32), ('v_ptr', ctypes.c_uint64), ('s_ptr', ctypes.c_uint64), ('flag_ptr', ctypes.c_uint64),
                ('v_ipc', ctypes.c_ubyte * 64), ('s_ipc', ctypes.c_ubyte * 64), ('flag_ipc', ctypes.c_ubyte * 64)]


class FdtdStats(ctypes.Structure):
    _fields_ = [('run_ms', ctypes.c_double), ('stress_ms', ctypes.c_double), ('particle_ms', ctypes.c_double),
                ('pml_ms', ctypes.c_double), ('other_ms', ctypes.c_double),
                ('stress_launches', ctypes.c_int64), ('particle_launches', ctypes.c_int64),
                ('pml_launches', ctypes.c_int64), ('other_launches', ctypes.c_int64),
                ('steps_done', ctypes.c_int64), ('cells_local', ctypes.c_int64),
                ('device_bytes', ctypes.c_int64), ('nsamples', ctypes.c_int64)]


# every symbol include/babelb200.h declares (checked by tests/test_capi.py)
SYMBOLS = ['bb_last_error', 'bb_version', 'bb_device_count', 'bb_device_name', 'bb_host_alloc', 'bb_host_free', 'bb_host_scatter_rows', 'bb_host_scatter_runs', 'bb_host_nonzero_u32', 'bb_host_lz4_decompress', 'bb_release_cached_memory', 'bb_fdtd_create',
           'bb_fdtd_destroy', 'bb_fdtd_set_stream', 'bb_fdtd_set_materials', 'bb_fdtd_set_maps',
           'bb_fdtd_set_source_cells', 'bb_fdtd_set_source_functions', 'bb_fdtd_set_source_functions_streamed', 'bb_fdtd_set_source_tones', 'bb_fdtd_set_sensors',
           'bb_fdtd_set_sensor_map', 'bb_fdtd_get_sensor_index',
           'bb_nccl_unique_id', 'bb_fdtd_comm_init', 'bb_fdtd_peer_export', 'bb_fdtd_peer_attach', 'bb_fdtd_run', 'bb_fdtd_reset', 'bb_fdtd_get_map',
           'bb_fdtd_get_sensors', 'bb_fdtd_get_sensors_runs', 'bb_fdtd_get_phase_data', 'bb_fdtd_get_stats', 'bb_fdtd_debug_cta_times', 'bb_rayleigh_forward', 'bb_bhte_run']

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BabelB200Error(
                'libbabelb200.so is not built (%s). Run `python -m babelbrain_b200.build`; '
                'there is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.bb_last_error.restype = ctypes.c_char_p
        L.bb_version.restype = ctypes.c_char_p
        vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.bb_device_name.argtypes = [i32, ctypes.c_char_p, i32]
        L.bb_host_alloc.argtypes = [i64, ctypes.POINTER(vp)]
        L.bb_host_free.argtypes = [vp]
        L.bb_host_scatter_rows.argtypes = [vp, vp, vp, i64, i64]
        L.bb_host_scatter_runs.argtypes = [vp, vp, vp, vp, vp, i64, i64]
        L.bb_host_nonzero_u32.argtypes = [vp, i64, vp, vp, i64, vp]
        L.bb_host_lz4_decompress.argtypes = [vp, i64, vp, i64]
        L.bb_host_lz4_decompress.restype = ctypes.c_longlong
        L.bb_fdtd_create.argtypes = [ctypes.POINTER(FdtdDesc), ctypes.POINTER(vp)]
        L.bb_fdtd_destroy.argtypes = [vp]
        L.bb_fdtd_destroy.restype = None
        L.bb_fdtd_set_stream.argtypes = [vp, vp]
        L.bb_fdtd_set_materials.argtypes = [vp, vp, vp]
        L.bb_fdtd_set_maps.argtypes = [vp, vp, vp]
        L.bb_fdtd_set_source_cells.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.bb_fdtd_set_source_functions.argtypes = [vp, vp, i32, i64]
        L.bb_fdtd_set_source_functions_streamed.argtypes = [vp, vp, i32, i64]
        L.bb_fdtd_set_source_tones.argtypes = [vp, vp, vp, vp, vp]
        L.bb_fdtd_set_sensors.argtypes = [vp, i64, vp]
        L.bb_fdtd_set_sensor_map.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_int64)]
        L.bb_fdtd_get_sensor_index.argtypes = [vp, vp, i32]
        L.bb_nccl_unique_id.argtypes = [ctypes.c_char_p]
        L.bb_fdtd_comm_init.argtypes = [vp, ctypes.c_char_p]
        L.bb_fdtd_peer_export.argtypes = [vp, ctypes.POINTER(PeerInfo)]
        L.bb_fdtd_peer_attach.argtypes = [vp, ctypes.POINTER(PeerInfo), ctypes.POINTER(PeerInfo)]
        L.bb_fdtd_run.argtypes = [vp, i64, i32]
        L.bb_fdtd_reset.argtypes = [vp]
        L.bb_fdtd_get_map.argtypes = [vp, i32, i32, vp]
        L.bb_fdtd_get_sensors.argtypes = [vp, i32, vp]
        L.bb_fdtd_get_sensors_runs.argtypes = [vp, i32, vp, i64, vp, vp, vp, i64, vp]
        L.bb_fdtd_get_phase_data.argtypes = [vp, i32, i32, i32, ctypes.c_float, vp, vp, vp]
        L.bb_fdtd_get_stats.argtypes = [vp, ctypes.POINTER(FdtdStats)]
        L.bb_fdtd_debug_cta_times.argtypes = [vp, vp, i64]
        L.bb_rayleigh_forward.argtypes = [ctypes.c_float, ctypes.c_float, i64, vp, vp, vp, i64, vp, vp,
                                          ctypes.c_float, i64, i32, ctypes.POINTER(ctypes.c_double)]
        L.bb_bhte_run.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i64, ctypes.c_float, ctypes.c_float, i32, i32, vp, vp,
                                  i64, vp, i32, ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().bb_last_error().decode('utf-8', 'replace')
        if rc == 1:
            raise ValueError('babelb200: ' + msg)
        raise BabelB200Error('babelb200 error %d: %s' % (rc, msg))


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def device_count():
    n = lib().bb_device_count()
    if n < 0:
        raise BabelB200Error(lib().bb_last_error().decode())
    return n


def device_names():
    out = []
    for d in range(device_count()):
        buf = ctypes.create_string_buffer(256)
        check(lib().bb_device_name(d, buf, 256))
        out.append(buf.value.decode())
    return out


def require_gpu():
    if device_count() == 0:
        raise BabelB200Error('no CUDA device visible: babelbrain_b200 has no CPU fallback')


class _PinnedBlock:
    """One page-locked host allocation; returns itself to the pool when the last numpy view dies."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        check(lib().bb_host_alloc(int(nbytes), ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)

    def release(self):
        if self.ptr:
            lib().bb_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


class PinnedPool:
    """Result buffers (RMS maps, sensor traces, index tables) handed to the caller as ordinary writable
    numpy arrays backed by page-locked memory.  A block goes back to the pool when the caller drops the
    array, so a worker that runs several simulations (forward, back-propagation, refocus:
    BabelIntegrationBASE.py:2338-2428) pays the page-locking once."""

    def __init__(self, max_idle_bytes=int(float(os.environ.get('BB_PINNED_POOL_GB', 32)) * (1 << 30)), block_type=None):
        import threading
        self.block_type = block_type or _PinnedBlock     # tests substitute an ordinary-memory block
        self.idle, self.max_idle, self.idle_bytes = [], max_idle_bytes, 0
        self.hits = self.misses = 0
        # the per-GPU threads of a slab-decomposed run allocate concurrently, and finalizers fire from any thread
        self._lock = threading.Lock()

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        with self._lock:
            blk = None
            for b in self.idle:               # smallest idle block that fits without wasting more than 2x
                if n <= b.nbytes <= max(2 * n, 1 << 20) and (blk is None or b.nbytes < blk.nbytes):
                    blk = b
            if blk is not None:
                self.idle.remove(blk)         # by identity (_PinnedBlock defines no __eq__)
                self.idle_bytes -= blk.nbytes
                self.hits += 1
            else:
                self.misses += 1
        if blk is None:
            blk = self.block_type(max(n, 16))
        buf = (ctypes.c_char * blk.nbytes).from_address(blk.ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
        weakref.finalize(buf, self._give_back, blk)   # buf lives as long as any view of it
        return arr

    def _give_back(self, blk):
        with self._lock:
            keep = self.idle_bytes + blk.nbytes <= self.max_idle
            if keep:
                self.idle.append(blk)
                self.idle_bytes += blk.nbytes
        if not keep:
            blk.release()


pinned = PinnedPool()
