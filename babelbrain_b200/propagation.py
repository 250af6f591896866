"""
Host-side mirror of ``BabelViscoFDTD.PropagationModel.PropagationModel`` for the path BabelBrain
drives (TranscranialModeling/BabelIntegrationBASE.py:43 ``PModel=PropagationModel()``;
:1799/:1801 ``CalculateMatricesForPropagation``; :2338/:2374/:2401
``StaggeredFDTD_3D_with_relaxation``).  Same names, argument meaning, return tuples and error
behaviour; the time loop runs in libbabelb200.so (CUDA, sm_100a) through the C ABI of
include/babelb200.h.  No CPU fallback: without the library or a GPU the call raises.
"""
import collections.abc
import ctypes
import os
import time
import weakref
import numpy as np

from . import _capi, hostprep
from .slab import SlabPlan
from .sources import CWSourceFunctions

_live_lastmaps = []


class FdtdSlab:
    """One slab [i0,i1) of a simulation on one GPU (the whole domain when nranks == 1)."""

    def __init__(self, MaterialMap, MaterialProperties, Frequency, SourceMap, SourceFunctions, SpatialStep,
                 DurationSimulation, SensorMap, Ox=np.array([1]), Oy=np.array([1]), Oz=np.array([1]),
                 AlphaCFL=0.99, NDelta=12, ReflectionLimit=1.0e-5, DT=None, QfactorCorrection=True,
                 QCorrection=1.0, TypeSource=0, SelRMSorPeak=1, SelMapsRMSPeakList=('ALLV',),
                 SelMapsSensorsList=('Vx', 'Vy', 'Vz'), SensorSubSampling=2, SensorStart=0,
                 ReflectorMask=None, device=0, rank=0, nranks=1, kernel_variant=0, steps=None,
                 origin=None, n1_global=None, global_sensor_table=True, MPMLRatio=None, stream_sources=False):
        self._h = None
        self._sf_keepalive = None
        _t_init = time.perf_counter()
        if not isinstance(MaterialMap, np.ndarray) or MaterialMap.ndim != 3:
            raise ValueError('MaterialMap must be a 3-D numpy array')
        for name, arr in (('MaterialMap', MaterialMap), ('SourceMap', SourceMap), ('SensorMap', SensorMap)):
            if not isinstance(arr, np.ndarray) or arr.dtype != np.uint32:
                raise TypeError('%s must be a numpy array of dtype uint32' % name)
            if arr.shape != MaterialMap.shape:
                raise ValueError('%s must have the shape of MaterialMap' % name)
        MP = np.atleast_2d(np.asarray(MaterialProperties, dtype=np.float64))
        # the volumes either cover the whole grid (origin None) or only this rank's planes
        # [origin, origin + shape[0]) = SlabPlan.with_halo(rank) of a grid with n1_global planes
        N1 = int(MaterialMap.shape[0] if origin is None else n1_global)
        N2, N3 = MaterialMap.shape[1:]
        org = 0 if origin is None else int(origin)
        self.shape = (N1, N2, N3)
        h = float(SpatialStep)
        table, self.analysis = hostprep.material_table(MP, Frequency, QfactorCorrection, h, QCorrection)
        dt_ideal = hostprep.stable_dt(MP, h, AlphaCFL)
        if DT is None:
            dt = dt_ideal
        else:
            dt = float(DT)
            # the reference only warns when the manual step exceeds the one it would choose; beyond the stability
            # limit of the scheme itself the run can only diverge, which is reported before any device work
            if dt > hostprep.hard_limit_dt(MP, h) * (1.0 + 1e-9):
                raise ValueError('Staggered:DT_INVALID: DT %g is above the stability limit %g of the O(2,4) scheme'
                                 % (dt, hostprep.hard_limit_dt(MP, h)))
            if dt > dt_ideal * (1.0 + 1e-9):
                import warnings
                warnings.warn('Staggered:DT_INVALID The specified manual step is larger than the minimal optimal size, '
                              'there is a risk of unstable calculation %g,%g' % (dt, dt_ideal))
        self.dt = dt
        self.steps = hostprep.number_of_steps(DurationSimulation, dt) if steps is None else int(steps)
        self.sub = int(SensorSubSampling)
        self.sensor_start = int(SensorStart)
        self.sample_steps = hostprep.sample_steps(self.steps, self.sub, self.sensor_start)
        self.rms_names = [n for n in hostprep.MAP_NAMES if n in SelMapsRMSPeakList]
        self.sensor_names = [n for n in hostprep.MAP_NAMES if n in SelMapsSensorsList]
        hostprep.maps_mask(SelMapsRMSPeakList), hostprep.maps_mask(SelMapsSensorsList)
        self.mpml_ratio = float(hostprep.MPML_RATIO if MPMLRatio is None else MPMLRatio)
        if not 0.0 <= self.mpml_ratio <= 1.0:
            raise ValueError('MPMLRatio must lie in [0, 1] (0 = classical split-field layer)')
        self.sel_rms_peak = int(SelRMSorPeak)
        if self.sel_rms_peak not in (1, 2, 3):
            raise ValueError('SelRMSorPeak must be 1 (RMS), 2 (peak) or 3 (both)')
        cw = SourceFunctions if isinstance(SourceFunctions, CWSourceFunctions) else None
        SF = SourceFunctions if cw is not None else np.asarray(SourceFunctions)
        if SF.ndim != 2:
            raise ValueError('SourceFunctions must be (Nsources, Ntime)')
        if cw is None and (SF.dtype not in (np.float64, np.float32) or SF.strides[1] != SF.itemsize):
            SF = np.ascontiguousarray(SF, dtype=np.float64)
        self.plan = SlabPlan(N1, nranks, NDelta)
        self.rank, self.nranks = int(rank), int(nranks)
        i0, i1 = self.plan.owned(rank)
        self.i0, self.i1 = i0, i1
        glo, ghi = self.plan.with_halo(rank)
        if origin is not None and (org != glo or MaterialMap.shape[0] != ghi - glo):
            raise ValueError('local volumes must cover planes [%d,%d) of the global grid' % (glo, ghi))
        top = int(MaterialMap[glo - org:ghi - org].max())      # this slab's planes only: every rank of a multi-GPU run checks its own
        if top >= MP.shape[0]:
            raise ValueError('MaterialMap holds label %d but MaterialProperties has %d rows' % (top, MP.shape[0]))

        # ---- sources owned by this slab (global C-order cell index; ids are 1-based rows)
        sm_slab = SourceMap[i0 - org:i1 - org]
        flat, ids = _nonzero_u32(sm_slab)
        rows = ids.astype(np.int64) - 1
        if flat.size and rows.max() >= SF.shape[0]:
            raise ValueError('SourceMap refers to source %d but SourceFunctions has %d rows' % (rows.max() + 1, SF.shape[0]))
        cells = flat.astype(np.int64) + np.int64(i0) * N2 * N3

        where = np.unravel_index(flat, sm_slab.shape)

        def weights(O):
            O = np.asarray(O)
            if O.size == 1:
                return np.full(cells.shape, float(O.reshape(-1)[0]), np.float32)
            if O.shape != MaterialMap.shape:
                raise ValueError('Ox/Oy/Oz must be single values or volumes of the MaterialMap shape')
            return np.ascontiguousarray(O[i0 - org:i1 - org][where], dtype=np.float32)   # no copy of a strided volume
        ox, oy, oz = weights(Ox), weights(Oy), weights(Oz)

        # ---- sensors: IndexSensorMap is the 1-based Fortran-order linear index (BASE.py:2503-2511).
        # With one slab, or slab-local volumes, the table is built on the device from the SensorMap planes
        # (bb_fdtd_set_sensor_map); only a multi-rank run fed with whole volumes needs the global table
        # on the host to place its rows.
        idx_dtype = np.uint32 if N1 * N2 * N3 < 2 ** 32 else np.uint64
        self._idx_dtype = idx_dtype
        self._sensor_planes = None
        if self.nranks > 1 and origin is None and global_sensor_table:
            sl, sj, sk = np.nonzero(SensorMap[i0 - org:i1 - org])
            findex = (sl.astype(np.int64) + i0) + sj.astype(np.int64) * N1 + sk.astype(np.int64) * N1 * N2
            order = np.argsort(findex, kind='stable')
            findex = findex[order]
            scell = (((sl[order].astype(np.int64) + i0) * N2 + sj[order]) * N3 + sk[order]).astype(np.int64)
            self.IndexSensorMapLocal = (findex + 1).astype(idx_dtype)
            fl = np.flatnonzero(SensorMap.reshape(-1, order='F'))
            self.IndexSensorMap = (fl + 1).astype(idx_dtype)
            self.sensor_rows = np.searchsorted(fl, findex)
            self.nsensors_total = self.IndexSensorMap.size
        else:
            scell = None
            self._sensor_planes = np.ascontiguousarray(SensorMap[i0 - org:i1 - org])

        _capi.require_gpu()
        if isinstance(device, tuple):       # (name substring, ordinal): resolved only now, after the argument checks
            device = _select_device(*device)
        L = _capi.lib()
        self._L = L
        d = _capi.FdtdDesc(n1=N1, n2=N2, n3=N3, i0=i0, i1=i1, pml=int(NDelta), nmat=MP.shape[0],
                           nsrc=SF.shape[0], nt_src=SF.shape[1], steps=self.steps, type_source=int(TypeSource),
                           sel_rms_peak=self.sel_rms_peak, sel_maps_rms=hostprep.maps_mask(self.rms_names),
                           sel_maps_sensor=hostprep.maps_mask(self.sensor_names),
                           sensor_subsampling=self.sub, sensor_start=self.sensor_start, device=int(device),
                           rank=self.rank, nranks=self.nranks, kernel_variant=int(kernel_variant), mpml_ratio=self.mpml_ratio,
                           dt=dt)
        _t = [time.perf_counter()]
        _marks = ['host prep %.3f' % (_t[0] - _t_init)]

        def _mark(name):          # host-side wall clock of the set-up stages, printed when BB_TIMING is set
            _t.append(time.perf_counter())
            _marks.append('%s %.3f' % (name, _t[-1] - _t[-2]))
        hp = ctypes.c_void_p()
        _capi.check(L.bb_fdtd_create(ctypes.byref(d), ctypes.byref(hp)))
        _mark('create')
        self._h = hp
        self._finalizer = weakref.finalize(self, L.bb_fdtd_destroy, hp)
        t32 = np.ascontiguousarray(table, dtype=np.float32)
        pml = np.ascontiguousarray(hostprep.pml_table(NDelta, h, dt, MP[:, 1].max(), ReflectionLimit), dtype=np.float32)
        _capi.check(L.bb_fdtd_set_materials(hp, _capi.ptr(t32), _capi.ptr(pml)))
        mm = np.ascontiguousarray(MaterialMap[glo - org:ghi - org])
        refl = None
        if ReflectorMask is not None:
            RM = np.asarray(ReflectorMask)
            if RM.shape != MaterialMap.shape:
                raise ValueError('ReflectorMask must have the shape of MaterialMap')
            refl = np.ascontiguousarray(RM[glo - org:ghi - org], dtype=np.uint32)
        _capi.check(L.bb_fdtd_set_maps(hp, _capi.ptr(mm), _capi.ptr(refl)))
        _mark('maps')
        rows32 = np.ascontiguousarray(rows, dtype=np.int32)
        _capi.check(L.bb_fdtd_set_source_cells(hp, cells.size, _capi.ptr(cells), _capi.ptr(rows32),
                                               _capi.ptr(ox), _capi.ptr(oy), _capi.ptr(oz)))
        sf_bytes = 0
        if cells.size and cw is not None:        # continuous-wave rows evaluated in the source kernel: no table anywhere
            tones = cw.tone_tables()
            _capi.check(L.bb_fdtd_set_source_tones(hp, *[_capi.ptr(a) for a in tones]))
            sf_bytes = sum(a.nbytes for a in tones)
        elif cells.size:
            # stream_sources: the table goes up in chunks of time samples while the time loop runs (the public call; SF is
            # kept alive here until the slab is closed); otherwise at once, so that a timed run() starts with its inputs resident
            setter = L.bb_fdtd_set_source_functions_streamed if stream_sources else L.bb_fdtd_set_source_functions
            _capi.check(setter(hp, _capi.ptr(SF), int(SF.dtype == np.float64), SF.strides[0] // SF.itemsize))
            self._sf_keepalive = SF if stream_sources else None
            sf_bytes = SF.nbytes
        _mark('sources')
        self.d2h_bytes = 0
        if scell is not None:
            _capi.check(L.bb_fdtd_set_sensors(hp, scell.size, _capi.ptr(scell)))
            sensor_bytes = scell.nbytes
        else:
            ns = ctypes.c_int64()
            _capi.check(L.bb_fdtd_set_sensor_map(hp, _capi.ptr(self._sensor_planes), ctypes.byref(ns)))
            sensor_bytes = self._sensor_planes.nbytes
            self._sensor_planes = None
            idx = _capi.pinned.empty((ns.value,), self._idx_dtype)
            _capi.check(L.bb_fdtd_get_sensor_index(hp, _capi.ptr(idx), idx.itemsize))
            self.d2h_bytes += idx.nbytes
            self.IndexSensorMap = self.IndexSensorMapLocal = idx
            self.sensor_rows = np.arange(ns.value)
            self.nsensors_total = ns.value
        _mark('sensors')
        if os.environ.get('BB_TIMING'):
            print('FdtdSlab setup: ' + ', '.join(_marks), flush=True)
        self.h2d_bytes = int(mm.nbytes + (refl.nbytes if refl is not None else 0) + sf_bytes
                             + cells.nbytes + rows32.nbytes + 3 * ox.nbytes + sensor_bytes + t32.nbytes + pml.nbytes)

    # ------------------------------------------------------------------
    def comm_init(self, unique_id=None):
        """Join the slab communicator; unique_id None re-attaches the one an earlier FdtdSlab of this process
        created for the same (device, rank, nranks)."""
        _capi.check(self._L.bb_fdtd_comm_init(self._h, unique_id))

    def peer_export(self):
        """Descriptor of this slab for its neighbours (bytes; may cross process boundaries)."""
        info = _capi.PeerInfo()
        _capi.check(self._L.bb_fdtd_peer_export(self._h, ctypes.byref(info)))
        return bytes(info)

    def peer_attach(self, lower, upper):
        """NVLink halo push: lower / upper = peer_export() of the ranks below / above (None at the ends).
        Replaces comm_init: no NCCL call remains in the time loop."""
        lo = _capi.PeerInfo.from_buffer_copy(lower) if lower is not None else None
        up = _capi.PeerInfo.from_buffer_copy(upper) if upper is not None else None
        _capi.check(self._L.bb_fdtd_peer_attach(self._h, ctypes.byref(lo) if lo is not None else None,
                                                ctypes.byref(up) if up is not None else None))

    @staticmethod
    def nccl_unique_id():
        buf = ctypes.create_string_buffer(128)
        _capi.check(_capi.lib().bb_nccl_unique_id(buf))
        return buf.raw

    def set_stream(self, cuda_stream_ptr):
        _capi.check(self._L.bb_fdtd_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def run(self, nsteps=-1, profile=False):
        _capi.check(self._L.bb_fdtd_run(self._h, int(nsteps), int(bool(profile))))
        if int(nsteps) < 0:
            self._sf_keepalive = None        # a run to the last step has streamed the whole source table
        return self.stats()

    def reset(self):
        _capi.check(self._L.bb_fdtd_reset(self._h))

    def stats(self):
        st = _capi.FdtdStats()
        _capi.check(self._L.bb_fdtd_get_stats(self._h, ctypes.byref(st)))
        return {k: getattr(st, k) for k, _ in _capi.FdtdStats._fields_}

    def get_map(self, which, name, out=None):
        """which: 0 RMS, 1 peak, 2 last field.  Returns / fills the (i1-i0, N2, N3) float32 slab."""
        N1, N2, N3 = self.shape
        if out is None:
            out = _capi.pinned.empty((self.i1 - self.i0, N2, N3), np.float32)
        assert out.flags.c_contiguous and out.dtype == np.float32
        _capi.check(self._L.bb_fdtd_get_map(self._h, int(which), _capi.MAP_ID[name], _capi.ptr(out)))
        self.d2h_bytes += out.nbytes
        return out

    def get_sensors(self, name):
        """(sensors of this slab, samples) float32 traces of one selected sensor map, in table order."""
        t0 = time.perf_counter()
        out = _capi.pinned.empty((self.sensor_rows.size, self.sample_steps.size), np.float32)
        t1 = time.perf_counter()
        _capi.check(self._L.bb_fdtd_get_sensors(self._h, _capi.MAP_ID[name], _capi.ptr(out)))
        if os.environ.get('BB_TIMING'):
            print('get_sensors: alloc %.3f s copy %.3f s (pool hits %d misses %d)'
                  % (t1 - t0, time.perf_counter() - t1, _capi.pinned.hits, _capi.pinned.misses), flush=True)
        self.d2h_bytes += out.nbytes
        return out

    def get_phase_data(self, name, bin_index, nsamples_used=None, scale=None, fourier=None, phase=None, peak=None):
        """Single-bin DFT, angle and largest sample of every sensor trace of this slab, computed on the device and
        scattered to (i1-i0, N2, N3) volumes (bb_fdtd_get_phase_data).  Arrays passed in are filled in place."""
        N1, N2, N3 = self.shape
        sh = (self.i1 - self.i0, N2, N3)
        n = int(self.sample_steps.size if nsamples_used is None else nsamples_used)
        if fourier is None:
            fourier = _capi.pinned.empty(sh, np.complex64)
        if phase is None:
            phase = _capi.pinned.empty(sh, np.float32)
        if peak is None:
            peak = _capi.pinned.empty(sh, np.float32)
        for a, dt in ((fourier, np.complex64), (phase, np.float32), (peak, np.float32)):
            assert a.shape == sh and a.dtype == dt and a.flags.c_contiguous
        _capi.check(self._L.bb_fdtd_get_phase_data(self._h, _capi.MAP_ID[name], int(bin_index), n,
                                                   ctypes.c_float(2.0 / n if scale is None else scale),
                                                   _capi.ptr(fourier), _capi.ptr(phase), _capi.ptr(peak)))
        self.d2h_bytes += fourier.nbytes + phase.nbytes + peak.nbytes
        return fourier, phase, peak

    def close(self):
        if self._h is not None:
            self._finalizer()
            self._h = None


class _LastMap(collections.abc.Mapping):
    """LastMap of the reference call, fetched from the device on first access of a key.  The caller
    discards it (BabelIntegrationBASE.py:2338 binds it to a throw-away local), so nothing is copied
    unless it is read.  Only the most recent simulation keeps its device state."""
    _keys = hostprep.MAP_NAMES[1:]

    def __init__(self, slab):
        self._slab = slab
        self._cache = {}

    def __getitem__(self, k):
        if k not in self._keys:
            raise KeyError(k)
        if k not in self._cache:
            if self._slab is None:
                raise RuntimeError('LastMap of an earlier simulation was released when a newer one started')
            self._cache[k] = self._slab.get_map(2, k)
        return self._cache[k]

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def _release(self):
        if self._slab is not None:
            self._slab.close()
            self._slab = None


class _LastMapSlabs(_LastMap):
    """LastMap of a slab-decomposed run: a key is assembled from every slab on first access."""

    def __init__(self, slabs):
        self._slabs, self._cache = list(slabs), {}

    def __getitem__(self, k):
        if k not in self._keys:
            raise KeyError(k)
        if k not in self._cache:
            if not self._slabs:
                raise RuntimeError('LastMap of an earlier simulation was released when a newer one started')
            out = _capi.pinned.empty(self._slabs[0].shape, np.float32)
            for s in self._slabs:
                s.get_map(2, k, out=out[s.i0:s.i1])
            self._cache[k] = out
        return self._cache[k]

    def _release(self):
        for s in self._slabs:
            s.close()
        self._slabs = []


def _nonzero_u32(volume):
    """(flat indices, values) of the nonzero entries of a C-contiguous uint32 volume, in order: np.flatnonzero takes 40-60 ms
    on the CTX-500 SourceMap (18.4 M cells, 46 656 sources), bb_host_nonzero_u32 scans it on a few threads in ~3."""
    a = np.ascontiguousarray(volume).reshape(-1)
    L = _capi.lib()
    cap = max(1024, a.size // 256)
    while True:
        idx, val = np.empty(cap, np.int64), np.empty(cap, np.uint32)
        n = ctypes.c_int64(0)
        _capi.check(L.bb_host_nonzero_u32(_capi.ptr(a), a.size, _capi.ptr(idx), _capi.ptr(val), cap, ctypes.byref(n)))
        if n.value <= cap:
            return idx[:n.value], val[:n.value]
        cap = n.value


def run_slabs_in_process(devices, args, kwargs, timeout=None):
    """One simulation cut into len(devices) slabs along axis 0, one thread and one bb_fdtd handle per GPU of
    this process, halos pushed over NVLink from the boundary CTAs (bb_fdtd_peer_attach).  The volumes in
    `args` / `kwargs` are the caller's whole-grid arrays; every thread uploads only its planes.  Returns
    (Sensor, RMS, Peak, InputParam, slabs, timing) with the maps and sensor rows of all slabs gathered into
    whole-grid arrays in the reference's order."""
    import threading
    from .slab import merge_sensor_runs
    n = len(devices)
    gate = threading.Barrier(n, timeout=timeout)
    slabs, exports, errors = [None] * n, [None] * n, []
    shared = {}
    marks = [dict() for _ in range(n)]

    def stage(r, body):
        try:
            body(r)
        except threading.BrokenBarrierError:
            pass
        except BaseException as e:  # noqa: BLE001 -- re-raised in the calling thread
            errors.append(e)
            gate.abort()

    def rank_main(r):
        t0 = time.perf_counter()
        s = slabs[r] = FdtdSlab(*args, device=devices[r], rank=r, nranks=n, global_sensor_table=False, **kwargs)
        exports[r] = s.peer_export()
        gate.wait()
        s.peer_attach(exports[r - 1] if r > 0 else None, exports[r + 1] if r < n - 1 else None)
        gate.wait()
        t1 = time.perf_counter()
        prep = None
        if r == 0:                      # whole-grid result arrays, page-locked, filled by every thread after the time loop
            def prepare():              # ... prepared while the GPUs run it (0.15 s of host work for eight CTX-500 slabs)
                try:
                    N1, N2, N3 = s.shape
                    # the slabs' tables merge into the whole-grid IndexSensorMap by runs (one per (j,k) line and slab)
                    ntot, runs = merge_sensor_runs([x.IndexSensorMapLocal for x in slabs], N1, N2 * N3)
                    shared['runs'] = runs
                    shared['index'] = _capi.pinned.empty((ntot,), s._idx_dtype)
                    shared['sensor'] = {k: _capi.pinned.empty((ntot, s.sample_steps.size), np.float32) for k in s.sensor_names}
                    shared['rms'] = {k: _capi.pinned.empty(s.shape, np.float32) for k in s.rms_names if s.sel_rms_peak & 1}
                    shared['peak'] = {k: _capi.pinned.empty(s.shape, np.float32) for k in s.rms_names if s.sel_rms_peak & 2}
                except BaseException as e:  # noqa: BLE001 -- re-raised in the calling thread
                    shared['error'] = e
            prep = threading.Thread(target=prepare, name='bb_gather_prep')
            prep.start()
        try:
            s.run()
        finally:
            if prep is not None:
                prep.join()
        if 'error' in shared:
            raise shared['error']
        gate.wait()                     # every slab has finished writing into its neighbours, and the result arrays exist
        t2 = time.perf_counter()
        for k, full in shared['rms'].items():
            s.get_map(0, k, out=full[s.i0:s.i1])
        for k, full in shared['peak'].items():
            s.get_map(1, k, out=full[s.i0:s.i1])
        dst, src, cnt = shared['runs'][r]

        def place(full, part):           # rows of this slab -> their runs in the whole-grid table (memcpy per run, no GIL)
            part = np.ascontiguousarray(part)
            row_bytes = part.itemsize * (part.shape[1] if part.ndim > 1 else 1)
            _capi.check(s._L.bb_host_scatter_runs(_capi.ptr(full), _capi.ptr(part), _capi.ptr(dst), _capi.ptr(src), _capi.ptr(cnt),
                                                  dst.size, row_bytes))
        place(shared['index'], s.IndexSensorMapLocal)
        for k, full in shared['sensor'].items():
            # the GPU transposes the slab's traces and writes them into their runs of the page-locked whole-grid table itself;
            # a pageable table (pool cap reached) takes the staged download + host placement instead
            done = ctypes.c_int(0)
            _capi.check(s._L.bb_fdtd_get_sensors_runs(s._h, _capi.MAP_ID[k], _capi.ptr(full), full.shape[0], _capi.ptr(dst), _capi.ptr(src),
                                                      _capi.ptr(cnt), dst.size, ctypes.byref(done)))
            if done.value:
                s.d2h_bytes += s.sensor_rows.size * full.shape[1] * 4
            else:
                place(full, s.get_sensors(k))     # (rows of this slab, samples), page-locked staging
        gate.wait()
        marks[r].update(setup_upload_s=t1 - t0, time_loop_s=t2 - t1, download_s=time.perf_counter() - t2)

    threads = [threading.Thread(target=stage, args=(r, rank_main), name='bb_slab_%d' % r) for r in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        for s in slabs:
            if s is not None:
                s.close()
        raise errors[0]
    s0 = slabs[0]
    N1, N2, N3 = s0.shape
    Sensor = {'time': s0.sample_steps.astype(np.float64) * s0.dt}
    Sensor.update(shared['sensor'])
    InputParam = {'IndexSensorMap': shared['index'], 'DT': s0.dt, 'N1': N1, 'N2': N2, 'N3': N3, 'TimeSteps': s0.steps,
                  'SensorSubSampling': s0.sub, 'SensorStart': s0.sensor_start}
    timing = {k: max(m[k] for m in marks) for k in marks[0]}
    timing['h2d_bytes'] = sum(s.h2d_bytes for s in slabs)
    timing['d2h_bytes'] = sum(s.d2h_bytes for s in slabs)
    return Sensor, shared['rms'], shared['peak'], InputParam, slabs, timing


def collect_results(slab):
    """Assemble the reference's return values from a finished single-slab run."""
    N1, N2, N3 = slab.shape
    Sensor = {'time': slab.sample_steps.astype(np.float64) * slab.dt}
    for name in slab.sensor_names:
        Sensor[name] = slab.get_sensors(name)
    RMS, Peak = {}, {}
    for name in slab.rms_names:
        if slab.sel_rms_peak & 1:
            RMS[name] = slab.get_map(0, name)
        if slab.sel_rms_peak & 2:
            Peak[name] = slab.get_map(1, name)
    InputParam = {'IndexSensorMap': slab.IndexSensorMap, 'DT': slab.dt, 'N1': N1, 'N2': N2, 'N3': N3,
                  'TimeSteps': slab.steps, 'SensorSubSampling': slab.sub, 'SensorStart': slab.sensor_start}
    return Sensor, RMS, Peak, InputParam


def release_device_state(purge_cache=False):
    """Free the device memory the most recent simulation still holds for LastMap / CalculatePhaseDataOnDevice (it is
    otherwise released when the next simulation starts or the process ends).  The arrays go back to the library's
    per-device cache for the next simulation of the same grid; purge_cache=True returns them to the driver."""
    for v in _live_lastmaps:
        v._release()
    del _live_lastmaps[:]
    if purge_cache:
        _capi.check(_capi.lib().bb_release_cached_memory(-1))


class PropagationModel:
    """Drop-in for BabelViscoFDTD.PropagationModel.PropagationModel (the two methods BabelBrain uses), plus
    CalculatePhaseDataOnDevice."""

    _last_slabs = ()
    ReleaseDeviceState = staticmethod(release_device_state)

    def CalculatePhaseDataOnDevice(self, Frequency, MapName='Pressure'):
        """What the caller's CalculatePhaseData (BabelIntegrationBASE.py:2460-2518, forward branch) derives on the
        host from Sensor[MapName] and InputParam['IndexSensorMap'] -- computed on the GPU(s) from the traces of the
        most recent simulation of this object, which are still resident: the DFT bin closest to Frequency of every
        sensor (times 2/Nsamples), its angle, and the largest sample, as whole-grid volumes with zeros where there
        is no sensor.  Returns {'PhaseMap': float32, 'PressMapFourier': complex64, 'PressMapPeak': float32,
        'IndSpectrum': int}.  Nothing but those three volumes crosses PCIe, and no host FFT runs."""
        slabs = [s for s in self._last_slabs if s._h is not None]
        if not slabs:
            raise RuntimeError('no finished simulation of this PropagationModel holds device state '
                               '(a newer simulation releases the previous one)')
        s0 = slabs[0]
        if MapName not in s0.sensor_names:
            raise ValueError('%s was not in SelMapsSensorsList' % MapName)
        t = s0.sample_steps.astype(np.float64) * s0.dt
        if t.size < 2:
            raise ValueError('phase data need at least two sensor samples')
        freqs = np.fft.fftfreq(t.size, np.diff(t).mean())           # BabelIntegrationBASE.py:2489, :2498-2499
        ind = int(np.argmin(np.abs(freqs - Frequency)))
        fourier = _capi.pinned.empty(s0.shape, np.complex64)
        phase = _capi.pinned.empty(s0.shape, np.float32)
        peak = _capi.pinned.empty(s0.shape, np.float32)

        def one(s):
            s.get_phase_data(MapName, ind, fourier=fourier[s.i0:s.i1], phase=phase[s.i0:s.i1], peak=peak[s.i0:s.i1])
        if len(slabs) == 1:
            one(s0)
        else:
            import threading
            errors = []

            def guarded(s):
                try:
                    one(s)
                except BaseException as e:  # noqa: BLE001
                    errors.append(e)
            th = [threading.Thread(target=guarded, args=(s,)) for s in slabs]
            for x in th:
                x.start()
            for x in th:
                x.join()
            if errors:
                raise errors[0]
        return {'PhaseMap': phase, 'PressMapFourier': fourier, 'PressMapPeak': peak, 'IndSpectrum': ind}

    def CalculateMatricesForPropagation(self, MaterialMap, MaterialProperties, Frequency, QfactorCorrection, h,
                                        AlphaCFL, QCorrection=1.0):
        """Returns the 10-tuple whose element 0 is the stable time step -- the only element the
        caller reads (BabelIntegrationBASE.py:1799,1801).  The others are the per-material arrays
        (rho, mu, lambda+2mu, lambda, tau_long, tau_shear, tau_sigma, Q_long, Q_shear)."""
        MP = np.atleast_2d(np.asarray(MaterialProperties, dtype=np.float64))
        T, A = hostprep.material_table(MP, Frequency, QfactorCorrection, h, QCorrection)
        dt = hostprep.stable_dt(MP, h, AlphaCFL)
        with np.errstate(divide='ignore'):
            tau_sigma = np.where(T[:, 6] > 0, 1.0 / np.where(T[:, 6] > 0, T[:, 6], 1.0), 0.0)
        return (dt, MP[:, 0].copy(), T[:, 1] * h, T[:, 0] * h, T[:, 2] * h, T[:, 4].copy(), T[:, 5].copy(),
                tau_sigma, A['QL'], A['QS'])

    def StaggeredFDTD_3D_with_relaxation(self, MaterialMap, MaterialProperties, Frequency, SourceMap,
                                         SourceFunctions, SpatialStep, DurationSimulation, SensorMap,
                                         Ox=np.array([1]), Oy=np.array([1]), Oz=np.array([1]),
                                         AlphaCFL=0.99, NDelta=12, ReflectionLimit=1.0000e-05,
                                         IntervalSnapshots=-1, COMPUTING_BACKEND=1, USE_SINGLE=True,
                                         SPP_ZONES=1, SPP_VolumeFraction=None, DT=None,
                                         QfactorCorrection=True, QCorrection=1.0, CheckOnlyParams=False,
                                         TypeSource=0, SelRMSorPeak=1, SelMapsRMSPeakList=['ALLV'],
                                         SelMapsSensorsList=['Vx', 'Vy', 'Vz'], SensorSubSampling=2,
                                         SensorStart=0, DefaultGPUDeviceName='B200', DefaultGPUDeviceNumber=0,
                                         ReflectorMask=None, SILENT=0, ManualGroupSize=None, ManualLocalSize=None,
                                         NumberGPUs=None, MPMLRatio=None, **_ignored):
        """Same call as the reference.  COMPUTING_BACKEND, DefaultGPUDeviceName, USE_SINGLE and the
        manual work-group sizes are accepted for compatibility; every backend value runs the
        sm_100a CUDA path in float32 (there is no multi-backend dispatch).

        NumberGPUs (extension; default: environment variable BABELB200_NGPUS, else 1) cuts the domain into that
        many slabs along axis 0, one per GPU of this box, with NVLink halo exchange; the return values are the
        same whole-grid arrays.  An unmodified BabelBrain enables it through the environment variable.

        MPMLRatio (extension; default 0 = the classical split-field layer of the reference scheme): > 0 selects the
        multi-axial layer for label maps that carry fluid-solid interfaces into the absorbing shell (hostprep.py)."""
        t0 = time.perf_counter()
        if IntervalSnapshots > 0:
            raise NotImplementedError('IntervalSnapshots is not supported (BabelBrain never passes it)')
        if SPP_ZONES != 1:
            raise NotImplementedError('superposition zones (SPP_ZONES>1) are not supported')
        release_device_state()
        if os.environ.get('BB_TIMING'):
            print('release of the previous simulation: %.3f s' % (time.perf_counter() - t0), flush=True)
        ngpu = int(NumberGPUs if NumberGPUs is not None else os.environ.get('BABELB200_NGPUS', 1))
        if ngpu < 1:
            raise ValueError('NumberGPUs must be >= 1')
        if ngpu > 1 and not CheckOnlyParams:
            return self._run_multi_gpu(ngpu, t0, (MaterialMap, MaterialProperties, Frequency, SourceMap, SourceFunctions,
                                                  SpatialStep, DurationSimulation, SensorMap),
                                       dict(Ox=Ox, Oy=Oy, Oz=Oz, AlphaCFL=AlphaCFL, NDelta=NDelta, ReflectionLimit=ReflectionLimit,
                                            DT=DT, QfactorCorrection=QfactorCorrection, QCorrection=QCorrection,
                                            TypeSource=TypeSource, SelRMSorPeak=SelRMSorPeak,
                                            SelMapsRMSPeakList=SelMapsRMSPeakList, SelMapsSensorsList=SelMapsSensorsList,
                                            SensorSubSampling=SensorSubSampling, SensorStart=SensorStart,
                                            ReflectorMask=ReflectorMask, MPMLRatio=MPMLRatio, stream_sources=True), DefaultGPUDeviceName, DefaultGPUDeviceNumber)
        slab = FdtdSlab(MaterialMap, MaterialProperties, Frequency, SourceMap, SourceFunctions, SpatialStep,
                        DurationSimulation, SensorMap, Ox=Ox, Oy=Oy, Oz=Oz, AlphaCFL=AlphaCFL, NDelta=NDelta,
                        ReflectionLimit=ReflectionLimit, DT=DT, QfactorCorrection=QfactorCorrection,
                        QCorrection=QCorrection, TypeSource=TypeSource, SelRMSorPeak=SelRMSorPeak,
                        SelMapsRMSPeakList=SelMapsRMSPeakList, SelMapsSensorsList=SelMapsSensorsList,
                        SensorSubSampling=SensorSubSampling, SensorStart=SensorStart, ReflectorMask=ReflectorMask,
                        MPMLRatio=MPMLRatio, stream_sources=True, device=(DefaultGPUDeviceName, DefaultGPUDeviceNumber))
        if CheckOnlyParams:
            slab.close()
            return None
        t1 = time.perf_counter()
        slab.run()
        t2 = time.perf_counter()
        self.last_stats = slab.stats()
        Sensor, RMS, Peak, InputParam = collect_results(slab)
        # host-side wall clock of the three phases of the call (seconds)
        self.last_timing = {'setup_upload_s': t1 - t0, 'time_loop_s': t2 - t1, 'download_s': time.perf_counter() - t2,
                            'h2d_bytes': slab.h2d_bytes, 'd2h_bytes': slab.d2h_bytes}
        last = _LastMap(slab)
        _live_lastmaps.append(last)
        self._last_slabs = (slab,)
        if SelRMSorPeak == 3:
            return Sensor, last, RMS, Peak, InputParam
        return Sensor, last, (RMS if SelRMSorPeak == 1 else Peak), InputParam

    def _run_multi_gpu(self, ngpu, t0, args, kwargs, name, number):
        _capi.require_gpu()
        devices = _select_devices(name, number, ngpu)
        Sensor, RMS, Peak, InputParam, slabs, timing = run_slabs_in_process(devices, args, kwargs)
        self.last_stats = [s.stats() for s in slabs]
        self.last_timing = dict(timing, total_s=time.perf_counter() - t0, devices=devices)
        last = _LastMapSlabs(slabs)
        _live_lastmaps.append(last)
        self._last_slabs = tuple(slabs)
        if kwargs['SelRMSorPeak'] == 3:
            return Sensor, last, RMS, Peak, InputParam
        return Sensor, last, (RMS if kwargs['SelRMSorPeak'] == 1 else Peak), InputParam


def _matching_devices(name):
    """Ordinals of the CUDA devices whose name contains `name`, like the reference's InitCuda / device selection
    (a substring such as 'A6000' or 'B200').  An empty name or None means "any device".  A non-empty name that matches
    nothing raises, as the reference does, unless BABELB200_ANY_DEVICE=1 asks for the old advisory behaviour (configs
    written for another GPU, e.g. BabelBrain/default.yaml)."""
    names = _capi.device_names()
    if not (isinstance(name, str) and name):
        return list(range(len(names)))
    hits = [n for n, s in enumerate(names) if name in s]
    if not hits:
        if os.environ.get('BABELB200_ANY_DEVICE') == '1':
            return list(range(len(names)))
        raise ValueError('no CUDA device whose name contains %r (visible: %s); set BABELB200_ANY_DEVICE=1 to run on any device'
                         % (name, names))
    return hits


def _select_devices(name, number, count):
    """`count` device ordinals for a slab-decomposed run: the devices whose name holds `name`, starting at the
    `number`-th of them."""
    hits = _matching_devices(name)
    first = min(int(number), max(len(hits) - 1, 0))
    hits = hits[first:] + hits[:first]
    if len(hits) < count:
        raise ValueError('NumberGPUs=%d but only %d matching CUDA device(s) are visible' % (count, len(hits)))
    return hits[:count]


def _select_device(name, number=0):
    """The `number`-th CUDA device whose name contains `name` (the last one if there are fewer)."""
    hits = _matching_devices(name)
    if not hits:
        raise _capi.BabelB200Error('no CUDA device visible: babelbrain_b200 has no CPU fallback')
    return hits[min(int(number), len(hits) - 1)]
