#!/usr/bin/env python
"""
bench.py -- FDTD Gcell-updates/s of the viscoelastic solver on BASELINE.json's CTX-500 case
(configs[1]: 500 kHz, PPW 6, 240x240x320 synthetic skull+brain label map, 2544 time steps).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (oracle port on the host cores)

Headline record.  A "step" is one whole simulation of the workload.  N>1 is weak scaling: every rank owns one
240-plane slab of a (240*N)x240x320 domain and exchanges velocity/stress halos each half-step (boundary CTAs store
into the neighbour's halo planes over NVLink; BB_HALO=nccl selects the NCCL send/recv exchange).  `value` = cell-updates
of all ranks / device time (CUDA events, max over ranks) with inputs resident in HBM; `e2e` = the same metric through
the public PropagationModel.StaggeredFDTD_3D_with_relaxation call with host buffers (uploads, run, downloads, and for
N>1 the whole-grid gather: rank 0 makes the call with NumberGPUs=N while the other ranks idle).

Sub-records of the same JSON line (each measured in this run, SURVEY.md 8e / BASELINE configs):
  strong        BASELINE configs[4]: the 1 MHz 1080^3 domain (1.26 G cells) cut into N slabs, --strong-periods periods of the
                11 250-step run with the RMS window and the sensors active in the last two; Gcell-updates/s, per-GPU
                per-class roofline fractions.  N=1 holds the whole domain on one GPU (158 GB).
  parity_vs_n1  (N>1) a small skull domain run as N slabs against the same domain on one GPU: relative L2 of the RMS
                pressure map and of the sensor traces, peak voxel.
  dome          (N = 2, 4) BASELINE configs[3]: DomeTx 650 kHz 700x700x500, 1024 volumetric stress sources, N slabs.
  rayleigh      the Rayleigh ForwardSimple source-field integral of the CTX-500 case (18.4 M field points), pairs/s.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'ctx500_skull'
DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed summary
    of the most recent `ncu --set full` capture (profiles/ncu_traffic.json, written by profiles/ncu_summary.py together
    with the source hash of the kernels it was taken from).  None when the capture is older than the kernels."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            t = json.load(f)
        return float(t['stress_tma']['dram_bytes_per_launch']), t.get('source'), t.get('kernel_sha16') == kernel_sha16()
    except Exception:
        return None, None, False


def kernel_sha16():
    import hashlib
    h = hashlib.sha256()
    for f in ('fdtd_tma.cuh', 'fdtd_cell.cuh', 'common.h'):
        with open(os.path.join(ROOT, 'babelbrain_b200', 'csrc', f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def traffic_model(MM, ML, pml, i0, i1, glo, n1g):
    """Algorithmic (compulsory) bytes of one launch of each half-step kernel over the owned planes, fp32
    state and uint8 labels.  Interior cells: stress reads 3 V, read-modify-writes
    the stresses and memory variables that exist for the cell's class (solid 6+6, attenuating fluid 3+3,
    lossless fluid 3+0); particle reads those stresses and read-modify-writes 3 V.  PML cells carry no
    memory variables but read-modify-write the damped split parts of each damped axis (3 normal + 2
    shear stress parts, 3 velocity parts).  The pressure accumulator (8 B/interior cell) is reported
    separately, as SURVEY.md 8(d) asks."""
    own = MM[i0 - glo:i1 - glo]
    n1, n2, n3 = own.shape
    gi = np.arange(i0, i1)
    di = ((gi < pml) | (gi >= n1g - pml)).astype(np.int8)[:, None, None]
    dj = ((np.arange(n2) < pml) | (np.arange(n2) >= n2 - pml)).astype(np.int8)[None, :, None]
    dk = ((np.arange(n3) < pml) | (np.arange(n3) >= n3 - pml)).astype(np.int8)[None, None, :]
    solid_m, att_m = (ML[:, 2] > 0), ((ML[:, 3] > 0) | (ML[:, 4] > 0))
    cls = {'solid': 0, 'att_fluid': 0, 'lossless': 0, 'pml': 0}
    nds = ndf = pml_s = pml_f = 0
    for a in range(0, n1, 32):                    # by groups of planes: the 1080^3 slab does not fit twice in host memory
        o = own[a:a + 32]
        nd = di[a:a + 32] + dj + dk
        solid, att, inner = solid_m[o], att_m[o], nd == 0
        cls['solid'] += int((inner & solid).sum())
        cls['att_fluid'] += int((inner & ~solid & att).sum())
        cls['lossless'] += int((inner & ~solid & ~att).sum())
        cls['pml'] += int((~inner).sum())
        nds += int(nd[~inner & solid].sum())
        ndf += int(nd[~inner & ~solid].sum())
        pml_s += int((~inner & solid).sum())
        pml_f += int((~inner & ~solid).sum())
    stress = (cls['solid'] * (12 + 48 + 48 + 1) + cls['att_fluid'] * (12 + 24 + 24 + 1) + cls['lossless'] * (12 + 24 + 1)
              + pml_s * (12 + 48 + 1) + nds * 40 + pml_f * (12 + 24 + 1) + ndf * 24)
    particle = (cls['solid'] * (24 + 24 + 1) + (cls['att_fluid'] + cls['lossless']) * (12 + 24 + 1)
                + pml_s * (24 + 24 + 1) + pml_f * (12 + 24 + 1) + (nds + ndf) * 24)
    pressure = 8 * (cls['solid'] + cls['att_fluid'] + cls['lossless'])
    return cls, {'stress': float(stress), 'particle': float(particle), 'pressure_accumulator': float(pressure)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for c, n in enumerate(names) if any(len(r) >= 7 and r[3 + c].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def build_weak_workload(nranks, planes=None):
    """The headline workload.  N == 1: the CTX-500 case itself.  N > 1: the same transverse size, 240*N planes along
    axis 0 (the base case placed periodically, sources and sensors on every non-PML plane).  planes=(lo, hi)
    materialises only those planes (one rank's slab + halo); None the whole grid."""
    from babelbrain_b200 import workloads
    w = workloads.make_workload(WORKLOAD)
    if nranks == 1:
        return w
    base = workloads.CONFIGS[WORKLOAD]['shape']
    n1g = base[0] * nranks
    glo, ghi = (0, n1g) if planes is None else planes
    MM, ML, f, SM, SF, h, T, SEN = w['args']
    pml = w['meta']['pml']
    idx = (np.arange(glo, ghi) % base[0])

    def tile_map(a):
        return np.ascontiguousarray(a[idx])
    MMl, SMl, SENl = tile_map(MM), tile_map(SM), tile_map(SEN)
    gi = np.arange(glo, ghi)
    inner = (gi >= pml) & (gi < n1g - pml)
    # the base case's water shell along axis 0 lands inside the big domain at the slab joints: harmless (water);
    # sources / sensors exist on every non-PML plane of the global grid
    src_plane = SM[pml:-pml][:, :, pml]
    fill = src_plane[(gi % base[0]).clip(0, src_plane.shape[0] - 1)]
    SMl[:, :, pml] = np.where(inner[:, None], fill, 0)
    SENl[:] = 0
    SENl[inner, pml:-pml, pml + 1:-pml] = 1
    kw = dict(w['kwargs'])
    for k in ('Ox', 'Oy', 'Oz'):
        kw[k] = np.ascontiguousarray(kw[k][idx])
    meta = dict(w['meta'], shape=(n1g, base[1], base[2]), cells=n1g * base[1] * base[2],
                cell_updates=n1g * base[1] * base[2] * w['meta']['steps'])
    return dict(args=(MMl, ML, f, SMl, SF, h, T, SENl), kwargs=kw, meta=meta)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(sample_steps=None, nslabs=1):
    """Oracle port (C/OpenMP float32 restatement) on all the host cores this process may use, bounded sample of the
    workload (nslabs > 1: the weak-scaled domain of that many slabs)."""
    nthreads = host_threads()
    os.environ['OMP_NUM_THREADS'] = str(nthreads)          # torch.distributed.run exports OMP_NUM_THREADS=1
    import oracle
    w = build_weak_workload(nslabs)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    lib = oracle.load(np.float32)
    if hasattr(lib, 'oracle_set_threads'):
        lib.oracle_set_threads(nthreads)
    cores = lib.oracle_num_threads()
    if sample_steps is None:
        t0 = time.time()
        oracle.run_c(*w['args'], steps_override=2, **kw)
        per = max((time.time() - t0) / 2, 1e-3)
        sample_steps = int(np.clip(15.0 / per, 4, 200))
    t0 = time.time()
    oracle.run_c(*w['args'], steps_override=sample_steps, **kw)
    el = time.time() - t0
    val = w['meta']['cells'] * sample_steps / el / 1e9
    return {'value': val, 'unit': 'Gcell-updates/s', 'cores': cores, 'kind': 'port',
            'sample': '%s, %d slab(s): first %d of %d time steps (%.1f s), C/OpenMP float32 oracle port on %d threads'
                      % (WORKLOAD, nslabs, sample_steps, w['meta']['steps'], el, cores)}, w['meta'], sample_steps


def workload_name(meta, world):
    return ('CTX-500 annular array 500 kHz, synthetic skull+brain label map PPW 6, %dx%dx%d, %d time steps per simulation (BASELINE configs[1]%s)'
            % (tuple(meta['shape']) + (meta['steps'], '' if world == 1 else '; %d slabs of 240 planes, weak scaling' % world)))


def run_reference(args):
    """The reference arm: the CPU implementation of the path (the oracle port -- BabelViscoFDTD's OpenMP backend is not
    installable here) on all host cores, on this arm's workload (the N-slab domain for --gpus N)."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    sample = max(3, 12 // max(args.gpus, 1))
    vals, meta = [], None
    for n in range(args.warmup + args.steps):
        cb, meta, sample = cpu_baseline(sample_steps=sample, nslabs=args.gpus)
        if n >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c['value'] for c in vals]))
    ms = meta['cells'] * sample / (v * 1e9) * 1e3
    cb = dict(vals[-1], value=v)
    out = {'impl': 'reference', 'metric': 'FDTD Gcell-updates/s', 'value': v, 'unit': 'Gcell-updates/s',
           'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': workload_name(meta, args.gpus), 'cells': meta['cells'], 'time_steps': meta['steps'],
                      'sample': 'each step times the first %d of the %d time steps of the simulation on the host cores' % (sample, meta['steps'])},
           'cpu_baseline': cb, 'e2e': {'value': v, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------------
class Ranks:
    """torch.distributed plumbing of one bench process (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        self.rank = int(os.environ.get('RANK', 0))
        self.local_rank = int(os.environ.get('LOCAL_RANK', 0))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: babelbrain_b200 has no CPU fallback')
        torch.cuda.set_device(self.local_rank)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            self.cpu_group = dist.new_group(backend='gloo')     # host-side waits that must not put a spinning kernel on the GPUs
        self.halo = os.environ.get('BB_HALO', 'peer')

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op='max'):
        t = self.torch.tensor([float(v) for v in values], device='cuda', dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op={'max': self.dist.ReduceOp.MAX, 'sum': self.dist.ReduceOp.SUM, 'min': self.dist.ReduceOp.MIN}[op])
        return [float(x) for x in t.tolist()]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def connect(self, slab, first=True):
        """Attach the slab to its neighbours: NVLink peer stores (CUDA IPC between the rank processes) or NCCL."""
        from babelbrain_b200.propagation import FdtdSlab
        if self.world == 1:
            return
        if self.halo == 'nccl':
            if first:
                ids = [FdtdSlab.nccl_unique_id() if self.rank == 0 else None]
                self.dist.broadcast_object_list(ids, src=0)
                slab.comm_init(ids[0])
            else:
                slab.comm_init()          # the process keeps its slab communicator between simulations
        else:
            exports = self.gather_objects(slab.peer_export())
            slab.peer_attach(exports[self.rank - 1] if self.rank > 0 else None,
                             exports[self.rank + 1] if self.rank < self.world - 1 else None)
            self.dist.barrier()

    def close(self, slab):
        self.barrier()                    # nobody frees memory a neighbour may still be writing to
        slab.close()
        self.barrier()


def kernel_fractions(st, alg, peak):
    """Per-launch time and algorithmic GB/s of the two half-step kernels of one rank (profile mode statistics)."""
    out = {}
    for name in ('stress', 'particle'):
        nl = max(st[name + '_launches'], 1)
        per_step = nl / max(st['steps_done'], 1)
        ms = st[name + '_ms'] / nl
        gbs = alg[name] / per_step / (ms * 1e-3) / 1e9 if ms > 0 else None
        out[name] = {'avg_launch_ms': ms, 'algorithmic_bytes_per_launch': alg[name] / per_step, 'achieved_GBs': gbs,
                     'frac': gbs / peak if gbs else None}
    return out


def strong_record(R, periods, n=1080):
    """BASELINE configs[4] cut into R.world slabs; every rank materialises only its own planes (+2 halo planes each side)."""
    from babelbrain_b200 import workloads
    from babelbrain_b200.propagation import FdtdSlab
    from babelbrain_b200.slab import SlabPlan
    plan = SlabPlan(n, R.world)
    glo, ghi = plan.with_halo(R.rank)
    t0 = time.time()
    w = workloads.make_workload('hires_1mhz', shape=(n, n, n), periods=periods, planes=(glo, ghi), lean=True, dense_sources=False)
    t_build = time.time() - t0
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    s = FdtdSlab(*w['args'], device=R.local_rank, rank=R.rank, nranks=R.world, origin=glo, n1_global=n, **kw)
    t_setup = time.time() - t0 - t_build
    R.connect(s)
    cls, alg = traffic_model(w['args'][0], w['args'][1], 12, s.i0, s.i1, glo, n)
    R.barrier()
    s.run(5)
    R.barrier()
    s.reset()
    R.barrier()
    st = s.run(profile=True)
    R.barrier()
    run_ms, = R.reduce([st['run_ms']])
    peak = measured_peak()[0]
    mine = kernel_fractions(st, alg, peak)
    allk = R.gather_objects({'rank': R.rank, 'stress_frac': mine['stress']['frac'], 'particle_frac': mine['particle']['frac'],
                             'run_ms': st['run_ms'], 'cells': cls})
    steps, cells = w['meta']['steps'], n ** 3
    rec = None
    if R.rank == 0:
        step_bytes = alg['stress'] + alg['particle'] + alg['pressure_accumulator']
        rec = {'workload': '1 MHz PPW 9 high-resolution domain %dx%dx%d (BASELINE configs[4]), %d of 11250 time steps (%d periods), RMS window '
                           'and sensors active in the last 2 periods, continuous-wave sources evaluated in the source kernel' % (n, n, n, steps, periods),
               'scaling': 'strong', 'n_gpus': R.world, 'planes_per_gpu': s.i1 - s.i0, 'time_steps': steps, 'run_ms': run_ms,
               'ms_per_time_step': run_ms / steps, 'value': cells * steps / run_ms / 1e6, 'unit': 'Gcell-updates/s',
               'per_gpu': cells * steps / run_ms / 1e6 / R.world,
               'nominal_158B_frac_of_peak_per_gpu': 158.0 * cells * steps / run_ms / 1e6 / R.world / peak,
               'rank0_kernels': mine, 'rank0_whole_step_algorithmic_frac': step_bytes * steps / (st['run_ms'] * 1e-3) / 1e9 / peak,
               'per_rank_frac': [{k: v for k, v in a.items() if k != 'cells'} for a in allk],
               'min_stress_frac_over_ranks': min(a['stress_frac'] for a in allk), 'min_particle_frac_over_ranks': min(a['particle_frac'] for a in allk),
               'device_GB_rank0': st['device_bytes'] / 1e9, 'host_build_s': t_build, 'setup_upload_s': t_setup,
               'halo_exchange': None if R.world == 1 else ('NVLink peer stores from the boundary CTAs' if R.halo == 'peer' else 'NCCL send/recv'),
               'cell_classes_rank0': cls}
    R.close(s)
    return rec


def parity_record(R):
    """A small skull domain as R.world slabs (the production exchange) against the same domain on one GPU."""
    from babelbrain_b200 import workloads
    from babelbrain_b200.propagation import FdtdSlab, collect_results
    n1 = max(64, 16 * R.world)
    w = workloads.make_workload(WORKLOAD, shape=(n1, 56, 72), periods=6, pml=8)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    one = FdtdSlab(*w['args'], device=R.local_rank, **kw)
    one.run()
    S1, R1, _, IP1 = collect_results(one)
    one.close()
    s = FdtdSlab(*w['args'], device=R.local_rank, rank=R.rank, nranks=R.world, **kw)
    R.connect(s)
    s.run()
    R.barrier()
    mine = s.get_map(0, 'Pressure')
    ref = R1['Pressure'][s.i0:s.i1].astype(np.float64)
    sens = s.get_sensors('Pressure').astype(np.float64)
    sref = S1['Pressure'][s.sensor_rows].astype(np.float64)
    pk = int(np.argmax(mine))
    num, den, snum, sden = R.reduce([((mine - ref) ** 2).sum(), (ref ** 2).sum(), ((sens - sref) ** 2).sum(), (sref ** 2).sum()], 'sum')
    peaks = R.gather_objects((float(mine.reshape(-1)[pk]), int(pk + s.i0 * mine.shape[1] * mine.shape[2])))
    idx_ok, = R.reduce([float(np.array_equal(s.IndexSensorMapLocal, IP1['IndexSensorMap'][s.sensor_rows]))], 'min')
    R.close(s)
    if R.rank != 0:
        return None
    best = max(peaks)
    return {'workload': '%s %dx56x72, 6 periods, PML 8: %d slabs vs one GPU' % (WORKLOAD, n1, R.world),
            'rms_pressure_rel_l2': float(np.sqrt(num / den)), 'sensor_rel_l2': float(np.sqrt(snum / max(sden, 1e-300))),
            'peak_voxel_identical': bool(best[1] == int(np.argmax(R1['Pressure']))), 'sensor_index_identical': bool(idx_ok == 1.0),
            'tolerance': 'same kernels and order of operations: <= 1e-6'}


def dome_record(R, periods=4):
    """BASELINE configs[3]: DomeTx full-dome domain, 1024 volumetric stress sources (TypeSource 2), R.world slabs."""
    from babelbrain_b200 import workloads
    from babelbrain_b200.propagation import FdtdSlab
    t0 = time.time()
    w = workloads.make_workload('dome_stress', periods=periods, dense_sources=False)
    t_build = time.time() - t0
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    s = FdtdSlab(*w['args'], device=R.local_rank, rank=R.rank, nranks=R.world, global_sensor_table=False, **kw)
    R.connect(s)
    glo = 0
    cls, alg = traffic_model(w['args'][0], w['args'][1], 12, s.i0, s.i1, glo, w['meta']['shape'][0])
    R.barrier()
    s.run(5)
    R.barrier()
    s.reset()
    R.barrier()
    st = s.run(profile=True)
    R.barrier()
    run_ms, = R.reduce([st['run_ms']])
    peak = measured_peak()[0]
    mine = kernel_fractions(st, alg, peak)
    fr = R.gather_objects((mine['stress']['frac'], mine['particle']['frac']))
    rec = None
    if R.rank == 0:
        m = w['meta']
        rec = {'workload': 'DomeTx 650 kHz PPW 6 full-dome domain %dx%dx%d (BASELINE configs[3]), 1024 volumetric stress sources, %d of ~6550 time steps'
                           % (tuple(m['shape']) + (m['steps'],)), 'n_gpus': R.world, 'time_steps': m['steps'], 'run_ms': run_ms,
               'value': m['cells'] * m['steps'] / run_ms / 1e6, 'unit': 'Gcell-updates/s', 'per_gpu': m['cells'] * m['steps'] / run_ms / 1e6 / R.world,
               'rank0_kernels': mine, 'per_rank_stress_particle_frac': fr, 'host_build_s': t_build, 'cell_classes_rank0': cls}
    R.close(s)
    return rec


def rayleigh_record(R):
    """ForwardSimple as BabelIntegrationSingle.py:290-297 calls it: a 64 mm / F = 63.2 mm bowl decomposed at lambda/6 into
    sub-elements, field = every voxel of the CTX-500 grid; on rank 0, sharded over the R.world GPUs of the box."""
    if R.rank != 0:
        return None
    from babelbrain_b200 import workloads, rayleigh
    import oracle
    cfg = workloads.CONFIGS[WORKLOAD]
    n1, n2, n3 = cfg['shape']
    f = cfg['frequency']
    h = workloads.SHEAR_FLOOR_SOS / f / cfg['ppw']
    Tx = rayleigh.GenerateFocusTx(f, cfg['focal'], cfg['aperture'], 1500.0, PPWSurface=6)
    center = Tx['center'].copy()
    center[:, 2] += cfg['focal'] - 0.01          # apex 10 mm behind the first grid plane
    x = (np.arange(n1) - n1 / 2 + 0.5) * h
    y = (np.arange(n2) - n2 / 2 + 0.5) * h
    z = np.arange(n3) * h
    X, Y, Z = np.meshgrid(x, y, z, indexing='ij')
    rf = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    u0 = np.ones(center.shape[0], np.complex64)
    k = np.array(2 * np.pi * f / 1500.0 + 0j).astype(np.complex64)
    rayleigh.InitCuda('')
    rayleigh.ForwardSimple(k, center, Tx['ds'], u0, rf[:200000], NumberGPUs=R.world)      # warm-up: buffers, module load
    t0 = time.perf_counter()
    out = rayleigh.ForwardSimple(k, center, Tx['ds'], u0, rf, NumberGPUs=R.world)
    wall = time.perf_counter() - t0
    kernel_ms = rayleigh._state['last_kernel_ms']
    pairs = float(rf.shape[0]) * center.shape[0]
    sub = np.linspace(0, rf.shape[0] - 1, 2000).astype(np.int64)
    ref = oracle.rayleigh_numpy(k, center, Tx['ds'], u0, rf[sub])
    err = float(np.linalg.norm(out[sub] - ref) / np.linalg.norm(ref))
    sfu_bound = 148 * 16 * 1.965e9 / 3 * R.world       # 3 MUFU ops per pair (rsqrt, sin, cos), 16 per SM per clock
    return {'workload': 'ForwardSimple: %d sub-elements (bowl 64 mm, F 63.2 mm, lambda/6) to the %d grid points of the CTX-500 domain' % (center.shape[0], rf.shape[0]),
            'n_gpus': R.world, 'pairs': pairs, 'kernel_ms_max_over_gpus': kernel_ms, 'pairs_per_s_kernel': pairs / (kernel_ms * 1e-3),
            'frac_of_sfu_bound': pairs / (kernel_ms * 1e-3) / sfu_bound, 'sfu_bound_pairs_per_s': sfu_bound,
            'call_s_host_buffers': wall, 'pairs_per_s_call': pairs / wall, 'rel_l2_vs_float64_on_2000_points': err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--variant', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--strong-periods', type=int, default=4, help='periods of the 1 MHz 1080^3 run timed by the strong-scaling sub-record (0: skip)')
    ap.add_argument('--strong-n', type=int, default=1080)
    ap.add_argument('--no-extras', action='store_true', help='headline record only (no strong / parity / dome / rayleigh sub-records)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    from babelbrain_b200.propagation import FdtdSlab, PropagationModel, release_device_state
    from babelbrain_b200.slab import SlabPlan
    R = Ranks()
    world, rank = R.world, R.rank
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit('--gpus %d needs torch.distributed.run with --nproc-per-node %d' % (args.gpus, args.gpus))

    base_n1 = 240
    n1g = base_n1 * world
    plan = SlabPlan(n1g, world)
    glo, ghi = plan.with_halo(rank)
    w = build_weak_workload(world, None if world == 1 else (glo, ghi))
    meta = w['meta']
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    origin, n1_global = (None, None) if world == 1 else (glo, n1g)
    slab = FdtdSlab(*w['args'], device=R.local_rank, rank=rank, nranks=world, kernel_variant=args.variant,
                    origin=origin, n1_global=n1_global, **kw)
    R.connect(slab, True)
    cls, alg = traffic_model(w['args'][0], w['args'][1], meta['pml'], slab.i0, slab.i1, 0 if origin is None else origin, meta['shape'][0])

    # ---- device-resident metric: reset + run, timed by CUDA events inside the library
    for _ in range(args.warmup):
        slab.reset()
        R.barrier()
        slab.run(profile=True)
    sampler = ClockSampler(R.local_rank)
    if rank == 0:
        sampler.start()
    ms, stats = [], []
    for _ in range(args.steps):
        slab.reset()     # also rewrites > 1 GB of state: the 126 MB L2 holds nothing of the next step
        R.barrier()
        st = slab.run(profile=True)
        R.barrier()
        ms.append(st['run_ms'])
        stats.append(st)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step, = R.reduce([float(np.mean(ms))])
    value = meta['cell_updates'] / (ms_per_step * 1e-3) / 1e9
    R.close(slab)

    # ---- end to end through the public API (host buffers in, numpy results out).  One process makes the call, as a user
    # does: at N > 1 rank 0 passes NumberGPUs=N (one thread per GPU inside the call, whole-grid arrays in and out) while
    # the other rank processes wait.
    e2e_ms, h2d, d2h, e2e_phases = [], 0, 0, None
    n_e2e = max(3, args.steps)        # one untimed call first (page-locked result pool, CUDA context), then n_e2e timed calls
    if rank == 0:
        wf = w if world == 1 else build_weak_workload(world)
        for n in range(1 + n_e2e):
            t0 = time.perf_counter()
            PM = PropagationModel()
            res = PM.StaggeredFDTD_3D_with_relaxation(*wf['args'], NumberGPUs=world, **wf['kwargs'])
            assert res[2]['Pressure'].shape == tuple(wf['meta']['shape'])
            dt_call = (time.perf_counter() - t0) * 1e3
            h2d, d2h = PM.last_timing['h2d_bytes'], PM.last_timing['d2h_bytes']
            e2e_phases = {k: v for k, v in PM.last_timing.items()}
            del res
            if n > 0:
                e2e_ms.append(dt_call)
        release_device_state()
        del wf
    R.host_barrier()                  # the idle ranks wait on the host: an NCCL barrier would spin on the GPUs rank 0 is timing
    e2e_val = None
    if rank == 0:
        e2e_med = float(np.median(e2e_ms))     # the host phases see the box's other tenants: median reported, all calls listed
        e2e_val = meta['cell_updates'] / (e2e_med * 1e-3) / 1e9

    out = None
    if rank == 0:
        st = stats[-1]
        peak, peak_kind = measured_peak()
        kf = kernel_fractions(st, alg, peak)
        step_bytes = alg['stress'] + alg['particle'] + alg['pressure_accumulator']
        step_gbs = step_bytes * meta['steps'] / (st['run_ms'] * 1e-3) / 1e9
        nominal = 158.0 * meta['cells'] / world * meta['steps'] / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_source, traffic_current = ncu_traffic()
        out = {
            'metric': 'FDTD Gcell-updates/s', 'value': value, 'unit': 'Gcell-updates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(meta, world),
                       'cells': meta['cells'], 'time_steps': meta['steps'], 'seconds_per_simulation': ms_per_step * 1e-3,
                       'l2_policy': 'state (>1.2 GB per GPU) is far larger than the 126 MB L2 and is rewritten by reset() between timed simulations',
                       'kernel_variant': args.variant, 'cell_classes_rank0': cls, 'pml_layer': 'classical split-field (MPMLRatio 0), water shell as the caller builds it',
                       'halo_exchange': None if world == 1 else ('NVLink peer stores from the boundary CTAs' if R.halo == 'peer' else 'NCCL send/recv')},
            'e2e': {'value': e2e_val, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_med, 'calls_ms': [round(x, 1) for x in e2e_ms], 'statistic': 'median of %d calls' % n_e2e,
                    'call': 'PropagationModel.StaggeredFDTD_3D_with_relaxation(*args, NumberGPUs=%d) on rank 0: whole-grid host arrays in, whole-grid numpy results out' % world,
                    'phases_last_call': e2e_phases},
            'gpu_launches': int(sum(s['stress_launches'] + s['particle_launches'] + s['pml_launches'] + s['other_launches'] for s in stats)),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'stress_tma (fused stress half-step: interior + PML shell, RMS folded in)', 'achieved': kf['stress']['achieved_GBs'], 'peak': peak,
                         'unit': 'GB/s', 'frac': kf['stress']['frac'], 'traffic': traffic, 'traffic_source': traffic_source,
                         'traffic_matches_current_kernels': traffic_current, 'peak_kind': peak_kind,
                         'algorithmic_bytes_per_launch': kf['stress']['algorithmic_bytes_per_launch'], 'avg_launch_ms': kf['stress']['avg_launch_ms'],
                         'kernel_share_of_step': st['stress_ms'] / st['run_ms'] if st['run_ms'] else None,
                         'particle_kernel': {'achieved': kf['particle']['achieved_GBs'], 'frac': kf['particle']['frac'],
                                             'algorithmic_bytes_per_launch': alg['particle'], 'avg_launch_ms': kf['particle']['avg_launch_ms']},
                         'whole_step_algorithmic_GBs': step_gbs, 'whole_step_algorithmic_frac': step_gbs / peak,
                         'pressure_accumulator_bytes_per_step': alg['pressure_accumulator'],
                         'whole_step_nominal_158B_GBs': nominal, 'whole_step_nominal_frac': nominal / peak,
                         'per_kernel_ms': {'stress': st['stress_ms'], 'particle': st['particle_ms'], 'pml': st['pml_ms'], 'other': st['other_ms'], 'run': st['run_ms']}},
        }
    del w, slab

    # ---- sub-records (each a measurement of this run; a failure is reported in place and does not void the headline)
    def guarded(name, fn):
        try:
            rec = fn()
        except Exception as e:  # noqa: BLE001
            rec = {'error': '%s: %s' % (type(e).__name__, e)} if rank == 0 else None
            if world > 1:
                raise            # a failed collective cannot be recovered rank by rank
        if rank == 0 and rec is not None:
            out[name] = rec
    if not args.no_extras:
        if world > 1:
            guarded('parity_vs_n1', lambda: parity_record(R))
        if args.strong_periods > 0:
            guarded('strong', lambda: strong_record(R, args.strong_periods, args.strong_n))
        if world in (2, 4):
            guarded('dome', lambda: dome_record(R))
        R.host_barrier()              # rank 0 shards the Rayleigh points over all GPUs: the others wait on the host, not in a spinning NCCL kernel
        guarded('rayleigh', lambda: rayleigh_record(R))
        R.host_barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                out['cpu_baseline'] = cpu_baseline()[0]
            except Exception as e:  # the baseline is a reported number, never a dependency of the GPU path
                out['cpu_baseline'] = {'value': None, 'unit': 'Gcell-updates/s', 'cores': None, 'kind': 'port', 'sample': 'failed: %r' % (e,)}
        print(json.dumps(out), flush=True)
    if world > 1:
        R.dist.destroy_process_group()


if __name__ == '__main__':
    main()
