#!/usr/bin/env python
"""
bench.py -- FDTD Gcell-updates/s of the viscoelastic solver on BASELINE.json's CTX-500 case
(configs[1]: 500 kHz, PPW 6, 240x240x320 synthetic skull+brain label map, 2544 time steps).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (oracle port on the host cores)

A "step" is one whole simulation of the workload.  N>1 is weak scaling: every rank owns one
240-plane slab of a (240*N)x240x320 domain and exchanges velocity/stress halos each half-step
(NCCL send/recv inside libbabelb200.so).  `value` = cell-updates of all ranks / device time (CUDA
events, max over ranks) with inputs resident in HBM; `e2e` = the same metric through the public
PropagationModel.StaggeredFDTD_3D_with_relaxation call with host buffers (uploads, run, downloads).
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
# dram__bytes_read.sum + dram__bytes_write.sum of one stress_tma launch from the committed ncu --set full capture
# (profiles/)
NCU_TRAFFIC_BYTES = 1352.5e6   # profiles/r1_ncu_stress_particle_summary.txt: 881.9 MB read + 470.6 MB written
DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')


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
    nd = di + dj + dk
    solid = (ML[:, 2] > 0)[own]
    att = ((ML[:, 3] > 0) | (ML[:, 4] > 0))[own]
    inner = nd == 0
    cls = {'solid': int((inner & solid).sum()), 'att_fluid': int((inner & ~solid & att).sum()),
           'lossless': int((inner & ~solid & ~att).sum()), 'pml': int((~inner).sum())}
    nds = int(nd[~inner & solid].sum())
    ndf = int(nd[~inner & ~solid].sum())
    pml_s, pml_f = int((~inner & solid).sum()), int((~inner & ~solid).sum())
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


def build_rank_workload(rank, nranks):
    """Arguments of this rank's slab.  N == 1: the CTX-500 case itself.  N > 1: the same transverse
    size, 240*N planes along axis 0; only this rank's planes (+halo) are materialised."""
    from babelbrain_b200 import workloads
    from babelbrain_b200.slab import SlabPlan
    if nranks == 1:
        w = workloads.make_workload(WORKLOAD)
        return w, None
    base = workloads.CONFIGS[WORKLOAD]['shape']
    n1g = base[0] * nranks
    plan = SlabPlan(n1g, nranks)
    glo, ghi = plan.with_halo(rank)
    # generate the base case and place it periodically along axis 0 (same physics in every slab)
    w = workloads.make_workload(WORKLOAD)
    MM, ML, f, SM, SF, h, T, SEN = w['args']
    pml = w['meta']['pml']
    idx = (np.arange(glo, ghi) % base[0])

    def tile_map(a):
        return np.ascontiguousarray(a[idx])
    MMl, SMl, SENl = tile_map(MM), tile_map(SM), tile_map(SEN)
    gi = np.arange(glo, ghi)
    inner = (gi >= pml) & (gi < n1g - pml)
    # sources / sensors exist on every non-PML plane of the global grid; lateral PML planes of the
    # base case that end up inside the big domain get the neighbouring plane's source rows
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
    return dict(args=(MMl, ML, f, SMl, SF, h, T, SENl), kwargs=kw, meta=meta), (glo, n1g)


def cpu_baseline(sample_steps=None, threads=None):
    """Oracle port (C/OpenMP float32 restatement) on the host cores, bounded sample of the workload."""
    import oracle
    from babelbrain_b200 import workloads
    w = workloads.make_workload(WORKLOAD)
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    lib = oracle.load(np.float32)
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
            'sample': '%s: first %d of %d time steps (%.1f s), C/OpenMP float32 oracle port' % (WORKLOAD, sample_steps, w['meta']['steps'], el)}, w['meta']


def workload_name(meta, world):
    return ('CTX-500 annular array 500 kHz, synthetic skull+brain label map PPW 6, %dx%dx%d, %d time steps per simulation (BASELINE configs[1]%s)'
            % (tuple(meta['shape']) + (meta['steps'], '' if world == 1 else '; %d slabs of 240 planes, weak scaling' % world)))


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    vals = []
    meta = None
    for n in range(args.warmup + args.steps):
        cb, meta = cpu_baseline(sample_steps=12)
        if n >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c['value'] for c in vals]))
    ms = meta['cells'] * 12 / (v * 1e9) * 1e3
    cb = dict(vals[-1], value=v)
    out = {'impl': 'reference', 'metric': 'FDTD Gcell-updates/s', 'value': v, 'unit': 'Gcell-updates/s',
           'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': workload_name(meta, 1), 'cells': meta['cells'], 'time_steps': meta['steps'],
                      'sample': 'each step times the first 12 of the %d time steps of the simulation on the host cores' % meta['steps']},
           'cpu_baseline': cb, 'e2e': {'value': v, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--variant', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from babelbrain_b200.propagation import FdtdSlab, PropagationModel, collect_results
    from babelbrain_b200 import _capi

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit('--gpus %d needs torch.distributed.run with --nproc-per-node %d' % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: babelbrain_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, local = build_rank_workload(rank, world)
    meta = w['meta']
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    origin, n1g = (None, None) if local is None else local
    slab = FdtdSlab(*w['args'], device=local_rank, rank=rank, nranks=world, kernel_variant=args.variant,
                    origin=origin, n1_global=n1g, **kw)
    halo = os.environ.get('BB_HALO', 'peer')     # 'peer': NVLink halo push from the boundary CTAs; 'nccl': send/recv exchange

    def connect(s, first):
        if world == 1:
            return
        if halo == 'nccl':
            if first:
                ids = [FdtdSlab.nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                s.comm_init(ids[0])
            else:
                s.comm_init()          # the process keeps its slab communicator between simulations
        else:
            exports = [None] * world
            dist.all_gather_object(exports, s.peer_export())
            s.peer_attach(exports[rank - 1] if rank > 0 else None, exports[rank + 1] if rank < world - 1 else None)
            dist.barrier()
    connect(slab, True)
    glo = 0 if origin is None else origin
    cls, alg = traffic_model(w['args'][0], w['args'][1], meta['pml'], slab.i0, slab.i1, glo, meta['shape'][0])

    # ---- device-resident metric: reset + run, timed by CUDA events inside the library
    for _ in range(args.warmup):
        slab.reset()
        barrier()
        slab.run(profile=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, stats = [], []
    for _ in range(args.steps):
        slab.reset()     # also rewrites > 1 GB of state: the 126 MB L2 holds nothing of the next step
        barrier()
        st = slab.run(profile=True)
        barrier()
        ms.append(st['run_ms'])
        stats.append(st)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([float(np.mean(ms))], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    value = meta['cell_updates'] / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API (host buffers in, numpy results out)
    e2e_ms = []
    h2d = d2h = 0
    e2e_phases = None
    n_e2e = max(3, args.steps)        # one untimed call first (page-locked result pool, CUDA context), then n_e2e timed calls
    for n in range(1 + n_e2e):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            PM = PropagationModel()
            res = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
            h2d, d2h = PM.last_timing['h2d_bytes'], PM.last_timing['d2h_bytes']
            e2e_phases = dict(PM.last_timing)
            del res
        else:
            s2 = FdtdSlab(*w['args'], device=local_rank, rank=rank, nranks=world, kernel_variant=args.variant,
                          origin=origin, n1_global=n1g, **kw)
            connect(s2, False)
            s2.run()
            collect_results(s2)
            h2d, d2h = s2.h2d_bytes, s2.d2h_bytes
            barrier()                   # nobody frees memory a neighbour may still be writing to
            s2.close()
        barrier()
        if n > 0:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
    # the host phases of a call (pageable uploads, result downloads) see the box's other tenants: the median of the
    # timed calls is reported, all of them are listed
    te = torch.tensor([float(np.median(e2e_ms))], device='cuda')
    tb = torch.tensor([float(h2d), float(d2h)], device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    e2e_val = meta['cell_updates'] / (float(te.item()) * 1e-3) / 1e9

    if rank == 0:
        st = stats[-1]
        peak, peak_kind = measured_peak()
        # the dominant kernel is the stress half-step: one launch per time step over the whole slab
        n_launch = max(st['stress_launches'], 1)
        launches_per_step_call = n_launch / max(st['steps_done'], 1)
        stress_bytes = alg['stress'] / launches_per_step_call
        avg_ms = st['stress_ms'] / n_launch
        achieved = stress_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None
        n_pl = max(st['particle_launches'], 1)
        part_ms = st['particle_ms'] / n_pl
        part_gbs = alg['particle'] / (n_pl / max(st['steps_done'], 1)) / (part_ms * 1e-3) / 1e9 if part_ms > 0 else None
        step_bytes = alg['stress'] + alg['particle'] + alg['pressure_accumulator']
        step_gbs = step_bytes * meta['steps'] / (st['run_ms'] * 1e-3) / 1e9
        nominal = 158.0 * meta['cells'] / world * meta['steps'] / (ms_per_step * 1e-3) / 1e9
        out = {
            'metric': 'FDTD Gcell-updates/s', 'value': value, 'unit': 'Gcell-updates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(meta, world),
                       'cells': meta['cells'], 'time_steps': meta['steps'], 'seconds_per_simulation': ms_per_step * 1e-3,
                       'l2_policy': 'state (>1.2 GB per GPU) is far larger than the 126 MB L2 and is rewritten by reset() between timed simulations',
                       'kernel_variant': args.variant, 'cell_classes_rank0': cls,
                       'halo_exchange': None if world == 1 else ('NVLink peer stores from the boundary CTAs' if halo == 'peer' else 'NCCL send/recv')},
            'e2e': {'value': e2e_val, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': int(tb[0].item()), 'd2h_bytes_per_step': int(tb[1].item()),
                    'ms_per_step': float(te.item()), 'calls_ms_rank0': [round(x, 1) for x in e2e_ms], 'statistic': 'median of %d calls' % n_e2e,
                    'phases_rank0_last_call': e2e_phases},
            'gpu_launches': int(sum(s['stress_launches'] + s['particle_launches'] + s['pml_launches'] + s['other_launches'] for s in stats)),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'stress_tma (fused stress half-step: interior + PML shell, RMS folded in)', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None, 'traffic': NCU_TRAFFIC_BYTES, 'peak_kind': peak_kind,
                         'algorithmic_bytes_per_launch': stress_bytes, 'avg_launch_ms': avg_ms,
                         'kernel_share_of_step': st['stress_ms'] / st['run_ms'] if st['run_ms'] else None,
                         'particle_kernel': {'achieved': part_gbs, 'frac': (part_gbs / peak) if part_gbs else None,
                                             'algorithmic_bytes_per_launch': alg['particle'], 'avg_launch_ms': part_ms},
                         'whole_step_algorithmic_GBs': step_gbs, 'whole_step_algorithmic_frac': step_gbs / peak,
                         'pressure_accumulator_bytes_per_step': alg['pressure_accumulator'],
                         'whole_step_nominal_158B_GBs': nominal, 'whole_step_nominal_frac': nominal / peak,
                         'per_kernel_ms': {'stress': st['stress_ms'], 'particle': st['particle_ms'], 'pml': st['pml_ms'], 'other': st['other_ms'], 'run': st['run_ms']}},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                out['cpu_baseline'], _ = cpu_baseline()
            except Exception as e:  # the baseline is a reported number, never a dependency of the GPU path
                out['cpu_baseline'] = {'value': None, 'unit': 'Gcell-updates/s', 'cores': None, 'kind': 'port', 'sample': 'failed: %r' % (e,)}
        print(json.dumps(out), flush=True)
    slab.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
