"""Strong scaling of the 1 MHz high-resolution domain (BASELINE configs[4]: 1080^3, 1.26 G cells) over the GPUs of one box,
a few periods of the real run (11250 steps) with the RMS window and the sensor sampling active in the second half:

   python profiles/run_strong_1mhz.py [periods] [n] [n1]                              one GPU, whole domain (n = grid size, default 1080)
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
          profiles/run_strong_1mhz.py [periods] [n]                                   N slabs, NVLink halo push

Every rank materialises only its own planes (+2 halo planes each side).  Time = CUDA events inside the library around the
whole time loop, max over ranks.  One JSON line per run is appended to gpurun_out/strong_1mhz.jsonl."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import FdtdSlab
from babelbrain_b200.slab import SlabPlan

periods = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
n1 = int(sys.argv[3]) if len(sys.argv) > 3 else n          # planes along the decomposed axis (default: cube)
world, rank, local_rank = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


shape = (n1, n, n)
plan = SlabPlan(n1, world)
glo, ghi = plan.with_halo(rank)
t0 = time.time()
w = workloads.make_workload('hires_1mhz', shape=shape, periods=periods, planes=(glo, ghi), lean=True, dense_sources=False)
t_build = time.time() - t0
kw = {k: v for k, v in w['kwargs'].items() if k not in bench.DROP}
s = FdtdSlab(*w['args'], device=local_rank, rank=rank, nranks=world, origin=glo, n1_global=n1, **kw)
t_setup = time.time() - t0 - t_build
if world > 1:
    exports = [None] * world
    dist.all_gather_object(exports, s.peer_export())
    s.peer_attach(exports[rank - 1] if rank > 0 else None, exports[rank + 1] if rank < world - 1 else None)
    dist.barrier()
cls, alg = bench.traffic_model(w['args'][0], w['args'][1], 12, s.i0, s.i1, glo, n1)
barrier()
s.run(5)
barrier()
s.reset()
barrier()
st = s.run(profile=True)
barrier()
t = torch.tensor([st['run_ms'], st['stress_ms'], st['particle_ms']], device='cuda', dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
run_ms, stress_ms, particle_ms = (float(x) for x in t.tolist())
steps = w['meta']['steps']
if rank == 0:
    peak = bench.measured_peak()[0]
    cells = n1 * n * n
    line = {'workload': '1 MHz PPW 9 high-resolution domain %dx%dx%d, %d of 11250 time steps (%d periods), RMS window and sensors in the last 2 periods'
                        % (n1, n, n, steps, periods),
            'n_gpus': world, 'planes_per_gpu': s.i1 - s.i0, 'run_ms': run_ms, 'ms_per_time_step': run_ms / steps,
            'gcell_updates_per_s': cells * steps / run_ms / 1e6, 'per_gpu': cells * steps / run_ms / 1e6 / world,
            'nominal_158B_frac_of_peak_per_gpu': 158.0 * cells * steps / run_ms / 1e6 / world / peak,
            'rank0_stress': {'ms_per_launch': st['stress_ms'] / steps, 'algorithmic_GBs': alg['stress'] * steps / st['stress_ms'] / 1e6,
                             'frac': alg['stress'] * steps / st['stress_ms'] / 1e6 / peak},
            'rank0_particle': {'ms_per_launch': st['particle_ms'] / steps, 'algorithmic_GBs': alg['particle'] * steps / st['particle_ms'] / 1e6,
                               'frac': alg['particle'] * steps / st['particle_ms'] / 1e6 / peak},
            'max_over_ranks_ms': {'stress': stress_ms, 'particle': particle_ms}, 'rank0_other_ms': st['other_ms'],
            'device_GB_rank0': st['device_bytes'] / 1e9, 'host_build_s': t_build, 'setup_upload_s': t_setup,
            'halo_exchange': None if world == 1 else 'NVLink peer stores from the boundary CTAs', 'cell_classes_rank0': cls}
    print(json.dumps(line), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/strong_1mhz.jsonl', 'a') as f:
        f.write(json.dumps(line) + '\n')
barrier()
s.close()
if world > 1:
    dist.destroy_process_group()
