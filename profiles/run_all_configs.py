"""Every BASELINE.json config end to end through the public call (PropagationModel.StaggeredFDTD_3D_with_relaxation, host
arrays in, numpy results out), whole simulations at the full sizes of SURVEY.md section 8(a):
   python profiles/run_all_configs.py single_water ctx500_skull h317_skull            (one GPU)
   python profiles/run_all_configs.py dome_stress:2                                   (name:NumberGPUs, GPUs of this process)
Sources are passed as CWSourceFunctions (no dense table) and Ox/Oy/Oz as broadcast views, so the host side stays small;
appends one JSON line per config to gpurun_out/configs.jsonl."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import workloads
from babelbrain_b200.propagation import PropagationModel
os.makedirs('gpurun_out', exist_ok=True)
for spec in sys.argv[1:]:
    name, _, ng = spec.partition(':')
    ng = int(ng or 1)
    t0 = time.perf_counter()
    w = workloads.make_workload(name, lean=(name != 'dome_stress'), dense_sources=False)
    m = w['meta']
    t1 = time.perf_counter()
    PM = PropagationModel()
    res = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], NumberGPUs=ng, **w['kwargs'])
    t2 = time.perf_counter()
    ph = PM.CalculatePhaseDataOnDevice(m['frequency'])
    t3 = time.perf_counter()
    st = PM.last_stats if isinstance(PM.last_stats, list) else [PM.last_stats]
    loop_s = max(s['run_ms'] for s in st) / 1e3
    rms = res[2]['Pressure']
    line = {'config': name, 'shape': list(m['shape']), 'cells': m['cells'], 'time_steps': m['steps'], 'n_gpus': ng,
            'maps': w['kwargs']['SelMapsRMSPeakList'], 'SelRMSorPeak': w['kwargs']['SelRMSorPeak'], 'TypeSource': w['kwargs']['TypeSource'],
            'sensors': int(res[-1]['IndexSensorMap'].size), 'samples': int(res[0]['time'].size), 'sources': m['nsrc'],
            'host_build_s': round(t1 - t0, 2), 'call_s': round(t2 - t1, 3), 'time_loop_device_s': round(loop_s, 3),
            'gcell_updates_per_s_device': round(m['cell_updates'] / loop_s / 1e9, 2), 'gcell_updates_per_s_call': round(m['cell_updates'] / (t2 - t1) / 1e9, 2),
            'phases': {k: (round(v, 3) if isinstance(v, float) else v) for k, v in PM.last_timing.items()},
            'phase_data_on_device_s': round(t3 - t2, 3), 'device_GB': [round(s['device_bytes'] / 1e9, 2) for s in st],
            'peak_voxel': [int(x) for x in np.unravel_index(int(np.argmax(rms)), rms.shape)], 'rms_max': float(rms.max()),
            'fourier_over_sqrt2_rms_at_peak': float(abs(ph['PressMapFourier'].reshape(-1)[int(np.argmax(rms))]) / (np.sqrt(2) * rms.max()))}
    print(json.dumps(line), flush=True)
    with open('gpurun_out/configs.jsonl', 'a') as f:
        f.write(json.dumps(line) + '\n')
    del res, ph, rms, w
