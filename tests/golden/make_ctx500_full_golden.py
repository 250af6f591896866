"""Generates tests/golden/ctx500_full.npz: BASELINE configs[1] (CTX-500, 240x240x320, 2544 steps) at FULL size,
computed once by the float64 C/OpenMP oracle (oracle/fdtd_oracle.c, ~10-20 min on 8 cores), reduced to what a
full-size parity check needs: the peak voxel, the three RMS planes through it, the norms of the whole map and of
the sensor table, every 97th sensor row, and the inputs' digest.

PARITY UNPINNED (see make_golden.py): the vectors come from the oracle restatement, not from BabelViscoFDTD.

    python tests/golden/make_ctx500_full_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
DROP = ('COMPUTING_BACKEND', 'USE_SINGLE', 'DefaultGPUDeviceName')
ROW_STRIDE = 97


def build_case():
    from babelbrain_b200 import workloads
    return workloads.make_workload('ctx500_skull')


def digest(w):
    """Cheap digest of the full-size inputs (label map, sizes, every 1009th source sample)."""
    import hashlib
    h = hashlib.sha256()
    MM, ML, f, SM, SF, hh, T, SEN = w['args']
    h.update(np.ascontiguousarray(MM).tobytes())
    h.update(np.ascontiguousarray(SM).tobytes())
    h.update(np.ascontiguousarray(SEN).tobytes())
    h.update(np.asarray(ML, np.float64).tobytes())
    h.update(np.ascontiguousarray(np.asarray(SF).reshape(-1)[::1009]).tobytes())
    h.update(repr((float(f), float(hh), float(T), w['kwargs']['SensorSubSampling'], w['kwargs']['SensorStart'],
                   w['kwargs']['NDelta'], float(w['kwargs']['DT']))).encode())
    return h.hexdigest()


def main():
    import oracle
    oracle.build()
    w = build_case()
    kw = {k: v for k, v in w['kwargs'].items() if k not in DROP}
    t0 = time.time()
    r = oracle.run_c(*w['args'], dtype=np.float64, **kw)
    rms = r['RMS']['Pressure']
    pk = np.unravel_index(int(np.argmax(rms)), rms.shape)
    sens = r['Sensor']['Pressure']
    out = dict(digest=np.array(digest(w)), peak_voxel=np.array(pk, np.int64), peak_value=np.array(rms[pk]),
               rms_norm=np.array(np.linalg.norm(rms.astype(np.float64))),
               plane_i=rms[pk[0]].astype(np.float32), plane_j=rms[:, pk[1]].astype(np.float32), plane_k=rms[:, :, pk[2]].astype(np.float32),
               line_k=rms[pk[0], pk[1]].astype(np.float64),
               sensor_norm=np.array(np.linalg.norm(sens.astype(np.float64))), sensor_rows=sens[::ROW_STRIDE].astype(np.float32),
               index_rows=r['IndexSensorMap'][::ROW_STRIDE], nsensors=np.array(r['IndexSensorMap'].size),
               time=r['Sensor']['time'], steps=np.array(r['steps']), oracle_seconds=np.array(time.time() - t0),
               oracle_threads=np.array(r['threads']))
    path = os.path.join(HERE, 'ctx500_full.npz')
    np.savez_compressed(path, **out)
    print('ctx500_full', os.path.getsize(path) // 1024, 'KiB; peak', pk, float(rms[pk]), 'in %.0f s' % (time.time() - t0))


if __name__ == '__main__':
    main()
