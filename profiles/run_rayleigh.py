"""Rayleigh integral throughput: python profiles/run_rayleigh.py [npts] [nsrc]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from babelbrain_b200 import _capi
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
rng = np.random.default_rng(0)
center = (rng.random((nsrc, 3)).astype(np.float32) - 0.5) * 0.06
center[:, 2] = -0.03 - 0.01 * rng.random(nsrc).astype(np.float32)
ds = np.full(nsrc, 1e-6, np.float32)
u0 = np.stack([rng.random(nsrc), rng.random(nsrc)], 1).astype(np.float32)
rf = (rng.random((npts, 3)).astype(np.float32) - 0.5) * 0.08
rf[:, 2] = rng.random(npts).astype(np.float32) * 0.1
out = np.zeros((npts, 2), np.float32)
L = _capi.lib()
for att in (0.0, -3.0):
    for rep in range(2):
        ms = ctypes.c_double()
        t0 = time.time()
        _capi.check(L.bb_rayleigh_forward(2 * np.pi * 5e5 / 1500, att, nsrc, _capi.ptr(center), _capi.ptr(ds), _capi.ptr(u0), npts, _capi.ptr(rf), _capi.ptr(out), -1.0, 0, 0, ctypes.byref(ms)))
        wall = time.time() - t0
    print('k_im=%g: %d x %d = %.2e pairs: kernel %.2f ms -> %.3e pairs/s (wall %.3f s)' % (att, npts, nsrc, npts * nsrc, ms.value, npts * nsrc / ms.value * 1e3, wall))
# accuracy against float64 on a subset
sub = slice(0, 2000)
d = rf[sub, None, :].astype(np.float64) - center[None, :, :].astype(np.float64)
R = np.sqrt((d * d).sum(-1))
k = 2 * np.pi * 5e5 / 1500 - 3.0j
u = u0[:, 0].astype(np.float64) + 1j * u0[:, 1]
ref = 1j * k * ((ds[None, :] * np.exp(k.imag * R) / R * u[None, :] * np.exp(-1j * k.real * R)).sum(1)) / (2 * np.pi)
got = out[sub, 0] + 1j * out[sub, 1]
print('rel L2 vs float64 (k_im=-3): %.2e' % (np.linalg.norm(got - ref) / np.linalg.norm(ref)))
