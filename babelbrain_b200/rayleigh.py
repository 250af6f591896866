"""
Host-side mirror of ``BabelViscoFDTD.tools.RayleighAndBHTE`` for the functions on BabelBrain's hot
path: ``ForwardSimple`` (19 call sites, e.g. TranscranialModeling/BabelIntegrationSingle.py:295),
``InitCuda``/``InitOpenCL``/``InitMetal``/``InitMLX`` (BabelIntegrationBASE.py:918-925) plus the two
small host helpers the transducer files import (``SpeedofSoundWater``, ``GenerateFocusTx``).
"""
import ctypes
import numpy as np

from . import _capi

_state = {'device': 0, 'devices': [0], 'last_kernel_ms': None}


def _init(deviceName='B200', **_kw):
    """Select the CUDA device(s) whose name contains deviceName, like the reference's InitCuda (an unknown name raises;
    BABELB200_ANY_DEVICE=1 makes the name advisory).  All matching devices are remembered: ForwardSimple shards its
    field points over the first BABELB200_NGPUS of them."""
    from .propagation import _matching_devices
    hits = _matching_devices(deviceName)
    if not hits:
        raise _capi.BabelB200Error('no CUDA device visible: babelbrain_b200 has no CPU fallback')
    _state['device'], _state['devices'] = hits[0], hits
    return _capi.device_names()[hits[0]]


InitCuda = _init
InitOpenCL = _init
InitMetal = _init
InitMLX = _init


def ForwardSimple(cwvnb, center, ds, u0, rf, MaxDistance=-1.0, u0step=0, MacOsPlatform='Metal', deviceMetal='B200',
                  NumberGPUs=None):
    """Rayleigh integral from N_src sub-elements to N_pts field points; returns complex64 (N_pts,).
    cwvnb complex wavenumber; center (N_src,3); ds (N_src,) or (N_src,1); u0 (N_src,) complex;
    rf (N_pts,3).  With u0step != 0, u0 holds one amplitude set per field point (N_pts*N_src).

    NumberGPUs (extension; default: environment variable BABELB200_NGPUS, else 1): the field points are independent, so
    they are cut into that many contiguous shares, one per GPU selected by InitCuda, computed concurrently (no
    exchange step: every GPU needs all sources and only its own points)."""
    import os
    _capi.require_gpu()
    k = complex(np.asarray(cwvnb).reshape(-1)[0])
    center = np.ascontiguousarray(center, dtype=np.float32)
    rf = np.ascontiguousarray(rf, dtype=np.float32)
    if center.ndim != 2 or center.shape[1] != 3 or rf.ndim != 2 or rf.shape[1] != 3:
        raise ValueError('center and rf must be (N,3) arrays')
    ds = np.ascontiguousarray(np.asarray(ds).reshape(-1), dtype=np.float32)
    u0 = np.ascontiguousarray(np.asarray(u0).reshape(-1), dtype=np.complex64)
    nsrc, npts = center.shape[0], rf.shape[0]
    if u0step != 0:
        if not (u0step == nsrc and u0.shape[0] == nsrc * npts and ds.shape[0] == nsrc):
            raise ValueError('u0step must equal the number of sources and u0 hold N_pts*N_src values')
    elif not (ds.shape[0] == nsrc and u0.shape[0] == nsrc):
        raise ValueError('center, ds and u0 must describe the same number of sources')
    out = np.empty(npts, np.complex64)
    ngpu = int(NumberGPUs if NumberGPUs is not None else os.environ.get('BABELB200_NGPUS', 1))
    devices = (_state['devices'] or [_state['device']])[:max(1, ngpu)]
    if ngpu > len(devices):
        raise ValueError('NumberGPUs=%d but only %d selected CUDA device(s)' % (ngpu, len(devices)))
    L = _capi.lib()

    def share(dev, lo, hi):
        ms = ctypes.c_double(0.0)
        u = u0 if u0step == 0 else u0[lo * nsrc:hi * nsrc]
        _capi.check(L.bb_rayleigh_forward(k.real, k.imag, nsrc, _capi.ptr(center), _capi.ptr(ds), _capi.ptr(u),
                                          hi - lo, _capi.ptr(rf[lo:hi]), _capi.ptr(out[lo:hi]), float(MaxDistance), int(u0step),
                                          dev, ctypes.byref(ms)))
        return ms.value
    if len(devices) == 1 or npts < 4096 * len(devices):
        _state['last_kernel_ms'] = share(devices[0], 0, npts)
        return out
    import threading
    cuts = [npts * r // len(devices) for r in range(len(devices) + 1)]
    times, errors = [0.0] * len(devices), []

    def work(r):
        try:
            times[r] = share(devices[r], cuts[r], cuts[r + 1])
        except BaseException as e:  # noqa: BLE001 -- re-raised in the calling thread
            errors.append(e)
    th = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise errors[0]
    _state['last_kernel_ms'] = max(times)
    return out


def SpeedofSoundWater(Temperature):
    """Speed of sound in water (m/s) vs temperature in Celsius (Marczak 1997 fifth-order fit)."""
    T = float(Temperature)
    return (1.402385e3 + 5.038813 * T - 5.799136e-2 * T ** 2 + 3.287156e-4 * T ** 3
            - 1.398845e-6 * T ** 4 + 2.787860e-9 * T ** 5)


def GenerateFocusTx(f, Foc, Diam, c, PPWSurface=4):
    """Spherical-cap transducer decomposed into sub-elements for the Rayleigh integral.  Returns a
    dict with 'center' (N,3), 'ds' (N,1), 'normal' (N,3), 'VertDisplay' (M,3), 'FaceDisplay',
    'elemcenter' (1,3) as the transducer files expect (BabelIntegrationSingle.py:241-247).
    The cap points towards +Z with its focus at the origin side: apex at z=-Foc, focus at z=0."""
    lam = c / f
    step = lam / PPWSurface
    Foc, Diam = float(Foc), float(Diam)
    betamax = np.arcsin(min(1.0, Diam / 2.0 / Foc))
    nring = max(1, int(np.ceil(Foc * betamax / step)))
    dbeta = betamax / nring
    centers, areas, normals, verts, faces = [], [], [], [], []
    for nr in range(nring):
        b0, b1 = nr * dbeta, (nr + 1) * dbeta
        bm = 0.5 * (b0 + b1)
        ring_area = 2.0 * np.pi * Foc ** 2 * (np.cos(b0) - np.cos(b1))
        nseg = max(1, int(np.ceil(2.0 * np.pi * Foc * np.sin(bm) / step))) if nr > 0 else max(3, int(np.ceil(2 * np.pi * Foc * np.sin(bm) / step)))
        th = (np.arange(nseg) + 0.5) * 2.0 * np.pi / nseg
        x = Foc * np.sin(bm) * np.cos(th)
        y = Foc * np.sin(bm) * np.sin(th)
        z = -Foc * np.cos(bm) * np.ones_like(th)
        c3 = np.stack([x, y, z], axis=1)
        centers.append(c3)
        areas.append(np.full(nseg, ring_area / nseg))
        normals.append(-c3 / Foc)
        t0 = np.arange(nseg) * 2.0 * np.pi / nseg
        for bb in (b0, b1):
            verts.append(np.stack([Foc * np.sin(bb) * np.cos(t0), Foc * np.sin(bb) * np.sin(t0), -Foc * np.cos(bb) * np.ones_like(t0)], axis=1))
    center = np.concatenate(centers).astype(np.float32)
    ds = np.concatenate(areas).astype(np.float32).reshape(-1, 1)
    Tx = {'center': center, 'ds': ds, 'normal': np.concatenate(normals).astype(np.float32),
          'VertDisplay': np.concatenate(verts).astype(np.float32),
          'FaceDisplay': np.zeros((0, 4), np.int64),
          'elemcenter': np.array([[0.0, 0.0, -Foc]], np.float32), 'elemdims': np.array([[center.shape[0]]])}
    return Tx


# the bio-heat functions of the same upstream module (thermal step)
from .thermal import BHTE, BHTEMultiplePressureFields  # noqa: E402,F401
