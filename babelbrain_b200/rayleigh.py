"""
Host-side mirror of ``BabelViscoFDTD.tools.RayleighAndBHTE`` for the functions on BabelBrain's hot
path: ``ForwardSimple`` (19 call sites, e.g. TranscranialModeling/BabelIntegrationSingle.py:295),
``InitCuda``/``InitOpenCL``/``InitMetal``/``InitMLX`` (BabelIntegrationBASE.py:918-925) plus the two
small host helpers the transducer files import (``SpeedofSoundWater``, ``GenerateFocusTx``).
"""
import ctypes
import numpy as np

from . import _capi

_state = {'device': 0, 'last_kernel_ms': None}


def _init(deviceName='B200', **_kw):
    """Select the CUDA device whose name contains deviceName (advisory: falls back to device 0)."""
    names = _capi.device_names()
    hits = [n for n, s in enumerate(names) if isinstance(deviceName, str) and deviceName in s]
    _state['device'] = hits[0] if hits else 0
    return names[_state['device']] if names else None


InitCuda = _init
InitOpenCL = _init
InitMetal = _init
InitMLX = _init


def ForwardSimple(cwvnb, center, ds, u0, rf, MaxDistance=-1.0, u0step=0, MacOsPlatform='Metal', deviceMetal='B200'):
    """Rayleigh integral from N_src sub-elements to N_pts field points; returns complex64 (N_pts,).
    cwvnb complex wavenumber; center (N_src,3); ds (N_src,) or (N_src,1); u0 (N_src,) complex;
    rf (N_pts,3).  With u0step != 0, u0 holds one amplitude set per field point (N_pts*N_src)."""
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
    ms = ctypes.c_double(0.0)
    _capi.check(_capi.lib().bb_rayleigh_forward(k.real, k.imag, nsrc, _capi.ptr(center), _capi.ptr(ds), _capi.ptr(u0),
                                                npts, _capi.ptr(rf), _capi.ptr(out), float(MaxDistance), int(u0step),
                                                _state['device'], ctypes.byref(ms)))
    _state['last_kernel_ms'] = ms.value
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


def _not_on_hot_path(name):
    def f(*a, **k):
        raise NotImplementedError('%s (bio-heat step) is outside the FDTD/Rayleigh hot path this package implements; '
                                  'see DESIGN.md "out of scope"' % name)
    f.__name__ = name
    return f


BHTE = _not_on_hot_path('BHTE')
BHTEMultiplePressureFields = _not_on_hot_path('BHTEMultiplePressureFields')
