"""
Synthetic solver-call arguments shaped like the ones BabelBrain builds (no imaging data needed).
Sizing follows the caller's own rules: wavelength from the 1102.5 m/s shear floor
(TranscranialModeling/BabelIntegrationBASE.py:170-182,1776-1793), PPP snapping (:1809-1828),
simulation length (:2082-2090), sensor sub-sampling and start (:2092-2109), material rows
(:100-167), source plane / CW pulse with a 4-cycle ramp (BabelIntegrationSingle.py:313-346),
sensor window (BabelIntegrationBASE.py:2283-2290) and particle-source weights (:2333-2335).
Used by bench.py, __graft_entry__.smoke() and the parity tests ("data": "synthetic").
"""
import numpy as np

SHEAR_FLOOR_SOS = 1102.515  # DensityToSSoSPichardo(1000.0), BabelIntegrationBASE.py:626-644


def _linfit(f, f0, v0, f1, v1):
    return np.round(v0 + (v1 - v0) * (f - f0) / (f1 - f0))


def material_rows(frequency):
    """[rho, cL, cS, attL, attS] rows for Water, Skin, Cortical, Trabecular, Brain at `frequency`."""
    f = float(frequency)
    att_shear = np.round((57.0 / .27 + 373 / 0.836) / 2 * (f / 1e6))
    cort = [1896.5, _linfit(f, 270e3, 2448.0, 836e3, 2516.0),
            _linfit(f, 270e3, np.mean([1577.0, 1498.0, 1313.0]), 836e3, np.mean([1758.0, 1674.0, 1545.0])),
            np.round(203.25090263 * (f / 1e6) * 0.8), att_shear]
    trab = [1738.0, _linfit(f, 270e3, 2140.0, 836e3, 2300.0),
            _linfit(f, 270e3, np.mean([1227.0, 1365.0, 1200.0]), 836e3, np.mean([1574.0, 1252.0, 1327.0])),
            np.round(202.76362433 * (f / 1e6) * 0.8), att_shear]
    return {'Water': [1000.0, 1500.0, 0.0, 0.0, 0.0], 'Skin': [1116.0, 1537.0, 0.0, 2.3 * f / 500e3, 0.0],
            'Cortical': cort, 'Trabecular': trab, 'Brain': [1041.0, 1562.0, 0.0, 3.45 * f / 500e3, 0.0]}


def snap_ppp(ppp):
    special = {31: 32, 34: 35, 23: 24, 71: 72, 74: 75, 79: 80, 47: 48}
    ppp = int(ppp)
    if ppp in special:
        return special[ppp]
    if ppp % 5 != 0:
        return (ppp // 5 + 1) * 5
    return ppp


def sizing(frequency, ppw, MaterialList, shape, pml=12, alpha_cfl=0.5, cycles_at_end=2, periods=None):
    """h, dt, PPP, steps, SensorSubSampling, SensorStart exactly as UpdateConditions derives them."""
    ML = np.asarray(MaterialList, float)
    speeds = np.sort(ML[:, 1:3].flatten())
    sos = min(speeds[speeds > 0][0], SHEAR_FLOOR_SOS)
    h = sos / frequency / ppw
    dt_ideal = alpha_cfl * np.sqrt(3.0) / 3.0 * h / ML[:, 1].max()
    ppp = snap_ppp(np.ceil(1.0 / frequency / dt_ideal))
    dt = 1.0 / frequency / ppp
    if periods is None:
        dims = (np.array(shape) - 2 * pml) * h
        tsim = np.floor(np.sqrt((dims ** 2).sum()) / ML[0, 1] / dt) * dt
        nt = np.arange(0.0, tsim, dt).shape[0]
        steps = (int(nt / ppp) + 1) * ppp
    else:
        steps = int(periods) * ppp
    divisors = np.array([x for x in range(1, ppp) if ppp % x == 0])
    divisors = divisors[ppp / divisors >= 4]
    sub = int(divisors[-1])
    sensor_start = int((steps - int(cycles_at_end * ppp)) / sub)
    return dict(h=h, dt=dt, ppp=ppp, steps=steps, sub=sub, sensor_start=sensor_start, dt_ideal=dt_ideal)


def _smooth_field(rng, n1, n2, amp):
    coarse = rng.standard_normal((6, 6))
    xi = np.linspace(0, 5, n1)
    yi = np.linspace(0, 5, n2)
    x0 = np.clip(np.floor(xi).astype(int), 0, 4)
    y0 = np.clip(np.floor(yi).astype(int), 0, 4)
    fx = (xi - x0)[:, None]
    fy = (yi - y0)[None, :]
    c = coarse
    f = (c[x0][:, y0] * (1 - fx) * (1 - fy) + c[x0 + 1][:, y0] * fx * (1 - fy)
         + c[x0][:, y0 + 1] * (1 - fx) * fy + c[x0 + 1][:, y0 + 1] * fx * fy)
    return amp * f / max(np.abs(f).max(), 1e-9)


def skull_labels(shape, h, pml, seed=1234, skin_depth_frac=0.14, planes=None, tissue_in_shell=False):
    """Labels 0 water, 1 skin, 2 cortical, 3 trabecular, 4 brain: spherical shell (outer radius
    85 mm; 1.5 mm skin; 2/2/2 mm cortical/trabecular/cortical) centred below the domain, radius
    perturbed by a smooth +-1 mm field so interfaces are not grid aligned.

    As in the caller, the absorbing shell itself is water: UpdateConditions starts from an all-zero map and writes
    the tissue mask only into [XLOffset:-XROffset, YLOffset:-YROffset, ZLOffset:-ZROffset] with every offset >= the PML
    thickness (BabelIntegrationBASE.py:1853-1862, :2110, :2154-2159).  tissue_in_shell=True keeps the labels in the
    shell (a map BabelBrain never builds; used to exercise the multi-axial layer, MPMLRatio)."""
    n1, n2, n3 = shape
    rng = np.random.default_rng(seed)
    x = (np.arange(n1) - n1 / 2 + 0.5) * h
    y = (np.arange(n2) - n2 / 2 + 0.5) * h
    z = np.arange(n3) * h
    r_out = 0.085
    zc = (pml + skin_depth_frac * n3) * h + r_out
    dr = _smooth_field(rng, n1, n2, 1e-3)
    if planes is not None:          # only planes [lo,hi) of axis 0 (one slab of a decomposed run)
        x, dr = x[planes[0]:planes[1]], dr[planes[0]:planes[1]]
    lab = np.zeros((x.size, n2, n3), np.uint32)
    rxy2 = (x[:, None] ** 2 + y[None, :] ** 2)
    for k0 in range(0, n3, 64):
        zz = z[k0:k0 + 64]
        r = np.sqrt(rxy2[:, :, None] + (zz[None, None, :] - zc) ** 2) + dr[:, :, None]
        l = np.zeros(r.shape, np.uint8)
        l[r < r_out] = 1
        l[r < r_out - 1.5e-3] = 2
        l[r < r_out - 3.5e-3] = 3
        l[r < r_out - 5.5e-3] = 2
        l[r < r_out - 7.5e-3] = 4
        lab[:, :, k0:k0 + 64] = l
    if not tissue_in_shell:
        glo = 0 if planes is None else planes[0]
        gi = np.arange(glo, glo + x.size)
        lab[(gi < pml) | (gi >= n1 - pml)] = 0
        lab[:, :pml] = 0
        lab[:, n2 - pml:] = 0
        lab[:, :, n3 - pml:] = 0
    lab[:, :, :pml + 1] = 0  # BabelIntegrationBASE.py:2201
    return lab


def cw_source_object(amp, phase, frequency, dt, steps, ramp_cycles=4):
    from .sources import CWSourceFunctions
    return CWSourceFunctions(amp, phase, frequency, dt, dt * steps, ramp_cycles)


def cw_sources(amp, phase, frequency, dt, steps, ramp_cycles=4):
    """(Nsrc, Nt) float64 pulse table, as CreateSources builds it (BabelIntegrationSingle.py:313-346; the formula lives
    in sources.CWSourceFunctions, pinned against the reference's method by tests/test_sources.py)."""
    return cw_source_object(amp, phase, frequency, dt, steps, ramp_cycles).dense()


CONFIGS = {
    # name: frequency, ppw, shape, medium, source kind, RMS maps, SelRMSorPeak
    'single_water': dict(frequency=250e3, ppw=6, shape=(120, 120, 160), medium='water', source='plane',
                         aperture=0.064, focal=0.0632),
    'ctx500_skull': dict(frequency=500e3, ppw=6, shape=(240, 240, 320), medium='skull', source='plane',
                         aperture=0.064, focal=0.0632),
    'h317_skull': dict(frequency=650e3, ppw=9, shape=(600, 600, 500), medium='skull', source='plane',
                       aperture=0.16, focal=0.135, rms_maps=['Pressure', 'Sigmaxx', 'Sigmayy', 'Sigmazz', 'Vx', 'Vy', 'Vz'],
                       sel_rms_peak=3),
    'dome_stress': dict(frequency=650e3, ppw=6, shape=(700, 700, 500), medium='skull', source='dome'),
    'hires_1mhz': dict(frequency=1e6, ppw=9, shape=(1080, 1080, 1080), medium='skull', source='plane',
                       aperture=0.064, focal=0.0632),
}


def make_workload(name='ctx500_skull', shape=None, periods=None, seed=1234, pml=12, amplitude=1e5, planes=None, lean=False,
                  dense_sources=True, tissue_in_shell=False):
    """Returns dict(args=tuple of the 8 positional arguments, kwargs=dict of the keyword arguments
    of StaggeredFDTD_3D_with_relaxation as BabelIntegrationBASE.py:2338-2365 passes them, meta=...).
    planes=(lo,hi): materialise only planes [lo,hi) of axis 0 of every volume (a slab with its halo, for
    FdtdSlab(origin=lo, n1_global=shape[0])); the source table then holds this slab's rows only.  lean=True
    makes Ox/Oy/Oz read-only broadcast views instead of dense float64 volumes (plane sources only).  dense_sources=False
    passes the sources as a sources.CWSourceFunctions object instead of the (Nsrc, Nt) float64 table."""
    cfg = dict(CONFIGS[name])
    if shape is not None:
        cfg['shape'] = tuple(int(s) for s in shape)
    f = cfg['frequency']
    rows = material_rows(f)
    if cfg['medium'] == 'water':
        ML = np.array([rows['Water']])
        qcorr = 1.0
    else:
        ML = np.array([rows[k] for k in ('Water', 'Skin', 'Cortical', 'Trabecular', 'Brain')])
        qcorr = np.array([1.0, 1.0, 3.0, 3.0, 1.0])
    n1, n2, n3 = cfg['shape']
    # the spatial step uses the full material set of the frequency (cortical/trabecular shear included)
    all_rows = np.array(list(rows.values()))
    S = sizing(f, cfg['ppw'], all_rows if cfg['medium'] != 'water' else ML, cfg['shape'], pml=pml, periods=periods)
    if cfg['medium'] == 'water':  # dt still limited by the fastest material BabelBrain would load
        S = sizing(f, cfg['ppw'], ML, cfg['shape'], pml=pml, periods=periods)
    h, dt, steps = S['h'], S['dt'], S['steps']
    lo, hi = (0, n1) if planes is None else (int(planes[0]), int(planes[1]))
    if planes is not None and cfg['source'] != 'plane':
        raise ValueError('planes= is implemented for plane sources')
    lshape = (hi - lo, n2, n3)
    MaterialMap = np.zeros(lshape, np.uint32) if cfg['medium'] == 'water' else skull_labels(cfg['shape'], h, pml, seed, planes=planes, tissue_in_shell=tissue_in_shell)
    SourceMap = np.zeros(lshape, np.uint32)
    kw = dict(NDelta=pml, DT=dt, ReflectionLimit=1e-5, COMPUTING_BACKEND=1, USE_SINGLE=True,
              SelMapsRMSPeakList=list(cfg.get('rms_maps', ['Pressure'])), SelMapsSensorsList=['Pressure'],
              SelRMSorPeak=cfg.get('sel_rms_peak', 1), DefaultGPUDeviceName='B200', AlphaCFL=1.0,
              QfactorCorrection=True, QCorrection=qcorr, SensorSubSampling=S['sub'], SensorStart=S['sensor_start'],
              ReflectorMask=None)
    kwater = 2 * np.pi * f / 1500.0
    x = (np.arange(n1) - n1 / 2 + 0.5) * h
    y = (np.arange(n2) - n2 / 2 + 0.5) * h
    if cfg['source'] == 'plane':
        ilo, ihi = max(pml, lo), max(min(n1 - pml, hi), max(pml, lo))
        ii, jj = np.meshgrid(np.arange(ilo, ihi), np.arange(pml, n2 - pml), indexing='ij')
        ids = np.arange(1, ii.size + 1, dtype=np.uint32).reshape(ii.shape)
        SourceMap[ilo - lo:ihi - lo, pml:n2 - pml, pml] = ids
        focal = min(cfg['focal'], 0.75 * (n3 - 2 * pml) * h)
        r2 = x[ii] ** 2 + y[jj] ** 2
        dist = np.sqrt(r2 + focal ** 2)
        ap = min(cfg['aperture'], 0.8 * (min(n1, n2) - 2 * pml) * h)
        amp = amplitude * (0.98 * np.exp(-(np.sqrt(r2) / (ap / 2)) ** 8) + 0.02) * focal / dist
        # phase lag growing with the distance to the focal point: with sin(wt + phase) the outer pixels fire late, so the
        # beam of the benchmark workloads diverges (the cost per cell-update is the same); tests/test_oracle.py flips the
        # sign for its focusing cross-check against the Rayleigh integral
        phase = -kwater * (dist - focal)
        cw = cw_source_object(amp.reshape(-1), phase.reshape(-1), f, dt, steps)
        SF = cw.dense() if dense_sources else cw
        if lean:
            Ox = Oy = np.broadcast_to(np.float64(0.0), lshape)
            Oz = np.broadcast_to(np.float64(1.0 / (1000.0 * 1500.0)), lshape)
        else:
            Ox = np.zeros(lshape)
            Oy = np.zeros(lshape)
            Oz = np.ones(lshape) / (1000.0 * 1500.0)
        kw.update(Ox=Ox, Oy=Oy, Oz=Oz, TypeSource=0)
        zsrc = pml
    else:  # hemispherical dome of small volumetric stress sources inside the domain
        rng = np.random.default_rng(seed)
        nel = 1024
        rad = 0.42 * (min(n1, n2) - 2 * pml) * h
        zc = (pml + 4) * h + rad
        u = rng.random(nel)
        th = rng.random(nel) * 2 * np.pi
        ce = np.stack([rad * np.sqrt(1 - u ** 2) * np.cos(th), rad * np.sqrt(1 - u ** 2) * np.sin(th), zc - rad * u], axis=1)
        z = np.arange(n3) * h
        for e in range(nel):
            ci = int(np.argmin(np.abs(x - ce[e, 0])))
            cj = int(np.argmin(np.abs(y - ce[e, 1])))
            ck = int(np.argmin(np.abs(z - ce[e, 2])))
            ci, cj, ck = (int(np.clip(c, pml + 1, n - pml - 3)) for c, n in ((ci, n1), (cj, n2), (ck, n3)))
            SourceMap[ci:ci + 2, cj:cj + 2, ck:ck + 2] = e + 1
        dist = np.sqrt(ce[:, 0] ** 2 + ce[:, 1] ** 2 + (ce[:, 2] - zc) ** 2)
        cw = cw_source_object(np.full(nel, amplitude), -kwater * (dist - rad), f, dt, steps)
        SF = cw.dense() if dense_sources else cw
        MaterialMap[SourceMap > 0] = 0
        kw.update(Ox=np.array([1]), Oy=np.array([1]), Oz=np.array([1]), TypeSource=2)
        kw['SelMapsRMSPeakList'] = ['Pressure']
        zsrc = pml
    SensorMap = np.zeros(lshape, np.uint32)
    slo, shi = max(pml, lo), max(min(n1 - pml, hi), max(pml, lo))
    SensorMap[slo - lo:shi - lo, pml:-pml, zsrc + 1:-pml] = 1
    args = (MaterialMap, ML, f, SourceMap, SF, h, dt * steps, SensorMap)
    meta = dict(name=name, shape=cfg['shape'], cells=n1 * n2 * n3, steps=steps, ppp=S['ppp'], dt=dt, h=h,
                sub=S['sub'], sensor_start=S['sensor_start'], nsrc=SF.shape[0], cell_updates=n1 * n2 * n3 * steps,
                frequency=f, ppw=cfg['ppw'], pml=pml, planes=(lo, hi), cw_sources=cw)
    return dict(args=args, kwargs=kw, meta=meta)


def cell_classes(MaterialMap, MaterialList, pml=12):
    """Cell counts by traffic class, for the roofline's algorithmic-byte accounting:
    pml shell / interior solid / interior attenuating fluid / interior lossless fluid."""
    MM = np.asarray(MaterialMap)
    ML = np.asarray(MaterialList, float)
    inner = MM[pml:-pml, pml:-pml, pml:-pml]
    counts = np.bincount(inner.reshape(-1), minlength=ML.shape[0])
    solid = counts[ML[:, 2] > 0].sum()
    att_fluid = counts[(ML[:, 2] == 0) & (ML[:, 3] > 0)].sum()
    lossless = counts[(ML[:, 2] == 0) & (ML[:, 3] == 0)].sum()
    return dict(pml=int(MM.size - inner.size), solid=int(solid), att_fluid=int(att_fluid), lossless=int(lossless))
