"""The unmodified reference caller driven through the BabelViscoFDTD shim:
RUN_SIM().RunCases -> Step1 (UpdateConditions, CalculateMatricesForPropagation) -> Step2 (ForwardSimple over the whole
grid) -> Step3 (CreateSources, CreateSensorMap) -> Step4 (StaggeredFDTD_3D_with_relaxation, dispersion correction,
RMS -> amplitude) -> Step5 (CalculatePhaseData) -> ReturnResults, on a synthetic water-only label volume.

Acceptance = the reference authors' own criterion for this configuration
(OfflineBatchExamples/CompareRayleightWithFDTD/SummaryAnalysis.xlsx, 309 water-only cases): the
FDTD pressure amplitude agrees with the Rayleigh integral -- peak difference -0.6 ... +3.9 %, L2 mean 5 % (max 24 %),
focal maximum within 0.3 mm on average.  The absolute scale of that comparison rests on the caller's dispersion-correction
polynomial (BabelIntegrationBASE.py:1674, applied at :2433-2440), which was fitted to BabelViscoFDTD's output: a solver
whose source gain or pressure definition differed from upstream's by more than a few per cent fails it.

  * CPU (here): the solver call and ForwardSimple are answered by the ORACLE, which pins the oracle's absolute scale
    and conventions to the reference's calibration; CalculateMatricesForPropagation is the product's host code.
  * GPU (-m gpu): the same run with nothing replaced: the CUDA path behind the shim.
Both need the reference tree (/root/reference here, baseline/_ref on the GPU box: python tests/make_ref_install.py).
"""
import numpy as np
import pytest

from tests import refcaller

needs_ref = pytest.mark.skipif(refcaller.reference_root() is None,
                               reason='no reference tree (/root/reference or baseline/_ref: python tests/make_ref_install.py)')
KARGS = dict(targets=['T'], ID='S', basedir='/nonexistent/', deviceName='B200', Frequencies=[250e3], basePPW=[6],
             bTightNarrowBeamDomain=True, bDoRefocusing=False, bWaterOnly=True, bMinimalSaving=True, bForceRecalc=True,
             bDisplay=False, Aperture=64e-3, FocalLength=63.2e-3)


def band(captured):
    """FDTD amplitude vs Rayleigh amplitude on the reference's own result volumes (ReturnResults)."""
    RayleighWater, _, Full, _, DataForSim, _, _, _, _ = captured['results']
    sim = captured['sim']
    a, r = np.asarray(Full, np.float64), np.asarray(RayleighWater, np.float64)
    sel = (a > 0) & (r > 0)
    assert sel.sum() > 0.3 * a.size
    # away from the source plane (the first wavelength holds the near field of the discretised source)
    zs = np.zeros(a.shape, bool)
    zs[:, :, :a.shape[2] - 10] = True           # volumes are back in the file's flipped-z convention: the source side is the far end
    sel &= zs
    pa, pr = np.unravel_index(np.argmax(np.where(sel, a, 0)), a.shape), np.unravel_index(np.argmax(np.where(sel, r, 0)), r.shape)
    peak_diff = a[pa] / r[pr] - 1.0
    core = sel & (r > 0.25 * r[pr])
    l2 = np.linalg.norm((a - r)[core]) / np.linalg.norm(r[core])
    dist_mm = np.linalg.norm((np.array(pa) - np.array(pr)) * sim._SpatialStep * 1e3)
    return dict(peak_diff=float(peak_diff), l2=float(l2), focal_distance_mm=float(dist_mm), peak_fdtd=float(a[pa]), peak_rayleigh=float(r[pr]),
                shape=(sim._N1, sim._N2, sim._N3), ppp=int(sim._PPP), steps=int(round(sim._TimeSimulation / sim._TemporalStep)),
                cfl_water=float(sim._TemporalStep / sim.DominantMediumTemporalStep))


def check(b):
    # reference data for this transducer / frequency / resolution (27 cases "Single_250kHz_6PPW" of SummaryAnalysis.xlsx):
    # peak difference +0.18 ... +2.53 % (mean +0.91 %), L2 1.9 ... 5 % (mean 3.2 %), focal maximum at the same voxel
    assert -0.006 <= b['peak_diff'] <= 0.026, b
    assert b['l2'] <= 0.05, b
    assert b['focal_distance_mm'] <= 0.8, b            # one voxel


@needs_ref
def test_unmodified_caller_with_the_oracle_behind_the_solver_call():
    import oracle
    oracle.build()
    calls = {}

    def solver(MaterialMap, MaterialList, Frequency, SourceMap, SourceFunctions, SpatialStep, TimeSimulation, SensorMap, **kw):
        calls['solver'] = dict(kw, shapes=(MaterialMap.shape, SourceFunctions.shape), dtypes=(MaterialMap.dtype, SourceMap.dtype, SensorMap.dtype))
        keep = ('Ox', 'Oy', 'Oz', 'NDelta', 'DT', 'ReflectionLimit', 'AlphaCFL', 'TypeSource', 'QfactorCorrection', 'QCorrection',
                'SelRMSorPeak', 'SelMapsRMSPeakList', 'SelMapsSensorsList', 'SensorSubSampling', 'SensorStart', 'ReflectorMask')
        r = oracle.run_c(MaterialMap, MaterialList, Frequency, SourceMap, SourceFunctions, SpatialStep, TimeSimulation, SensorMap,
                         **{k: v for k, v in kw.items() if k in keep})
        return r['Sensor'], {}, r['RMS'], {'IndexSensorMap': r['IndexSensorMap']}

    def forward_simple(cwvnb, center, ds, u0, rf, **kw):
        calls['rayleigh'] = (center.shape, rf.shape, np.asarray(u0).dtype)
        return oracle.rayleigh_c(cwvnb, center, ds, u0, rf)

    cap = {}

    def patch(base, tx):
        base.PModel.StaggeredFDTD_3D_with_relaxation = solver        # instance attribute: CalculateMatricesForPropagation stays the product's
        tx.ForwardSimple = forward_simple
    refcaller.run_cases(refcaller.water_mask(), cap, patch=patch, COMPUTING_BACKEND=0, **KARGS)
    # what the caller handed over: the contract of SURVEY.md 8(a) F0
    s = calls['solver']
    assert s['dtypes'] == (np.uint32, np.uint32, np.uint32) and s['NDelta'] == 12 and s['SelRMSorPeak'] == 1
    assert s['SelMapsSensorsList'] == ['Pressure'] and s['TypeSource'] == 0 and s['USE_SINGLE'] is True
    assert calls['rayleigh'][1][0] == int(np.prod(s['shapes'][0]))          # the Rayleigh field covers every grid point
    b = band(cap)
    print('unmodified caller + oracle:', b)
    assert b['ppp'] == 30 and b['steps'] % b['ppp'] == 0
    check(b)
    # the dictionary the caller would write as *_DataForSim.h5 (ReturnResults' DataForSim, BASE.py:2815-2884; Step10 saves it
    # with SaveToH5py, the thermal step reads it back with ReadFromH5py): through the shim's HDF5 path of this image
    import os, tempfile
    from BabelViscoFDTD.H5pySimple import SaveToH5py, ReadFromH5py
    data = cap['results'][4]
    assert isinstance(data, dict) and 'p_amp' in data and 'MaterialMap' in data
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'Single_DataForSim.h5')
        SaveToH5py(data, path)
        back = ReadFromH5py(path)
    assert sorted(back) == sorted(str(k) for k in data)
    for k, v in data.items():
        if isinstance(v, np.ndarray) and v.dtype.kind != 'O':
            assert np.array_equal(back[str(k)], v), k


@needs_ref
@pytest.mark.gpu
def test_unmodified_caller_through_the_cuda_path():
    cap = {}
    refcaller.run_cases(refcaller.water_mask(), cap, COMPUTING_BACKEND=1, **KARGS)
    b = band(cap)
    print('unmodified caller + CUDA path:', b)
    check(b)
    # the same run with the skull materials (homogeneous attenuating medium test mode of the caller, BASE.py:904-912,:1306-1312)
    cap2 = {}
    refcaller.run_cases(refcaller.water_mask(), cap2, COMPUTING_BACKEND=1, **dict(KARGS, bWaterOnly=False, bForceHomogenousMedium=True))
    full = np.asarray(cap2['results'][2])
    assert np.isfinite(full).all() and full.max() > 0 and full.max() < np.asarray(cap['results'][2]).max()      # 5 Np/m of attenuation lowers the focus


CTX500 = dict(targets=['T'], ID='S', basedir='/nonexistent/', deviceName='B200', Frequencies=[500e3], basePPW=[6],
              bTightNarrowBeamDomain=True, bDoRefocusing=False, bWaterOnly=True, bMinimalSaving=True, bForceRecalc=True, bDisplay=False,
              ZSteering=0.0, Aperture=64.0e-3, FocalLength=62.94e-3,                                   # BabelBrain/Babel_CTX500/default.yaml
              InDiameters=np.array([0.0, 31.6988e-3, 44.2688e-3, 53.6688e-3]), OutDiameters=np.array([31.14e-3, 43.71e-3, 53.11e-3, 60.83e-3]))


@needs_ref
@pytest.mark.gpu
def test_unmodified_annular_array_caller_ctx500_through_the_cuda_path():
    """BASELINE configs[1]'s own caller: BabelIntegrationANNULAR_ARRAY.py (CTX-500: four rings, 500 kHz, 6 PPW) unmodified --
    per-ring phase programming with ForwardSimple on a single point (:376-395), the whole-grid Rayleigh field (:403-420),
    sources, the solver call, the dispersion correction -- on the CUDA path, water only.  Reference data for this
    configuration (12 cases "CTX_500_500kHz_6PPW" of SummaryAnalysis.xlsx): peak difference -0.15 ... +1.18 % (mean
    +0.56 %), L2 mean 3.2 %."""
    cap = {}
    refcaller.run_cases(refcaller.water_mask(shape=(150, 150, 190), h_mm=0.3675, skin_z=90, focus=(75, 75, 120)), cap,
                        transducer='BabelIntegrationANNULAR_ARRAY', COMPUTING_BACKEND=1, **CTX500)
    b = band(cap)
    print('unmodified CTX-500 caller + CUDA path:', b)
    assert b['ppp'] == 30
    assert -0.01 <= b['peak_diff'] <= 0.02, b
    assert b['l2'] <= 0.05, b
    assert b['focal_distance_mm'] <= 0.4, b            # one voxel
