"""Test harness: drives the UNMODIFIED reference caller (TranscranialModeling/BabelIntegrationBASE.py RUN_SIM_BASE.RunCases
-> Step1..Step5 -> ReturnResults, with BabelIntegrationSingle.py on top) headless through this repo's BabelViscoFDTD shim.

The reference sources are imported from where they lie -- /root/reference in the build container, or baseline/_ref
(git-ignored, travels to the GPU box; populated by `python tests/make_ref_install.py`) -- never copied into the
repository history.  The reference's other dependencies are absent from this image (nibabel, SimpleITK, h5py, linetimer,
pwlf, matplotlib, numpy-stl, trimesh: SURVEY.md Appendix B), so sys.modules gets minimal stand-ins: only what the code
paths Step1-Step5 + ReturnResults touch (a NIfTI object with get_fdata / header.get_zooms / affine, a no-op CodeTimer).
Steps 9-10 of RunCases (plots, NIfTI/HDF5 files: needs the real I/O libraries) are replaced at run time by a function
that calls the reference's own ReturnResults and keeps what it returns; no reference file is edited.
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ('/root/reference', os.path.join(ROOT, 'baseline', '_ref'))


def reference_root():
    for c in ((os.environ['BB_REF_ROOT'],) if os.environ.get('BB_REF_ROOT') else CANDIDATES):
        if os.path.isfile(os.path.join(c, 'TranscranialModeling', 'BabelIntegrationBASE.py')):
            return c
    return None


class FakeHeader:
    def __init__(self, zooms):
        self._z = tuple(float(z) for z in zooms)

    def get_zooms(self):
        return self._z


class FakeNifti:
    """What UpdateConditions / Step1 read of a nibabel image (BabelIntegrationBASE.py:1160-1166, :1844-1848)."""

    def __init__(self, data, zooms_mm):
        self._data = np.asarray(data, dtype=np.float64)
        self.header = FakeHeader(zooms_mm)
        self.affine = np.diag([zooms_mm[0], zooms_mm[1], zooms_mm[2], 1.0])

    def get_fdata(self):
        return self._data


class _CodeTimer(contextlib.ContextDecorator):
    def __init__(self, name='', unit='s', **_kw):
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install_stubs(mask):
    """sys.modules stand-ins for the reference's absent dependencies; nibabel.load returns `mask` whatever the path."""
    def unavailable(*a, **k):
        raise RuntimeError('not available in the headless harness')
    plt = _module('matplotlib.pyplot', figure=unavailable, imshow=unavailable, plot=unavailable, cm=types.SimpleNamespace(jet=None, gray=None))
    ticker = _module('matplotlib.ticker')
    mpl = _module('matplotlib', pyplot=plt, ticker=ticker)
    stubs = {
        'matplotlib': mpl, 'matplotlib.pyplot': plt, 'matplotlib.ticker': ticker,
        'nibabel': _module('nibabel', load=lambda path: mask, Nifti1Image=unavailable),
        'SimpleITK': _module('SimpleITK'), 'h5py': _module('h5py'), 'pwlf': _module('pwlf'),
        'linetimer': _module('linetimer', CodeTimer=_CodeTimer),
        'stl': _module('stl', mesh=_module('stl.mesh')), 'stl.mesh': _module('stl.mesh'),
        'trimesh': _module('trimesh', creation=_module('trimesh.creation')), 'trimesh.creation': _module('trimesh.creation'),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    for k, v in stubs.items():
        try:
            importlib.import_module(k)          # the genuine package wins when it is installed
        except Exception:
            sys.modules[k] = v
    return saved


def pichardo_table(path):
    """Stand-in for ReadFromH5py('MapPichardo.h5') (BabelIntegrationBASE.py:61-69): the density -> speed / attenuation
    table of the CT mapping, read at import time and not used by the runs of this harness (no CT)."""
    rho = np.linspace(1000.0, 3000.0, 5)
    freq = np.linspace(0.2, 1.2, 4)
    sos = 1500.0 + (rho[:, None] - 1000.0) * 0.9 + 0.0 * freq[None, :]
    att = 5.0 + (rho[:, None] - 1000.0) * 0.05 * freq[None, :]
    return {'rho': rho, 'freq': freq, 'MapSoS': sos, 'MapAtt': att}


def water_mask(shape=(100, 100, 110), h_mm=0.735, skin_z=44, focus=(50, 50, 60)):
    """A label volume like *_BabelViscoInput.nii.gz (labels 0 water, 4 brain, exactly one voxel 5 = target,
    BabelDatasetPreps.py:766-772): tissue from plane skin_z on, in the flipped-z convention of the file."""
    D = np.zeros(shape, np.float64)
    D[:, :, skin_z:] = 4.0
    D[focus] = 5.0
    return FakeNifti(np.flip(D, axis=2), (h_mm, h_mm, h_mm))


def load_reference(mask, transducer='BabelIntegrationSingle'):
    """Import the unmodified reference modules with the shim first on sys.path.  Returns (BASE module, Tx module)."""
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError('no reference tree: neither /root/reference nor baseline/_ref (python tests/make_ref_install.py)')
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    install_stubs(mask)
    import BabelViscoFDTD.H5pySimple as H5
    if not getattr(H5, '_harness_patched', False):
        genuine = H5.ReadFromH5py

        def ReadFromH5py(f, *a, **k):
            if os.path.basename(str(f)) == 'MapPichardo.h5':
                try:
                    return genuine(f, *a, **k)
                except Exception:
                    return pichardo_table(f)
            return genuine(f, *a, **k)
        H5.ReadFromH5py = ReadFromH5py
        H5._harness_patched = True
    for k in [k for k in sys.modules if k == 'TranscranialModeling' or k.startswith('TranscranialModeling.')]:
        del sys.modules[k]
    if ref not in sys.path:
        sys.path.append(ref)                   # after the repo root: BabelViscoFDTD resolves to the shim
    pkg = types.ModuleType('TranscranialModeling')     # the package's own __init__ may import GUI-side modules
    pkg.__path__ = [os.path.join(ref, 'TranscranialModeling')]
    sys.modules['TranscranialModeling'] = pkg
    old = np.geterr()
    base = importlib.import_module('TranscranialModeling.BabelIntegrationBASE')
    tx = importlib.import_module('TranscranialModeling.' + transducer)
    sys.modules['nibabel'].load = lambda path: mask
    return base, tx, old


def run_cases(mask, captured, transducer='BabelIntegrationSingle', patch=None, **kargs):
    """RUN_SIM().RunCases(**kargs) of the unmodified reference.  Steps 9-10 are replaced by a capture of the reference's
    own ReturnResults(); `captured` receives 'sim' (the SimulationConditions object) and 'results'.  patch(base, tx) may
    replace names inside the freshly imported reference modules (the CPU test answers the solver calls with the oracle)."""
    base, tx, old_err = load_reference(mask, transducer)
    if patch is not None:
        patch(base, tx)

    def step9(self):
        return None

    def step10(self, FILENAMES, subsamplingFactor=1, bMinimalSaving=False, bUseRayleighForWater=False, FILENAMESWater=None):
        captured['sim'] = self._SIM_SETTINGS
        captured['results'] = self._SIM_SETTINGS.ReturnResults(bDoRefocusing=self._bDoRefocusing, bUseRayleighForWater=bUseRayleighForWater)
        return FILENAMES['DataForSim']
    base.BabelFTD_Simulations_BASE.Step9_PrepAndPlotData = step9
    base.BabelFTD_Simulations_BASE.Step10_GetResults = step10
    base.bGPU_INITIALIZED = False
    try:
        return tx.RUN_SIM().RunCases(**kargs)
    finally:
        np.seterr(**old_err)                   # the reference sets np.seterr(divide='raise') at import (BASE.py:11)


def load_thermal():
    """Import the unmodified ThermalModeling/CalculateTemperatureEffects.py (the driver of the thermal step) with the shim
    first on sys.path and the stand-ins above for matplotlib / linetimer."""
    ref = reference_root()
    if ref is None or not os.path.isfile(os.path.join(ref, 'ThermalModeling', 'CalculateTemperatureEffects.py')):
        raise FileNotFoundError('no reference ThermalModeling: neither /root/reference nor baseline/_ref (python tests/make_ref_install.py)')
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    install_stubs(None)
    for k in [k for k in sys.modules if k == 'ThermalModeling' or k.startswith('ThermalModeling.')]:
        del sys.modules[k]
    if ref not in sys.path:
        sys.path.append(ref)
    pkg = types.ModuleType('ThermalModeling')
    pkg.__path__ = [os.path.join(ref, 'ThermalModeling')]
    sys.modules['ThermalModeling'] = pkg
    return importlib.import_module('ThermalModeling.CalculateTemperatureEffects')
