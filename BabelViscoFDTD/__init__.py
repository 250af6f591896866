"""Drop-in import surface: the names BabelBrain imports from the BabelViscoFDTD package
(SURVEY.md section 8b), re-exported from babelbrain_b200 (B200 CUDA path behind a C ABI).

This shim replaces the hot path only -- PropagationModel.StaggeredFDTD_3D_with_relaxation /
CalculateMatricesForPropagation, ForwardSimple, Init*, ListDevices.  Everything else the package
offers (the bio-heat functions BHTE / BHTEMultiplePressureFields used by
ThermalModeling/CalculateTemperatureEffects.py:14, the real H5pySimple) is taken from a genuine
BabelViscoFDTD installation found further down sys.path, when there is one: `upstream()` returns
that package (or None), loaded under the private name _upstream_BabelViscoFDTD."""
import importlib.util
import os
import sys

__version__ = '1.2.4+b200'
_HERE = os.path.dirname(os.path.abspath(__file__))
_upstream = False


def upstream():
    """The genuine BabelViscoFDTD package shadowed by this shim, or None."""
    global _upstream
    if _upstream is not False:
        return _upstream
    _upstream = None
    for entry in sys.path:
        cand = os.path.join(entry or '.', 'BabelViscoFDTD')
        init = os.path.join(cand, '__init__.py')
        if os.path.isfile(init) and os.path.abspath(cand) != _HERE:
            spec = importlib.util.spec_from_file_location('_upstream_BabelViscoFDTD', init, submodule_search_locations=[cand])
            mod = importlib.util.module_from_spec(spec)
            sys.modules['_upstream_BabelViscoFDTD'] = mod
            try:
                spec.loader.exec_module(mod)
                _upstream = mod
            except Exception:          # a broken install must not take the hot path down with it
                sys.modules.pop('_upstream_BabelViscoFDTD', None)
            break
    return _upstream


def upstream_attr(submodule, name):
    """getattr(<upstream>.<submodule>, name) or None when there is no genuine install."""
    if upstream() is None:
        return None
    try:
        mod = importlib.import_module('_upstream_BabelViscoFDTD.' + submodule)
        return getattr(mod, name, None)
    except Exception:
        return None
