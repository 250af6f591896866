"""Continuous-wave sources without the dense table.  CWSourceFunctions.dense() is pinned bit for
bit against the reference's own CreateSources (tests/golden/sources_ref.npz, produced by tests/golden/make_sources_golden.py
from /root/reference); the in-kernel evaluation is checked against a simulation fed with that dense table."""
import os

import numpy as np
import pytest

from babelbrain_b200 import workloads
from babelbrain_b200.sources import CWSourceFunctions

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'sources_ref.npz')


def test_dense_table_and_row_numbering_equal_the_reference():
    g = np.load(GOLD)
    cw, mask = CWSourceFunctions.from_source_plane(g['plane'], float(g['frequency']), float(g['dt']), float(g['tsim']))
    assert cw.shape == g['PulseSource'].shape and cw.dtype == g['PulseSource'].dtype
    assert np.array_equal(cw.dense(), g['PulseSource'])
    assert np.array_equal(np.asarray(cw), g['PulseSource'])
    sm = g['SourceMap']
    assert np.array_equal(mask, sm[:, :, int(g['zsrc'])]) and sm.sum() == mask.sum()


def test_tone_tables_reproduce_the_table_to_float32_rounding():
    g = np.load(GOLD)
    cw, _ = CWSourceFunctions.from_source_plane(g['plane'], float(g['frequency']), float(g['dt']), float(g['tsim']))
    ac, asn, es, ec = cw.tone_tables()
    assert all(a.dtype == np.float32 for a in (ac, asn, es, ec)) and ac.size == cw.shape[0] and es.size == cw.shape[1]
    syn = es[None, :].astype(np.float64) * ac[:, None] + ec[None, :].astype(np.float64) * asn[:, None]
    ref = g['PulseSource']
    assert np.abs(syn - ref).max() <= 3e-7 * np.abs(ref).max()
    assert workloads.cw_sources(cw.amplitude, cw.phase, cw.Frequency, cw.TemporalStep, float(g['tsim']) / cw.TemporalStep).shape == ref.shape
    with pytest.raises(ValueError):
        CWSourceFunctions(np.ones(3), np.ones(4), 5e5, 1e-8, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('name,shape,pml', [('ctx500_skull', (56, 48, 72), 8), ('dome_stress', (64, 64, 48), 8)])
def test_in_kernel_sources_match_the_dense_table(name, shape, pml):
    from babelbrain_b200.propagation import PropagationModel
    w = workloads.make_workload(name, shape=shape, periods=6, pml=pml)
    MM, ML, f, SM, SF, h, T, SEN = w['args']
    cw = w['meta']['cw_sources']
    assert np.array_equal(cw.dense(), SF)
    PM = PropagationModel()
    S1, _, R1, _ = PM.StaggeredFDTD_3D_with_relaxation(*w['args'], **w['kwargs'])
    up1 = PM.last_timing['h2d_bytes']
    S2, _, R2, _ = PM.StaggeredFDTD_3D_with_relaxation(MM, ML, f, SM, cw, h, T, SEN, **w['kwargs'])
    assert PM.last_timing['h2d_bytes'] < up1 - SF.nbytes + 16 * sum(cw.shape)
    n = np.linalg.norm(R1['Pressure'].astype(np.float64))
    assert np.linalg.norm(R2['Pressure'].astype(np.float64) - R1['Pressure']) <= 1e-5 * n
    assert int(np.argmax(R2['Pressure'])) == int(np.argmax(R1['Pressure']))
    ns = np.linalg.norm(S1['Pressure'].astype(np.float64))
    assert np.linalg.norm(S2['Pressure'].astype(np.float64) - S1['Pressure']) <= 1e-5 * ns
