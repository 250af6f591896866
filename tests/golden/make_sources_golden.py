"""Generates tests/golden/sources_ref.npz from the REFERENCE's own code: CreateSources of
/root/reference/TranscranialModeling/BabelIntegrationSingle.py (the continuous-wave source table every transducer model
builds) is cut out with `ast` and executed, unmodified, on a stub object with a seeded complex source plane.  Only the
inputs and the arrays it produced are stored.

    python tests/golden/make_sources_golden.py        (needs /root/reference; run in the build container only)
"""
import ast
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/TranscranialModeling/BabelIntegrationSingle.py'


def reference_method():
    tree = ast.parse(open(REF).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == 'CreateSources':
            ns = {'np': np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, 'exec'), ns)
            return ns['CreateSources'], (node.lineno, node.end_lineno)
    raise RuntimeError('CreateSources not found in the reference')


def main():
    fn, lines = reference_method()
    rng = np.random.default_rng(11)
    n1, n2, n3, pml, zsrc = 14, 12, 10, 3, 3
    f, ppp, periods = 500e3, 48, 9
    dt = 1.0 / f / ppp
    tsim = dt * ppp * periods
    plane = np.zeros((n1, n2), np.complex64)
    amp = rng.uniform(1e3, 1e5, (n1 - 2 * pml, n2 - 2 * pml))
    plane[pml:-pml, pml:-pml] = (amp * np.exp(1j * rng.uniform(-np.pi, np.pi, amp.shape))).astype(np.complex64)
    plane[pml + 2, pml + 1] = 0                                   # a hole: rows are numbered over non-zero pixels only
    me = types.SimpleNamespace(_TimeSimulation=tsim, _Frequency=f, _TemporalStep=dt, _N1=n1, _N2=n2, _N3=n3,
                               _ZSourceLocation=zsrc, _SourceMapRayleigh=plane.copy(), _bDisplay=False)
    fn(me)
    out = os.path.join(HERE, 'sources_ref.npz')
    np.savez_compressed(out, plane=plane, frequency=f, dt=dt, tsim=tsim, zsrc=zsrc, shape=np.array([n1, n2, n3]),
                        PulseSource=me._PulseSource, SourceMap=me._SourceMap, reference_lines=np.array(lines))
    print('wrote', out, me._PulseSource.shape, 'from lines %d-%d' % lines)


if __name__ == '__main__':
    main()
