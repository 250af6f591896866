"""ReadFromH5py / SaveToH5py (TranscranialModeling/BabelIntegrationBASE.py:17,61,1586): nested
dict <-> HDF5.  Needs h5py, which this image does not ship; the import is deferred so that the
solver path works without it."""
import numpy as np


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:  # pragma: no cover
        raise ImportError('h5py is required for ReadFromH5py/SaveToH5py') from e


def SaveToH5py(MyDict, f, group=None):
    h5py = _h5py()
    own = isinstance(f, str)
    fh = h5py.File(f, 'w') if own else f
    g = fh if group is None else fh.create_group(group)
    for k, v in MyDict.items():
        _save(g, str(k), v)
    if own:
        fh.close()


def _save(g, name, v):
    if isinstance(v, dict):
        sg = g.create_group(name)
        sg.attrs['type'] = 'dict'
        for k, x in v.items():
            _save(sg, str(k), x)
    elif isinstance(v, (list, tuple)):
        sg = g.create_group(name)
        sg.attrs['type'] = 'list' if isinstance(v, list) else 'tuple'
        for n, x in enumerate(v):
            _save(sg, 'item_%d' % n, x)
    elif isinstance(v, str):
        ds = g.create_dataset(name, data=np.bytes_(v))
        ds.attrs['type'] = 'str'
    elif v is None:
        ds = g.create_dataset(name, data=0)
        ds.attrs['type'] = 'None'
    else:
        a = np.asarray(v)
        kw = {'compression': 'gzip'} if a.ndim > 0 and a.size > 1024 else {}
        ds = g.create_dataset(name, data=a, **kw)
        ds.attrs['type'] = 'ndarray' if isinstance(v, np.ndarray) else 'scalar'


def ReadFromH5py(f, group=None):
    h5py = _h5py()
    own = isinstance(f, str)
    fh = h5py.File(f, 'r') if own else f
    g = fh if group is None else fh[group]
    out = {k: _read(g[k]) for k in g.keys()}
    if own:
        fh.close()
    return out


def _read(o):
    t = o.attrs.get('type', None)
    if hasattr(o, 'keys'):
        if t in ('list', 'tuple'):
            items = [_read(o['item_%d' % n]) for n in range(len(o.keys()))]
            return items if t == 'list' else tuple(items)
        return {k: _read(o[k]) for k in o.keys()}
    v = o[()]
    if t == 'str':
        return v.decode() if isinstance(v, bytes) else str(v)
    if t == 'None':
        return None
    if t == 'scalar' and np.ndim(v) == 0:
        return v.item() if hasattr(v, 'item') else v
    return v


# prefer the genuine implementation when a BabelViscoFDTD install is shadowed by this shim (same on-disk format)
from BabelViscoFDTD import upstream_attr as _up  # noqa: E402
ReadFromH5py = _up('H5pySimple', 'ReadFromH5py') or ReadFromH5py
SaveToH5py = _up('H5pySimple', 'SaveToH5py') or SaveToH5py
