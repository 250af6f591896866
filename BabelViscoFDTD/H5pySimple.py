"""ReadFromH5py / SaveToH5py (TranscranialModeling/BabelIntegrationBASE.py:17,61,1586): nested
dict <-> HDF5.  With h5py installed the files go through it; without it (this image) through
babelbrain_b200.h5mini, a reader / writer of the HDF5 subset these files use.  The reader is
checked against the two HDF5 files of the reference tree (MapPichardo.h5, written by the genuine
SaveToH5py with Blosc-LZ4 chunks, and the k-Plan CT calibration file); the conventions -- attribute
'type' = 'ndarray' / 'dict' / 'list' / 'tuple' / 'str' / 'scalar' / 'None', list items as
'item_<n>' -- are those of MapPichardo.h5 plus the published package."""
import numpy as np


def _h5py():
    try:
        import h5py
    except ImportError:
        return None
    return h5py if hasattr(h5py, 'File') else None        # a placeholder module (test harnesses) is not h5py


def _tree(v):
    """nested dict -> h5mini.Group / Dataset with the 'type' attributes"""
    from babelbrain_b200 import h5mini
    if isinstance(v, dict):
        g = h5mini.Group()
        g.attrs['type'] = 'dict'
        for k, x in v.items():
            g[str(k)] = _tree(x)
        return g
    if isinstance(v, (list, tuple)):
        g = h5mini.Group()
        g.attrs['type'] = 'list' if isinstance(v, list) else 'tuple'
        for n, x in enumerate(v):
            g['item_%d' % n] = _tree(x)
        return g
    if isinstance(v, str):
        return h5mini.Dataset(np.bytes_(v), {'type': 'str'})
    if v is None:
        return h5mini.Dataset(np.int64(0), {'type': 'None'})
    a = np.asarray(v)
    if a.dtype.kind == 'c':           # HDF5 has no complex type: h5py stores a compound (r, i); this writer two planes
        return h5mini.Dataset(np.stack([a.real, a.imag], -1), {'type': 'ndarray' if isinstance(v, np.ndarray) else 'scalar', 'complex': np.int8(1)})
    return h5mini.Dataset(a, {'type': 'ndarray' if isinstance(v, np.ndarray) else 'scalar'})


def _untree(o):
    t = o.attrs.get('type', None)
    if o.is_group:
        if t in ('list', 'tuple'):
            items = [_untree(o['item_%d' % n]) for n in range(len(o.keys()))]
            return items if t == 'list' else tuple(items)
        return {k: _untree(o[k]) for k in o.keys()}
    v = o.read()
    if 'complex' in o.attrs:
        v = v[..., 0] + 1j * v[..., 1]
    if t == 'str':
        return v.decode() if isinstance(v, (bytes, np.bytes_)) else str(v)
    if t == 'None':
        return None
    if t == 'scalar' and np.ndim(v) == 0:
        return v.item() if hasattr(v, 'item') else v
    return v


def SaveToH5py(MyDict, f, group=None):
    h5py = _h5py()
    if h5py is None:
        from babelbrain_b200 import h5mini
        if not isinstance(f, str) or group is not None:
            raise NotImplementedError('without h5py only whole files can be written (f = file name)')
        root = _tree(dict(MyDict))
        root.attrs.clear()                # the root group carries no marker (MapPichardo.h5)
        h5mini.write_file(f, root)
        return
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
    if h5py is None:
        from babelbrain_b200 import h5mini
        if not isinstance(f, str):
            raise NotImplementedError('without h5py only file names can be read')
        root = h5mini.H5File(f).root
        if group is not None:
            for part in str(group).strip('/').split('/'):
                root = root[part]
        return {k: _untree(root[k]) for k in root.keys()}
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
