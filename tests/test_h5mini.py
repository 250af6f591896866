"""HDF5 without h5py: babelbrain_b200/h5mini.py and the H5pySimple shim on top of it.

Pinned against the reference tree's own HDF5 files where a reference tree is present (/root/reference, or the git-ignored
copy under baseline/_ref): MapPichardo.h5 -- written by the genuine BabelViscoFDTD.H5pySimple.SaveToH5py (Blosc filter 32001,
LZ4 codec, byte shuffle, chunked B-tree layout, variable-length string attribute 'type' = 'ndarray') and read at import time
by BabelIntegrationBASE.py:61 -- and the k-Plan CT calibration file (new-style compact group, contiguous layout).  Two of
MapPichardo's arrays are exact linspaces, which checks the whole chunk -> Blosc -> LZ4 -> unshuffle path bit for bit
without needing libhdf5.  Everything else (writer, conventions of nested containers) is checked by round trips."""
import os

import numpy as np
import pytest

from babelbrain_b200 import h5mini
from BabelViscoFDTD.H5pySimple import ReadFromH5py, SaveToH5py
from tests import refcaller


def ref_file(name):
    root = refcaller.reference_root()
    path = None if root is None else os.path.join(root, 'TranscranialModeling', name)
    if path is None or not os.path.isfile(path):
        pytest.skip('no reference tree with %s (python tests/make_ref_install.py)' % name)
    return path


def lz4_compress(data):
    """A plain greedy LZ4 block compressor (test helper: produces literals, matches, long lengths and overlapping matches)."""
    data = bytes(data)
    n, out, anchor, i, table = len(data), bytearray(), 0, 0, {}

    def emit(lit, off, ml):
        ll = len(lit)
        tok = (min(ll, 15) << 4) | (min(ml - 4, 15) if ml else 0)
        out.append(tok)
        if ll >= 15:
            r = ll - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out.extend(lit)
        if ml:
            out.extend((off & 255, off >> 8))
            if ml - 4 >= 15:
                r = ml - 4 - 15
                while r >= 255:
                    out.append(255); r -= 255
                out.append(r)
    while i + 4 <= n - 5:
        key = data[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 65535:
            ml = 4
            while i + ml < n - 5 and data[cand + ml] == data[i + ml]:
                ml += 1
            emit(data[anchor:i], i - cand, ml)
            i += ml
            anchor = i
        else:
            i += 1
    emit(data[anchor:], 0, 0)
    return bytes(out)


@pytest.mark.parametrize('use_c', [False, True])
def test_lz4_block_decoder(use_c):
    rng = np.random.default_rng(0)
    cases = [b'', b'a', b'abcd' * 3, bytes(1000), b'xyz' * 5000, rng.integers(0, 4, 70000, dtype=np.uint8).tobytes(),
             rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), (b'0123456789' * 40 + bytes(range(256))) * 30]
    saved = h5mini._LZ4[0]
    try:
        if not use_c:
            h5mini._LZ4[0] = None
        elif h5mini._lz4_helper() is None:
            pytest.skip('libbabelb200.so is not built')
        for data in cases:
            comp = lz4_compress(data)
            assert h5mini.lz4_block_decompress(comp, len(data)) == data
            if len(data) > 100:
                assert len(comp) < len(data) or data is cases[6]
        with pytest.raises(h5mini.H5Error):
            h5mini.lz4_block_decompress(lz4_compress(b'xyz' * 500)[:-3], 1500)
    finally:
        h5mini._LZ4[0] = saved


def test_blosc_container_round_trip():
    """Blosc 1.x chunk as c-blosc lays it out (header, block starts, one stream per byte plane after the shuffle), built here
    from the format description with the LZ4 helper above, for sizes with and without split blocks and a leftover block."""
    rng = np.random.default_rng(1)
    for count, typesize, blocksize in ((500, 8, 4000), (2016, 8, 16128), (5000, 4, 4096), (300, 2, 4096), (77, 8, 1 << 15)):
        a = np.cumsum(rng.integers(0, 3, count)).astype({8: np.float64, 4: np.float32, 2: np.int16}[typesize])
        raw = a.tobytes()
        nbytes = len(raw)
        nblocks = (nbytes + blocksize - 1) // blocksize
        body, starts = bytearray(), []
        for k in range(nblocks):
            blk = raw[k * blocksize:(k + 1) * blocksize]
            ne = len(blk) // typesize
            sh = np.frombuffer(blk[:ne * typesize], np.uint8).reshape(ne, typesize).T.tobytes() + blk[ne * typesize:]
            split = typesize if (len(blk) == blocksize and len(blk) // typesize >= 128) else 1
            starts.append(16 + 4 * nblocks + len(body))
            part = len(blk) // split
            for s in range(split):
                piece = sh[s * part:(s + 1) * part]
                comp = lz4_compress(piece)
                if len(comp) >= len(piece):
                    comp = piece
                body += np.int32(len(comp)).tobytes() + comp
        chunk = bytes([2, 1, 0x01 | (1 << 5), typesize]) + np.array([nbytes, blocksize, 16 + 4 * nblocks + len(body)], np.uint32).tobytes() \
            + np.array(starts, np.uint32).tobytes() + bytes(body)
        assert h5mini.blosc_decompress(chunk) == raw


def test_map_pichardo_written_by_the_genuine_package():
    path = ref_file('MapPichardo.h5')
    f = h5mini.H5File(path)
    assert sorted(f.root.keys()) == ['MapAtt', 'MapSoS', 'freq', 'rho']
    for k in f.root.keys():
        o = f.root[k]
        assert o.attrs == {'type': 'ndarray'} and o.filters[0][0] == 32001 and o.filters[0][1][6] == 1      # Blosc, LZ4
    d = ReadFromH5py(path)
    assert np.array_equal(d['rho'], np.linspace(1242.0, 2900.0, 500))          # exact: every byte of the chunk path is right
    assert np.array_equal(d['freq'], np.linspace(0.1, 1.0, 500))
    sos, att = d['MapSoS'], d['MapAtt']
    assert sos.shape == att.shape == (500, 500) and sos.dtype == np.float64
    assert 1700 < sos.min() < 1720 and 3700 < sos.max() < 3800 and np.all(np.diff(sos, axis=1) > 0)   # speed of sound grows with density
    assert 6 < att.min() < 7 and 300 < att.max() < 320 and np.isfinite(att).all()
    # both decoders give the same arrays
    saved = h5mini._LZ4[0]
    try:
        h5mini._LZ4[0] = None
        d2 = ReadFromH5py(path)
    finally:
        h5mini._LZ4[0] = saved
    assert all(np.array_equal(d[k], d2[k]) for k in d)


def test_kplan_calibration_file():
    f = h5mini.H5File(ref_file('ct-calibration-low-dose-30-March-2023-v1.h5'))
    assert f.root.attrs['application_name'] == 'k-Plan' and f.root.attrs['file_type'] == 'k-Plan CT Calibration'
    assert int(f.root.attrs['major_version'][0]) == 1
    cal = f.root['ct_calibration'].read()
    assert cal.shape == (1, 10, 2) and cal.dtype == np.float32
    assert np.all(np.diff(cal[0, :, 0]) > 0) and np.all(np.diff(cal[0, :, 1]) > 0)          # HU -> density, both increasing
    assert cal[0, 0, 1] == pytest.approx(1.2) and cal[0, -1, 1] == pytest.approx(2150.0)


def test_nested_dict_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    out = {'p_amp': rng.random((5, 6, 7)).astype(np.float32), 'MaterialMap': np.arange(24, dtype=np.uint8).reshape(2, 3, 4),
           'x_vec': np.linspace(0, 1, 5), 'big': rng.random((40, 50, 60)), 'name': 'CTX-500', 'n': 7, 'f': 2.5, 'none': None,
           'lst': [1, 'two', np.array([3.0]), {'deep': (1, 2.0), 'more': [np.int16(4)]}], 'cplx': np.array([1 + 2j, 3 - 4j], np.complex64),
           'empty': np.zeros((0, 3)), 'i64': np.array([-2 ** 62, 2 ** 62]), 'f16': np.array([1.5, -2.25], np.float16),
           'many': [float(n) for n in range(300)]}               # 300 entries: three symbol-table nodes under one B-tree node
    path = str(tmp_path / 'DataForSim.h5')
    SaveToH5py(out, path)
    back = ReadFromH5py(path)
    assert sorted(back) == sorted(out)
    for k, v in out.items():
        if isinstance(v, np.ndarray):
            assert back[k].dtype == v.dtype and np.array_equal(back[k], v), k
        else:
            assert back[k] == v or (k == 'lst' and back[k][:2] == v[:2]), k
    assert back['lst'][2][0] == 3.0 and back['lst'][3]['deep'] == (1, 2.0) and back['lst'][3]['more'] == [4]
    assert isinstance(back['n'], int) and isinstance(back['f'], float) and isinstance(back['name'], str)
    f = h5mini.H5File(path)
    assert f.root.attrs == {} and f.root['lst'].attrs == {'type': 'list'} and f.root['p_amp'].attrs == {'type': 'ndarray'}
    assert ReadFromH5py(path, group='lst/item_3') == {'deep': (1, 2.0), 'more': [4]}


def test_unsupported_features_are_named(tmp_path):
    p = tmp_path / 'x.h5'
    p.write_bytes(b'not hdf5 at all')
    with pytest.raises(h5mini.H5Error):
        h5mini.H5File(str(p))
    p.write_bytes(b'\x89HDF\r\n\x1a\n' + bytes([2]) + bytes(100))
    with pytest.raises(NotImplementedError, match='superblock version 2'):
        h5mini.H5File(str(p))
