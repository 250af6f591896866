"""
A small reader for the HDF5 files BabelBrain writes and reads through ``BabelViscoFDTD.H5pySimple`` (SURVEY.md section 8f
row 3), for installations without h5py / hdf5plugin (this image has neither).  It covers what those files use -- the
reference tree holds two of them, TranscranialModeling/MapPichardo.h5 (read at import time, BabelIntegrationBASE.py:61)
and ct-calibration-low-dose-30-March-2023-v1.h5:

  superblock version 0-1, object headers version 1 (with continuation blocks), old-style groups (symbol table: B-tree v1 +
  local heap), datasets with compact / contiguous / chunked (B-tree v1) layout, the deflate, shuffle and Blosc (filter 32001:
  blosclz / LZ4 / zlib codecs, byte shuffle) filters, fixed-point, floating-point, fixed-length and variable-length string
  types (global heap), scalar / simple dataspaces, attributes.

Anything else raises NotImplementedError naming the feature.  Not a general HDF5 library.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def _u(fmt, buf, off):
    return struct.unpack_from('<' + fmt, buf, off)


# ------------------------------------------------------------------------------------------
# Blosc 1.x container (https://github.com/Blosc/c-blosc/blob/main/README_CHUNK_FORMAT.rst) and LZ4 block format
# ------------------------------------------------------------------------------------------
def lz4_block_decompress(src, out_size):
    """LZ4 block format: sequences of (token, literals, offset, match)."""
    L = _lz4_helper()
    if L is not None:                       # the library's host helper when it is built (C)
        out = np.empty(max(out_size, 1), np.uint8)
        a = np.frombuffer(bytes(src) or b'\0', np.uint8)
        n = L.bb_host_lz4_decompress(a.ctypes.data, len(src), out.ctypes.data, out_size)
        if n != out_size:
            raise H5Error('corrupt LZ4 block (%d of %d bytes)' % (n, out_size))
        return out[:out_size].tobytes()
    return _lz4_block_decompress_py(src, out_size)


_LZ4 = [False]


def _lz4_helper():
    if _LZ4[0] is False:
        try:
            from . import _capi
            L = _capi.lib()
            _LZ4[0] = L if hasattr(L, 'bb_host_lz4_decompress') else None
        except Exception:                   # library not built: the Python decoder below
            _LZ4[0] = None
    return _LZ4[0]


def _lz4_block_decompress_py(src, out_size):
    src = bytes(src)
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        tok = src[i]; i += 1
        ll = tok >> 4
        if ll == 15:
            while True:
                b = src[i]; i += 1
                ll += b
                if b != 255:
                    break
        out += src[i:i + ll]; i += ll
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8); i += 2
        if off == 0:
            raise H5Error('corrupt LZ4 block (zero offset)')
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]; i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        if start < 0:
            raise H5Error('corrupt LZ4 block (offset before start)')
        if off >= ml:
            out += out[start:start + ml]
        else:                               # overlapping match: the pattern repeats
            pat = bytes(out[start:])
            out += (pat * (ml // off + 1))[:ml]
    if len(out) != out_size:
        raise H5Error('corrupt LZ4 block (%d of %d bytes)' % (len(out), out_size))
    return bytes(out)


def blosclz_decompress(src, out_size):
    """blosclz (a FastLZ level-1/2 derivative) as c-blosc 1.x writes it."""
    src = bytes(src)
    out = bytearray()
    ip, n = 0, len(src)
    ctrl = src[ip] & 31; ip += 1
    while True:
        if ctrl >= 32:
            ln = (ctrl >> 5) - 1
            ofs = (ctrl & 31) << 8
            if ln == 7 - 1:
                while True:
                    code = src[ip]; ip += 1
                    ln += code
                    if code != 255:
                        break
            code = src[ip]; ip += 1
            ln += 3
            ref = len(out) - ofs - code
            if code == 255 and ofs == (31 << 8):
                ofs = (src[ip] << 8) + src[ip + 1]; ip += 2
                ref = len(out) - ofs - 8191
            ref -= 1
            if ref < 0:
                raise H5Error('corrupt blosclz stream')
            dist = len(out) - ref
            if dist >= ln:
                out += out[ref:ref + ln]
            else:
                pat = bytes(out[ref:])
                out += (pat * (ln // dist + 1))[:ln]
        else:
            ctrl += 1
            out += src[ip:ip + ctrl]; ip += ctrl
        if ip >= n:
            break
        ctrl = src[ip]; ip += 1
    if len(out) != out_size:
        raise H5Error('corrupt blosclz stream (%d of %d bytes)' % (len(out), out_size))
    return bytes(out)


def blosc_decompress(chunk):
    chunk = bytes(chunk)
    version, versionlz, flags, typesize = chunk[0], chunk[1], chunk[2], chunk[3]
    nbytes, blocksize, cbytes = _u('III', chunk, 4)
    if flags & 0x2:                          # memcpyed
        return chunk[16:16 + nbytes]
    doshuffle, dobitshuffle = bool(flags & 0x1), bool(flags & 0x4)
    if dobitshuffle:
        raise NotImplementedError('Blosc bit-shuffle')
    codec = flags >> 5
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = _u('%dI' % nblocks, chunk, 16)
    dont_split = bool(flags & 0x10)
    out = bytearray()
    for k in range(nblocks):
        bsize = min(blocksize, nbytes - k * blocksize)
        leftover = bsize != blocksize
        # c-blosc splits a block into `typesize` streams (one per byte plane) when it is shuffled and big enough
        nsplits = typesize if (not dont_split and typesize <= 16 and bsize // typesize >= 128 and not leftover and bsize % typesize == 0) else 1
        if versionlz and codec in (1, 3, 4) and version >= 2 and dont_split:
            nsplits = 1
        neblock = bsize // nsplits
        ip = bstarts[k]
        block = bytearray()
        for _ in range(nsplits):
            cb, = _u('i', chunk, ip); ip += 4
            if cb == neblock:
                piece = chunk[ip:ip + cb]
            elif codec == 0:
                piece = blosclz_decompress(chunk[ip:ip + cb], neblock)
            elif codec == 1:
                piece = lz4_block_decompress(chunk[ip:ip + cb], neblock)
            elif codec == 3:
                piece = zlib.decompress(chunk[ip:ip + cb])
            else:
                raise NotImplementedError('Blosc codec %d (snappy / zstd)' % codec)
            ip += cb
            block += piece
        if doshuffle and typesize > 1:
            ne = bsize // typesize
            a = np.frombuffer(bytes(block[:ne * typesize]), np.uint8).reshape(typesize, ne).T
            block = bytearray(np.ascontiguousarray(a).tobytes()) + block[ne * typesize:]
        out += block
    return bytes(out[:nbytes])


def hdf5_unshuffle(buf, typesize):
    if typesize <= 1:
        return buf
    ne = len(buf) // typesize
    a = np.frombuffer(buf[:ne * typesize], np.uint8).reshape(typesize, ne).T
    return np.ascontiguousarray(a).tobytes() + buf[ne * typesize:]


# ------------------------------------------------------------------------------------------
# file structures
# ------------------------------------------------------------------------------------------
class Datatype:
    def __init__(self, buf, off):
        cv, b0, b1, b2, size = _u('BBBBI', buf, off)
        self.cls, self.version, self.size = cv & 15, cv >> 4, size
        self.vlen_string = False
        self.nbytes_header = 8
        if self.cls == 0:                    # fixed point
            signed = bool(b0 & 8)
            self.dtype = np.dtype(('>' if b0 & 1 else '<') + ('i' if signed else 'u') + str(size))
            self.nbytes_header += 4
        elif self.cls == 1:                  # floating point
            self.dtype = np.dtype(('>' if b0 & 1 else '<') + 'f' + str(size))
            self.nbytes_header += 12
        elif self.cls == 3:                  # fixed-length string
            self.dtype = np.dtype('S%d' % size)
        elif self.cls == 9:                  # variable length
            if (b0 & 15) != 1:
                raise NotImplementedError('variable-length sequences')
            self.vlen_string = True
            self.dtype = np.dtype('O')
        elif self.cls == 6:
            raise NotImplementedError('compound datatypes')
        else:
            raise NotImplementedError('datatype class %d' % self.cls)


class Dataspace:
    def __init__(self, buf, off):
        version, rank, flags = buf[off], buf[off + 1], buf[off + 2]
        if version == 1:
            p = off + 8
        elif version == 2:
            p = off + 4
            if buf[off + 3] == 2:            # null dataspace
                self.shape = None
                return
        else:
            raise NotImplementedError('dataspace version %d' % version)
        self.shape = tuple(_u('%dQ' % rank, buf, p)) if rank else ()


class H5Object:
    """A group (has .links) or a dataset (has .read())."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = f._object_header(addr)
        self.attrs = {}
        self.dtype = self.space = self.layout = None
        self.filters = []
        self.links = None
        for t, body in self.msgs:
            buf, off = body
            if t == 0x01:
                self.space = Dataspace(buf, off)
            elif t == 0x03:
                self.dtype = Datatype(buf, off)
            elif t == 0x08:
                self.layout = (buf, off)
            elif t == 0x0B:
                self.filters = f._filter_pipeline(buf, off)
            elif t == 0x0C:
                k, v = f._attribute(buf, off)
                self.attrs[k] = v
            elif t == 0x11:
                btree, heap = _u('QQ', buf, off)
                self.links = f._symbol_table(btree, heap)
            elif t == 0x02:                                      # link info: compact groups only (links as 0x06 messages)
                flags = buf[off + 1]
                p = off + 2 + (8 if flags & 1 else 0)
                fheap, = _u('Q', buf, p)
                if fheap != UNDEF:
                    raise NotImplementedError('dense groups (fractal heap); write the file with libver="earliest"')
                if self.links is None:
                    self.links = {}
            elif t == 0x06:                                      # link message
                flags = buf[off + 1]
                p = off + 2
                ltype = 0
                if flags & 8:
                    ltype = buf[p]; p += 1
                if flags & 4:
                    p += 8
                if flags & 16:
                    p += 1
                nl = (flags & 3)
                ln = int.from_bytes(bytes(buf[p:p + (1 << nl)]), 'little'); p += 1 << nl
                name = bytes(buf[p:p + ln]).decode('utf-8'); p += ln
                if ltype != 0:
                    raise NotImplementedError('soft / external links')
                if self.links is None:
                    self.links = {}
                self.links[name], = _u('Q', buf, p)

    @property
    def is_group(self):
        return self.links is not None

    def keys(self):
        return list(self.links)

    def __getitem__(self, name):
        return H5Object(self.f, self.links[name])

    def read(self):
        f = self.f
        if self.space is None or self.dtype is None or self.layout is None:
            raise H5Error('not a dataset')
        shape = self.space.shape
        if shape is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        esize = self.dtype.size
        buf, off = self.layout
        version = buf[off]
        if version != 3:
            raise NotImplementedError('data layout message version %d' % version)
        cls = buf[off + 1]
        if cls == 0:                                             # compact
            size, = _u('H', buf, off + 2)
            raw = bytes(buf[off + 4:off + 4 + size])
        elif cls == 1:                                           # contiguous
            addr, size = _u('QQ', buf, off + 2)
            raw = b'\0' * (count * esize) if addr == UNDEF else f.buf[addr:addr + size]
        elif cls == 2:                                           # chunked
            rank = buf[off + 2]
            btree, = _u('Q', buf, off + 3)
            cdims = _u('%dI' % rank, buf, off + 11)              # chunk dims + element size
            raw = self._read_chunks(btree, cdims[:-1], shape, esize)
        else:
            raise NotImplementedError('data layout class %d' % cls)
        if self.dtype.vlen_string:
            out = np.empty(count, object)
            for n in range(count):
                ln, gaddr, gidx = _u('IQI', raw, n * 16)
                out[n] = f._global_heap_object(gaddr, gidx)[:ln].decode('utf-8', 'replace')
            return out.reshape(shape) if shape else out[0]
        a = np.frombuffer(raw, self.dtype.dtype, count)
        if a.dtype.byteorder == '>':
            a = a.astype(a.dtype.newbyteorder('<'))
        a = a.reshape(shape).copy()
        return a if shape else a[()]

    def _read_chunks(self, btree, cdims, shape, esize):
        f = self.f
        rank = len(shape)
        out = np.zeros(shape, np.dtype('V%d' % esize))
        chunk_bytes = int(np.prod(cdims)) * esize
        for offs, addr, size, mask in f._chunk_btree(btree, rank):
            data = f.buf[addr:addr + size]
            for n, (fid, cd) in reversed(list(enumerate(self.filters))):
                if mask & (1 << n):
                    continue
                if fid == 1:
                    data = zlib.decompress(data)
                elif fid == 2:
                    data = hdf5_unshuffle(bytes(data), cd[0] if cd else esize)
                elif fid == 32001:
                    data = blosc_decompress(data)
                elif fid == 3:                                   # fletcher32: checksum behind the data
                    data = data[:-4]
                else:
                    raise NotImplementedError('HDF5 filter %d' % fid)
            block = np.frombuffer(bytes(data[:chunk_bytes]), np.dtype('V%d' % esize)).reshape(cdims)
            sl_out, sl_in = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + cdims[d], shape[d])
                sl_out.append(slice(lo, hi)); sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = block[tuple(sl_in)]
        return out.tobytes()


class H5File:
    def __init__(self, path):
        with open(path, 'rb') as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b'\x89HDF\r\n\x1a\n':
            raise H5Error('%s is not an HDF5 file' % path)
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError('superblock version %d (write the file with libver="earliest")' % ver)
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError('offsets / lengths that are not 8 bytes')
        p = 24 + (4 if ver == 1 else 0)
        base, _, _, _ = _u('QQQQ', b, p)
        if base != 0:
            raise NotImplementedError('non-zero base address')
        p += 32
        _, oh, cache = _u('QQI', b, p)
        self.root = H5Object(self, oh)

    # ---- object header (version 1)
    def _object_header(self, addr):
        b = self.buf
        if b[addr:addr + 4] == b'OHDR':
            raise NotImplementedError('object header version 2 (write the file with libver="earliest")')
        version, _, nmsg, _, hsize = _u('BBHII', b, addr)
        if version != 1:
            raise H5Error('object header version %d at %d' % (version, addr))
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                t, sz, flags = _u('HHB', b, p)
                body = p + 8
                if t == 0x10:                                    # continuation
                    caddr, clen = _u('QQ', b, body)
                    blocks.append((caddr, clen))
                    msgs.append((t, (b, body)))
                elif flags & 2:                                  # shared message
                    raise NotImplementedError('shared object header messages')
                else:
                    msgs.append((t, (b, body)))
                p = body + sz
        return [m for m in msgs if m[0] not in (0x00, 0x10)]

    # ---- old-style group: B-tree v1 of symbol-table nodes + local heap of names
    def _local_heap(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b'HEAP':
            raise H5Error('no local heap at %d' % addr)
        size, _, data = _u('QQQ', b, addr + 8)
        return data

    def _symbol_table(self, btree, heap):
        b = self.buf
        hdata = self._local_heap(heap)
        links = {}

        def node(addr):
            if b[addr:addr + 4] == b'TREE':
                ntype, level, n = _u('BBH', b, addr + 4)
                p = addr + 8 + 16
                for k in range(n):
                    p += 8                                       # key
                    child, = _u('Q', b, p); p += 8
                    node(child)
            elif b[addr:addr + 4] == b'SNOD':
                n, = _u('H', b, addr + 6)
                p = addr + 8
                for k in range(n):
                    name_off, oh = _u('QQ', b, p)
                    e = b.index(b'\0', hdata + name_off)
                    links[b[hdata + name_off:e].decode('utf-8')] = oh
                    p += 40
            else:
                raise H5Error('unexpected group node at %d' % addr)
        node(btree)
        return links

    # ---- chunk index: B-tree v1, node type 1
    def _chunk_btree(self, addr, rank):
        b = self.buf
        if addr == UNDEF:
            return
        if b[addr:addr + 4] != b'TREE':
            raise H5Error('no chunk B-tree at %d' % addr)
        ntype, level, n = _u('BBH', b, addr + 4)
        p = addr + 8 + 16
        ksize = 8 + 8 * (rank + 1)
        for k in range(n):
            size, mask = _u('II', b, p)
            offs = _u('%dQ' % rank, b, p + 8)
            child, = _u('Q', b, p + ksize)
            if level == 0:
                yield offs, child, size, mask
            else:
                yield from self._chunk_btree(child, rank)
            p += ksize + 8

    def _filter_pipeline(self, buf, off):
        version, nf = buf[off], buf[off + 1]
        if version == 1:
            p = off + 8
        elif version == 2:
            p = off + 2
        else:
            raise NotImplementedError('filter pipeline version %d' % version)
        out = []
        for _ in range(nf):
            fid, = _u('H', buf, p)
            if version == 1 or fid >= 256:
                nlen, flags, ncd = _u('HHH', buf, p + 2)
                p += 8
                if version == 1:
                    nlen = (nlen + 7) & ~7
                p += nlen
            else:
                flags, ncd = _u('HH', buf, p + 2)
                p += 6
            cd = _u('%dI' % ncd, buf, p)
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _global_heap_object(self, addr, index):
        b = self.buf
        if b[addr:addr + 4] != b'GCOL':
            raise H5Error('no global heap collection at %d' % addr)
        csize, = _u('Q', b, addr + 8)
        p = addr + 16
        while p < addr + csize:
            idx, refc, _, size = _u('HHIQ', b, p)
            if idx == 0:
                break
            if idx == index:
                return bytes(b[p + 16:p + 16 + size])
            p += 16 + ((size + 7) & ~7)
        raise H5Error('global heap object %d not found' % index)

    def _attribute(self, buf, off):
        version = buf[off]
        if version == 1:
            nsize, tsize, ssize = _u('HHH', buf, off + 2)
            p = off + 8
            pad = lambda x: (x + 7) & ~7
        elif version in (2, 3):
            nsize, tsize, ssize = _u('HHH', buf, off + 2)
            p = off + 8 + (1 if version == 3 else 0)
            pad = lambda x: x
        else:
            raise NotImplementedError('attribute message version %d' % version)
        name = bytes(buf[p:p + nsize]).split(b'\0')[0].decode('utf-8')
        p += pad(nsize)
        dt = Datatype(buf, p)
        p += pad(tsize)
        sp = Dataspace(buf, p)
        p += pad(ssize)
        if sp.shape is None:
            return name, None
        count = int(np.prod(sp.shape)) if sp.shape else 1
        if dt.vlen_string:
            vals = []
            for n in range(count):
                ln, gaddr, gidx = _u('IQI', buf, p + 16 * n)
                vals.append(self._global_heap_object(gaddr, gidx)[:ln].decode('utf-8', 'replace'))
            return name, (vals[0] if not sp.shape else np.array(vals, object).reshape(sp.shape))
        a = np.frombuffer(bytes(buf[p:p + count * dt.size]), dt.dtype, count)
        if dt.cls == 3:
            a = np.array([x.split(b'\0')[0].decode('utf-8', 'replace') for x in a], object)
        return name, (a.reshape(sp.shape).copy() if sp.shape else a[0])


# ------------------------------------------------------------------------------------------
# writer: the same subset (superblock 0, object headers 1, old-style groups, contiguous datasets, variable-length string
# attributes in a global heap).  No compression.  Checked by reading the files back with the reader above; this image
# has no libhdf5 to check them against.
# ------------------------------------------------------------------------------------------
class Group(dict):
    """A group to write: name -> Group | numpy array / scalar / bytes; .attrs for its attributes."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.attrs = {}


class Dataset:
    def __init__(self, data, attrs=None):
        self.data = data
        self.attrs = dict(attrs or {})


LEAF_K, INTERNAL_K = 64, 512         # symbol-table nodes of up to 128 entries, B-tree nodes of up to 1024 children


class _Writer:
    def __init__(self):
        self.buf = bytearray()
        self.gheap = []              # global heap objects (bytes) of the single collection, index = position + 1
        self.gheap_fixups = []       # (position of the 8-byte collection address)

    def alloc(self, n, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b'\0' * pad
        at = len(self.buf)
        self.buf += b'\0' * n
        return at

    def put(self, at, data):
        self.buf[at:at + len(data)] = data

    # ---- messages
    @staticmethod
    def _dtype_msg(dt):
        dt = np.dtype(dt)
        if dt.kind in 'iu':
            return struct.pack('<BBBBIHH', 0x10, (8 if dt.kind == 'i' else 0), 0, 0, dt.itemsize, 0, dt.itemsize * 8)
        if dt.kind == 'f':
            if dt.itemsize == 4:
                return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
            if dt.itemsize == 8:
                return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
            if dt.itemsize == 2:
                return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 15, 0, 2, 0, 16, 10, 5, 0, 10, 15)
        if dt.kind == 'S':
            return struct.pack('<BBBBI', 0x13, 0, 0, 0, max(dt.itemsize, 1))
        if dt.kind == 'b':
            return struct.pack('<BBBBIHH', 0x10, 8, 0, 0, 1, 0, 8)
        raise NotImplementedError('cannot write dtype %s' % dt)

    @staticmethod
    def _vlen_str_dtype_msg():
        # class 9 (variable length), type = string, padding null-terminated, character set UTF-8; base type: 1-byte string
        base = struct.pack('<BBBBI', 0x13, 0, 0, 0, 1)
        return struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + base

    @staticmethod
    def _space_msg(shape):
        if shape == ():
            return struct.pack('<BBBB4x', 1, 0, 0, 0)
        return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + struct.pack('<%dQ' % len(shape), *shape)

    def _attr_msg(self, name, value):
        nm = name.encode('utf-8') + b'\0'
        pad = lambda b: b + b'\0' * ((-len(b)) % 8)
        if isinstance(value, str):
            self.gheap.append(value.encode('utf-8'))
            idx = len(self.gheap)
            dt, sp = self._vlen_str_dtype_msg(), self._space_msg(())
            data = struct.pack('<IQI', len(self.gheap[-1]), 0, idx)
            fix = True
        else:
            a = np.asarray(value)
            if a.dtype.kind in 'OU':
                raise NotImplementedError('attribute %r of type %s' % (name, a.dtype))
            dt, sp = self._dtype_msg(a.dtype), self._space_msg(a.shape)
            data = np.ascontiguousarray(a).astype(a.dtype.newbyteorder('<')).tobytes()
            fix = False
        head = struct.pack('<BxHHH', 1, len(nm), len(dt), len(sp))
        body = head + pad(nm) + pad(dt) + pad(sp)
        return body + data, (len(body) + 4 if fix else None)     # offset of the heap address inside the message body

    def _object_header(self, msgs):
        """msgs: list of (type, body, fixup offset in body or None).  Returns the header address."""
        blob = bytearray()
        fixups = []
        for t, body, fix in msgs:
            body = bytes(body) + b'\0' * ((-len(body)) % 8)
            if fix is not None:
                fixups.append(len(blob) + 8 + fix)
            blob += struct.pack('<HHB3x', t, len(body), 0) + body
        at = self.alloc(16 + len(blob))
        self.put(at, struct.pack('<BxHII4x', 1, len(msgs), 1, len(blob)))
        self.put(at + 16, blob)
        for f in fixups:
            self.gheap_fixups.append(at + 16 + f)
        return at

    # ---- objects
    def dataset(self, ds):
        data = ds.data
        if isinstance(data, (bytes, np.bytes_)):
            a = np.array(bytes(data), dtype='S%d' % max(len(data), 1))
        else:
            a = np.asarray(data)
        if a.dtype.kind == 'U':
            a = np.char.encode(a, 'utf-8')
        if a.dtype == np.bool_:
            a = a.astype(np.int8)
        raw = np.ascontiguousarray(a).astype(a.dtype.newbyteorder('<') if a.dtype.kind in 'iuf' else a.dtype).tobytes()
        daddr = self.alloc(len(raw)) if raw else UNDEF
        if raw:
            self.put(daddr, raw)
        msgs = [(0x01, self._space_msg(a.shape), None), (0x03, self._dtype_msg(a.dtype), None),
                (0x05, struct.pack('<BBBB', 2, 2, 2, 0), None),             # fill value: version 2, late allocation, none defined
                (0x08, struct.pack('<BBQQ', 3, 1, daddr, len(raw)), None)]  # layout: version 3, contiguous
        for k, v in ds.attrs.items():
            body, fix = self._attr_msg(k, v)
            msgs.append((0x0C, body, fix))
        return self._object_header(msgs)

    def group(self, g):
        entries = []
        for name in g:
            child = g[name]
            addr = self.group(child) if isinstance(child, Group) else self.dataset(child if isinstance(child, Dataset) else Dataset(child))
            entries.append((name.encode('utf-8'), addr))
        entries.sort(key=lambda e: e[0])                          # strcmp order
        # local heap: offset 0 holds the empty string
        heap_data = bytearray(b'\0' * 8)
        offs = []
        for nm, _ in entries:
            offs.append(len(heap_data))
            heap_data += nm + b'\0'
            heap_data += b'\0' * ((-len(heap_data)) % 8)
        free_off = len(heap_data)
        heap_data += b'\0' * 16                                   # one free block at the end (its own header)
        struct.pack_into('<QQ', heap_data, free_off, 1, 16)       # next free = 1 (none), size
        hdata = self.alloc(len(heap_data))
        self.put(hdata, heap_data)
        heap = self.alloc(32)
        self.put(heap, b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), free_off, hdata))
        # symbol-table nodes
        per = 2 * LEAF_K
        snods, keys = [], [0]
        for s in range(0, max(len(entries), 1), per):
            part = list(zip(entries[s:s + per], offs[s:s + per]))
            at = self.alloc(8 + per * 40)
            self.put(at, b'SNOD' + struct.pack('<BxH', 1, len(part)))
            p = at + 8
            for (nm, addr), o in part:
                self.put(p, struct.pack('<QQII16x', o, addr, 0, 0))
                p += 40
            snods.append(at)
            keys.append(part[-1][1] if part else 0)
        if len(snods) > 2 * INTERNAL_K:
            raise NotImplementedError('groups with more than %d entries' % (2 * INTERNAL_K * per))
        bt = self.alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8)
        self.put(bt, b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF))
        p = bt + 24
        for n, sn in enumerate(snods):
            self.put(p, struct.pack('<QQ', keys[n], sn))
            p += 16
        self.put(p, struct.pack('<Q', keys[len(snods)]))
        msgs = [(0x11, struct.pack('<QQ', bt, heap), None)]
        for k, v in g.attrs.items():
            body, fix = self._attr_msg(k, v)
            msgs.append((0x0C, body, fix))
        return self._object_header(msgs), bt, heap

    def finish_global_heap(self):
        if not self.gheap:
            return
        body = bytearray()
        for n, obj in enumerate(self.gheap):
            body += struct.pack('<HHIQ', n + 1, 0, 0, len(obj)) + obj + b'\0' * ((-len(obj)) % 8)
        size = max(4096, 16 + len(body) + 16)
        body += struct.pack('<HHIQ', 0, 0, 0, size - 16 - len(body))          # free space object
        at = self.alloc(size)
        self.put(at, b'GCOL' + struct.pack('<B3xQ', 1, size))
        self.put(at + 16, body)
        for f in self.gheap_fixups:
            struct.pack_into('<Q', self.buf, f, at)


def write_file(path, root):
    """root: Group.  Writes an HDF5 file of the subset described above."""
    w = _Writer()
    w.alloc(96)                                                   # superblock (56 bytes) + root symbol table entry (40)
    orig_group = w.group

    def group(g):                                                 # nested groups return only their header address
        r = orig_group(g)
        return r[0]
    w.group = group
    oh, bt, heap = orig_group(root)
    w.finish_global_heap()
    eof = len(w.buf)
    sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    sb += struct.pack('<QQII', 0, oh, 1, 0) + struct.pack('<QQ', bt, heap)
    w.put(0, sb)
    with open(path, 'wb') as fh:
        fh.write(bytes(w.buf))
