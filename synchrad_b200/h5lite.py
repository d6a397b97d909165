"""Minimal pure-Python HDF5 reader/writer for the two file layouts of the path:
the tracks file produced by the reference's converters (`converters.py:102-127`) and the spectrum file
written/read by `SynchRad` (`calc.py:274-290, 648-666`).

Used only when h5py is not importable (it is absent from the build image); `trackio.py` prefers h5py.

Supported subset of the HDF5 file format (classic "earliest" layout, what h5py writes by default):
  superblock v0/v1; old-style groups (symbol-table message, v1 B-tree + SNOD + local heap, any depth);
  v1 object headers incl. continuation blocks; datasets with contiguous or compact layout;
  datatypes: little-endian IEEE float32/64, signed/unsigned integers of 1/2/4/8 bytes, fixed-length
  strings, variable-length strings (global heap; read only); simple and scalar dataspaces (v1/v2).
Not supported (raises): chunked/compressed datasets, new-style groups (v2 object headers), references.
The writer emits superblock v0, one symbol-table group per group, contiguous datasets, fixed-length
UTF-8 strings.
"""
import mmap
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b'\x89HDF\r\n\x1a\n'
LEAF_K, INTERNAL_K = 4, 16


# =============================================================================== reader
class _Reader:
    def __init__(self, path):
        self.f = open(path, 'rb')
        self.base = 0
        self._heaps, self._gheaps, self._kids = {}, {}, {}
        try:        # metadata (tens of small reads per object) through a mapping: no seek/read system calls
            self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        except (ValueError, OSError):
            self.mm = None
        try:
            self._superblock()
        except Exception:
            self.close()
            raise

    def close(self):
        if self.mm is not None:
            self.mm.close()
            self.mm = None
        self.f.close()

    def rd(self, addr, n):
        a = self.base + addr
        if self.mm is not None:
            b = self.mm[a:a + n]
        else:
            self.f.seek(a)
            b = self.f.read(n)
        if len(b) != n:
            raise IOError('h5lite: truncated file')
        return b

    def _superblock(self):
        off = 0
        while True:                                  # the superblock may sit at 0, 512, 1024, ...
            self.f.seek(off)
            if self.f.read(8) == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off > (1 << 24):
                raise IOError('h5lite: not an HDF5 file')
        b = self.rd(off, 96)
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError(f'h5lite: superblock version {ver} (only the classic v0/v1 layout)')
        so, sl = b[13], b[14]
        if so != 8 or sl != 8:
            raise NotImplementedError('h5lite: only 8-byte offsets/lengths')
        p = 24 if ver == 0 else 28
        self.base = struct.unpack_from('<Q', b, p)[0]
        p += 32                                      # base, free-space, eof, driver-info addresses
        # root symbol table entry: link name offset, object header address, cache type, reserved, scratch
        _, ohdr, cache = struct.unpack_from('<QQI', b, p)
        self.root = ohdr

    # ---- object headers
    def messages(self, addr):
        b = self.rd(addr, 16)
        if b[:4] == b'OHDR':
            raise NotImplementedError('h5lite: version-2 object headers (file written with libver="latest")')
        if b[0] != 1:
            raise IOError('h5lite: bad object header')
        nmsg, = struct.unpack_from('<H', b, 2)
        size, = struct.unpack_from('<I', b, 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            a, n = blocks.pop(0)
            buf = self.rd(a, n)
            p = 0
            while p + 8 <= n and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from('<HHB', buf, p)
                body = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                   # continuation
                    ca, cl = struct.unpack_from('<QQ', body, 0)
                    blocks.append((ca, cl))
                out.append((mtype, body))
        return out

    # ---- groups
    def _heap_name(self, heap_addr, off):
        cache, key = self._heaps, heap_addr
        if key not in cache:
            h = self.rd(heap_addr, 32)
            if h[:4] != b'HEAP':
                raise IOError('h5lite: bad local heap')
            size, _, data_addr = struct.unpack_from('<QQQ', h, 8)
            cache[key] = self.rd(data_addr, size)
        d = cache[key]
        return d[off:d.index(b'\x00', off)].decode()

    def _btree_entries(self, addr, heap_addr, out):
        h = self.rd(addr, 24)
        if h[:4] == b'SNOD':
            n, = struct.unpack_from('<H', h, 6)
            b = self.rd(addr + 8, n * 40)
            for i in range(n):
                name_off, ohdr = struct.unpack_from('<QQ', b, i * 40)
                out[self._heap_name(heap_addr, name_off)] = ohdr
            return
        if h[:4] != b'TREE' or h[4] != 0:
            raise IOError('h5lite: bad group B-tree node')
        n, = struct.unpack_from('<H', h, 6)
        b = self.rd(addr + 24, (2 * n + 1) * 8)
        for i in range(n):
            child, = struct.unpack_from('<Q', b, 8 + 16 * i)
            self._btree_entries(child, heap_addr, out)

    def children(self, addr):
        """name -> object-header address of a group's members (None if `addr` is not a group); cached, so that
        `f['tracks/123/x']` does not re-walk the B-tree of a group with 1e5 members on every access."""
        if addr not in self._kids:
            self._kids[addr] = self._children(addr)
        return self._kids[addr]

    def _children(self, addr):
        for mtype, body in self.messages(addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from('<QQ', body, 0)
                out = {}
                self._btree_entries(btree, heap, out)
                return out
            if mtype in (0x0002, 0x0006):
                raise NotImplementedError('h5lite: new-style (link message) groups')
        return None                                  # not a group

    # ---- datasets
    @staticmethod
    def _dtype(body):
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size, = struct.unpack_from('<I', body, 4)
        if bits & 1 and cls in (0, 1):
            raise NotImplementedError('h5lite: big-endian data')
        if cls == 0:
            return np.dtype(('<i' if bits & 0x08 else '<u') + str(size)), None
        if cls == 1:
            return np.dtype('<f' + str(size)), None
        if cls == 3:
            return np.dtype('S' + str(size)), 'str'
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise NotImplementedError('h5lite: variable-length sequences')
            return np.dtype('V16'), 'vlen_str'
        raise NotImplementedError(f'h5lite: datatype class {cls}')

    @staticmethod
    def _shape(body):
        ver, rank = body[0], body[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:
                return None                           # null dataspace
            p = 4
        else:
            raise NotImplementedError('h5lite: dataspace version')
        return tuple(struct.unpack_from('<Q', body, p + 8 * i)[0] for i in range(rank))

    def _gheap_obj(self, addr, index):
        cache, key = self._gheaps, addr
        if key not in cache:
            h = self.rd(addr, 16)
            if h[:4] != b'GCOL':
                raise IOError('h5lite: bad global heap')
            size, = struct.unpack_from('<Q', h, 8)
            buf = self.rd(addr, size)
            objs, p = {}, 16
            while p + 16 <= size:
                idx, _, _, osz = struct.unpack_from('<HHIQ', buf, p)
                if idx == 0:
                    break
                objs[idx] = buf[p + 16:p + 16 + osz]
                p += 16 + ((osz + 7) & ~7)
            cache[key] = objs
        return cache[key][index]

    def dataset_info(self, addr):
        """(numpy dtype, kind, shape, layout) of the dataset whose object header sits at `addr`."""
        dt = kind = shape = None
        layout = None
        for mtype, body in self.messages(addr):
            if mtype == 0x0003:
                dt, kind = self._dtype(body)
            elif mtype == 0x0001:
                shape = self._shape(body)
            elif mtype == 0x0008:
                ver = body[0]
                if ver != 3:
                    raise NotImplementedError(f'h5lite: data layout version {ver}')
                if body[1] == 0:                       # compact
                    n, = struct.unpack_from('<H', body, 2)
                    layout = ('compact', body[4:4 + n])
                elif body[1] == 1:                     # contiguous
                    a, n = struct.unpack_from('<QQ', body, 2)
                    layout = ('contig', a, n)
                else:
                    raise NotImplementedError('h5lite: chunked datasets are not supported')
        if dt is None or layout is None:
            raise KeyError('h5lite: object is not a dataset')
        return dt, kind, shape, layout

    def read_direct(self, addr, dest):
        """Read a numeric dataset straight into the C-contiguous array `dest` (same element count): one
        `readinto` from the file when the stored type is dest's, a converting copy otherwise."""
        dt, kind, shape, layout = self.dataset_info(addr)
        count = (int(np.prod(shape)) if len(shape) else 1) if shape is not None else 0
        if count != dest.size:
            raise ValueError(f'h5lite: dataset holds {count} elements, destination {dest.size}')
        if (kind is None and dt == dest.dtype and layout[0] == 'contig' and layout[1] != UNDEF and count
                and dest.flags.c_contiguous):
            self.f.seek(self.base + layout[1])
            if self.f.readinto(memoryview(dest).cast('B')) != count * dt.itemsize:
                raise IOError('h5lite: truncated file')
        elif count:
            dest[...] = np.asarray(self.read_dataset(addr)).reshape(dest.shape)

    def read_dataset(self, addr):
        dt, kind, shape, layout = self.dataset_info(addr)
        if shape is None:
            return np.zeros((0,), dtype=dt)
        count = int(np.prod(shape)) if len(shape) else 1
        nbytes = count * dt.itemsize
        if layout[0] == 'compact':
            raw = layout[1][:nbytes]
        elif layout[1] == UNDEF or nbytes == 0:
            raw = b'\x00' * nbytes
        else:
            raw = self.rd(layout[1], nbytes)
        arr = np.frombuffer(raw, dtype=dt, count=count).reshape(shape)
        if kind == 'vlen_str':
            vals = []
            for rec in arr.reshape(-1):
                ln, ga, gi = struct.unpack('<IQI', rec.tobytes())
                vals.append(self._gheap_obj(ga, gi)[:ln] if ga not in (0, UNDEF) else b'')
            if not shape:
                return vals[0]                        # bytes, like h5py >= 3
            return np.array(vals, dtype=object).reshape(shape)
        if kind == 'str':
            if not shape:
                return arr.reshape(-1)[0].split(b'\x00')[0]
            return arr.copy()
        return arr[()] if not shape else arr.copy()


class _RNode:
    def __init__(self, rd, addr):
        self._rd, self._addr, self._kids = rd, addr, None

    def _children(self):
        if self._kids is None:
            self._kids = self._rd.children(self._addr)
            if self._kids is None:
                raise KeyError('h5lite: not a group')
        return self._kids

    def keys(self):
        return list(self._children().keys())

    @property
    def shape(self):
        """Dataset shape, from the header only (h5py: `dset.shape`)."""
        shape = self._rd.dataset_info(self._addr)[2]
        return tuple(int(v) for v in shape) if shape is not None else (0,)

    @property
    def dtype(self):
        """Stored element type (h5py: `dset.dtype`)."""
        return self._rd.dataset_info(self._addr)[0]

    def read_direct(self, dest):
        """h5py's `dset.read_direct(dest)`: fill `dest` without an intermediate array."""
        self._rd.read_direct(self._addr, dest)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        if path == ():
            return self._rd.read_dataset(self._addr)
        if not isinstance(path, str):       # h5py's `dset[::2, 3, :]`: whole dataset read, then NumPy indexing
            return self._rd.read_dataset(self._addr)[path]
        node = self
        for part in [p for p in path.split('/') if p]:
            kids = node._children()
            if part not in kids:
                raise KeyError(path)
            node = _RNode(self._rd, kids[part])
        return node


# =============================================================================== writer
def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _msg(mtype, body):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), 0) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        exp_bits, man_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        head = struct.pack('<BBBBI', 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        return head + struct.pack('<HHBBBBI', 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in 'iu':
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack('<HH', 0, dt.itemsize * 8)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x10, 0, 0, dt.itemsize)     # null-terminated, UTF-8
    raise TypeError(f'h5lite: cannot store dtype {dt}')


def _space_msg(shape):
    if shape == ():
        return struct.pack('<BBBB4x', 1, 0, 0, 0)
    return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + b''.join(struct.pack('<Q', s) for s in shape)


class _Writer:
    def __init__(self, path):
        self.f = open(path, 'wb')
        self.pos = 0
        self.tree = {}                                # nested dict: name -> dict | ndarray

    def alloc(self, n):
        a = self.pos
        self.pos += (n + 7) & ~7
        return a

    def put(self, addr, data):
        self.f.seek(addr)
        self.f.write(data)

    def set(self, path, value):
        parts = [p for p in path.split('/') if p]
        node = self.tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise ValueError(f'h5lite: {p} is a dataset')
        if isinstance(value, str):
            value = np.array(value.encode('utf-8') + b'\x00')
        elif isinstance(value, bytes):
            value = np.array(value + b'\x00')
        else:
            value = np.asarray(value)
            if value.dtype == object or value.dtype.kind == 'U':
                raise TypeError('h5lite: arrays of Python strings are not supported')
            if value.dtype.kind == 'b':
                value = value.astype(np.uint8)
            if value.dtype.byteorder == '>':
                value = value.astype(value.dtype.newbyteorder('<'))
        node[parts[-1]] = value

    # ---- serialisation
    def _write_dataset(self, arr):
        arr = np.asarray(arr)
        data = arr.tobytes(order='C')
        daddr = self.alloc(len(data)) if data else UNDEF
        if data:
            self.put(daddr, data)
        msgs = (_msg(0x0001, _space_msg(arr.shape)) + _msg(0x0003, _dtype_msg(arr.dtype)) +
                _msg(0x0005, struct.pack('<BBBB', 2, 2, 0, 0)) +
                _msg(0x0008, struct.pack('<BBQQ', 3, 1, daddr, len(data))))
        hdr = struct.pack('<BBHII4x', 1, 0, 4, 1, len(msgs)) + msgs
        a = self.alloc(len(hdr))
        self.put(a, hdr)
        return a

    def _write_group(self, kids):
        """kids: dict name -> object header address.  Returns (ohdr, btree, heap) addresses."""
        names = sorted(kids, key=lambda s: s.encode())
        heap = bytearray(b'\x00' * 8)                  # offset 0: empty string (key 0)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b'\x00')
        heap_data = self.alloc(len(heap) + 16)
        self.put(heap_data, bytes(heap) + struct.pack('<QQ', 1, 16))            # trailing free block
        heap_addr = self.alloc(32)
        self.put(heap_addr, b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap) + 16, len(heap), heap_data))

        # leaves: symbol table nodes with up to 2*LEAF_K entries
        level = []                                     # (address, key offset of largest name)
        per = 2 * LEAF_K
        for i in range(0, max(len(names), 1), per):
            chunk = names[i:i + per]
            body = b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk))
            for n in chunk:
                body += struct.pack('<QQII16x', offs[n], kids[n], 0, 0)
            body += b'\x00' * (40 * (per - len(chunk)))
            a = self.alloc(len(body))
            self.put(a, body)
            level.append((a, offs[chunk[-1]] if chunk else 0))
        depth = 0
        per_i = 2 * INTERNAL_K
        while True:
            nxt = []
            for i in range(0, len(level), per_i):
                chunk = level[i:i + per_i]
                body = b'TREE' + struct.pack('<BBHQQ', 0, depth, len(chunk), UNDEF, UNDEF)
                body += struct.pack('<Q', 0)           # key 0
                for (a, k) in chunk:
                    body += struct.pack('<QQ', a, k)
                body += b'\x00' * (16 * (per_i - len(chunk)))
                na = self.alloc(len(body))
                nxt.append((na, chunk[-1][1], body))
            # sibling pointers within the level
            for j, (na, k, body) in enumerate(nxt):
                left = nxt[j - 1][0] if j > 0 else UNDEF
                right = nxt[j + 1][0] if j + 1 < len(nxt) else UNDEF
                body = body[:8] + struct.pack('<QQ', left, right) + body[24:]
                self.put(na, body)
            level = [(na, k) for (na, k, _) in nxt]
            depth += 1
            if len(level) == 1:
                break
        btree = level[0][0]
        msgs = _msg(0x0011, struct.pack('<QQ', btree, heap_addr))
        hdr = struct.pack('<BBHII4x', 1, 0, 1, 1, len(msgs)) + msgs
        a = self.alloc(len(hdr))
        self.put(a, hdr)
        return a, btree, heap_addr

    def _emit(self, node):
        if isinstance(node, dict):
            kids = {name: self._emit(child)[0] for name, child in node.items()}
            return self._write_group(kids)
        return (self._write_dataset(node), None, None)

    def close(self):
        self.pos = 96                                  # superblock v0 + root symbol table entry
        root, btree, heap = self._emit(self.tree)
        sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, self.pos, UNDEF)
        sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap)
        assert len(sb) == 96, len(sb)
        self.put(0, sb)
        self.f.seek(0, 2)
        if self.f.tell() < self.pos:
            self.f.truncate(self.pos)
        self.f.close()


# =============================================================================== public API
class File:
    """Tiny h5py.File look-alike: `f[path][()]`, `f[path] = value`, `name in f[group]`, `.keys()`."""

    def __init__(self, path, mode='r'):
        self.mode = mode
        if mode == 'r':
            self._r = _Reader(path)
            self._root = _RNode(self._r, self._r.root)
        elif mode == 'w':
            self._w = _Writer(path)
        else:
            raise ValueError("h5lite.File: mode must be 'r' or 'w'")

    def __getitem__(self, path):
        if self.mode != 'r':
            raise IOError('h5lite: file is open for writing')
        return self._root[path]

    def __setitem__(self, path, value):
        if self.mode != 'w':
            raise IOError('h5lite: file is open read-only')
        self._w.set(path, value)

    def __contains__(self, path):
        return path in self._root

    def keys(self):
        return self._root.keys()

    def close(self):
        if self.mode == 'r':
            self._r.close()
        else:
            self._w.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
