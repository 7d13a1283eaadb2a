"""Keras `.h5` weight files without h5py.  Host-side only: no arithmetic, no device code.

The reference checkpoints through Keras besides `tf.train.Saver`: `model.save_weights(path + '.h5')` and
`model.load_weights(path)` (experiments_antisymmetric_resnet_v6.ipynb cells 8 / 11, v7 the same; the Saver side is
`training/training.py:848-872` -> `tf_bundle.py`).  h5py / HDF5 are not in this image, so this module restates the
subset of the HDF5 file format those files use (HDF5 File Format Specification, version 0 superblock, the layout
h5py writes with its default `libver='earliest'`):

    superblock v0/v1 -> root symbol-table entry -> version-1 object headers (with continuation blocks)
    old-style groups: symbol-table message -> v1 B-tree ('TREE') -> symbol nodes ('SNOD') + local heap ('HEAP')
    datasets: dataspace v1/v2, datatype classes 0 (integer) / 1 (float) / 3 (fixed string) / 9 (variable-length
    string through the global heap 'GCOL'), layout v1-v3 compact / contiguous / chunked-without-filters
    attributes: message versions 1-3

and the Keras layout on top of it (`keras/engine/saving.py: save_weights_to_hdf5_group`):

    /            attrs layer_names [S], backend, keras_version
    /<layer>     attr  weight_names [S]   e.g. b'res_1_1_branch2/a:0'
    /<layer>/<weight_name>                one float32 dataset per variable (the '/' in the name nests a group)

The READER is pinned on a file the HDF5 C library itself wrote (the MATLAB 7.3 fixture shipped with scipy's test
data, copied to tests/golden/hdf5_library_written.mat: 512-byte user block, TREE / SNOD / HEAP groups, attributes);
the WRITER is pinned only through the reader (no HDF5 library exists here to open its files): "parity unpinned"
for files going TO Keras, stated in DESIGN.md §7.
"""
import struct
from collections import OrderedDict

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------------------------
class H5Dataset:
    def __init__(self, name, array, attrs):
        self.name, self.array, self.attrs = name, array, attrs


class H5Group:
    def __init__(self, name, attrs):
        self.name, self.attrs, self.children = name, attrs, OrderedDict()

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, H5Group) or part not in node.children:
                raise KeyError(path)
            node = node.children[part]
        return node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def visit(self, prefix=""):
        """Yield (path, node) depth-first, link-name order."""
        for k, v in self.children.items():
            p = prefix + "/" + k if prefix else k
            yield p, v
            if isinstance(v, H5Group):
                yield from v.visit(p)


class _Reader:
    def __init__(self, data):
        self.d = data
        sb = -1
        off = 0
        while off < len(data):                    # the superblock sits at 0 or at a power of two >= 512 (user block)
            if data[off:off + 8] == SIGNATURE:
                sb = off
                break
            off = 512 if off == 0 else off * 2
        if sb < 0:
            raise H5FormatError("not an HDF5 file (no signature at 0, 512, 1024, ...)")
        ver = data[sb + 8]
        if ver > 1:
            raise H5FormatError("superblock version %d (libver='latest' files) is not supported; "
                                "Keras / h5py write version 0 by default" % ver)
        self.O, self.L = data[sb + 13], data[sb + 14]
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5FormatError("offset / length sizes %d / %d" % (self.O, self.L))
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", data, sb + 16)
        p = sb + 24 + (4 if ver == 1 else 0)
        self.base = self._off(p)
        if self.base == 0 and sb != 0:
            self.base = sb            # files whose base address was left 0 below a user block
        self.eof = self._off(p + 2 * self.O)
        self.root_entry = p + 4 * self.O
        self._oh_cache = {}

    # -- primitives ------------------------------------------------------------------------------------------
    def _off(self, p):
        v = int.from_bytes(self.d[p:p + self.O], "little")
        return None if v == (1 << (8 * self.O)) - 1 else v

    def _len(self, p):
        return int.from_bytes(self.d[p:p + self.L], "little")

    def _at(self, addr):
        a = addr + self.base
        if a < 0 or a >= len(self.d):
            raise H5FormatError("address %d outside the file (%d bytes)" % (a, len(self.d)))
        return a

    # -- object headers --------------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, payload bytes)] of a version-1 object header, continuation blocks followed."""
        if addr in self._oh_cache:
            return self._oh_cache[addr]
        p = self._at(addr)
        d = self.d
        if d[p:p + 4] == b"OHDR":
            raise H5FormatError("version-2 object headers (libver='latest') are not supported")
        if d[p] != 1:
            raise H5FormatError("object header version %d at %d" % (d[p], addr))
        nmsg, = struct.unpack_from("<H", d, p + 2)
        hsize, = struct.unpack_from("<I", d, p + 8)
        blocks = [(p + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", d, q)
                body = d[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x10:
                    blocks.append((self._at(self._off_b(body, 0)), self._len_b(body, self.O)))
                out.append((mtype, mflags, body))
        self._oh_cache[addr] = out
        return out

    def _off_b(self, b, p):
        return int.from_bytes(b[p:p + self.O], "little")

    def _len_b(self, b, p):
        return int.from_bytes(b[p:p + self.L], "little")

    # -- datatype / dataspace --------------------------------------------------------------------------------
    def datatype(self, b):
        """-> (kind, numpy dtype or None, element size, total bytes of the message)."""
        cls, ver = b[0] & 0x0F, b[0] >> 4
        bits = b[1] | (b[2] << 8) | (b[3] << 16)
        size, = struct.unpack_from("<I", b, 4)
        if cls == 0:
            order = ">" if bits & 1 else "<"
            return "num", np.dtype("%s%s%d" % (order, "i" if bits & 8 else "u", size)), size, 12
        if cls == 1:
            order = ">" if bits & 1 else "<"
            return "num", np.dtype("%sf%d" % (order, size)), size, 20
        if cls == 3:
            return "str", np.dtype("S%d" % size), size, 8
        if cls == 9:
            base = self.datatype(b[8:])
            if bits & 0x0F != 1:
                raise H5FormatError("variable-length sequences are not supported (only strings)")
            return "vlen_str", None, size, 8 + base[3]
        raise H5FormatError("datatype class %d (version %d) is not supported" % (cls, ver))

    def dataspace(self, b):
        ver, rank, flags = b[0], b[1], b[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if b[3] == 2:                        # null dataspace
                return None
            p = 4
        else:
            raise H5FormatError("dataspace version %d" % ver)
        return tuple(self._len_b(b, p + i * self.L) for i in range(rank))

    def _global_heap_object(self, addr, index):
        p = self._at(addr)
        if self.d[p:p + 4] != b"GCOL":
            raise H5FormatError("no global heap collection at %d" % addr)
        size = self._len(p + 8)
        q, end = p + 8 + self.L, p + size
        while q + 8 + self.L <= end:
            idx, = struct.unpack_from("<H", self.d, q)
            osize = self._len(q + 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(self.d[q + 8 + self.L:q + 8 + self.L + osize])
            q += 8 + self.L + ((osize + 7) & ~7)
        raise H5FormatError("global heap object %d not found in the collection at %d" % (index, addr))

    def decode(self, dt, shape, raw):
        kind, npdt, esize, _ = dt
        n = 1
        for s in (shape or ()):
            n *= s
        if shape is None:
            return None
        if kind in ("num", "str"):
            if len(raw) < n * esize:
                raise H5FormatError("%d bytes of data for %d elements of %d bytes" % (len(raw), n, esize))
            return np.frombuffer(bytes(raw[:n * esize]), dtype=npdt).reshape(shape).copy()
        out = []
        for i in range(n):                           # variable-length strings: (length, heap address, object index)
            q = i * esize
            ln, = struct.unpack_from("<I", raw, q)
            addr = self._off_b(raw, q + 4)
            idx, = struct.unpack_from("<I", raw, q + 4 + self.O)
            out.append(self._global_heap_object(addr, idx)[:ln] if ln else b"")
        arr = np.empty(n, dtype=object)
        arr[:] = out
        return arr.reshape(shape)

    # -- attributes ------------------------------------------------------------------------------------------
    def attribute(self, b):
        ver = b[0]
        nsize, tsize, ssize = struct.unpack_from("<HHH", b, 2)
        p = 8
        if ver == 3:
            p = 9
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        if ver not in (1, 2, 3):
            raise H5FormatError("attribute message version %d" % ver)
        name = bytes(b[p:p + nsize]).split(b"\0", 1)[0].decode("utf8")
        p += pad(nsize)
        dt = self.datatype(b[p:p + tsize])
        p += pad(tsize)
        shape = self.dataspace(b[p:p + ssize])
        p += pad(ssize)
        return name, self.decode(dt, shape, b[p:])

    # -- datasets --------------------------------------------------------------------------------------------
    def _chunked(self, btree, chunk_dims, shape, dt):
        esize = dt[2]
        rank = len(shape)
        if int(np.prod(shape, dtype=object)) * esize > 2 * len(self.d) + (1 << 20):
            raise H5FormatError("chunked dataset of shape %s cannot fit this %d-byte file" % (shape, len(self.d)))
        out = np.zeros(shape, dtype=dt[1])

        def walk(addr):
            p = self._at(addr)
            if self.d[p:p + 4] != b"TREE" or self.d[p + 4] != 1:
                raise H5FormatError("no chunk B-tree node at %d" % addr)
            level = self.d[p + 5]
            used, = struct.unpack_from("<H", self.d, p + 6)
            q = p + 8 + 2 * self.O
            ksize = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize, fmask = struct.unpack_from("<II", self.d, q)
                offs = struct.unpack_from("<%dQ" % rank, self.d, q + 8)
                child = self._off(q + ksize)
                q += ksize + self.O
                if level:
                    walk(child)
                    continue
                if fmask or csize != int(np.prod(chunk_dims)) * esize:
                    raise H5FormatError("filtered (compressed) chunks are not supported")
                a = self._at(child)
                chunk = np.frombuffer(bytes(self.d[a:a + csize]), dtype=dt[1]).reshape(chunk_dims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk_dims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if btree is not None:
            walk(btree)
        return out

    def dataset(self, msgs):
        dt = shape = None
        layout = None
        for mtype, _, b in msgs:
            if mtype == 0x01:
                shape = self.dataspace(b)
            elif mtype == 0x03:
                dt = self.datatype(b)
            elif mtype == 0x08:
                layout = b
            elif mtype == 0x0B:
                raise H5FormatError("filtered (compressed) datasets are not supported")
        if dt is None or layout is None:
            raise H5FormatError("dataset without datatype / layout message")
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 0:
                size, = struct.unpack_from("<H", layout, 2)
                return self.decode(dt, shape, layout[4:4 + size])
            if cls == 1:
                addr = self._off_b(layout, 2)
                if addr == (1 << (8 * self.O)) - 1:                 # never written: fill value 0
                    return self.decode(dt, shape, bytes(n * dt[2]))
                a = self._at(addr)
                return self.decode(dt, shape, self.d[a:a + n * dt[2]])
            if cls == 2:
                nd = layout[2]
                bt = self._off_b(layout, 3)
                dims = struct.unpack_from("<%dI" % nd, layout, 3 + self.O)
                if dt[0] != "num":
                    raise H5FormatError("chunked string datasets are not supported")
                return self._chunked(None if bt == (1 << (8 * self.O)) - 1 else bt, dims[:-1], shape, dt)
            raise H5FormatError("layout class %d" % cls)
        if ver in (1, 2):
            nd, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr = self._off_b(layout, p)
                p += self.O
            dims = struct.unpack_from("<%dI" % nd, layout, p)
            p += 4 * nd
            if cls == 1:
                a = self._at(addr)
                return self.decode(dt, shape, self.d[a:a + n * dt[2]])
            if cls == 2:
                return self._chunked(addr, dims[:-1], shape, dt)
            size, = struct.unpack_from("<I", layout, p)
            return self.decode(dt, shape, layout[p + 4:p + 4 + size])
        raise H5FormatError("data layout message version %d" % ver)

    # -- groups ----------------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, off):
        p = self._at(heap_addr)
        if self.d[p:p + 4] != b"HEAP":
            raise H5FormatError("no local heap at %d" % heap_addr)
        seg = self._at(self._off(p + 8 + 2 * self.L))
        q = seg + off
        e = self.d.index(b"\0", q)
        return bytes(self.d[q:e]).decode("utf8")

    def _symbol_entries(self, btree, heap):
        out = []

        def walk(addr):
            p = self._at(addr)
            sig = self.d[p:p + 4]
            if sig == b"TREE":
                if self.d[p + 4] != 0:
                    raise H5FormatError("group B-tree node of type %d" % self.d[p + 4])
                used, = struct.unpack_from("<H", self.d, p + 6)
                q = p + 8 + 2 * self.O + self.L            # first child follows key 0
                for _ in range(used):
                    walk(self._off(q))
                    q += self.O + self.L
            elif sig == b"SNOD":
                cnt, = struct.unpack_from("<H", self.d, p + 6)
                q = p + 8
                for _ in range(cnt):
                    name = self._heap_string(heap, self._off(q))
                    out.append((name, self._off(q + self.O)))
                    q += 2 * self.O + 24
            else:
                raise H5FormatError("expected TREE or SNOD at %d, found %r" % (addr, bytes(sig)))
        walk(btree)
        return out

    def node(self, addr, name, depth=0):
        if depth > 64:
            raise H5FormatError("group nesting deeper than 64 (a link cycle?)")
        msgs = self.messages(addr)
        attrs = OrderedDict()
        stab = None
        links = []
        for mtype, _, b in msgs:
            if mtype == 0x0C:
                k, v = self.attribute(b)
                attrs[k] = v
            elif mtype == 0x11:
                stab = (self._off_b(b, 0), self._off_b(b, self.O))
            elif mtype == 0x06:                         # new-style compact group: one link message per child
                flags = b[1]
                p = 2
                ltype = 0
                if flags & 8:
                    ltype = b[p]
                    p += 1
                if flags & 4:
                    p += 8
                if flags & 16:
                    p += 1
                w = 1 << (flags & 3)
                ln = int.from_bytes(b[p:p + w], "little")
                p += w
                lname = bytes(b[p:p + ln]).decode("utf8")
                p += ln
                if ltype == 0:
                    links.append((lname, self._off_b(b, p)))
        if stab is None and not links and any(t in (0x03, 0x08) for t, _, _ in msgs):
            return H5Dataset(name, self.dataset(msgs), attrs)
        g = H5Group(name, attrs)
        entries = links + (self._symbol_entries(*stab) if stab else [])
        for cname, caddr in entries:
            g.children[cname] = self.node(caddr, cname, depth + 1)
        return g

    def root(self):
        return self.node(self._off(self.root_entry + self.O), "/")


def read_h5(path_or_bytes):
    """Parse an HDF5 file (the subset in the module docstring) into an H5Group tree of numpy arrays."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    try:
        return _Reader(bytes(data)).root()
    except H5FormatError:
        raise
    except (struct.error, IndexError, ValueError, KeyError, OverflowError, RecursionError, MemoryError, TypeError) as e:
        raise H5FormatError("truncated or corrupt HDF5 file: %s: %s" % (type(e).__name__, e))


# ----------------------------------------------------------------------------------------------------------------
# writer (superblock v0, version-1 object headers, symbol-table groups, contiguous datasets)
# ----------------------------------------------------------------------------------------------------------------
_LEAF_K, _INTERNAL_K = 32, 16          # 64 links per symbol node, 32 nodes per B-tree node: 2048 links per group


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 4:
            exp_loc, exp_size, man_size, bias = 23, 8, 23, 127
        elif dt.itemsize == 8:
            exp_loc, exp_size, man_size, bias = 52, 11, 52, 1023
        else:
            raise H5FormatError("float%d" % (8 * dt.itemsize))
        bits = 0x20 | ((8 * dt.itemsize - 1) << 8)                     # little-endian, implied mantissa msb, sign bit
        return struct.pack("<BBBBI", 0x11, bits & 0xFF, (bits >> 8) & 0xFF, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, exp_loc, exp_size, 0, man_size, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 8 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))     # null-padded ASCII
    raise H5FormatError("dtype %s cannot be written" % dt)


def _space_message(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


class VlenStr:
    """A scalar attribute stored as an HDF5 variable-length string in a global heap collection -- what h5py writes
    for a Python `bytes` / `str` value (Keras' `backend` and `keras_version` attributes)."""

    def __init__(self, value):
        self.value = value.encode("utf8") if isinstance(value, str) else bytes(value)


def _attr_message(name, value, writer=None):
    if isinstance(value, VlenStr):
        nm = name.encode("utf8") + b"\0"
        # class 9 (variable length), type = string, null-terminated ASCII; base type H5T_C_S1 (class 3 string of size 1)
        dt = struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)
        sp = _space_message(())
        data = struct.pack("<IQI", len(value.value), writer.global_heap(value.value), 1)
        return (0x0C, struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + data)
    if isinstance(value, (bytes, str)):
        value = np.array(value.encode("utf8") if isinstance(value, str) else value, dtype="S")
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = np.char.encode(value, "utf8")
    if value.dtype.kind == "S" and value.dtype.itemsize == 0:
        value = value.astype("S1")
    nm = name.encode("utf8") + b"\0"
    dt, sp = _dtype_message(value.dtype), _space_message(value.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + \
        np.ascontiguousarray(value).tobytes()
    if len(body) > 0xFFF0:
        raise H5FormatError("attribute %r is %d bytes: HDF5 object-header messages hold < 64 KiB "
                            "(Keras splits such lists into weight_names0, weight_names1, ...)" % (name, len(body)))
    return (0x0C, body)


def _object_header(msgs):
    body = b"".join(struct.pack("<HHBBBB", t, len(_pad8(b)), 0, 0, 0, 0) + _pad8(b) for t, b in msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                  # superblock v0 with 8-byte offsets: 56 + 40-byte root entry

    def alloc(self, b):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += b
        return addr

    def global_heap(self, payload):
        """One global heap collection ('GCOL', 4096 bytes, the library's minimum) holding `payload` as object 1."""
        obj = struct.pack("<HHIQ", 1, 0, 0, len(payload)) + _pad8(payload)
        free = 4096 - 16 - len(obj)
        if free < 16:
            raise H5FormatError("variable-length string of %d bytes" % len(payload))
        col = b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, 4096) + obj + struct.pack("<HHIQ", 0, 0, 0, free)
        return self.alloc(col + bytes(4096 - len(col)))

    def dataset(self, arr, attrs):
        arr = np.asarray(arr, order="C")              # (ascontiguousarray would turn a scalar into shape (1,))
        if arr.dtype.kind == "f" and arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        daddr = self.alloc(raw) if raw else UNDEF
        msgs = [(0x01, _space_message(arr.shape)), (0x03, _dtype_message(arr.dtype)),
                (0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                       # fill value v2: late alloc, undefined
                (0x08, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]
        msgs += [_attr_message(k, v, self) for k, v in attrs.items()]
        return self.alloc(_object_header(msgs))

    def group(self, children, attrs):
        """children: {name: ('g', children, attrs) | ('d', array, attrs)} -> (header address, btree, heap)."""
        entries = []
        for name, (kind, payload, a) in children.items():
            if not name or "/" in name:
                raise H5FormatError("bad link name %r" % name)
            if kind == "g":
                entries.append((name.encode("utf8"), self.group(payload, a)))
            else:
                entries.append((name.encode("utf8"), (self.dataset(payload, a), None, None)))
        entries.sort(key=lambda e: e[0])                       # symbol nodes are ordered by strcmp of the link names
        if len(entries) > 2 * _LEAF_K * 2 * _INTERNAL_K:
            raise H5FormatError("%d links in one group (this writer holds %d)" % (len(entries), 4 * _LEAF_K * _INTERNAL_K))
        heap = bytearray(8)                                     # offset 0: the empty string (key 0 of the B-tree)
        offs = []
        for nm, _ in entries:
            offs.append(len(heap))
            heap += nm + b"\0"
            heap += b"\0" * (-len(heap) % 8)
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)                      # one free block: next = 1 (none), size 16
        seg = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free_off, seg))
        per = 2 * _LEAF_K
        snods, keys = [], [0]
        for i in range(0, len(entries), per):
            part = entries[i:i + per]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for j, (nm, (haddr, bt, hp)) in enumerate(part):
                if bt is None:
                    body += struct.pack("<QQII", offs[i + j], haddr, 0, 0) + bytes(16)
                else:
                    body += struct.pack("<QQII", offs[i + j], haddr, 1, 0) + struct.pack("<QQ", bt, hp)
            body += bytes(40 * (per - len(part)))
            snods.append(self.alloc(body))
            keys.append(offs[i + len(part) - 1])
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, s in enumerate(snods):
            tree += struct.pack("<QQ", keys[i], s)
        tree += struct.pack("<Q", keys[len(snods)])
        tree += bytes((2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - (len(tree) - 24))
        bt_addr = self.alloc(tree)
        msgs = [(0x11, struct.pack("<QQ", bt_addr, heap_addr))] + [_attr_message(k, v, self) for k, v in attrs.items()]
        return self.alloc(_object_header(msgs)), bt_addr, heap_addr

    def finish(self, root):
        haddr, bt, hp = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, haddr, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, tree, attrs=None):
    """tree: nested {name: ndarray | dict | (ndarray_or_dict, attrs)}; attrs: root attributes."""
    def conv(t):
        out = OrderedDict()
        for k, v in t.items():
            a = {}
            if isinstance(v, tuple):
                v, a = v
            out[k] = ("g", conv(v), a) if isinstance(v, dict) else ("d", np.asarray(v), a)
        return out
    w = _Writer()
    data = w.finish(w.group(conv(tree), attrs or {}))
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data


# ----------------------------------------------------------------------------------------------------------------
# the Keras layout
# ----------------------------------------------------------------------------------------------------------------
def save_keras_weights(path, layers, backend="tensorflow", keras_version="2.2.4-tf"):
    """`layers`: ordered {layer name: ordered {weight name (e.g. 'res_1_1_branch2/a:0'): ndarray}} -- the structure
    `save_weights_to_hdf5_group` writes for `model.layers` / `layer.weights` (layers without weights included, with
    an empty weight list, as Keras does)."""
    tree = OrderedDict()
    names = []
    for lname, weights in layers.items():
        names.append(lname.encode("utf8"))
        sub = OrderedDict()
        wnames = []
        for wname, arr in weights.items():
            wnames.append(wname.encode("utf8"))
            node = sub
            parts = wname.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, OrderedDict())
            node[parts[-1]] = np.asarray(arr, dtype=np.float32)
        wn = np.array(wnames, dtype="S") if wnames else np.zeros((0,), dtype="S1")
        tree[lname] = (sub, {"weight_names": wn})
    ln = np.array(names, dtype="S") if names else np.zeros((0,), dtype="S1")
    return write_h5(path, tree, {"layer_names": ln, "backend": VlenStr(backend), "keras_version": VlenStr(keras_version)})


def _names(v):
    out = []
    for s in np.asarray(v).reshape(-1):
        s = s if isinstance(s, bytes) else bytes(s)
        out.append(s.rstrip(b"\0").decode("utf8"))
    return out


def _chunked_attr(attrs, name):
    """Keras `load_attributes_from_hdf5_group`: `name`, or `name0`, `name1`, ... when the list passed 64 KiB."""
    if name in attrs:
        return _names(attrs[name])
    out, i = [], 0
    while "%s%d" % (name, i) in attrs:
        out += _names(attrs["%s%d" % (name, i)])
        i += 1
    if i == 0:
        raise H5FormatError("no %r attribute: not a Keras weights file" % name)
    return out


def load_keras_weights(path):
    """-> ordered {layer name: ordered {weight name: ndarray}} of a Keras weights file; a full-model file
    (`model.save`) keeps the same structure under its 'model_weights' group and is accepted too."""
    root = read_h5(path)
    if "layer_names" not in root.attrs and "model_weights" in root.children:
        root = root["model_weights"]
    out = OrderedDict()
    for lname in _chunked_attr(root.attrs, "layer_names"):
        g = root[lname]
        ws = OrderedDict()
        for wname in _chunked_attr(g.attrs, "weight_names"):
            node = g[wname]
            if not isinstance(node, H5Dataset):
                raise H5FormatError("%s/%s is not a dataset" % (lname, wname))
            ws[wname] = node.array
        out[lname] = ws
    return out


def _net_layers(net):
    """EulerNet -> the Keras {layer: {weight name: array}} structure (reference variable names + ':0')."""
    from .checkpoint import export_reference_variables
    layers = OrderedDict()
    for key, arr in export_reference_variables(net).items():
        lname = key.rsplit("/", 1)[0]
        layers.setdefault(lname, OrderedDict())[key + ":0"] = arr
    return layers


def save_weights_h5(net, path):
    """`model.save_weights(path)` of the reference's notebooks for an EulerNet: one group per weighted layer, the
    C+4 variables of every antisymmetric layer under the reference's names."""
    save_keras_weights(path, _net_layers(net))


def load_weights_h5(net, path):
    """`model.load_weights(path)`: by layer and weight NAME (Keras `by_name` semantics; topological order is not
    assumed), shapes checked by `import_reference_variables`."""
    from .checkpoint import import_reference_variables
    variables = {}
    for ws in load_keras_weights(path).values():
        for wname, arr in ws.items():
            variables[wname] = arr
    import_reference_variables(net, variables)
