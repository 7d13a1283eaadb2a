"""Reader / writer of TensorFlow's V2 checkpoint format ("tensor bundle"), the files `tf.train.Saver.save` writes in
the reference trainer (`training/training.py:848-865` -> `<dir>/variables.index`, `<dir>/variables.data-00000-of-00001`,
`<dir>/checkpoint`) and `Saver.restore` reads (`training/training.py:867-872`).  Pure Python + NumPy: TensorFlow is not
installable in this image, so the format is implemented from its public specification:

  * `<prefix>.index`  is a LevelDB-style sorted string table (tensorflow/core/lib/io/table*.cc, format.cc): data blocks
    of prefix-compressed (key, value) entries with restart points, every block followed by a 5-byte trailer
    (compression type, masked CRC-32C), a meta-index block, an index block of BlockHandles and a 48-byte footer ending
    in the magic 0xdb4775248b80fb57.  Key "" holds a `BundleHeaderProto` (num_shards, endianness, version), every other
    key is a tensor name and its value a `BundleEntryProto` (dtype, shape, shard_id, offset, size, crc32c)
    (tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/util/tensor_bundle/tensor_bundle.cc).
  * `<prefix>.data-SSSSS-of-NNNNN`  are the raw little-endian tensor bytes at the recorded offsets.

The reader accepts uncompressed and Snappy-compressed blocks (TensorFlow's own writer uses none for the index) and any
number of data shards; the writer produces one shard, uncompressed blocks, restart interval 16, like BundleWriter.
No file written by TensorFlow itself is available here to cross-check against: tests pin the CRC-32C check value, the
masking formula, the footer magic, the varint / proto encodings and the write -> read round trip.

Variable names follow the reference graph: `<layer>/<variable>` (`res2_0_branch2/a`, `conv1/kernel`, ...) and, for the
optimiser, TF's slot naming `<variable>/Adam`, `<variable>/Adam_1`, `beta1_power`, `beta2_power`, `global_step`."""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
MASK_DELTA = 0xA282EAD8
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64, DT_BOOL, DT_HALF, DT_BFLOAT16 = 1, 2, 3, 4, 9, 10, 19, 14
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_UINT8: np.dtype("u1"),
           DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("?"), DT_HALF: np.dtype("<f2")}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("uint8"): DT_UINT8,
          np.dtype("int64"): DT_INT64, np.dtype("bool"): DT_BOOL, np.dtype("float16"): DT_HALF}

# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_POLY = 0x82F63B78
_T = np.zeros((8, 256), dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _T[0, _i] = _c
for _k in range(1, 8):
    _T[_k] = (_T[_k - 1] >> 8) ^ _T[0][_T[_k - 1] & 0xFF]
_T0 = [int(v) for v in _T[0]]


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C of `data` (check value crc32c(b'123456789') == 0xE3069283); slicing-by-8 in NumPy for large buffers."""
    crc ^= 0xFFFFFFFF
    mv = memoryview(data)
    n = len(mv)
    i = 0
    if n >= 64:
        # process 8-byte words with the slicing-by-8 tables, vectorised over independent lanes is not possible (serial
        # dependence), so walk words in Python over NumPy-extracted bytes: ~10 MB/s, fine for checkpoints of a few MB
        a = np.frombuffer(mv[: n - n % 8], dtype=np.uint8).reshape(-1, 8)
        t = [[int(v) for v in _T[k]] for k in range(8)]
        for row in a.tolist():
            c = crc ^ (row[0] | (row[1] << 8) | (row[2] << 16) | (row[3] << 24))
            crc = (t[7][c & 0xFF] ^ t[6][(c >> 8) & 0xFF] ^ t[5][(c >> 16) & 0xFF] ^ t[4][c >> 24] ^
                   t[3][row[4]] ^ t[2][row[5]] ^ t[1][row[6]] ^ t[0][row[7]])
        i = n - n % 8
    for b in mv[i:].tobytes():
        crc = _T0[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """leveldb / TensorFlow CRC masking: rotate right by 15 bits and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf
def _put_varint(out: bytearray, v: int):
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7


def _proto_fields(buf):
    """Yield (field number, wire type, value) of a serialized protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _encode_shape(shape):
    """TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }"""
    out = bytearray()
    for d in shape:
        dim = bytearray()
        dim.append(0x08); _put_varint(dim, int(d))
        out.append(0x12); _put_varint(out, len(dim)); out += dim
    return bytes(out)


def _decode_shape(buf):
    shape = []
    for f, wt, v in _proto_fields(buf):
        if f == 2 and wt == 2:
            size = 0
            for f2, wt2, v2 in _proto_fields(v):
                if f2 == 1 and wt2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            shape.append(size)
        elif f == 3 and wt == 0 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(shape)


def encode_entry(dtype, shape, shard_id, offset, size, crc):
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32)."""
    out = bytearray()
    out.append(0x08); _put_varint(out, dtype)
    sh = _encode_shape(shape)
    out.append(0x12); _put_varint(out, len(sh)); out += sh
    if shard_id:
        out.append(0x18); _put_varint(out, shard_id)
    if offset:
        out.append(0x20); _put_varint(out, offset)
    out.append(0x28); _put_varint(out, size)
    out.append(0x35); out += struct.pack("<I", crc)
    return bytes(out)


def decode_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, wt, v in _proto_fields(buf):
        if f == 1: e["dtype"] = v
        elif f == 2: e["shape"] = _decode_shape(v)
        elif f == 3: e["shard_id"] = v
        elif f == 4: e["offset"] = v
        elif f == 5: e["size"] = v
        elif f == 6: e["crc32c"] = v
        elif f == 7: e["sliced"] = True
    return e


def encode_header(num_shards=1):
    """BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0, omitted), version = 3 { producer = 1 }."""
    out = bytearray()
    out.append(0x08); _put_varint(out, num_shards)
    ver = bytes([0x08, 0x01])
    out.append(0x1A); _put_varint(out, len(ver)); out += ver
    return bytes(out)


def decode_header(buf):
    h = {"num_shards": 1, "endianness": 0, "producer": 0}
    for f, wt, v in _proto_fields(buf):
        if f == 1: h["num_shards"] = v
        elif f == 2: h["endianness"] = v
        elif f == 3:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1: h["producer"] = v2
    return h


# ---------------------------------------------------------------------------------------------- Snappy (reader only)
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        t = tag & 3
        if t == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if t == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif t == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt Snappy block")
    return bytes(out)


# ---------------------------------------------------------------------------------------------- table (index file)
def _block_bytes(entries, restart_interval=16):
    """One table block from sorted (key, value) byte pairs: prefix-compressed entries + restart array."""
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        _put_varint(out, shared); _put_varint(out, len(k) - shared); _put_varint(out, len(v))
        out += k[shared:]; out += v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _parse_block(block):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared]); pos += unshared
        out.append((key, bytes(block[pos:pos + vlen]))); pos += vlen
    return out


def _read_block(data, offset, size, verify=True):
    raw, ctype = data[offset:offset + size], data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(bytes(data[offset:offset + size + 1])):
        raise ValueError("index block at %d fails its CRC-32C check" % offset)
    if ctype == 0:
        return bytes(raw)
    if ctype == 1:
        return _snappy_uncompress(bytes(raw))
    raise ValueError("unknown block compression type %d" % ctype)


def write_table(path, items, block_size=4096):
    """Sorted string table of (key bytes, value bytes) pairs (keys strictly increasing)."""
    blob, index, pending, psize = bytearray(), [], [], 0

    def flush():
        nonlocal pending, psize
        if not pending:
            return
        b = _block_bytes(pending)
        off = len(blob)
        blob.extend(b); blob.append(0)
        blob.extend(struct.pack("<I", mask_crc(crc32c(b + b"\x00"))))
        h = bytearray(); _put_varint(h, off); _put_varint(h, len(b))
        index.append((pending[-1][0], bytes(h)))        # separator = the block's last key (a valid choice: >= every key in it)
        pending, psize = [], 0

    last = None
    for k, v in items:
        if last is not None and k <= last:
            raise ValueError("table keys must be strictly increasing")
        last = k
        pending.append((k, v)); psize += len(k) + len(v) + 3
        if psize >= block_size:
            flush()
    flush()

    def emit(b):
        off = len(blob)
        blob.extend(b); blob.append(0)
        blob.extend(struct.pack("<I", mask_crc(crc32c(b + b"\x00"))))
        h = bytearray(); _put_varint(h, off); _put_varint(h, len(b))
        return bytes(h)
    meta_h = emit(_block_bytes([]))
    index_h = emit(_block_bytes(index, restart_interval=1))
    footer = bytearray(meta_h + index_h)
    footer.extend(b"\x00" * (40 - len(footer)))
    footer.extend(struct.pack("<Q", TABLE_MAGIC))
    blob.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(blob))


def read_table(path, verify=True):
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow / LevelDB table (bad magic)" % path)
    foot = data[-48:]
    pos = 0
    _, pos = _get_varint(foot, pos); _, pos = _get_varint(foot, pos)          # meta-index handle
    ioff, pos = _get_varint(foot, pos); isz, pos = _get_varint(foot, pos)     # index handle
    out = []
    for _, handle in _parse_block(_read_block(data, ioff, isz, verify)):
        off, p = _get_varint(handle, 0)
        sz, p = _get_varint(handle, p)
        out += _parse_block(_read_block(data, off, sz, verify))
    return out


# ---------------------------------------------------------------------------------------------- bundle API
def _shard_name(prefix, i, n):
    return "%s.data-%05d-of-%05d" % (prefix, i, n)


def write_bundle(prefix, tensors):
    """tensors: mapping name -> ndarray.  Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` (names sorted, as
    BundleWriter does) and the `checkpoint` state file next to them."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", encode_header(1))]
    offset = 0
    with open(_shard_name(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])
            if a.ndim and not a.flags.c_contiguous:      # (ascontiguousarray would turn a scalar into shape (1,))
                a = np.ascontiguousarray(a)
            if a.dtype not in _DT_OF:
                raise ValueError("unsupported dtype %s for %s" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"), encode_entry(_DT_OF[a.dtype], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index", items)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def read_bundle(prefix, verify=True):
    """-> OrderedDict name -> ndarray (sorted by name, as stored).  Every tensor's masked CRC-32C is checked."""
    items = read_table(prefix + ".index", verify)
    if not items or items[0][0] != b"":
        raise ValueError("bundle index has no header entry")
    header = decode_header(items[0][1])
    if header["endianness"] != 0:
        raise ValueError("big-endian bundles are not supported")
    n = header["num_shards"]
    shards = {}
    out = OrderedDict()
    for key, val in items[1:]:
        e = decode_entry(val)
        if e["sliced"]:
            raise ValueError("partitioned variable %s (tensor slices) is not supported" % key.decode())
        if e["dtype"] not in _DTYPES:
            raise ValueError("tensor %s has unsupported dtype enum %d" % (key.decode(), e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open(_shard_name(prefix, sid, n), "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError("tensor %s runs past the end of its data shard" % key.decode())
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("tensor %s fails its CRC-32C check" % key.decode())
        out[key.decode("utf-8")] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ---------------------------------------------------------------------------------------------- EulerNet <-> Saver checkpoint
def save_tf_checkpoint(net, directory, name="variables", with_optimizer=True):
    """`Training.save(saver='train_saver')` (`training/training.py:848-865`): `<directory>/<name>.index|.data-*` with the
    reference graph's variable names; with_optimizer adds TF's Adam slots (`<var>/Adam` = m, `<var>/Adam_1` = v),
    `beta1_power`, `beta2_power` and `global_step`."""
    from . import checkpoint as ck
    tensors = OrderedDict(ck.export_reference_variables(net))
    if with_optimizer:
        for slot, buf in (("Adam", net.adam_m), ("Adam_1", net.adam_v)):
            saved = net.theta
            try:
                net.theta = buf
                for k, v in ck.export_reference_variables(net).items():
                    tensors[k + "/" + slot] = v
            finally:
                net.theta = saved
        t = int(net.step_counter) - 1                       # updates applied so far
        tensors["beta1_power"] = np.float32(0.9 ** (t + 1))  # TF keeps beta^(t+1): the power used by the NEXT update
        tensors["beta2_power"] = np.float32(0.999 ** (t + 1))
        tensors["global_step"] = np.int64(t)
    write_bundle(os.path.join(directory, name), tensors)


def load_tf_checkpoint(net, prefix, restore_optimizer=True):
    """`Training.load_variables` (`training/training.py:867-872`) for a checkpoint written by `tf.train.Saver` (or by
    save_tf_checkpoint): parameters by the reference's variable names; Adam slots and global_step when present."""
    import torch
    from . import checkpoint as ck
    tensors = read_bundle(prefix)
    ck.import_reference_variables(net, tensors)
    if restore_optimizer and "global_step" in tensors:
        saved = net.theta
        for slot, buf in (("Adam", net.adam_m), ("Adam_1", net.adam_v)):
            sl = {k[:-len(slot) - 1]: v for k, v in tensors.items() if k.endswith("/" + slot)}
            if not sl:
                continue
            try:
                net.theta = buf
                ck.import_reference_variables(net, sl)
            finally:
                net.theta = saved
        with torch.no_grad():
            net.step_counter.fill_(int(tensors["global_step"]) + 1)
    return tensors
