"""Keras `.h5` weight files without h5py (`differential_equations_resnet_b200/keras_h5.py`; the reference's use:
experiments_antisymmetric_resnet_v6.ipynb `model.save_weights(...h5)` / `model.load_weights(...)`).

The reader is pinned on tests/golden/hdf5_library_written.mat: a MATLAB 7.3 file = an HDF5 file written by the HDF5 C
library itself (512-byte user block, version-0 superblock, TREE / SNOD / HEAP groups, an attribute), taken from scipy's
test data (scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD-3); its one dataset is MATLAB's
`testdouble = 0:pi/4:2*pi`.  The writer is checked through the reader and structurally (field by field)."""
import os
import struct

import numpy as np
import pytest

from differential_equations_resnet_b200 import keras_h5 as kh

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hdf5_library_written.mat")


def test_reader_on_a_library_written_file():
    root = kh.read_h5(GOLDEN)
    assert list(root.children) == ["testdouble"]
    ds = root["testdouble"]
    assert isinstance(ds, kh.H5Dataset) and ds.array.dtype == np.float64 and ds.array.shape == (9, 1)
    assert np.array_equal(ds.array.reshape(-1), np.arange(9) * (np.pi / 4))          # bit-exact 0:pi/4:2*pi
    assert bytes(ds.attrs["MATLAB_class"]) == b"double"


def test_reader_rejects_garbage_and_truncation():
    with pytest.raises(kh.H5FormatError):
        kh.read_h5(b"not an hdf5 file" * 100)
    data = open(GOLDEN, "rb").read()
    with pytest.raises(kh.H5FormatError):
        kh.read_h5(data[:2000])
    bad = bytearray(data)
    bad[512 + 8] = 2                       # superblock version 2 (libver='latest'): refused, not misread
    with pytest.raises(kh.H5FormatError, match="superblock version"):
        kh.read_h5(bytes(bad))


def test_writer_superblock_fields_match_the_library_written_file():
    """Same superblock layout as the library's: signature, versions, 8-byte offsets / lengths, root symbol-table entry
    with cached B-tree / heap addresses (cache type 1), EOF address = file size."""
    ours = kh.write_h5(None, {"x": np.arange(4, dtype=np.float32)})
    ref = open(GOLDEN, "rb").read()[512:]
    assert ours[:8] == ref[:8] == kh.SIGNATURE
    assert ours[8:13] == ref[8:13] and ours[13:16] == ref[13:16] == bytes([8, 8, 0])
    assert struct.unpack_from("<Q", ours, 40)[0] == len(ours)                              # EOF address
    assert struct.unpack_from("<Q", ref, 40)[0] == os.path.getsize(GOLDEN)       # (this 2008 library wrote it absolute)
    assert struct.unpack_from("<I", ours, 72)[0] == struct.unpack_from("<I", ref, 72)[0] == 1   # root entry cache type
    bt, hp = struct.unpack_from("<QQ", ours, 80)
    assert ours[bt:bt + 4] == b"TREE" and ours[hp:hp + 4] == b"HEAP"
    oh, = struct.unpack_from("<Q", ours, 64)
    assert ours[oh] == 1 and oh % 8 == 0                                                   # version-1 object header


def test_float32_datatype_message_is_the_ieee_le_encoding():
    """The 20 header + property bytes the HDF5 library writes for H5T_IEEE_F32LE / F64LE (format spec, class 1)."""
    assert kh._dtype_message(np.float32) == bytes.fromhex("11201f0004000000" "0000200017080017" "7f000000")
    assert kh._dtype_message(np.float64) == bytes.fromhex("11203f0008000000" "00004000340b0034" "ff030000")
    # and the library-written file holds exactly the F64LE message for its dataset
    assert kh._dtype_message(np.float64) in open(GOLDEN, "rb").read()


def test_roundtrip_tree_attrs_and_many_links(tmp_path):
    rng = np.random.default_rng(0)
    big = {"v%03d" % i: rng.standard_normal((3, 3, 1 + i % 5)).astype(np.float32) for i in range(300)}   # > one symbol node
    tree = {"g": ({"a:0": np.float32(2.5) * np.ones((1, 1, 1, 16), np.float32), "sub": {"k": np.arange(6, dtype=np.int32).reshape(2, 3)}},
                  {"weight_names": np.array([b"g/a:0", b"g/sub/k"], dtype="S")}),
            "big": big, "empty": np.zeros((0,), np.float32), "scalar": np.float64(3.25)}
    p = str(tmp_path / "t.h5")
    kh.write_h5(p, tree, {"backend": "tensorflow", "n": np.int64(7)})
    r = kh.read_h5(p)
    assert bytes(r.attrs["backend"]) == b"tensorflow" and int(r.attrs["n"]) == 7
    assert [bytes(x) for x in r["g"].attrs["weight_names"]] == [b"g/a:0", b"g/sub/k"]
    assert np.array_equal(r["g/a:0"].array, tree["g"][0]["a:0"]) and r["g/a:0"].array.dtype == np.float32
    assert np.array_equal(r["g/sub/k"].array, np.arange(6).reshape(2, 3))
    assert list(r["big"].children) == sorted(big)                      # symbol nodes are in strcmp order
    for k, v in big.items():
        assert np.array_equal(r["big"][k].array, v)
    assert r["empty"].array.shape == (0,) and float(r["scalar"].array) == 3.25
    assert "nope" not in r and "g/sub/k" in r


def test_keras_layout_roundtrip_with_reference_variable_names(tmp_path):
    from differential_equations_resnet_b200.checkpoint import variable_shapes_3by3
    rng = np.random.default_rng(1)
    layers = {"conv1": {"conv1/kernel:0": rng.standard_normal((3, 3, 3, 16)).astype(np.float32), "conv1/bias:0": np.zeros(16, np.float32)},
              "activation": {}}
    for C, name in ((16, "res2_0_branch2"), (64, "res4_1_branch2")):
        layers[name] = {"%s/%s:0" % (name, v): rng.standard_normal(shp).astype(np.float32) for v, shp in variable_shapes_3by3(C)}
    p = str(tmp_path / "w.h5")
    kh.save_keras_weights(p, layers)
    root = kh.read_h5(p)
    assert [bytes(x).decode() for x in root.attrs["layer_names"]] == list(layers)          # model.layers order, not sorted
    # h5py stores Python bytes as VARIABLE-length strings (global heap): what Keras' backend / keras_version are
    assert root.attrs["backend"].item() == b"tensorflow" and root.attrs["keras_version"].item() == b"2.2.4-tf"
    assert open(p, "rb").read().count(b"GCOL") == 2
    assert isinstance(root["res2_0_branch2/res2_0_branch2/a:0"], kh.H5Dataset)            # '/' in the weight name nests a group
    back = kh.load_keras_weights(p)
    assert list(back) == list(layers) and back["activation"] == {}
    for l, ws in layers.items():
        assert list(back[l]) == list(ws)                                                     # creation order kept by weight_names
        for k, v in ws.items():
            assert back[l][k].dtype == np.float32 and np.array_equal(back[l][k], v)
    assert len(back["res4_1_branch2"]) == 64 + 4


def test_keras_chunked_name_attributes_and_model_weights_group(tmp_path):
    """Keras splits name lists above 64 KiB into name0, name1, ...; `model.save` nests everything under 'model_weights'."""
    w = {"d/kernel:0": np.ones((2, 2), np.float32)}
    tree = {"model_weights": ({"d": ({"d": {"kernel:0": w["d/kernel:0"]}}, {"weight_names0": np.array([b"d/kernel:0"], dtype="S")})},
                              {"layer_names0": np.array([b"d"], dtype="S"), "layer_names1": np.zeros((0,), "S1")})}
    p = str(tmp_path / "m.h5")
    kh.write_h5(p, tree)
    back = kh.load_keras_weights(p)
    assert list(back) == ["d"] and np.array_equal(back["d"]["d/kernel:0"], w["d/kernel:0"])
    with pytest.raises(kh.H5FormatError, match="weight_names|layer_names"):
        kh.write_h5(p, {"x": np.zeros(3, np.float32)})
        kh.load_keras_weights(p)
    with pytest.raises(kh.H5FormatError, match="64 KiB"):
        kh.write_h5(None, {}, {"layer_names": np.array([b"x" * 100] * 700, dtype="S")})


def test_general_layer_variable_names_follow_the_tf1_uniquifier():
    from differential_equations_resnet_b200.layers.tfkeras_layer_Conv2DAntisymmetric import Conv2DAntisymmetric, diag_slots
    lay = Conv2DAntisymmetric.__new__(Conv2DAntisymmetric)
    lay.num_channels, lay.kernel_size, lay.antisymmetric, lay.use_bias = 3, 3, True, True
    names = lay._variable_names()
    nd = len(diag_slots(3, True))
    assert names[:nd] == ["centro_sym_0_0", "centro_sym_0_1", "centro_sym_0_2", "centro_sym_1_2"]
    assert names[0] == "centro_sym_0_0" and names[nd] == "input_kernels_for_output_kernel_0"
    assert names[nd + 1] == "centro_sym_0_0_1" and names[-1] == "bias" and names[-2] == "centro_sym_1_2_2"
    assert len(names) == len(lay._variable_shapes()) == 3 * nd + 2 + 1 and len(set(names)) == len(names)


@pytest.mark.gpu
def test_eulernet_and_model_h5_weights(tmp_path):
    import torch
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(2, 1, 1), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
    a, b = EulerNet(NetSpec(**kw), seed=3), EulerNet(NetSpec(**kw), seed=99)
    p = str(tmp_path / "weights.h5")
    a.save_weights(p)
    saved = kh.load_keras_weights(p)
    assert saved["res2_0_branch2"]["res2_0_branch2/a:0"].shape == (1, 1, 1, 16) and len(saved["res2_0_branch2"]) == 20
    assert saved["fc"]["fc/kernel:0"].shape == (64, 10)
    assert not torch.equal(a.theta, b.theta)
    b.load_weights(p)
    assert torch.equal(a.theta, b.theta)

    from differential_equations_resnet_b200.models import get_single_block_resnet_build_function
    kw = dict(kernel_type='antisymmetric', precision='strict', h=0.5, gamma=-0.1, num_stages=4, blocks_per_stage=[2, 1, 1],
              filters_per_block=[16, 32, 64], strides=[(1, 1), (2, 2), (2, 2)], num_classes=10, use_batch_norm=True)
    x = torch.rand(2, 16, 16, 3, device="cuda")
    m1 = get_single_block_resnet_build_function(seed=1, **kw)(x)
    m2 = get_single_block_resnet_build_function(seed=2, **kw)(x)
    y1, y2 = m1(x, training=False), m2(x, training=False)
    assert not torch.equal(y1, y2)
    q = str(tmp_path / "model.h5")
    m1.save_weights(q)
    saved = kh.load_keras_weights(q)
    assert list(saved)[:3] == ["conv1", "bn_conv1", "res2_0_branch2"] or list(saved)[0] == "conv1"
    assert any(k.endswith("moving_variance:0") for ws in saved.values() for k in ws)
    m2.load_weights(q)
    assert torch.equal(m1(x, training=False), m2(x, training=False))


def test_reader_survives_random_corruption():
    """Byte flips in a valid file end in a parse or in H5FormatError -- no other exception type, no hang."""
    good = kh.save_keras_weights(None, {"l%d" % i: {"l%d/k:0" % i: np.ones((3, 3, 4), np.float32)} for i in range(4)})
    gold = open(GOLDEN, "rb").read()
    rng = np.random.default_rng(0)
    for base in (good, gold):
        for _ in range(400):
            b = bytearray(base)
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            try:
                kh.read_h5(bytes(b))
            except kh.H5FormatError:
                pass


def test_writer_messages_equal_the_library_written_ones():
    """The same dataset (9x1 float64, attribute MATLAB_class = 'double') written by this module: its dataspace and datatype
    messages are byte-identical to the HDF5 library's, the attribute message differs only in the string padding bits
    (MATLAB wrote a null-terminated string, numpy 'S' arrays are null-padded as h5py writes them); B-tree / heap / symbol
    node layout follows the library's (key 0 = offset 0 = the empty string, key 1 = offset of the last name)."""
    ref = kh._Reader(open(GOLDEN, "rb").read())
    ours = kh._Reader(kh.write_h5(None, {"testdouble": (np.arange(9.0).reshape(9, 1), {"MATLAB_class": b"double"})}))

    def dataset_messages(r):
        root = r._off(r.root_entry + r.O)
        stab = [b for t, _, b in r.messages(root) if t == 0x11][0]
        (name, addr), = r._symbol_entries(r._off_b(stab, 0), r._off_b(stab, r.O))
        assert name == "testdouble"
        bt = r._at(r._off_b(stab, 0))
        keys = (r._len(bt + 24), r._len(bt + 24 + r.L + r.O))
        return {t: bytes(b) for t, _, b in r.messages(addr)}, keys

    (mref, kref), (mours, kours) = dataset_messages(ref), dataset_messages(ours)
    assert mours[0x01] == mref[0x01] and mours[0x03] == mref[0x03]
    a, b = bytearray(mours[0x0C]), bytearray(mref[0x0C])
    assert a[25] == 0x01 and b[25] == 0x00                 # string padding type: null-padded vs null-terminated
    a[25] = b[25]
    assert a == b
    assert kours == kref == (0, 8)


def test_reader_chunked_and_compact_layouts_hand_built():
    """Layouts Keras does not write but h5py users may (chunks without filters, compact): files assembled here from the
    format specification's structures (v1 B-tree of type 1 with (size, filter mask, offsets) keys; layout message v3
    classes 2 and 0) around the module's own object-header writer."""
    import struct
    w = kh._Writer()
    full = np.arange(5 * 6, dtype=np.float32).reshape(5, 6)
    cdims = (2, 4)
    chunks = []
    for i in range(0, 5, 2):
        for j in range(0, 6, 4):
            c = np.zeros(cdims, np.float32)
            blk = full[i:i + 2, j:j + 4]
            c[:blk.shape[0], :blk.shape[1]] = blk
            chunks.append(((i, j), w.alloc(c.tobytes())))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(chunks), kh.UNDEF, kh.UNDEF)
    for (i, j), addr in chunks:
        node += struct.pack("<II", 32, 0) + struct.pack("<QQQ", i, j, 0) + struct.pack("<Q", addr)
    node += struct.pack("<II", 0, 0) + struct.pack("<QQQ", 6, 8, 0)                      # final key
    bt = w.alloc(node)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", 2, 4, 4)
    msgs = [(0x01, kh._space_message(full.shape)), (0x03, kh._dtype_message(np.float32)), (0x08, layout)]
    chunked_addr = w.alloc(kh._object_header(msgs))
    small = np.array([1.5, -2.0, 3.25], np.float64)
    compact = struct.pack("<BBH", 3, 0, small.nbytes) + small.tobytes()
    compact_addr = w.alloc(kh._object_header([(0x01, kh._space_message(small.shape)), (0x03, kh._dtype_message(np.float64)), (0x08, compact)]))
    # a root group holding the two hand-built datasets: reuse the group writer with placeholder datasets, then patch the entries
    root = w.group({"a_chunked": ("d", np.zeros(1, np.float32), {}), "b_compact": ("d", np.zeros(1, np.float32), {})}, {})
    data = bytearray(w.finish(root))
    snod = data.index(b"SNOD")
    struct.pack_into("<Q", data, snod + 8 + 8, chunked_addr)
    struct.pack_into("<Q", data, snod + 8 + 40 + 8, compact_addr)
    r = kh.read_h5(bytes(data))
    assert list(r.children) == ["a_chunked", "b_compact"]
    assert np.array_equal(r["a_chunked"].array, full) and np.array_equal(r["b_compact"].array, small)
    # a filtered (compressed) chunk is refused, not misread
    bad = bytearray(data)
    p = bytes(bad).index(node[:8]) + 24
    struct.pack_into("<II", bad, p, 20, 1)
    with pytest.raises(kh.H5FormatError, match="filtered"):
        kh.read_h5(bytes(bad))
