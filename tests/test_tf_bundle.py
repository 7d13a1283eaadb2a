"""TensorFlow V2 checkpoint ("tensor bundle") reader / writer (differential_equations_resnet_b200/tf_bundle.py): the
file format `tf.train.Saver` writes in the reference trainer (training/training.py:848-872).  No TensorFlow-written file
exists in this image; pinned here: the CRC-32C check value and masking, the LevelDB table footer / block structure,
protobuf encodings of BundleHeaderProto / BundleEntryProto, corruption detection and the write -> read round trip."""
import os
import struct

import numpy as np
import pytest

from differential_equations_resnet_b200 import tf_bundle as tb


def test_crc32c_known_answers():
    assert tb.crc32c(b"123456789") == 0xE3069283            # the standard CRC-32C check value
    assert tb.crc32c(b"") == 0
    assert tb.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43      # RFC 3720 B.4: 32 bytes of ones
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E        # RFC 3720 B.4: incrementing bytes
    rng = np.random.default_rng(0)
    blob = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
    slow = 0xFFFFFFFF
    for b in blob:                                          # bitwise reference
        slow ^= b
        for _ in range(8):
            slow = (slow >> 1) ^ (0x82F63B78 if slow & 1 else 0)
    assert tb.crc32c(blob) == slow ^ 0xFFFFFFFF             # slicing-by-8 path == bitwise definition
    # leveldb's documented masking example: unmask(mask(c)) == c, mask changes the value, double masking differs
    c = tb.crc32c(b"foo")
    assert tb.unmask_crc(tb.mask_crc(c)) == c and tb.mask_crc(c) != c and tb.mask_crc(tb.mask_crc(c)) != c
    assert tb.mask_crc(0) == 0xA282EAD8


def test_table_roundtrip_and_layout(tmp_path):
    items = [(b"", b"header")] + [(("layer%03d/var_%d" % (i // 7, i % 7)).encode(), os.urandom(5 + i % 40)) for i in range(700)]
    items.sort()
    path = str(tmp_path / "t.index")
    tb.write_table(path, items, block_size=512)             # many blocks -> multi-entry index block
    data = open(path, "rb").read()
    assert struct.unpack("<Q", data[-8:])[0] == 0xDB4775248B80FB57 and len(data) > 48
    assert tb.read_table(path) == items
    with pytest.raises(ValueError):
        tb.write_table(str(tmp_path / "bad.index"), [(b"b", b"1"), (b"a", b"2")])
    # a flipped byte inside a data block is caught by the block CRC
    bad = bytearray(data); bad[10] ^= 0x40
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="CRC"):
        tb.read_table(path)
    open(path, "wb").write(data[:-1] + b"\x00")
    with pytest.raises(ValueError, match="magic"):
        tb.read_table(path)


def test_proto_encodings():
    e = tb.encode_entry(tb.DT_FLOAT, (3, 3, 15), 0, 4096, 540, 0xDEADBEEF)
    # field 1 varint 1 | field 2 len-delimited shape {dim{size:3} x2, dim{size:15}} | offset | size | fixed32 crc
    assert e[:2] == b"\x08\x01" and e[2] == 0x12 and e.endswith(b"\x35" + struct.pack("<I", 0xDEADBEEF))
    d = tb.decode_entry(e)
    assert d["dtype"] == 1 and d["shape"] == (3, 3, 15) and d["offset"] == 4096 and d["size"] == 540 and d["crc32c"] == 0xDEADBEEF
    assert tb.decode_entry(tb.encode_entry(tb.DT_INT64, (), 0, 0, 8, 1))["shape"] == ()
    h = tb.decode_header(tb.encode_header(1))
    assert h == {"num_shards": 1, "endianness": 0, "producer": 1}
    assert tb.encode_header(1) == b"\x08\x01\x1a\x02\x08\x01"


def test_snappy_reader():
    # literal "abcd", then a copy of 8 bytes at offset 4 (tag type 1), then literal "!"
    comp = bytes([13]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([0 << 2]) + b"!"
    assert tb._snappy_uncompress(comp) == b"abcdabcdabcd!"


def test_bundle_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        "conv1/kernel": rng.standard_normal((3, 3, 3, 16)).astype(np.float32),
        "res2_0_branch2/a": rng.standard_normal((1, 1, 1, 16)).astype(np.float32),
        "res2_0_branch2/input_kernels_for_output_kernel_0": rng.standard_normal((3, 3, 15)).astype(np.float32),
        "res2_0_branch2/a/Adam": np.zeros((1, 1, 1, 16), np.float32),
        "beta1_power": np.float32(0.9), "global_step": np.int64(1563),
        "big": rng.standard_normal(200_000).astype(np.float32),
    }
    prefix = str(tmp_path / "model" / "variables")
    tb.write_bundle(prefix, tensors)
    assert sorted(os.listdir(str(tmp_path / "model"))) == ["checkpoint", "variables.data-00000-of-00001", "variables.index"]
    assert 'model_checkpoint_path: "variables"' in open(str(tmp_path / "model" / "checkpoint")).read()
    back = tb.read_bundle(prefix)
    assert list(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v)
    # data shard = the raw little-endian bytes in name order
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    assert raw[:4] == np.float32(0.9).tobytes() and len(raw) == sum(np.asarray(v).nbytes for v in tensors.values())
    blob = bytearray(raw); blob[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(blob))
    with pytest.raises(ValueError, match="CRC"):
        tb.read_bundle(prefix)


@pytest.mark.gpu
def test_eulernet_saver_checkpoint_resume(tmp_path):
    """EulerNet -> Saver-format checkpoint under the reference's variable names (+ Adam slots, beta powers, global_step)
    -> a fresh net resumes bit-identically."""
    import torch
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(2, 2, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
    gen = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8).cuda()
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float().cuda()
    a = EulerNet(NetSpec(**kw), seed=3)
    a.train_step(img, lab); a.train_step(img, lab)
    tb.save_tf_checkpoint(a, str(tmp_path / "ck"))
    t = tb.read_bundle(str(tmp_path / "ck" / "variables"))
    assert "res2_0_branch2/a" in t and "res2_0_branch2/a/Adam" in t and "res2_1_branch2/input_kernels_for_output_kernel_14/Adam_1" in t
    assert t["res2_0_branch2/a"].shape == (1, 1, 1, 16) and int(t["global_step"]) == 2
    assert abs(float(t["beta1_power"]) - 0.9 ** 3) < 1e-7
    b = EulerNet(NetSpec(**kw), seed=99)
    tb.load_tf_checkpoint(b, str(tmp_path / "ck" / "variables"))
    assert torch.equal(a.theta, b.theta) and torch.equal(a.adam_m, b.adam_m) and torch.equal(a.adam_v, b.adam_v)
    assert float(a.train_step(img, lab)) == float(b.train_step(img, lab)) and torch.equal(a.theta, b.theta)
