"""tf.train.Saver's tensor-bundle checkpoint format written and read without TensorFlow (scenedepthestimation_b200/tf_checkpoint.py).
The reference writes such checkpoints in train.py:116,152-156 and restores them in process_functional.py:25-33; mc_cnn_brunch.py:51-66
holds the .npy side. Parity unpinned (no file made by TensorFlow exists in this image): round trips, the format's own checksums
and the documented layout of the files."""
import struct

import numpy as np
import pytest

from scenedepthestimation_b200 import synthetic as syn
from scenedepthestimation_b200 import tf_checkpoint as tfc


def test_tf_checkpoint_round_trip(tmp_path):
    w = syn.glorot_weights()
    prefix = str(tmp_path / "ckpt" / "model.ckpt-14")
    extra = {"conv1/weights/Momentum": np.zeros((3, 3, 1, 64), np.float32), "global_step": np.array(1234, np.int64),
             "lr": np.array([0.5, 0.25], np.float64)}
    tensors = {k[:-2]: v for k, v in w.items()}
    tensors.update(extra)
    tfc.write_checkpoint(prefix, tensors)
    assert tfc.is_checkpoint_prefix(prefix) and not tfc.is_checkpoint_prefix(str(tmp_path / "nope"))
    # layout: 48-byte footer ending in LevelDB's table magic; the data file is the tensors' bytes back to back in key order
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xDB4775248B80FB57
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(data) == sum(np.asarray(v).nbytes for v in tensors.values())
    back = tfc.read_checkpoint(prefix)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v), k
    # the reference's dict layout: ':0' names, conv variables only (optimizer slots and counters dropped)
    got = tfc.checkpoint_to_weights(prefix)
    assert sorted(got) == sorted(w) and all(np.array_equal(got[k], w[k]) for k in w)
    tfc.weights_to_checkpoint(w, str(tmp_path / "again"))
    assert all(np.array_equal(tfc.checkpoint_to_weights(str(tmp_path / "again"))[k], w[k]) for k in w)
    # CLI: checkpoint -> .npy -> checkpoint
    tfc.main([prefix, str(tmp_path / "w.npy")])
    npy = np.load(tmp_path / "w.npy", allow_pickle=True).item()
    assert all(np.array_equal(npy[k], w[k]) for k in w)
    tfc.main([str(tmp_path / "w.npy"), str(tmp_path / "c2")])
    assert sorted(tfc.read_checkpoint(str(tmp_path / "c2"))) == sorted(k[:-2] for k in w)
    # corruption is caught by the format's checksums
    bad = bytearray(data)
    bad[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tfc.read_checkpoint(prefix)
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    badi = bytearray(idx)
    badi[10] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(badi))
    with pytest.raises(ValueError):
        tfc.read_checkpoint(prefix)


def test_tf_checkpoint_primitives():
    """CRC-32C known answers (RFC 3720 B.4), LevelDB's mask, Snappy decoding (literal, short and long copies, an overlapping
    copy), prefix-compressed blocks with restart points, a Snappy-compressed block inside a table."""
    assert tfc.crc32c(b"") == 0 and tfc.crc32c(b"123456789") == 0xE3069283
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA and tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    c = tfc.crc32c(b"foo")
    assert tfc.masked_crc(b"foo") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    # "abcd" x 5 then 70 x "z": literal 'abcd', overlapping copies of 8 (1-byte-offset form) and 8 (4-byte-offset form) at offset 4,
    # literal 'z', overlapping copies of 64 + 5 at offset 1 (2-byte-offset form)
    stream = (bytes([90]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([((8 - 1) << 2) | 3, 4, 0, 0, 0])
              + bytes([0 << 2]) + b"z" + bytes([((64 - 1) << 2) | 2, 1, 0]) + bytes([((5 - 1) << 2) | 2, 1, 0]))
    assert tfc._snappy_decompress(stream) == b"abcd" * 5 + b"z" * 70
    items = sorted((f"conv{i}/weights".encode(), bytes([i]) * (i + 1)) for i in range(1, 40))
    block = tfc._build_block(items, restart_interval=4)
    assert list(tfc._block_entries(block)) == items
    table = tfc._table_write(items)
    assert list(tfc._table_read(table)) == items
    # the same table with its data block stored Snappy-compressed (as TensorFlow's writer does when it pays): one literal element
    raw = tfc._build_block(items)
    n = len(raw) - 1
    comp = tfc._put_varint(len(raw)) + bytes([61 << 2]) + n.to_bytes(2, "little") + raw
    out = bytearray(comp + b"\x01" + struct.pack("<I", tfc.masked_crc(comp + b"\x01")))
    data_handle = tfc._put_varint(0) + tfc._put_varint(len(comp))
    meta = tfc._build_block([])
    moff = len(out)
    out += meta + b"\x00" + struct.pack("<I", tfc.masked_crc(meta + b"\x00"))
    index = tfc._build_block([(items[-1][0] + b"\x00", data_handle)])
    ioff = len(out)
    out += index + b"\x00" + struct.pack("<I", tfc.masked_crc(index + b"\x00"))
    footer = tfc._put_varint(moff) + tfc._put_varint(len(meta)) + tfc._put_varint(ioff) + tfc._put_varint(len(index))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tfc.MAGIC)
    assert list(tfc._table_read(bytes(out))) == items
