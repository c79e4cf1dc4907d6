"""TensorFlow checkpoint (tf.train.Saver, "V2" / tensor-bundle format)  <->  the reference's .npy weight dict, without TensorFlow.

The reference trains with TF 1.x and keeps its weights in two forms: tf.train.Saver checkpoints (train.py:116,152-156, restored
by process_functional.compute_feature, :25-33) and the .npy dict written by Net.save_weights_dict (mc_cnn_brunch.py:61-66,
{'conv1/weights:0': HWIO, 'conv1/biases:0', ...}), which is what this repo's kernels load. TensorFlow is not installed here,
so the bundle format is read and written by hand:

  <prefix>.index                 an SSTable in LevelDB's table format: sorted key -> value records in prefix-compressed blocks
                                 (restart arrays, 1-byte compression tag + masked CRC-32C trailer), a metaindex block, an index
                                 block of block handles and a 48-byte footer ending in the magic 0xdb4775248b80fb57. Key ""
                                 holds a BundleHeaderProto, every other key a BundleEntryProto (dtype, shape, shard, offset,
                                 size, masked CRC-32C of the bytes).
  <prefix>.data-00000-of-00001   the tensors' raw little-endian bytes at the recorded offsets.

Only what the reference's checkpoints need is implemented: one shard, dense tensors without slices, dtypes float32 / float64 /
int32 / int64, blocks stored raw or Snappy-compressed (TF's table writer compresses a block when that saves 12.5 %). No file
written by TensorFlow itself could be tested in this image: PARITY UNPINNED, checked by round trips, by the format's own
checksums and against the published layout (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, leveldb
table_format.md).
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}   # DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64
_DTYPE_ENUM = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ CRC-32C (Castagnoli), masked
def _crc_table():
    t = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC = _crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    tab = _CRC
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf wire format
def _put_varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _get_varint(buf, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _proto_fields(buf: bytes):
    """Yield (field number, wire type, value) of one serialized message; length-delimited values as bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


def _field(num: int, wt: int, payload) -> bytes:
    head = _put_varint((num << 3) | wt)
    if wt == 0:
        return head + _put_varint(payload)
    if wt == 2:
        return head + _put_varint(len(payload)) + payload
    if wt == 5:
        return head + struct.pack("<I", payload)
    raise ValueError(wt)


def _parse_shape(buf: bytes):
    dims = []
    for f, _, v in _proto_fields(buf):
        if f == 2:   # TensorShapeProto.dim
            size = 0
            for g, _, w in _proto_fields(v):
                if g == 1:
                    size = w if w < (1 << 63) else w - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _proto_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif f == 7:
            e["slices"] += 1
    return e


def _serialize_entry(dtype_enum: int, shape, offset: int, size: int, crc: int) -> bytes:
    shp = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in shape)
    out = _field(1, 0, dtype_enum) + _field(2, 2, shp)
    if offset:
        out += _field(4, 0, offset)
    return out + _field(5, 0, size) + _field(6, 5, crc)


# ------------------------------------------------------------------------------------------------ Snappy (decompression only)
def _snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt Snappy stream")
        for _ in range(ln):   # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt Snappy stream (length)")
    return bytes(out)


# ------------------------------------------------------------------------------------------------ LevelDB table
def _read_block(data: bytes, offset: int, size: int) -> bytes:
    raw, kind = data[offset:offset + size], data[offset + size]
    crc = struct.unpack("<I", data[offset + size + 1:offset + size + 5])[0]
    if masked_crc(raw + bytes([kind])) != crc:
        raise ValueError("checkpoint index: block checksum mismatch")
    if kind == 0:
        return raw
    if kind == 1:
        return _snappy_decompress(raw)
    raise ValueError(f"checkpoint index: unknown block compression {kind}")


def _block_entries(block: bytes):
    num_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _build_block(items, restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    return bytes(out + struct.pack("<I", len(restarts)))


def _table_read(data: bytes):
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (table magic missing)")
    footer = data[-48:]
    _, p = _get_varint(footer, 0)          # metaindex handle
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)       # index handle
    isize, p = _get_varint(footer, p)
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, q = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, q)
        yield from _block_entries(_read_block(data, boff, bsize))


def _table_write(items) -> bytes:
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", masked_crc(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = emit(_build_block(items))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block([(items[-1][0] + b"\x00" if items else b"", data_handle)]))
    footer = meta_handle + index_handle
    return bytes(out) + footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)


# ------------------------------------------------------------------------------------------------ public API
def read_checkpoint(prefix: str) -> dict:
    """Every tensor of the checkpoint `prefix` (the path given to Saver.save / Saver.restore) -> {variable name: ndarray}."""
    with open(prefix + ".index", "rb") as f:
        index = f.read()
    entries, shards = {}, 1
    for key, value in _table_read(index):
        if key == b"":
            for fnum, _, v in _proto_fields(value):
                if fnum == 1:
                    shards = v
                elif fnum == 2 and v != 0:
                    raise ValueError("big-endian checkpoints are not supported")
            continue
        entries[key.decode()] = _parse_entry(value)
    out, files = {}, {}
    for name, e in entries.items():
        if e["slices"]:
            raise ValueError(f"{name}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in _DTYPES:
            continue   # strings etc. (not weights)
        sid = e["shard_id"]
        if sid not in files:
            with open(f"{prefix}.data-{sid:05d}-of-{shards:05d}", "rb") as f:
                files[sid] = f.read()
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError(f"{name}: data file too short")
        if e["crc32c"] is not None and masked_crc(raw) != e["crc32c"]:
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_checkpoint(prefix: str, tensors: dict) -> None:
    """{variable name: ndarray} -> `prefix`.index + `prefix`.data-00000-of-00001 (one shard, raw blocks)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1))   # num_shards = 1, little-endian (default), version.producer = 1
    items, blob = [(b"", header)], bytearray()
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        shape = a.shape                      # (np.ascontiguousarray would turn a scalar into shape (1,))
        a = np.ascontiguousarray(a)
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        if np.dtype(dt) not in _DTYPE_ENUM:
            raise TypeError(f"{name}: dtype {a.dtype} is not supported")
        raw = a.astype(dt, copy=False).tobytes()
        items.append((name.encode(), _serialize_entry(_DTYPE_ENUM[np.dtype(dt)], shape, len(blob), len(raw), masked_crc(raw))))
        blob += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(blob))
    with open(prefix + ".index", "wb") as f:
        f.write(_table_write(items))


def checkpoint_to_weights(prefix: str) -> dict:
    """The reference's weight dict layout (mc_cnn_brunch.py:61-66: keys are tf variable names, 'conv1/weights:0') from a Saver
    checkpoint; optimizer slots ('.../Momentum') and bookkeeping variables are dropped."""
    out = {}
    for name, a in read_checkpoint(prefix).items():
        parts = name.split("/")
        if len(parts) == 2 and parts[1] in ("weights", "biases") and (parts[0].startswith("conv") or parts[0].startswith("fc")):
            out[name + ":0"] = a.astype(np.float32)
    if not out:
        raise ValueError(f"{prefix}: no conv*/weights, conv*/biases variables found")
    return out


def weights_to_checkpoint(weights: dict, prefix: str) -> None:
    """The .npy dict layout -> a Saver checkpoint (variable names without the ':0' output suffix)."""
    write_checkpoint(prefix, {(k[:-2] if k.endswith(":0") else k): np.asarray(v, np.float32) for k, v in weights.items()})


def is_checkpoint_prefix(path: str) -> bool:
    return os.path.exists(path + ".index")


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description="convert between a tf.train.Saver checkpoint and the reference's .npy weight dict")
    ap.add_argument("src", help="checkpoint prefix (e.g. check_points_11_11/model.ckpt-14) or .npy file")
    ap.add_argument("dst", help=".npy file or checkpoint prefix")
    a = ap.parse_args(argv)
    if a.src.endswith(".npy"):
        w = np.load(a.src, encoding="bytes", allow_pickle=True).item()
        weights_to_checkpoint({(k.decode() if isinstance(k, bytes) else k): v for k, v in w.items()}, a.dst)
    else:
        np.save(a.dst, checkpoint_to_weights(a.src))


if __name__ == "__main__":
    main()
