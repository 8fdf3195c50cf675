"""Reader / writer for TensorFlow "tensor bundle" checkpoints (the V2 format of ``tf.train.Saver``), without TensorFlow.

The reference saves ``snapshots/snap-<step>`` with ``tf.train.Saver(GLOBAL_VARIABLES)`` (PointSegment/RandLANet.py:101-102,
180-184) and restores it in test mode (testPancreas.py:129-132).  A snapshot is two files:

    <prefix>.index                  an SSTable (LevelDB table format, tensorflow/core/lib/io/table*.cc): sorted
                                    key -> value entries; key "" holds a BundleHeaderProto, every other key is a variable
                                    name whose value is a BundleEntryProto (dtype, shape, shard, offset, size, crc32c)
    <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, addressed by (offset, size)

TensorFlow 1.11 is not installable in this environment and the reference ships no checkpoint, so the format is restated
from its published definition (tensor_bundle.proto, table_format.txt); tests round-trip through the writer below, which
emits what ``BundleWriter`` emits for single-shard checkpoints (uncompressed blocks, prefix-compressed keys with restart
interval 16, masked CRC32C block trailers).  Snappy-compressed blocks are rejected with a clear error (TF's bundle writer
does not produce them).
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_FOOTER = 48
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---- CRC32C (Castagnoli), table driven; LevelDB "masks" stored CRCs ------------------------------------------------
def _crc_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return np.array(tbl, dtype=np.uint32)


_TBL = _crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC32C of ``data``; large buffers go through the library's host helper ``pu_crc32c`` when it is built."""
    if len(data) >= 4096:
        try:
            import ctypes
            from . import _lib
            L = _lib.lib()
            L.pu_crc32c.restype = ctypes.c_uint
            L.pu_crc32c.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint]
            return int(L.pu_crc32c(bytes(data), len(data), crc))
        except (OSError, AttributeError, RuntimeError):
            pass
    c = crc ^ 0xFFFFFFFF
    tbl = _TBL
    for b in data:
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf ---------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf) -> dict:
    """{field number: [values]}: varints as int, length-delimited as bytes, fixed32/64 as int."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _field(num: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((num << 3) | wt) + payload


# ---- SSTable -----------------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if _mask(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError("checkpoint index: block checksum mismatch")
    if ctype != 0:
        raise ValueError("checkpoint index: compressed block (snappy) -- not produced by TensorFlow's BundleWriter, unsupported")
    return block


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_table(path: str, verify: bool = True) -> dict:
    data = open(path, "rb").read()
    if len(data) < _FOOTER or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError(f"{path} is not a TensorFlow checkpoint index (bad table magic)")
    foot = data[len(data) - _FOOTER:]
    pos = 0
    _, pos = _get_varint(foot, pos)          # metaindex handle: offset, size (unused)
    _, pos = _get_varint(foot, pos)
    ioff, pos = _get_varint(foot, pos)
    isize, pos = _get_varint(foot, pos)
    out = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        for k, v in _block_entries(_read_block(data, boff, bsize, verify)):
            out[bytes(k)] = bytes(v)
    return out


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.ri = bytearray(), [0], 0, b"", restart_interval

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count < self.ri:
            m = min(len(self.last), len(key))
            while shared < m and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _write_table(path: str, entries: list, block_size: int = 4096) -> None:
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.append(0)                                          # kNoCompression
        out.extend(struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return off, len(block)

    bb, last_key = _BlockBuilder(), None
    for key, value in entries:
        bb.add(key, value)
        last_key = key
        if len(bb.buf) >= block_size:
            off, size = emit(bb.finish())
            index.add(last_key, _put_varint(off) + _put_varint(size))
            bb = _BlockBuilder()
    if bb.buf or last_key is None:
        off, size = emit(bb.finish())
        index.add(last_key if last_key is not None else b"", _put_varint(off) + _put_varint(size))
    moff, msize = emit(_BlockBuilder().finish())               # empty metaindex block
    ioff, isize = emit(index.finish())
    foot = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", _MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- bundle --------------------------------------------------------------------------------------------------------------
def _shape_of(entry: dict) -> tuple:
    dims = []
    for shp in entry.get(2, []):
        for dim in _parse_proto(shp).get(2, []):
            d = _parse_proto(dim).get(1, [0])[0]
            dims.append(d - (1 << 64) if d >= 1 << 63 else d)
    return tuple(dims)


def list_variables(prefix: str) -> dict:
    """{name: (numpy dtype, shape)} of every tensor in the checkpoint ``prefix`` (``.index`` is appended)."""
    out = {}
    for key, val in _read_table(prefix + ".index").items():
        if key == b"":
            continue
        e = _parse_proto(val)
        dt = _DTYPES.get(e.get(1, [0])[0])
        out[key.decode()] = (dt, _shape_of(e))
    return out


def read_checkpoint(prefix: str, verify_crc: bool = True) -> dict:
    """{variable name: numpy array} of a TF V2 checkpoint (``<prefix>.index`` + ``<prefix>.data-XXXXX-of-YYYYY``)."""
    table = _read_table(prefix + ".index", verify_crc)
    header = _parse_proto(table.get(b"", b""))
    num_shards = header.get(1, [1])[0] or 1
    if header.get(2, [0])[0] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for key, val in table.items():
        if key == b"":
            continue
        e = _parse_proto(val)
        dtype_id = e.get(1, [0])[0]
        if 7 in e:
            raise ValueError(f"{key.decode()}: partitioned (sliced) variables are not supported")
        if dtype_id not in _DTYPES:
            continue  # strings / resources etc.: not part of the network's variables
        shard = e.get(3, [0])[0]
        if shard not in shards:
            shards[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), "rb").read()
        off, size = e.get(4, [0])[0], e.get(5, [0])[0]
        raw = shards[shard][off:off + size]
        if len(raw) != size:
            raise ValueError(f"{key.decode()}: data shard is truncated")
        if verify_crc and 6 in e and _mask(crc32c(raw)) != e[6][0]:
            raise ValueError(f"{key.decode()}: tensor checksum mismatch")
        out[key.decode()] = np.frombuffer(raw, dtype=_DTYPES[dtype_id]).reshape(_shape_of(e)).copy()
    return out


def write_checkpoint(prefix: str, tensors: dict) -> None:
    """Write ``{name: array}`` as a single-shard TF V2 checkpoint (what ``tf.train.Saver`` would restore by name)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(1))))
    entries = [(b"", header)]
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = np.asarray(tensors[name])   # (ascontiguousarray would turn a scalar into shape (1,); tobytes is C-order anyway)
        if a.dtype not in _DTYPE_IDS:
            raise ValueError(f"{name}: unsupported dtype {a.dtype}")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        shape = b"".join(_field(2, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(int(d))))) for d in a.shape)
        e = _field(1, 0, _put_varint(_DTYPE_IDS[a.dtype])) + _field(2, 2, _put_varint(len(shape)) + shape)
        if len(data):
            e += _field(4, 0, _put_varint(len(data)))
        e += _field(5, 0, _put_varint(len(raw))) + _field(6, 5, struct.pack("<I", _mask(crc32c(raw))))
        entries.append((name.encode(), e))
        data.extend(raw)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    _write_table(prefix + ".index", entries)
