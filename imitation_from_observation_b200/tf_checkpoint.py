"""Reader / writer for TensorFlow checkpoints — what `tf.train.Saver()` writes in the reference:
`saver.save(sess, 'model/.../model_...')` (scripts/train_script.py:189), `tf.train.Saver(...).restore(sess, modelfile)`
(rllab/sampler/base.py:144-145) and the slim `inception_v3.ckpt` restored at base.py:130-132 / train_script.py:113-115.

Two on-disk formats exist and both are read here:

  V2 "tensor bundle" (Saver default since TF 0.12):
    <prefix>.index                    an SSTable (TensorFlow's copy of the LevelDB table format) mapping
                                      "" -> BundleHeaderProto and  <variable name> -> BundleEntryProto
    <prefix>.data-00000-of-00001      the raw little-endian tensor bytes, addressed by (shard, offset, size)
  V1 (one file, e.g. slim's published `inception_v3.ckpt`):
    <path>                            an SSTable mapping "" -> SavedTensorSlices{meta} (names, shapes, dtypes) and one
                                      ordered-code key per stored slice -> SavedTensorSlices{data: SavedSlice{name, slice,
                                      TensorProto}} with the values inline (tensor_content or the typed *_val fields)

No TensorFlow here and no checkpoint ships with the reference, so this module is written from the published formats
(tensorflow/core/util/{tensor_bundle,saved_tensor_slice.proto,tensor_slice_writer}, tensorflow/core/lib/io/{table,block,
format}, the protobuf wire format, Snappy's format description) and is **not validated against a TF-written file**;
tests/test_host_logic.py holds it to its own writers, to hand-assembled tables with prefix-compressed keys, several data
blocks and Snappy-compressed blocks, and to the CRC-32C check values of the published algorithm.  Only what a Saver
checkpoint of dense float variables needs is implemented: DT_FLOAT / DT_DOUBLE / DT_INT32 / DT_INT64 / DT_UINT8 / DT_HALF /
DT_BFLOAT16-as-uint16 tensors, little-endian files; V2 entries must be unsliced, V1 slices are assembled.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           14: np.uint16, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}     # tensorflow DataType enum
_DTYPE_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.uint8): 4,
             np.dtype(np.int64): 9, np.dtype(np.float16): 19}


class CheckpointError(ValueError):
    pass


# ---- CRC-32C, masked as in tensorflow/core/lib/hash/crc32c.h ------------------------------------------------------------
def crc32c(data: bytes, crc: int = 0) -> int:
    try:
        from . import _lib
        buf = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        return int(_lib.load().ctx_crc32c(bytes(buf), len(buf), crc)) & 0xFFFFFFFF
    except Exception:                                          # library not built: bitwise fallback (slow, small inputs)
        c = ~crc & 0xFFFFFFFF
        for b in bytes(data):
            c ^= b
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        return ~c & 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / protobuf wire format -----------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise CheckpointError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message; value = int (varint / fixed) or bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            if len(v) != n:
                raise CheckpointError("truncated protobuf field")
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(buf: bytes) -> dict:
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto: dim=2 {size=1}), shard_id=3, offset=4, size=5, crc32c=6 (fixed32),
    slices=7."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for num, wt, v in _fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            for n2, _, v2 in _fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _fields(v2):
                        if n3 == 1:
                            size = _signed64(v3)
                    e["shape"].append(size)
                elif n2 == 3 and v2:
                    raise CheckpointError("tensor of unknown rank in a checkpoint")
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
        elif num == 7:
            e["sliced"] = True
    return e


def _entry_bytes(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ---- Snappy (raw format) decompression: index blocks may be stored compressed ---------------------------------------------------
def snappy_decompress(buf: bytes) -> bytes:
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                         # literal
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
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy stream")
        for _ in range(ln):                                   # byte by byte: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy stream decodes to %d bytes, header says %d" % (len(out), n))
    return bytes(out)


# ---- table (SSTable) ---------------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    """Block contents + 1-byte compression type + 4-byte masked CRC-32C of (contents + type)."""
    raw = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise CheckpointError("block handle points outside the index file")
    if verify:
        want = struct.unpack("<I", trailer[1:])[0]
        if mask_crc(crc32c(raw + trailer[:1])) != want:
            raise CheckpointError("index block checksum mismatch")
    if trailer[0] == 0:
        return raw
    if trailer[0] == 1:
        return snappy_decompress(raw)
    raise CheckpointError("unknown block compression type %d" % trailer[0])


def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise CheckpointError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    if end < 0:
        raise CheckpointError("corrupt block restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > end:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise CheckpointError("%s is not a TensorFlow / LevelDB table (bad magic)" % path)
    footer = data[-48:]
    _, p = _varint(footer, 0)                                  # metaindex handle (unused)
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)
    isize, p = _varint(footer, p)
    out: List[Tuple[bytes, bytes]] = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _varint(handle, 0)
        bsize, q = _varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, entries: List[Tuple[bytes, bytes]], block_size: int = 4096):
    """Uncompressed table; `entries` must be sorted by key (the table format requires it)."""
    keys = [k for k, _ in entries]
    if keys != sorted(keys) or len(set(keys)) != len(keys):
        raise CheckpointError("table keys must be unique and sorted")
    out = bytearray()
    index: List[Tuple[bytes, bytes]] = []

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    cur: List[Tuple[bytes, bytes]] = []
    size = 0
    for k, v in entries:
        cur.append((k, v))
        size += len(k) + len(v) + 3
        if size >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, size = [], 0
    if cur or not index:
        index.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- the bundle ----------------------------------------------------------------------------------------------------------------
def _is_v2(prefix: str) -> bool:
    return os.path.exists(prefix + ".index")


def _is_v1(path: str) -> bool:
    """A single-file (Saver V1 / slim) checkpoint: a table file at `path` itself."""
    if not os.path.isfile(path) or os.path.getsize(path) < 48:
        return False
    with open(path, "rb") as f:
        f.seek(-8, os.SEEK_END)
        return struct.unpack("<Q", f.read(8))[0] == TABLE_MAGIC


def checkpoint_exists(prefix: str) -> bool:
    return _is_v2(prefix) or _is_v1(prefix)


def list_variables(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...]]]:
    """{variable name: (dtype, shape)} of a V2 bundle `<prefix>.index` or a V1 single-file checkpoint `<prefix>`."""
    if _is_v2(prefix):
        return _list_variables_v2(prefix)
    if _is_v1(prefix):
        return {k: (dt, shp) for k, (dt, shp, _) in _v1_meta(read_table(prefix)).items()}
    raise CheckpointError("%s is neither a TensorFlow V2 bundle (<prefix>.index) nor a V1 checkpoint file" % prefix)


def read_checkpoint(prefix: str, names: Optional[List[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """{variable name: ndarray} of a Saver checkpoint: V2 bundle `<prefix>.index` + `<prefix>.data-*`, or the V1
    single file `<prefix>` (slim's `inception_v3.ckpt`)."""
    if _is_v2(prefix):
        return _read_checkpoint_v2(prefix, names, verify)
    if _is_v1(prefix):
        return _read_checkpoint_v1(prefix, names, verify)
    raise CheckpointError("%s is neither a TensorFlow V2 bundle (<prefix>.index) nor a V1 checkpoint file" % prefix)


# ---- V1: SavedTensorSlices in one table ------------------------------------------------------------------------------------
def _parse_shape(buf: bytes) -> List[int]:
    dims = []
    for n2, _, v2 in _fields(buf):                             # TensorShapeProto: dim = 2 { size = 1 }
        if n2 == 2:
            size = 0
            for n3, _, v3 in _fields(v2):
                if n3 == 1:
                    size = _signed64(v3)
            dims.append(size)
    return dims


def _parse_slice(buf: bytes) -> List[Tuple[int, Optional[int]]]:
    """TensorSliceProto: extent = 1 { start = 1, length = 2 (absent = the full dimension) } -> [(start, length | None)]."""
    out = []
    for n, _, v in _fields(buf):
        if n == 1:
            start, length = 0, None
            for n2, _, v2 in _fields(v):
                if n2 == 1:
                    start = _signed64(v2)
                elif n2 == 2:
                    length = _signed64(v2)
            out.append((start, length))
    return out


def _v1_meta(table) -> Dict[str, Tuple[np.dtype, Tuple[int, ...], int]]:
    if not table or table[0][0] != b"":
        raise CheckpointError("V1 checkpoint has no SavedTensorSlices meta entry")
    out = {}
    for n, _, v in _fields(table[0][1]):
        if n != 1:                                             # SavedTensorSlices.meta
            continue
        for n2, _, v2 in _fields(v):
            if n2 != 1:                                        # SavedTensorSliceMeta.tensor
                continue
            name, shape, dtype = "", [], 0
            for n3, _, v3 in _fields(v2):                      # SavedSliceMeta: name = 1, shape = 2, type = 3
                if n3 == 1:
                    name = v3.decode("utf-8")
                elif n3 == 2:
                    shape = _parse_shape(v3)
                elif n3 == 3:
                    dtype = v3
            out[name] = (np.dtype(_DTYPES.get(dtype, np.uint8)), tuple(shape), dtype)
    return out


_V1_VAL_FIELDS = {5: "<f4", 6: "<f8", 7: "<i4", 10: "<i8", 11: "u1", 13: "<i4"}   # TensorProto *_val (13 = half_val as int32)


def _tensor_proto_values(buf: bytes, dt: np.dtype, count: int) -> np.ndarray:
    """Flat values of a TensorProto: tensor_content (4) or the typed, packed or repeated, *_val field."""
    content, packed, scalars, field = None, [], [], None
    for n, wt, v in _fields(buf):
        if n == 4:
            content = v
        elif n in _V1_VAL_FIELDS:
            field = n
            if wt == 2:
                packed.append(v)
            else:
                scalars.append(v)
    if content is not None and len(content):
        return np.frombuffer(content, dtype=dt.newbyteorder("<")).astype(dt)
    if field is None:
        return np.zeros(count, dtype=dt)
    wire = np.dtype(_V1_VAL_FIELDS[field])
    if packed:
        raw = b"".join(packed)
        if field in (7, 10, 11, 13):                           # varint-packed integer fields
            vals, pos = [], 0
            while pos < len(raw):
                x, pos = _varint(raw, pos)
                vals.append(_signed64(x))
            arr = np.asarray(vals, dtype=np.int64)
        else:
            arr = np.frombuffer(raw, dtype=wire)
    else:
        if field in (5,):
            arr = np.asarray(scalars, dtype=np.uint32).view("<f4")
        elif field in (6,):
            arr = np.asarray(scalars, dtype=np.uint64).view("<f8")
        else:
            arr = np.asarray([_signed64(x) for x in scalars], dtype=np.int64)
    if field == 13:                                            # half_val: the 16 bits of each value in an int32
        arr = arr.astype(np.uint16).view(np.float16)
    if arr.size == 1 and count > 1:                            # TensorProto's "all elements equal" compression
        arr = np.repeat(arr, count)
    return arr.astype(dt)


def _read_checkpoint_v1(path: str, names: Optional[List[str]], verify: bool) -> Dict[str, np.ndarray]:
    table = read_table(path, verify)
    meta = _v1_meta(table)
    want = None if names is None else set(names)
    out: Dict[str, np.ndarray] = {}
    filled: Dict[str, int] = {}
    for k, v in table[1:]:
        for n, _, sv in _fields(v):
            if n != 2:                                         # SavedTensorSlices.data = SavedSlice
                continue
            name, sl, tp = "", [], b""
            for n2, _, v2 in _fields(sv):                      # SavedSlice: name = 1, slice = 2, data = 3
                if n2 == 1:
                    name = v2.decode("utf-8")
                elif n2 == 2:
                    sl = _parse_slice(v2)
                elif n2 == 3:
                    tp = v2
            if want is not None and name not in want:
                continue
            if name not in meta:
                raise CheckpointError("V1 checkpoint %s: slice of unknown tensor %s" % (path, name))
            dt, shape, code = meta[name]
            if code not in _DTYPES:
                raise CheckpointError("variable %s has unsupported dtype enum %d" % (name, code))
            if len(sl) not in (0, len(shape)):
                raise CheckpointError("variable %s: slice rank %d does not match shape %s" % (name, len(sl), shape))
            idx = tuple(slice(s, None if ln is None or ln < 0 else s + ln) for s, ln in sl) if sl else tuple(slice(None) for _ in shape)
            if name not in out:
                out[name] = np.zeros(shape, dtype=dt)
                filled[name] = 0
            view = out[name][idx] if shape else out[name]
            vals = _tensor_proto_values(tp, dt, int(view.size))
            if vals.size != view.size:
                raise CheckpointError("variable %s: slice holds %d values, %d expected" % (name, vals.size, view.size))
            if shape:
                out[name][idx] = vals.reshape(view.shape)
            else:
                out[name] = vals.reshape(())
            filled[name] += int(vals.size)
    for name, arr in out.items():
        if filled[name] != arr.size:
            raise CheckpointError("variable %s: %d of %d values present in %s" % (name, filled[name], arr.size, path))
    if want is not None and want - set(out):
        raise KeyError("not in checkpoint %s: %s" % (path, sorted(want - set(out))))
    return out


def _ordered_name_slice_key(name: str, rank: int) -> bytes:
    """Key of a full-extent slice as TensorSliceWriter encodes it (OrderedCode: 0, name, rank, (start, length) per
    dimension with length = -1 for "full").  Readers here never decode keys; the writer needs them unique and sorted."""
    esc = b"".join(b"\x00\xff" if c == 0 else b"\xff\x00" if c == 255 else bytes([c]) for c in name.encode("utf-8"))
    num = lambda v: b"\x00" if v == 0 else bytes([len(v.to_bytes((v.bit_length() + 7) // 8, "big"))]) + v.to_bytes((v.bit_length() + 7) // 8, "big")
    return num(0) + esc + b"\x00\x01" + num(rank) + (b"\x80" + b"\x7f") * rank


def write_checkpoint_v1(path: str, variables: Dict[str, np.ndarray], use_tensor_content: bool = False):
    """Single-file V1 checkpoint holding `variables` as full slices (float values in the packed float_val field, as
    TensorSliceWriter::SaveData does, or in tensor_content)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)

    def ld(num: int, payload: bytes) -> bytes:                 # length-delimited field
        return _put_varint(num << 3 | 2) + _put_varint(len(payload)) + payload

    def shape_proto(shape) -> bytes:
        return b"".join(ld(2, _put_varint(1 << 3) + _put_varint(int(d))) for d in shape)

    metas, entries = b"", []
    for name in sorted(variables, key=lambda s: s.encode("utf-8")):
        a = np.ascontiguousarray(variables[name]) if np.ndim(variables[name]) else np.asarray(variables[name])
        if a.dtype not in _DTYPE_OF:
            raise CheckpointError("cannot store dtype %s (variable %s)" % (a.dtype, name))
        code = _DTYPE_OF[a.dtype]
        full = b"".join(ld(1, b"") for _ in a.shape)           # TensorSliceProto: one empty Extent per dimension
        metas += ld(1, ld(1, name.encode("utf-8")) + ld(2, shape_proto(a.shape)) + _put_varint(3 << 3) + _put_varint(code) +
                    ld(4, full))
        if use_tensor_content or a.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            body = ld(4, a.astype(a.dtype.newbyteorder("<")).tobytes())
        else:
            body = ld(5 if a.dtype == np.float32 else 6, a.astype(a.dtype.newbyteorder("<")).tobytes())
        tensor = _put_varint(1 << 3) + _put_varint(code) + ld(2, shape_proto(a.shape)) + body
        entries.append((_ordered_name_slice_key(name, a.ndim),
                        ld(2, ld(1, name.encode("utf-8")) + ld(2, full) + ld(3, tensor))))
    entries.sort(key=lambda kv: kv[0])
    write_table(path, [(b"", ld(1, metas))] + entries)


def _list_variables_v2(prefix: str) -> Dict[str, Tuple[np.dtype, Tuple[int, ...]]]:
    out = {}
    for k, v in read_table(prefix + ".index"):
        if k:
            e = _parse_entry(v)
            out[k.decode("utf-8")] = (np.dtype(_DTYPES.get(e["dtype"], np.uint8)), tuple(e["shape"]))
    return out


def _read_checkpoint_v2(prefix: str, names: Optional[List[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """{variable name: ndarray} of a Saver V2 checkpoint `<prefix>.index` + `<prefix>.data-*`."""
    table = read_table(prefix + ".index", verify)
    if not table or table[0][0] != b"":
        raise CheckpointError("%s.index has no bundle header" % prefix)
    num_shards, endian = 1, 0
    for num, _, v in _fields(table[0][1]):
        if num == 1:
            num_shards = v
        elif num == 2:
            endian = v
    if endian != 0:
        raise CheckpointError("big-endian tensor bundles are not supported")
    shards: Dict[int, np.memmap] = {}
    out: Dict[str, np.ndarray] = {}
    want = None if names is None else set(names)
    for k, v in table[1:]:
        name = k.decode("utf-8")
        if want is not None and name not in want:
            continue
        e = _parse_entry(v)
        if e["sliced"]:
            raise CheckpointError("variable %s is stored in slices (partitioned variable): not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise CheckpointError("variable %s has unsupported dtype enum %d" % (name, e["dtype"]))
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError("variable %s: %d bytes stored, shape %s needs %d" % (name, e["size"], e["shape"], count * dt.itemsize))
        sid = e["shard_id"]
        if sid not in shards:
            path = "%s.data-%05d-of-%05d" % (prefix, sid, num_shards)
            if not os.path.exists(path):
                raise CheckpointError("missing data shard %s" % path)
            shards[sid] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise CheckpointError("variable %s lies outside its data shard" % name)
        raw = np.ascontiguousarray(raw)
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw.tobytes())) != e["crc32c"]:
            raise CheckpointError("variable %s: checksum mismatch" % name)
        out[name] = raw.view(dt).reshape(e["shape"]).copy()
    if want is not None and want - set(out):
        raise KeyError("not in checkpoint %s: %s" % (prefix, sorted(want - set(out))))
    return out


def write_checkpoint(prefix: str, variables: Dict[str, np.ndarray]):
    """Write `<prefix>.index` + `<prefix>.data-00000-of-00001` holding `variables` (dense, one shard, little endian)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries: List[Tuple[bytes, bytes]] = []
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(variables, key=lambda s: s.encode("utf-8")):
            a = np.asarray(variables[name])
            shape = a.shape                                    # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in _DTYPE_OF:
                raise CheckpointError("cannot store dtype %s (variable %s)" % (a.dtype, name))
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            raw = np.ascontiguousarray(a).tobytes()
            f.write(raw)
            entries.append((name.encode("utf-8"),
                            _entry_bytes(_DTYPE_OF[a.dtype], shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    # BundleHeaderProto: num_shards = 1, endianness = LITTLE (default, omitted), version { producer = 1 }
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"
    write_table(prefix + ".index", [(b"", header)] + entries)


# ---- one restore path for every caller (compat.Saver.restore, sampler.RewardFeaturizer.initialize) ---------------------------
def load_variables(path: str, names: List[str], prefixes=("",), verify: bool = True) -> Tuple[Dict[str, np.ndarray], str]:
    """Values of `names` from whatever `path` is — a TensorFlow checkpoint (V2 bundle `<path>.index` or V1 single file),
    else the `.npz` keyed by the same variable names that compat.Saver.save writes by default — under the first prefix of
    `prefixes` for which every name is present.  The reference builds the model under "contextmodel/" in the trainer
    (train_script.py:120) but not always in the sampler (base.py:134-138), hence the probing.
    Returns ({prefix + name: array}, prefix)."""
    if checkpoint_exists(path):
        have = set(list_variables(path))
        read = lambda keys: read_checkpoint(path, keys, verify)
    else:
        npz = path if path.endswith(".npz") else path + ".npz"
        if not os.path.exists(npz):
            raise FileNotFoundError("no checkpoint at %s (looked for %s.index, a V1 checkpoint file %s, and %s)"
                                    % (path, path, path, npz))
        data = np.load(npz)
        have = set(data.files)
        read = lambda keys: {k: data[k] for k in keys}
    for pre in dict.fromkeys(prefixes):
        if all(pre + n in have for n in names):
            return read([pre + n for n in names]), pre
    pre0 = next(iter(prefixes))
    missing = [n for n in names if pre0 + n not in have]
    raise KeyError("checkpoint %s lacks %d variables of the model, e.g. %s" % (path, len(missing), missing[:3]))
