"""TensorFlow checkpoint V2 ("tensor bundle": ``prefix.index`` + ``prefix.data-0000N-of-0000M``) reader / writer without
TensorFlow, so that the released ``model.ckpt`` (example.py:121-125, ``tf.train.Saver.restore``) can be dropped in:

    weights = ckpt.read_checkpoint('model/model.ckpt')      # {variable name: np.ndarray}
    variables.load_weights(weights)

SURVEY.md 8(f) row N3.  The format is restated from its public layout - ``.index`` is a LevelDB-style sorted table
(prefix-compressed blocks + index block + 48-byte footer, magic 0xdb4775248b80fb57) whose values are protobufs
(``BundleHeaderProto`` under the empty key, one ``BundleEntryProto`` {dtype, shape, shard_id, offset, size, crc32c} per
tensor); the data files hold the raw little-endian tensor bytes.  NOT validated against a TensorFlow-written file here (the
checkpoint is absent from the reference tree, `.MISSING_LARGE_BLOBS`): tests round-trip through the writer below and
check the table / protobuf / CRC-32C arithmetic against fixed known answers.  Snappy-compressed index blocks (not what
the TF bundle writer emits) are rejected."""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
# tensorflow DataType enum -> numpy
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in DTYPES.items()}

# ------------------------------------------------------------------ CRC-32C (Castagnoli), masked as in leveldb / TF
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (0x82F63B78 if _c & 1 else 0)
    _CRC_TABLE.append(_c)


def crc32c(data, crc=0):
    crc ^= 0xffffffff
    for b in bytes(data):
        crc = _CRC_TABLE[(crc ^ b) & 0xff] ^ (crc >> 8)
    return crc ^ 0xffffffff


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xffffffff


# ------------------------------------------------------------------ varints / protobuf wire format
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> list of (field number, wire type, value) with value int (varint / fixed) or bytes (length-delimited)."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _parse_shape(buf):
    dims = []
    for fn, wt, v in _parse_proto(buf):
        if fn == 2 and wt == 2:                       # TensorShapeProto.dim
            size = 0
            for f2, w2, v2 in _parse_proto(v):
                if f2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for fn, wt, v in _parse_proto(buf):
        if fn == 1:
            e['dtype'] = v
        elif fn == 2:
            e['shape'] = _parse_shape(v)
        elif fn == 3:
            e['shard_id'] = v
        elif fn == 4:
            e['offset'] = v
        elif fn == 5:
            e['size'] = v
        elif fn == 6:
            e['crc32c'] = v
        elif fn == 7:
            e['sliced'] = True
    return e


# ------------------------------------------------------------------ sorted-table reader
def _read_block(buf, offset, size, verify=True):
    body = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from('<I', buf, offset + size + 1)[0]
    if verify and mask_crc(crc32c(buf[offset:offset + size + 1])) != stored:
        raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise NotImplementedError("checkpoint index: compressed block (type %d)" % ctype)
    return body


def _block_entries(block):
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix, verify=True):
    """-> (header dict, {name: entry dict}) of ``prefix.index``."""
    buf = open(prefix + '.index', 'rb').read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s.index is not a TensorFlow V2 checkpoint index (bad magic)" % prefix)
    foot = buf[len(buf) - 48:]
    pos = 0
    _, pos = _get_varint(foot, pos)            # metaindex handle
    _, pos = _get_varint(foot, pos)
    ioff, pos = _get_varint(foot, pos)
    isize, pos = _get_varint(foot, pos)
    header, entries = dict(num_shards=1, endianness=0), {}
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        for key, val in _block_entries(_read_block(buf, boff, bsize, verify)):
            if key == b'':
                for fn, wt, v in _parse_proto(val):
                    if fn == 1:
                        header['num_shards'] = v
                    elif fn == 2:
                        header['endianness'] = v
            else:
                entries[key.decode('utf-8')] = _parse_entry(val)
    if header['endianness'] != 0:
        raise NotImplementedError("big-endian checkpoint")
    return header, entries


def _fast_crc():
    """CRC-32C over whole tensors through libatvs.so's host helper (slice-by-8); None when the library is not built."""
    try:
        from . import _lib
        lib = _lib.load()
        return lambda raw: int(lib.atvs_crc32c(raw, len(raw), 0))
    except Exception:
        return None


def read_checkpoint(prefix, names=None, verify_crc_below=1 << 16, required=()):
    """{variable name: np.ndarray} from a V2 checkpoint.  ``names``: optional filter (iterable or predicate).  Every
    tensor's CRC-32C is verified (libatvs.so's host helper; without the library only tensors smaller than
    ``verify_crc_below`` bytes, in pure Python).  Partitioned (sliced) variables are not assembled: one that is listed in
    ``required`` raises, others are skipped."""
    header, entries = read_index(prefix)
    want = (lambda n: True) if names is None else (names if callable(names) else set(names).__contains__)
    files, out = {}, {}
    fast = _fast_crc()
    required = set(required)
    for name, e in entries.items():
        if e['sliced'] and name in required:
            raise ValueError("checkpoint variable %r is stored as partitioned slices, which this reader does not assemble" % name)
        if not want(name) or e['sliced']:
            continue
        if e['dtype'] not in DTYPES:
            continue                                   # strings / resources: not weights
        dt = np.dtype(DTYPES[e['dtype']])
        path = '%s.data-%05d-of-%05d' % (prefix, e['shard_id'], header['num_shards'])
        if path not in files:
            files[path] = open(path, 'rb')
        f = files[path]
        f.seek(e['offset'])
        raw = f.read(e['size'])
        n = int(np.prod(e['shape'])) if e['shape'] else 1
        if len(raw) != e['size'] or e['size'] != n * dt.itemsize:
            raise ValueError("checkpoint tensor %r: %d bytes for shape %s %s" % (name, len(raw), e['shape'], dt))
        if e['crc32c'] is not None and (fast is not None or e['size'] < verify_crc_below):
            got = fast(raw) if fast is not None else crc32c(raw)
            if mask_crc(got) != e['crc32c']:
                raise ValueError("checkpoint tensor %r: CRC mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).copy()
    for f in files.values():
        f.close()
    return out


# ------------------------------------------------------------------ writer (export / tests)
def _pb_varint_field(fn, v):
    return _put_varint(fn << 3) + _put_varint(v)


def _pb_bytes_field(fn, b):
    return _put_varint((fn << 3) | 2) + _put_varint(len(b)) + b


def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_checkpoint(prefix, tensors, entries_per_block=24):
    """write {name: ndarray} as a single-shard V2 checkpoint (sorted keys, several data blocks, prefix compression)."""
    names = sorted(tensors)
    data, items = bytearray(), []
    header = _pb_varint_field(1, 1) + _pb_varint_field(2, 0) + _pb_bytes_field(3, _pb_varint_field(1, 1))
    items.append((b'', header))
    for n in names:
        a = np.require(np.asarray(tensors[n]), requirements='C')      # keeps 0-d tensors 0-d
        if a.dtype not in _DT_OF:
            raise ValueError("write_checkpoint: dtype %s" % a.dtype)
        raw = a.tobytes()
        shape = b''.join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in a.shape)
        entry = _pb_varint_field(1, _DT_OF[a.dtype]) + _pb_bytes_field(2, shape)
        if len(data):
            entry += _pb_varint_field(4, len(data))
        entry += _pb_varint_field(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack('<I', mask_crc(crc32c(raw)))
        items.append((n.encode('utf-8'), entry))
        data += raw
    table, index_items = bytearray(), []

    def emit(block):
        off = len(table)
        trailer = bytes([0])
        table.extend(block + trailer + struct.pack('<I', mask_crc(crc32c(block + trailer))))
        return _put_varint(off) + _put_varint(len(block))

    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        index_items.append((chunk[-1][0], emit(_build_block(chunk))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index_items, restart_interval=1))
    footer = meta + idx
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    open(prefix + '.index', 'wb').write(bytes(table) + footer)
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
