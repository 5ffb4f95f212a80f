"""TF-1 checkpoint import / export: the files `tf.train.Saver().save(sess, path)` writes at
src/train.py:121 and `saver.restore` reads at src/train.py:93-94 / src/eval_embed_reason.py:27 --
`<path>.index` (a leveldb-format table of BundleEntryProto records) and `<path>.data-00000-of-00001` (raw tensor
bytes) -- read and written without TensorFlow, plus the mapping between the reference graph's variable names and this
library's canonical parameter names.

PROVENANCE (same caveat as SURVEY section 8c): TensorFlow is not installed here and the reference ships no
checkpoint, so nothing in this module could be checked against a file TensorFlow wrote.  The container format
(table blocks, footer, masked crc32c, snappy, the two protos) follows the published leveldb / tensor_bundle formats and
is exercised by round-trip tests; the VARIABLE NAMES and the cuDNN parameter layouts below are restated from the
TF-1.x sources as recalled and are marked as assumptions where they are.  `load_reference_checkpoint` therefore lists
every tensor it could not place instead of guessing.
"""
import os
import struct

import numpy as np

# ------------------------------------------------------------------------------------------------ crc32c, varints

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of leveldb tables and tensor bundles."""
    t = _crc_table()
    c = (~crc) & 0xFFFFFFFF
    for b in bytes(data):
        c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
    return (~c) & 0xFFFFFFFF


def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def _zeros_operator(nbytes):
    """the GF(2) matrix that advances a (finalised) CRC-32C over `nbytes` zero bytes (zlib's crc32_combine scheme)."""
    op = None
    odd = [0x82F63B78] + [1 << (n - 1) for n in range(1, 32)]     # one zero bit
    even = _gf2_square(odd)                                         # two
    odd = _gf2_square(even)                                         # four
    cur, n = odd, nbytes
    while n:
        cur = _gf2_square(cur)                                      # 8, 16, 32, ... bits = 1, 2, 4, ... bytes
        if n & 1:
            op = cur if op is None else [_gf2_times(cur, col) for col in op]
        n >>= 1
    return op


def crc32c_combine(crc1, crc2, len2):
    """CRC-32C of A || B from crc32c(A), crc32c(B) and len(B)."""
    return crc1 if len2 <= 0 else _gf2_times(_zeros_operator(len2), crc1) ^ crc2


def _crc32c_np(buf, chunk=8192):
    """CRC-32C of a large buffer: numpy lanes (one per `chunk`-byte piece, all advanced together byte by byte), then
    the pieces are stitched with the zero-bytes operator -- 16 MB in well under a second instead of ~10 s."""
    a = np.frombuffer(bytes(buf), np.uint8)
    k = a.size // chunk
    if k < 4:
        return crc32c(a.tobytes())
    t = _crc_table()
    lanes = a[:k * chunk].reshape(k, chunk)
    c = np.full(k, 0xFFFFFFFF, np.uint32)
    for j in range(chunk):
        c = t[(c ^ lanes[:, j]) & 0xFF] ^ (c >> 8)
    c = (~c).astype(np.uint32)
    op = _zeros_operator(chunk)
    acc = int(c[0])
    for i in range(1, k):
        acc = _gf2_times(op, acc) ^ int(c[i])
    tail = a[k * chunk:].tobytes()
    return crc32c_combine(acc, crc32c(tail), len(tail)) if tail else acc


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


def put_varint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def get_varint(buf, pos):
    n = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        n |= (b & 0x7F) << shift
        if b < 0x80:
            return n, pos
        shift += 7


# ------------------------------------------------------------------------------------------------ snappy (decoder only)

def snappy_decompress(src):
    """raw snappy block format (leveldb compresses table blocks with it when that saves >= 12.5 %)."""
    n, pos = get_varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:                                   # copy, 1-byte offset
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:                                 # copy, 2-byte offset
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 2], 'little')
            pos += 2
        else:                                           # copy, 4-byte offset
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy stream')
        for _ in range(ln):                             # may overlap: byte by byte
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy length mismatch')
    return bytes(out)


# ------------------------------------------------------------------------------------------------ minimal protobuf

def pb_fields(buf):
    """yields (field number, wire type, value) of one message; value is int (varint / fixed) or bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = get_varint(buf, pos)
        f, w = key >> 3, key & 7
        if w == 0:
            v, pos = get_varint(buf, pos)
        elif w == 1:
            v = int.from_bytes(buf[pos:pos + 8], 'little')
            pos += 8
        elif w == 2:
            ln, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif w == 5:
            v = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % w)
        yield f, w, v


def _pb_varint(f, v):
    return put_varint(f << 3) + put_varint(v)


def _pb_bytes(f, b):
    return put_varint((f << 3) | 2) + put_varint(len(b)) + b


# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


def parse_entry(buf):
    """BundleEntryProto: dtype=1, shape=2 {dim=2 {size=1}}, shard_id=3, offset=4, size=5, crc32c=6 (fixed32), slices=7."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, w, v in pb_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, v2 in pb_fields(v):
                if f2 == 2:
                    sz = 0
                    for f3, _, v3 in pb_fields(v2):
                        if f3 == 1:
                            sz = v3
                    e['shape'].append(sz)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['sliced'] = True
    return e


def build_entry(dtype_id, shape, offset, size, crc):
    shp = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)
    return (_pb_varint(1, dtype_id) + _pb_bytes(2, shp) + _pb_varint(4, offset) + _pb_varint(5, size) +
            put_varint((6 << 3) | 5) + struct.pack('<I', crc))


# ------------------------------------------------------------------------------------------------ leveldb table

TABLE_MAGIC = 0xDB4775248B80FB57


def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError('truncated table block')
    data, ctype, crc = raw[:size], raw[size], struct.unpack('<I', raw[size + 1:])[0]
    if verify and unmask_crc(crc) != crc32c(raw[:size + 1]):
        raise ValueError('table block checksum mismatch')
    if ctype == 1:
        data = snappy_decompress(data)
    elif ctype != 0:
        raise ValueError('unknown block compression %d' % ctype)
    return data


def _block_entries(data):
    nrestart = struct.unpack('<I', data[-4:])[0]
    end = len(data) - 4 - 4 * nrestart
    pos, key = 0, b''
    while pos < end:
        shared, pos = get_varint(data, pos)
        unshared, pos = get_varint(data, pos)
        vlen, pos = get_varint(data, pos)
        key = key[:shared] + data[pos:pos + unshared]
        pos += unshared
        yield key, data[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """all (key, value) pairs of a leveldb-format table file, in key order."""
    with open(path, 'rb') as f:
        f.seek(0, os.SEEK_END)
        n = f.tell()
        if n < 48:
            raise ValueError('not a table file: ' + path)
        f.seek(n - 48)
        foot = f.read(48)
        if struct.unpack('<Q', foot[40:])[0] != TABLE_MAGIC:
            raise ValueError('bad table magic: ' + path)
        pos = 0
        _, pos = get_varint(foot, pos)          # metaindex handle
        _, pos = get_varint(foot, pos)
        ioff, pos = get_varint(foot, pos)
        isize, pos = get_varint(foot, pos)
        out = []
        for _, handle in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, p = get_varint(handle, 0)
            bsize, p = get_varint(handle, p)
            out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
        return out


def _make_block(pairs, restart_interval=16):
    buf, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        buf += put_varint(shared) + put_varint(len(k) - shared) + put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', len(restarts))
    return bytes(buf)


def write_table(path, pairs, block_bytes=4096):
    """writes sorted (key, value) pairs as an uncompressed leveldb-format table."""
    pairs = sorted(pairs)
    with open(path, 'wb') as f:
        def emit(block):
            off = f.tell()
            f.write(block + b'\x00' + struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
            return put_varint(off) + put_varint(len(block))
        index, cur, size = [], [], 0
        for k, v in pairs:
            cur.append((k, v))
            size += len(k) + len(v) + 8
            if size >= block_bytes:
                index.append((cur[-1][0], emit(_make_block(cur))))
                cur, size = [], 0
        if cur or not index:
            index.append((cur[-1][0] if cur else b'', emit(_make_block(cur))))
        meta = emit(_make_block([]))
        idx = emit(_make_block(index, restart_interval=1))
        foot = meta + idx
        f.write(foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', TABLE_MAGIC))


# ------------------------------------------------------------------------------------------------ tensor bundle

def load_bundle(prefix, verify_tensors=False):
    """{variable name: ndarray} of a V2 checkpoint `<prefix>.index` + `<prefix>.data-*`."""
    entries = read_table(prefix + '.index')
    shards, out = {}, {}
    nshards = 1
    for k, v in entries:
        if k == b'':
            for f, _, val in pb_fields(v):          # BundleHeaderProto: num_shards=1, endianness=2, version=3
                if f == 1:
                    nshards = val
                elif f == 2 and val != 0:
                    raise ValueError('big-endian bundles are not supported')
            continue
        e = parse_entry(v)
        if e['sliced']:
            raise ValueError('%s: partitioned variables are not supported' % k.decode())
        if e['dtype'] not in DTYPES:
            continue                                  # strings etc.: nothing this model saves
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap('%s.data-%05d-of-%05d' % (prefix, sid, nshards), dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if verify_tensors and e['crc32c'] is not None and unmask_crc(e['crc32c']) != _crc32c_np(raw):
            raise ValueError('%s: tensor checksum mismatch' % k.decode())
        out[k.decode()] = np.frombuffer(bytes(raw), dtype=DTYPES[e['dtype']]).reshape(e['shape'])
    return out


def save_bundle(prefix, tensors, checksum=True):
    """writes {name: ndarray} as `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard, little endian)."""
    pairs = [(b'', _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))]      # num_shards=1, version{producer=1}
    off = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in DTYPE_IDS:
                raise ValueError('%s: dtype %s cannot be saved' % (name, a.dtype))
            raw = a.tobytes()
            f.write(raw)
            crc = mask_crc(_crc32c_np(raw)) if checksum else 0
            pairs.append((name.encode(), build_entry(DTYPE_IDS[a.dtype], a.shape, off, len(raw), crc)))
            off += len(raw)
    write_table(prefix + '.index', pairs)


# ------------------------------------------------------------------------------------------------ name / layout mapping
#
# Reference graph (src/model.py) -> variables tf.train.Saver() saves.  ASSUMPTIONS (unverifiable here, see header):
#   * tf.layers.dense(name=n) under scope s                -> s/n/kernel (in,out), s/n/bias            [= canonical]
#   * tf.contrib.cudnn_rnn.CudnnGRU(L, H, name=n) in scope s saves through CudnnOpaqueParamsSaveable in the
#     "cudnn-compatible" canonical form, per layer l:
#       s/n/[cudnn_gru/]rnn/multi_rnn_cell/cell_l/cudnn_compatible_gru_cell/gates/kernel      (in+H, 2H)  cols [r | u]
#       .../gates/bias (2H) = bW + bR of r,u   .../candidate/input_projection/{kernel (in,H), bias}  = W_n^T, bW_n
#       .../candidate/hidden_projection/{kernel (H,H), bias} = R_n^T, bR_n
#   * its Adam slots are slots of the OPAQUE variable s/n/opaque_kernel and are saved raw, in cuDNN's blob layout:
#       all weight matrices first -- per layer W_r, W_u, W_n (H,in) then R_r, R_u, R_n (H,H), row-major -- then all biases
#       -- per layer bW_r, bW_u, bW_n, bR_r, bR_u, bR_n.
#   * global step: step/global_step (int64); Adam: train/beta1_power, train/beta2_power; slots train/<var>/Adam(_1).

def _gru_scopes(cfg):
    """[(tf scope of a CudnnGRU layer, [canonical prefix per layer])] for the graph `cfg` builds"""
    L = cfg.get('rnn_layers', 3)
    out = []
    bidir, stacked = cfg.get('bidirectional', True), cfg.get('bidir_stacked', True)
    if bidir and stacked:
        for i in range(1, L + 1):
            for d in ('fwd', 'bwd'):
                out.append(('encode/rnn%d/%s' % (i, d), ['encode/rnn%d/%s/' % (i, d)]))
    elif bidir:
        for d in ('fwd', 'bwd'):
            out.append(('encode/rnn/%s' % d, ['encode/rnn/%s/l%d/' % (d, j) for j in range(L)]))
    else:
        out.append(('encode/rnn', ['encode/rnn/l%d/' % j for j in range(L)]))
    out.append(('decode/rnn', ['decode/rnn/l%d/' % j for j in range(L)]))
    return out


# The reference names its layers (CudnnGRU(..., name='fwd'), src/model.py:15,118-131), so the layer scope is the user's
# name and the saveable most likely writes <scope>/rnn/multi_rnn_cell/...; 'cudnn_gru' is only the DEFAULT layer name and
# would appear as an extra scope level if the saveable prefixes it regardless.  No TF-written file exists to settle it:
# the reader accepts both forms (whichever is present), the writer emits the first.  UNVERIFIED either way.
_CELL = 'rnn/multi_rnn_cell/cell_%d/cudnn_compatible_gru_cell/'
_CELL_FORMS = (_CELL, 'cudnn_gru/' + _CELL)


def canonical_to_tf(params, cfg, step=0, adam=None):
    """{canonical name: array} (+ optional {name: (m, v)} Adam slots, step) -> {TF variable name: array}"""
    H = cfg.get('dim_emb', 512)
    out, done = {}, set()
    for scope, layers in _gru_scopes(cfg):
        blobs = {'p': [[], []], 'm': [[], []], 'v': [[], []]}
        for l, pre in enumerate(layers):
            W, R, bW, bR = (np.asarray(params[pre + k], np.float32) for k in ('W', 'R', 'bW', 'bR'))
            c = scope + '/' + _CELL % l
            out[c + 'gates/kernel'] = np.concatenate([W[:2 * H].T, R[:2 * H].T], 0)
            out[c + 'gates/bias'] = bW[:2 * H] + bR[:2 * H]
            out[c + 'candidate/input_projection/kernel'] = np.ascontiguousarray(W[2 * H:].T)
            out[c + 'candidate/input_projection/bias'] = bW[2 * H:].copy()
            out[c + 'candidate/hidden_projection/kernel'] = np.ascontiguousarray(R[2 * H:].T)
            out[c + 'candidate/hidden_projection/bias'] = bR[2 * H:].copy()
            done.update(pre + k for k in ('W', 'R', 'bW', 'bR'))
            if adam is not None:
                for key, idx in (('m', 0), ('v', 1)):
                    blobs[key][0] += [adam[pre + 'W'][idx].ravel(), adam[pre + 'R'][idx].ravel()]
                    blobs[key][1] += [adam[pre + 'bW'][idx].ravel(), adam[pre + 'bR'][idx].ravel()]
        if adam is not None:
            out['train/' + scope + '/opaque_kernel/Adam'] = np.concatenate(blobs['m'][0] + blobs['m'][1]).astype(np.float32)
            out['train/' + scope + '/opaque_kernel/Adam_1'] = np.concatenate(blobs['v'][0] + blobs['v'][1]).astype(np.float32)
    for k, v in params.items():
        if k in done:
            continue
        out[k] = np.asarray(v, np.float32)
        if adam is not None:
            out['train/' + k + '/Adam'] = np.asarray(adam[k][0], np.float32)
            out['train/' + k + '/Adam_1'] = np.asarray(adam[k][1], np.float32)
    out['step/global_step'] = np.asarray(step, np.int64)
    if adam is not None:
        out['train/beta1_power'] = np.asarray(0.9 ** (int(step) + 1), np.float32)     # beta^(updates + 1): initialised to beta
        out['train/beta2_power'] = np.asarray(0.999 ** (int(step) + 1), np.float32)
    return out


def _split_blob(blob, layers, shapes):
    """cuDNN GRU parameter blob -> {canonical name: array}; weights of all layers first, then biases (assumption)."""
    out, pos = {}, 0
    for pre in layers:
        for k in ('W', 'R'):
            n = int(np.prod(shapes[pre + k]))
            out[pre + k] = blob[pos:pos + n].reshape(shapes[pre + k])
            pos += n
    for pre in layers:
        for k in ('bW', 'bR'):
            n = int(np.prod(shapes[pre + k]))
            out[pre + k] = blob[pos:pos + n].reshape(shapes[pre + k])
            pos += n
    if pos != blob.size:
        raise ValueError('opaque cuDNN blob has %d values, the model expects %d' % (blob.size, pos))
    return out


def tf_to_canonical(tensors, cfg, shapes):
    """{TF variable name: array} -> (params, adam {name: [m, v]} or None, step, unplaced names).  `shapes` =
    {canonical name: shape} of the model the checkpoint is loaded into."""
    H = cfg.get('dim_emb', 512)
    params, adam, used = {}, {}, set()

    def take(name):
        used.add(name)
        return np.asarray(tensors[name], np.float32)

    for scope, layers in _gru_scopes(cfg):
        opaque = scope + '/opaque_kernel'
        form = next((f for f in _CELL_FORMS if scope + '/' + f % 0 + 'gates/kernel' in tensors), None)
        if form is not None:
            for l, pre in enumerate(layers):
                c = scope + '/' + form % l
                gk, gb = take(c + 'gates/kernel'), take(c + 'gates/bias')
                cin = gk.shape[0] - H
                params[pre + 'W'] = np.concatenate([gk[:cin].T, take(c + 'candidate/input_projection/kernel').T], 0)
                params[pre + 'R'] = np.concatenate([gk[cin:].T, take(c + 'candidate/hidden_projection/kernel').T], 0)
                # the canonical form holds only the SUM of the two r/u biases: all of it goes to bW (same function)
                params[pre + 'bW'] = np.concatenate([gb, take(c + 'candidate/input_projection/bias')])
                params[pre + 'bR'] = np.concatenate([np.zeros(2 * H, np.float32), take(c + 'candidate/hidden_projection/bias')])
        elif opaque in tensors:
            params.update(_split_blob(take(opaque).ravel(), layers, shapes))
        for sfx, idx in (('/Adam', 0), ('/Adam_1', 1)):
            for cand in ('train/' + opaque + sfx, opaque + sfx):
                if cand in tensors:
                    for k, v in _split_blob(take(cand).ravel(), layers, shapes).items():
                        adam.setdefault(k, [None, None])[idx] = v
                    break
    for name, shp in shapes.items():
        if name in params:
            continue
        if name in tensors:
            params[name] = take(name).reshape(shp)
        for sfx, idx in (('/Adam', 0), ('/Adam_1', 1)):
            for cand in ('train/' + name + sfx, name + sfx):
                if cand in tensors:
                    adam.setdefault(name, [None, None])[idx] = take(cand).reshape(shp)
                    break
    step = None
    for cand in ('step/global_step', 'global_step'):
        if cand in tensors:
            step = int(np.asarray(tensors[cand]).ravel()[0])
            used.add(cand)
            break
    used.update(k for k in tensors if k.endswith('beta1_power') or k.endswith('beta2_power'))
    complete = adam and all(k in adam and adam[k][0] is not None and adam[k][1] is not None for k in shapes)
    return params, (adam if complete else None), step, sorted(set(tensors) - used)


def load_reference_checkpoint(handle, prefix, cfg, strict=True):
    """restores a checkpoint the reference's `saver.save` wrote (src/train.py:121) into an argsim_b200 handle:
    parameters, Adam slots when all are present, and the global step.  Returns the list of tensors it did not use."""
    shapes = handle.param_shapes()
    params, adam, step, unused = tf_to_canonical(load_bundle(prefix), cfg, shapes)
    missing = [k for k in shapes if k not in params]
    if missing and strict:
        raise ValueError('checkpoint %s lacks %s (tensors not placed: %s)' % (prefix, missing[:4], unused[:6]))
    for k, v in params.items():
        if tuple(v.shape) != tuple(shapes[k]):
            raise ValueError('%s: checkpoint shape %r, model %r' % (k, v.shape, shapes[k]))
        handle.set_param(k, v)
    if adam is not None:
        for k, (m, v) in adam.items():
            handle.set_opt_state(k, m, v)
    if step is not None:
        handle.step = step
    return unused


def save_reference_checkpoint(handle, prefix, cfg):
    """writes the handle's state under the reference's variable names (see the assumptions above)."""
    params = handle.get_params()
    adam = {k: handle.get_opt_state(k) for k in params}
    save_bundle(prefix, canonical_to_tf(params, cfg, step=handle.step, adam=adam))
