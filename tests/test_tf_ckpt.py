"""TF-1 checkpoint container (argsim_b200/tf_ckpt.py): known answers for the checksum, the table
and bundle formats round-trip, the snappy decoder on a hand-built stream, and the variable-name / cuDNN-layout mapping
for every encoder branch.  No TensorFlow here: what a real tf.train.Saver file looks like is restated, not observed."""
import numpy as np
import pytest

from argsim_b200 import tf_ckpt as T
from oracle import vae_oracle as O


def test_crc32c_known_answers_and_combine():
    assert T.crc32c(b'123456789') == 0xE3069283                       # the standard check value of CRC-32C
    assert T.crc32c(b'\x00' * 32) == 0x8A9136AA                       # RFC 3720 B.4
    assert T.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert T.unmask_crc(T.mask_crc(0x12345678)) == 0x12345678
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, 1000, np.uint8).tobytes(), rng.integers(0, 256, 777, np.uint8).tobytes()
    assert T.crc32c_combine(T.crc32c(a), T.crc32c(b), len(b)) == T.crc32c(a + b)
    big = rng.integers(0, 256, 8192 * 5 + 123, np.uint8).tobytes()
    assert T._crc32c_np(big) == T.crc32c(big)


def test_varint_and_snappy():
    for n in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 40 + 5):
        assert T.get_varint(T.put_varint(n), 0) == (n, len(T.put_varint(n)))
    # "abcdabcdabcdX": literal "abcd", copy(offset 4, length 8), literal "X"
    stream = T.put_varint(13) + bytes([3 << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([0 << 2]) + b'X'
    assert T.snappy_decompress(stream) == b'abcdabcdabcdX'
    with pytest.raises(ValueError):
        T.snappy_decompress(T.put_varint(5) + bytes([((4 - 4) << 2) | 1, 9]))


def test_table_and_bundle_roundtrip(tmp_path):
    pairs = [(('key%05d' % i).encode(), ('value-%d' % (i * i)).encode() * (1 + i % 7)) for i in range(1500)]
    T.write_table(str(tmp_path / 't.index'), pairs, block_bytes=512)
    assert T.read_table(str(tmp_path / 't.index')) == sorted(pairs)
    raw = bytearray(open(tmp_path / 't.index', 'rb').read())
    raw[100] ^= 0x40
    open(tmp_path / 'bad.index', 'wb').write(raw)
    with pytest.raises(ValueError):
        T.read_table(str(tmp_path / 'bad.index'))
    rng = np.random.default_rng(1)
    tensors = {'embed/embedding': rng.standard_normal((50, 8)).astype(np.float32), 'step/global_step': np.asarray(12345, np.int64),
               'a/b/bias': np.zeros(0, np.float32), 'train/beta1_power': np.asarray(0.5, np.float32),
               'ints': rng.integers(-5, 5, (3, 4, 5)).astype(np.int32)}
    T.save_bundle(str(tmp_path / 'ck'), tensors)
    back = T.load_bundle(str(tmp_path / 'ck'), verify_tensors=True)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape
        np.testing.assert_array_equal(back[k], tensors[k])


@pytest.mark.parametrize('extra', [dict(), dict(bidirectional=True, bidir_stacked=False), dict(bidirectional=False),
                                   dict(logit_use_embed=False), dict(attentive=True)])
def test_name_and_layout_mapping_roundtrip(tmp_path, extra):
    cfg = dict(dim_tgt=40, dim_emb=8, dim_rep=16, rnn_layers=2, accelerate=1e-4, learn_rate=1e-3, bos=2, eos=1, **extra)
    P = O.init_params(cfg, seed=3, dtype=np.float32, bias_scale=0.2)
    shapes = {k: v.shape for k, v in P.items()}
    rng = np.random.default_rng(4)
    adam = {k: (rng.standard_normal(v.shape).astype(np.float32), rng.random(v.shape).astype(np.float32)) for k, v in P.items()}
    tf_vars = T.canonical_to_tf(P, cfg, step=777, adam=adam)
    H = cfg['dim_emb']
    assert tf_vars['step/global_step'] == 777 and 'embed/embedding' in tf_vars and 'train/embed/embedding/Adam_1' in tf_vars
    gk = [k for k in tf_vars if k.endswith('cell_0/cudnn_compatible_gru_cell/gates/kernel') and k.startswith('decode/rnn/')]
    assert len(gk) == 1 and tf_vars[gk[0]].shape == (cfg['dim_emb'] + H, 2 * H)
    T.save_bundle(str(tmp_path / 'ref'), tf_vars)
    params, adam2, step, unused = T.tf_to_canonical(T.load_bundle(str(tmp_path / 'ref')), cfg, shapes)
    assert step == 777 and unused == [] and set(params) == set(P) and adam2 is not None
    for k in P:
        if k.endswith('/bW') or k.endswith('/bR'):
            continue
        np.testing.assert_array_equal(params[k], P[k])
        np.testing.assert_array_equal(adam2[k][0], adam[k][0])
        np.testing.assert_array_equal(adam2[k][1], adam[k][1])
    # the canonical cuDNN form only keeps bW + bR of the r / u gates: the FUNCTION is what must survive the trip
    src = np.array([[5, 6, 7, 1], [8, 9, 1, 1], [3, 4, 5, 6]], np.int32)
    a, _ = O.forward(P, cfg, src, src, 'valid')
    b, _ = O.forward({k: v.astype(np.float32) for k, v in params.items()}, cfg, src, src, 'valid')
    assert abs(a['loss'] - b['loss']) < 1e-5 * abs(a['loss'])
    # both spellings of the canonical cell scope are read: '<scope>/rnn/multi_rnn_cell/...' (written) and
    # '<scope>/cudnn_gru/rnn/multi_rnn_cell/...' (the default layer name as an extra level)
    assert not any('/cudnn_gru/' in k for k in tf_vars)
    alt = {k.replace('/rnn/multi_rnn_cell/', '/cudnn_gru/rnn/multi_rnn_cell/'): v for k, v in tf_vars.items()}
    assert any('/cudnn_gru/' in k for k in alt)
    params_alt, adam_alt, step_alt, unused_alt = T.tf_to_canonical(alt, cfg, shapes)
    assert step_alt == 777 and unused_alt == [] and adam_alt is not None
    for k in P:
        np.testing.assert_array_equal(params_alt[k], params[k])
    # a checkpoint that only has the opaque blobs (weights in cuDNN order) is placed too
    blob_only = {k: v for k, v in tf_vars.items() if 'cudnn_compatible_gru_cell' not in k}
    for scope, layers in T._gru_scopes(cfg):
        ws = [P[pre + k].ravel() for pre in layers for k in ('W', 'R')] + [P[pre + k].ravel() for pre in layers for k in ('bW', 'bR')]
        blob_only[scope + '/opaque_kernel'] = np.concatenate(ws)
    params3, _, _, unused3 = T.tf_to_canonical(blob_only, cfg, shapes)
    assert unused3 == []
    for k in P:
        np.testing.assert_array_equal(params3[k], P[k])
