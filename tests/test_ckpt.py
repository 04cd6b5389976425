"""CPU: TensorFlow checkpoint V2 reader (a-tvsnet_b200/ckpt.py).  No TF-written file exists offline, so
the reader is exercised on files produced by the module's own writer (prefix-compressed multi-block tables) and its leaf
arithmetic on published known answers (CRC-32C check value, leveldb CRC mask, varints, protobuf fields)."""
import importlib.util
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('atvs_ckpt', os.path.join(ROOT, 'a-tvsnet_b200', 'ckpt.py'))
K = importlib.util.module_from_spec(spec)
spec.loader.exec_module(K)


def test_crc32c_and_varint_known_answers():
    assert K.crc32c(b'123456789') == 0xE3069283                      # the standard CRC-32C check value
    assert K.crc32c(b'\x00' * 32) == 0x8A9136AA                       # RFC 3720 B.4 test vector
    assert K.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert K.mask_crc(0) == 0xa282ead8
    assert K._put_varint(300) == b'\xac\x02' and K._get_varint(b'\xac\x02', 0) == (300, 2)
    assert K._get_varint(K._put_varint(2 ** 40 + 5), 0)[0] == 2 ** 40 + 5
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size:3} dim{size:4}}, offset: 48, size: 48, crc32c: fixed32}
    e = bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x04, 0x20, 0x30, 0x28, 0x30, 0x35]) + struct.pack('<I', 7)
    p = K._parse_entry(e)
    assert p['dtype'] == 1 and p['shape'] == (3, 4) and p['offset'] == 48 and p['size'] == 48 and p['crc32c'] == 7


def test_round_trip_with_checkpoint_variable_names(tmp_path):
    import sys
    sys.path.insert(0, ROOT)
    import atvsnet_b200 as A
    w = A.variables.synthetic_weights(seed=3)
    w.update(A.variables.synthetic_refine_weights())
    w['global_step'] = np.array(150000, np.int64)
    w['flag'] = np.array([True, False])
    w['half'] = np.arange(6, dtype=np.float16).reshape(2, 3)
    prefix = str(tmp_path / 'model.ckpt')
    K.write_checkpoint(prefix, w)
    header, entries = K.read_index(prefix)
    assert header['num_shards'] == 1 and set(entries) == set(w)
    assert entries['conv_b0_1_0/conv3d/kernel']['shape'] == (3, 3, 3, 64, 16)
    back = K.read_checkpoint(prefix)
    assert set(back) == set(w)
    for k in w:
        assert back[k].dtype == np.asarray(w[k]).dtype and np.array_equal(back[k], w[k]), k
    # the package-level restore (variables.load_checkpoint) skips non-float entries and optimizer slots
    w2 = dict(w)
    w2['conv_b0_1_0/conv3d/kernel/Adam'] = np.zeros((3, 3, 3, 64, 16), np.float32)
    K.write_checkpoint(prefix + '2', w2)
    names = A.variables.load_checkpoint(prefix + '2', device='cpu', parts=('crm',))
    assert 'global_step' not in names and 'flag' not in names and 'conv_b0_1_0/conv3d/kernel/Adam' not in names
    assert np.array_equal(A.variables.get_variable('conv_b2_6_2/kernel').numpy(), w['conv_b2_6_2/kernel'])
    only = K.read_checkpoint(prefix, names=lambda n: n.startswith('attention_aggregate/'))
    assert sorted(only) == ['attention_aggregate/attention_activation/weight_shared',
                            'attention_aggregate/attention_activation/weight_unique']


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'c.ckpt')
    K.write_checkpoint(prefix, {'a/kernel': np.arange(12, dtype=np.float32).reshape(3, 4), 'a/bias': np.ones(4, np.float32)})
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[5] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='CRC'):
        K.read_checkpoint(prefix)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x01
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError, match='checksum'):
        K.read_index(prefix)
    open(prefix + '.index', 'wb').write(b'not a table' * 10)
    with pytest.raises(ValueError, match='magic'):
        K.read_index(prefix)


def test_load_checkpoint_validates_the_variable_set_and_large_tensor_crc(tmp_path):
    """a partial / mis-shaped checkpoint is refused with the list of offending names, and a flipped bit in a LARGE
    tensor (beyond what the pure-Python CRC covers) is caught through libatvs.so's host CRC-32C."""
    import __graft_entry__ as ge
    ge.build()
    import atvsnet_b200 as A
    w = A.variables.synthetic_weights(seed=3)
    prefix = str(tmp_path / 'm.ckpt')
    K.write_checkpoint(prefix, w)
    with pytest.raises(ValueError, match='missing'):
        A.variables.load_checkpoint(prefix, device='cpu')                       # FEM / refinement variables absent
    assert 'conv_b2_6_2/kernel' in A.variables.load_checkpoint(prefix, device='cpu', parts=('crm',))
    bad = dict(w)
    bad['conv_b1_2_1/conv3d/kernel'] = np.zeros((3, 3, 3, 32, 16), np.float32)
    del bad['attention_prob_vol/kernel']
    K.write_checkpoint(prefix + 'b', bad)
    with pytest.raises(ValueError) as ei:
        A.variables.load_checkpoint(prefix + 'b', device='cpu', parts=('crm',))
    assert 'attention_prob_vol/kernel' in str(ei.value) and 'conv_b1_2_1/conv3d/kernel' in str(ei.value)
    # corruption inside a 442 KB kernel
    header, entries = K.read_index(prefix)
    e = entries['conv_b0_3_1/conv3d/kernel']
    assert e['size'] > (1 << 16)
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[e['offset'] + e['size'] // 2] ^= 0x10
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='CRC'):
        A.ckpt.read_checkpoint(prefix)           # the packaged module reaches libatvs.so's host CRC
    lib = A._lib.load()
    assert lib.atvs_crc32c(b'123456789', 9, 0) == 0xE3069283                     # the published CRC-32C check value
