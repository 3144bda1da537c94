"""T5X checkpoint layout (msgpack index + zarr v2 arrays) read without t5x/tensorstore.

No real checkpoint exists offline (parity unpinned, see t5x_checkpoint.py); these tests pin the
reader against hand-built fixtures of the published layout and against its own writer."""
import gzip
import json
import os

import msgpack
import numpy as np
import pytest

from music_spectrogram_diffusion_b200 import config, t5x_checkpoint as tc, weights


def test_zarr_round_trip_edge_chunks(tmp_path):
  rng = np.random.default_rng(0)
  a = rng.standard_normal((37, 10)).astype(np.float32)
  tc.write_zarr_array(str(tmp_path / 'a'), a, chunks=(16, 4))   # ragged edge chunks on both axes
  assert sorted(os.listdir(tmp_path / 'a'))[:3] == ['.zarray', '0.0', '0.1']
  np.testing.assert_array_equal(tc.read_zarr_array(str(tmp_path / 'a')), a)
  v = rng.standard_normal(5).astype(np.float32)
  tc.write_zarr_array(str(tmp_path / 'v'), v)
  np.testing.assert_array_equal(tc.read_zarr_array(str(tmp_path / 'v')), v)


def test_zarr_hand_built_fixture(tmp_path):
  """Bytes laid out by hand per the zarr v2 spec: '/' separator, zlib, missing chunk = fill."""
  d = tmp_path / 'arr'
  os.makedirs(d / '0')
  meta = {'zarr_format': 2, 'shape': [2, 4], 'chunks': [2, 2], 'dtype': '<i4', 'order': 'C',
          'compressor': {'id': 'zlib', 'level': 1}, 'fill_value': 7, 'filters': None,
          'dimension_separator': '/'}
  (d / '.zarray').write_text(json.dumps(meta))
  import zlib
  (d / '0' / '0').write_bytes(zlib.compress(np.array([[1, 2], [3, 4]], '<i4').tobytes()))
  got = tc.read_zarr_array(str(d))
  np.testing.assert_array_equal(got, [[1, 2, 7, 7], [3, 4, 7, 7]])


def test_zarr_bfloat16_and_fortran_order(tmp_path):
  d = tmp_path / 'b'
  os.makedirs(d)
  vals = np.array([[1.0, -2.5, 0.15625], [3.0, 65536.0, -0.0]], np.float32)
  bits = (vals.view(np.uint32) >> 16).astype('<u2')
  meta = {'zarr_format': 2, 'shape': [2, 3], 'chunks': [2, 3], 'dtype': 'bfloat16', 'order': 'F',
          'compressor': None, 'fill_value': None, 'filters': None}
  (d / '.zarray').write_text(json.dumps(meta))
  (d / '0.0').write_bytes(np.asfortranarray(bits).tobytes(order='F'))
  np.testing.assert_array_equal(tc.read_zarr_array(str(d)), vals)


def test_unsupported_compressor_is_loud(tmp_path):
  d = tmp_path / 'c'
  os.makedirs(d)
  meta = {'zarr_format': 2, 'shape': [2], 'chunks': [2], 'dtype': '<f4', 'order': 'C',
          'compressor': {'id': 'blosc'}, 'fill_value': None, 'filters': None}
  (d / '.zarray').write_text(json.dumps(meta))
  (d / '0').write_bytes(b'xxxx')
  with pytest.raises(tc.CheckpointError, match='blosc'):
    tc.read_zarr_array(str(d))


def test_checkpoint_round_trip_tiny_model(tmp_path):
  t5 = config.t5_tiny()
  params = weights.synthetic_params(t5, 128, 128, 128, 128, seed=3)
  path = tc.save_t5x_checkpoint(str(tmp_path / 'checkpoint_7'), params, step=7, inline_below=200,
                                chunk_rows=48)
  assert os.path.isfile(os.path.join(path, 'checkpoint'))
  assert os.path.isdir(os.path.join(path, 'target.decoder.spec_out_dense.kernel'))
  for p in (path, os.path.join(path, 'checkpoint')):      # directory or the msgpack file inside it
    got = tc.load_t5x_checkpoint(p)
    assert set(got) == set(params)
    for k in params:
      np.testing.assert_array_equal(got[k], params[k], err_msg=k)
  weights.check_params(got, t5, 128, 128, 128, 128)


def test_msgpack_index_hand_built(tmp_path):
  """Index written with plain msgpack calls (flax ext type 1, chunked array dict, TensorStore
  spec with an absolute training-time path, legacy top-level 'target')."""
  ck = tmp_path / 'checkpoint_1'
  os.makedirs(ck)
  big = np.arange(12, dtype=np.float32).reshape(3, 4)
  small = np.array([1.5, 2.5], np.float32)
  chunked = np.arange(6, dtype=np.float32)

  def nd(a):
    return msgpack.ExtType(1, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))

  tc.write_zarr_array(str(ck / 'target.decoder.big.kernel'), big)
  tree = {'target': {'decoder': {
      'big': {'kernel': {'driver': 'zarr', 'kvstore': {'driver': 'gfile',
                                                       'path': '/cns/train/run1/checkpoint_1.tmp-123/target.decoder.big.kernel'},
                         'metadata': {'shape': [3, 4], 'chunks': [3, 4], 'compressor': {'id': 'gzip'}}}},
      'norm': {'scale': nd(small)},
      'chunky': {'__msgpack_chunked_array__': True, 'shape': [2, 3],
                 'chunks': {'0': nd(chunked[:4]), '1': nd(chunked[4:])}}}},
          'state': {'step': 1}}
  (ck / 'checkpoint').write_bytes(msgpack.packb(tree, use_bin_type=True))
  got = tc.load_t5x_checkpoint(str(ck))
  np.testing.assert_array_equal(got['decoder/big/kernel'], big)
  np.testing.assert_array_equal(got['decoder/norm/scale'], small)
  np.testing.assert_array_equal(got['decoder/chunky'], chunked.reshape(2, 3))


def test_missing_array_and_shape_mismatch_are_loud(tmp_path):
  t5 = config.t5_tiny()
  params = weights.synthetic_params(t5, 128, 128, 128, 128, seed=1)
  path = tc.save_t5x_checkpoint(str(tmp_path / 'ck'), params)
  import shutil
  shutil.rmtree(os.path.join(path, 'target.decoder.decoder_norm.scale'))
  with pytest.raises(tc.CheckpointError, match='decoder_norm'):
    tc.load_t5x_checkpoint(path)
  bad = dict(params)
  bad['decoder/decoder_norm/scale'] = np.zeros(3, np.float32)
  del bad['decoder/spec_out_dense/kernel']
  with pytest.raises(ValueError, match='2 problems'):
    weights.check_params(bad, t5, 128, 128, 128, 128)
  with pytest.raises(tc.CheckpointError):
    tc.load_t5x_checkpoint(str(tmp_path / 'nope'))


def test_gcs_style_spec_and_dict_shaped_chunks(tmp_path):
  """Checkpoints written to GCS (the published ones) carry kvstore = {'driver': 'gcs', 'bucket':
  ...} plus a top-level 'path' instead of kvstore.path; flax writes a chunked array's shape as
  {'0': n, '1': m}.  Both forms restore; a spec without any usable path falls back to the
  conventional 'target.<name>' directory; a missing directory is a CheckpointError, not a KeyError."""
  import msgpack
  from music_spectrogram_diffusion_b200 import t5x_checkpoint as X
  params = {'decoder/decoder_norm/scale': np.arange(8, dtype=np.float32),
            'decoder/spec_out_dense/kernel': np.arange(32, dtype=np.float32).reshape(8, 4)}
  ck = X.save_t5x_checkpoint(str(tmp_path / 'checkpoint_7'), params, step=7)
  idx = os.path.join(ck, 'checkpoint')
  state = msgpack.unpackb(open(idx, 'rb').read(), ext_hook=X._ext_hook, raw=False, strict_map_key=False)

  def rewrite(node):
    for k, v in list(node.items()):
      if isinstance(v, dict) and v.get('driver') == 'zarr':
        path = v['kvstore']['path']
        v['kvstore'] = {'driver': 'gcs', 'bucket': 't5x-dummy-bucket'}
        if 'scale' in k:
          v['path'] = path          # gcs form: top-level path
        # the kernel keeps no path at all -> 'target.<name>' fallback
      elif isinstance(v, dict):
        rewrite(v)
  rewrite(state)
  # an inline chunked array with flax's dict-shaped 'shape'
  tgt = X._target_tree(state)
  tgt['extra'] = {X._CHUNK_MARK: True, 'shape': {'0': 2, '1': 3},
                  'chunks': {'0': np.arange(4, dtype=np.float32), '1': np.arange(4, 6, dtype=np.float32)}}
  open(idx, 'wb').write(msgpack.packb(state, default=X._default, use_bin_type=True))
  got = X.load_t5x_checkpoint(ck)
  for k, v in params.items():
    np.testing.assert_array_equal(got[k], v)
  np.testing.assert_array_equal(got['extra'], np.arange(6, dtype=np.float32).reshape(2, 3))
  # remove an array directory: loud CheckpointError
  import shutil
  shutil.rmtree(os.path.join(ck, 'target.decoder.spec_out_dense.kernel'))
  with pytest.raises(X.CheckpointError, match='not found'):
    X.load_t5x_checkpoint(ck)
