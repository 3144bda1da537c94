"""Reader (and test-fixture writer) for T5X checkpoints, without t5x / flax / tensorstore.

What the reference does: `InferenceModel.__init__` restores `checkpoint_path` through
`t5x.utils.RestoreCheckpointConfig(path=..., mode='specific')` + `TrainStateInitializer`
(msd/inference.py:95-111, 171-181); the colab passes `.../base_with_context/checkpoint_500000`
(ipynb:203-229).  The format itself lives in third-party code (t5x @ 2e05ad4 `checkpoints.py`,
flax `serialization.py`, tensorstore's zarr driver), not in /root/reference, so this module
restates the published layout:

  <checkpoint_dir>/checkpoint          msgpack of {'version': 3, 'optimizer': {'target': tree,
                                       'state': ...}} (older: {'target': tree, ...}).  Leaves
                                       are either inline arrays (flax ext type 1 = packed
                                       (shape, dtype, bytes); large ones as a
                                       '__msgpack_chunked_array__' dict) or a TensorStore spec
                                       {'driver': 'zarr', 'kvstore': {'path': 'target.a.b.c'},
                                       'metadata': {'shape', 'chunks', 'compressor'}}.
  <checkpoint_dir>/target.a.b.c/       one zarr v2 array per large leaf: `.zarray` JSON + chunk
                                       files named by chunk index joined with '.', gzip
                                       compressed, C order.

The parameter names ('/'-joined tree path below `target`) are the flax names listed in
`weights.param_shapes`.  **Parity unpinned**: no real checkpoint is available offline; the
round trip against `save_t5x_checkpoint` and hand-built fixtures is what `tests/` pin.
"""

from __future__ import annotations

import gzip
import itertools
import json
import os
import zlib
from typing import Any, Dict, Iterable, Mapping, Optional, Tuple

import msgpack
import numpy as np

ParamDict = Dict[str, np.ndarray]

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3   # flax.serialization ext type codes
_CHUNK_MARK = '__msgpack_chunked_array__'


class CheckpointError(ValueError):
  pass


# ---------------------------------------------------------------------------------------------
# msgpack side (flax.serialization restated)
# ---------------------------------------------------------------------------------------------
def _dtype_from_name(name: str) -> Tuple[np.dtype, bool]:
  """numpy dtype for a flax / zarr dtype name; second value: payload is bfloat16 bits."""
  if name in ('bfloat16', '<V2', 'V2'):
    return np.dtype('<u2'), True
  try:
    return np.dtype(name), False
  except TypeError as e:
    raise CheckpointError(f'unsupported array dtype {name!r}') from e


def _bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
  return (bits.astype(np.uint32) << 16).view(np.float32)


def _ext_hook(code: int, data: bytes) -> Any:
  if code == _EXT_NDARRAY:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
    dt, is_bf16 = _dtype_from_name(dtype_name)
    arr = np.frombuffer(buf, dtype=dt).reshape(shape)
    return _bf16_bits_to_f32(arr) if is_bf16 else arr
  if code == _EXT_COMPLEX:
    re, im = msgpack.unpackb(data, raw=False)
    return complex(re, im)
  if code == _EXT_NPSCALAR:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
    dt, is_bf16 = _dtype_from_name(dtype_name)
    arr = np.frombuffer(buf, dtype=dt)
    arr = _bf16_bits_to_f32(arr) if is_bf16 else arr
    return arr.reshape(())[()]
  return msgpack.ExtType(code, data)


def _unchunk(tree: Any) -> Any:
  """Reassemble flax's chunked arrays ({'__msgpack_chunked_array__': True, 'shape', 'chunks'})."""
  if isinstance(tree, dict):
    if tree.get(_CHUNK_MARK):
      sh = tree['shape']
      # flax serialises the tuple as {'0': n, '1': m, ...} (serialization._tuple_to_dict); a
      # plain list is accepted as well
      shape = (tuple(int(sh[str(i)]) for i in range(len(sh))) if isinstance(sh, dict)
               else tuple(int(v) for v in sh))
      chunks = tree['chunks']
      parts = [np.asarray(chunks[str(i)]).reshape(-1) for i in range(len(chunks))]
      return np.concatenate(parts).reshape(shape)
    return {k: _unchunk(v) for k, v in tree.items()}
  return tree


def _pack_ndarray(a: np.ndarray) -> msgpack.ExtType:
  a = np.ascontiguousarray(a)
  return msgpack.ExtType(_EXT_NDARRAY,
                         msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))


def _default(o: Any) -> Any:
  if isinstance(o, np.ndarray):
    return _pack_ndarray(o)
  if isinstance(o, np.generic):
    a = np.asarray(o)
    return msgpack.ExtType(_EXT_NPSCALAR,
                           msgpack.packb(([], a.dtype.name, a.tobytes()), use_bin_type=True))
  raise TypeError(f'cannot serialise {type(o)}')


# ---------------------------------------------------------------------------------------------
# zarr v2 side (tensorstore's `zarr` driver as T5X configures it)
# ---------------------------------------------------------------------------------------------
def _decompress(raw: bytes, compressor: Optional[Mapping[str, Any]]) -> bytes:
  if compressor is None:
    return raw
  cid = compressor.get('id')
  if cid == 'gzip':
    return gzip.decompress(raw)
  if cid == 'zlib':
    return zlib.decompress(raw)
  raise CheckpointError(f'unsupported zarr compressor {cid!r} (gzip, zlib and none are readable)')


def read_zarr_array(path: str) -> np.ndarray:
  """One zarr v2 array directory -> numpy (bfloat16 payloads are widened to float32)."""
  meta_path = os.path.join(path, '.zarray')
  if not os.path.isfile(meta_path):
    raise CheckpointError(f'{path}: no .zarray metadata (not a zarr v2 array)')
  with open(meta_path) as f:
    meta = json.load(f)
  if meta.get('zarr_format', 2) != 2:
    raise CheckpointError(f'{path}: zarr_format {meta.get("zarr_format")} is not supported')
  if meta.get('filters'):
    raise CheckpointError(f'{path}: zarr filters are not supported')
  shape = tuple(int(x) for x in meta['shape'])
  chunks = tuple(int(x) for x in meta['chunks'])
  order = meta.get('order', 'C')
  sep = meta.get('dimension_separator', '.')
  dt, is_bf16 = _dtype_from_name(meta['dtype'])
  fill = meta.get('fill_value')
  out = np.empty(shape, dtype=dt)
  if fill is not None and not isinstance(fill, str):
    out[...] = fill
  else:
    out[...] = 0
  grid = [range((s + c - 1) // c) for s, c in zip(shape, chunks)] if shape else []
  n_chunk_elems = int(np.prod(chunks)) if chunks else 1
  for idx in itertools.product(*grid):
    key = sep.join(str(i) for i in idx) if idx else '0'
    cpath = os.path.join(path, key)
    if not os.path.isfile(cpath):
      continue  # missing chunk = fill value (zarr semantics)
    with open(cpath, 'rb') as f:
      buf = _decompress(f.read(), meta.get('compressor'))
    block = np.frombuffer(buf, dtype=dt)
    if block.size != n_chunk_elems:
      raise CheckpointError(f'{cpath}: {block.size} elements, expected {n_chunk_elems}')
    block = block.reshape(chunks, order=order) if chunks else block.reshape(())
    sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
    trim = tuple(slice(0, sl.stop - sl.start) for sl in sel)
    out[sel] = block[trim]
  return _bf16_bits_to_f32(out) if is_bf16 else out


def write_zarr_array(path: str, a: np.ndarray, chunks: Optional[Tuple[int, ...]] = None) -> None:
  """Writes `a` the way T5X's tensorstore spec does (gzip, C order, '.'-joined chunk keys)."""
  a = np.ascontiguousarray(a)
  chunks = tuple(chunks) if chunks is not None else tuple(max(1, s) for s in a.shape)
  os.makedirs(path, exist_ok=True)
  meta = {'chunks': list(chunks), 'compressor': {'id': 'gzip', 'level': 1}, 'dtype': a.dtype.str,
          'fill_value': None, 'filters': None, 'order': 'C', 'shape': list(a.shape),
          'zarr_format': 2}
  with open(os.path.join(path, '.zarray'), 'w') as f:
    json.dump(meta, f)
  grid = [range((s + c - 1) // c) for s, c in zip(a.shape, chunks)]
  for idx in itertools.product(*grid):
    block = np.zeros(chunks, dtype=a.dtype)
    sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, a.shape))
    trim = tuple(slice(0, sl.stop - sl.start) for sl in sel)
    block[trim] = a[sel]
    key = '.'.join(str(i) for i in idx) if idx else '0'
    with open(os.path.join(path, key), 'wb') as f:
      f.write(gzip.compress(block.tobytes(), compresslevel=1))


# ---------------------------------------------------------------------------------------------
# checkpoint level
# ---------------------------------------------------------------------------------------------
def _is_ts_spec(leaf: Any) -> bool:
  return isinstance(leaf, dict) and 'kvstore' in leaf and 'driver' in leaf


def _flatten(tree: Mapping[str, Any], prefix: str = '') -> Iterable[Tuple[str, Any]]:
  for k, v in tree.items():
    name = f'{prefix}/{k}' if prefix else str(k)
    if isinstance(v, dict) and not _is_ts_spec(v):
      yield from _flatten(v, name)
    else:
      yield name, v


def _target_tree(state: Mapping[str, Any]) -> Mapping[str, Any]:
  if 'optimizer' in state and isinstance(state['optimizer'], dict) and 'target' in state['optimizer']:
    return state['optimizer']['target']
  if 'target' in state:
    return state['target']
  raise CheckpointError(f'no parameter tree ("optimizer/target" or "target") in checkpoint; '
                        f'top-level keys: {sorted(state)}')


def resolve_checkpoint_dir(path: str) -> str:
  """Accepts the checkpoint directory or the msgpack file inside it."""
  if os.path.isdir(path):
    if not os.path.isfile(os.path.join(path, 'checkpoint')):
      raise CheckpointError(f'{path}: directory has no "checkpoint" msgpack file')
    return path
  if os.path.isfile(path) and os.path.basename(path) == 'checkpoint':
    return os.path.dirname(path) or '.'
  raise CheckpointError(f'{path}: not a T5X checkpoint directory')


def is_t5x_checkpoint(path: str) -> bool:
  try:
    resolve_checkpoint_dir(path)
    return True
  except CheckpointError:
    return False


def load_t5x_checkpoint(path: str, dtype=np.float32, threads: int = 8) -> ParamDict:
  """Flat {'decoder/layers_0/self_attention/query/kernel': array, ...} of the `target` tree.
  The zarr arrays are read by `threads` workers (gzip releases the GIL): base_with_context is
  1.6 GB of float32 in ~600 arrays."""
  ckpt_dir = resolve_checkpoint_dir(path)
  with open(os.path.join(ckpt_dir, 'checkpoint'), 'rb') as f:
    state = msgpack.unpackb(f.read(), ext_hook=_ext_hook, raw=False, strict_map_key=False)
  state = _unchunk(state)
  leaves = list(_flatten(_target_tree(state)))

  def read(item):
    name, leaf = item
    return name, _read_leaf(ckpt_dir, name, leaf, dtype)

  if threads > 1 and len(leaves) > 1:
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as pool:
      return dict(pool.map(read, leaves))
  return dict(map(read, leaves))


def _read_leaf(ckpt_dir: str, name: str, leaf: Any, dtype) -> np.ndarray:
  if _is_ts_spec(leaf):
    if leaf.get('driver') != 'zarr':
      raise CheckpointError(f'{name}: TensorStore driver {leaf.get("driver")!r} is not supported')
    # Local-file checkpoints keep the array directory in kvstore.path; checkpoints written to GCS
    # (e.g. the published base_with_context/checkpoint_500000) carry kvstore = {'driver': 'gcs',
    # 'bucket': ...} and a top-level 'path'.  Either may be absent or hold the training job's
    # absolute path, so the basename and the conventional 'target.<name>' directory are tried too.
    kv = leaf.get('kvstore')
    rel = (kv.get('path') if isinstance(kv, dict) else (str(kv) if kv else None)) or leaf.get('path') or ''
    rel = str(rel)
    cand = [os.path.join(ckpt_dir, 'target.' + name.replace('/', '.'))]
    if rel:
      cand = [os.path.join(ckpt_dir, rel), os.path.join(ckpt_dir, os.path.basename(rel.rstrip('/')))] + cand
    apath = next((c for c in cand if os.path.isdir(c)), None)
    if apath is None:
      raise CheckpointError(f'{name}: array directory {rel!r} not found under {ckpt_dir}')
    arr = read_zarr_array(apath)
    want = leaf.get('metadata', {}).get('shape')
    if want is not None and tuple(want) != arr.shape:
      raise CheckpointError(f'{name}: zarr shape {arr.shape} != spec shape {tuple(want)}')
  elif isinstance(leaf, np.ndarray):
    arr = leaf
  elif isinstance(leaf, (int, float, np.generic)):
    arr = np.asarray(leaf)
  else:
    raise CheckpointError(f'{name}: unexpected leaf of type {type(leaf).__name__}')
  return np.ascontiguousarray(arr, dtype=dtype)


def save_t5x_checkpoint(path: str, params: Mapping[str, np.ndarray], step: int = 0,
                        inline_below: int = 0, chunk_rows: Optional[int] = None) -> str:
  """Writes `params` in the layout above (fixture generator; also lets `.npz` / synthetic trees
  be exported for tools that expect a T5X directory).  Arrays with fewer than `inline_below`
  elements are stored inline in the msgpack; `chunk_rows` splits the first axis into chunks."""
  os.makedirs(path, exist_ok=True)
  tree: Dict[str, Any] = {}
  for name, a in params.items():
    a = np.asarray(a)
    node = tree
    parts = name.split('/')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    if a.size < inline_below:
      node[parts[-1]] = a
      continue
    dirname = 'target.' + name.replace('/', '.')
    chunks = tuple(a.shape)
    if chunk_rows and a.ndim >= 1:
      chunks = (min(chunk_rows, a.shape[0]),) + tuple(a.shape[1:])
    write_zarr_array(os.path.join(path, dirname), a, chunks)
    node[parts[-1]] = {'driver': 'zarr', 'dtype': a.dtype.name,
                       'kvstore': {'driver': 'file', 'path': dirname},
                       'metadata': {'chunks': list(chunks), 'compressor': {'id': 'gzip'},
                                    'shape': list(a.shape)}}
  state = {'version': 3, 'optimizer': {'target': tree, 'state': {'step': np.int32(step)}}}
  with open(os.path.join(path, 'checkpoint'), 'wb') as f:
    f.write(msgpack.packb(state, default=_default, use_bin_type=True))
  return path
