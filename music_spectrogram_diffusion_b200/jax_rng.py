"""numpy restatement of the `jax.random` stream the reference draws its noise from.

Call sites in the reference: `inference.py:203` (`jax.random.PRNGKey(seed)`),
`diffusion_utils.py:462` (`init_z = jax.random.normal(rng, target_shape)`) and
`diffusion_utils.py:389-390` (`jax.random.normal(jax.random.fold_in(rng, i), shape)`).  The
generator itself is third-party (jax==0.3.16 in the colab, default `threefry2x32`
implementation, `jax_threefry_partitionable` off), so its published algorithm is restated:

  PRNGKey(seed)        -> key = [seed >> 32, seed & 0xffffffff]
  threefry_2x32(key,c) -> counters c (padded to even length) split into halves (x0, x1), 20 rounds
                          of Threefry-2x32 (Salmon et al. 2011), halves concatenated back
  fold_in(key, i)      -> threefry_2x32(key, [0, i])
  random_bits(key, n)  -> threefry_2x32(key, arange(n))
  normal(key, shape)   -> u = max(lo, f * (1 - lo) + lo), f = bitcast((bits >> 9) | 0x3f800000) - 1,
                          lo = nextafter(-1, 0); sqrt(2) * erfinv(u) with XLA's float32 erfinv
                          (Giles' single-precision polynomial)

Pinned (tests/test_jax_rng.py): the Threefry-2x32 known-answer vectors of the Random123
distribution, and the values the JAX documentation prints for `PRNGKey(0)`:
`normal(key, (3,)) = [1.8160863, -0.48262316, 0.33988908]`, `normal(key, ()) = -0.20584226`,
`split(key) = [[4146024105, 967050713], [2718843009, 1272950319]]`, `normal(subkey, ()) =
-1.2515389`.  The `fold_in` composition has no published vector: **unverified** against JAX.
The CUDA sampler carries the same generator (`rng_kind = 1`); this module is its CPU twin and
the reference for its test.
"""

from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

_ROTATIONS = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
  return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(key: Sequence[int], x0, x1) -> Tuple[np.ndarray, np.ndarray]:
  """20-round Threefry-2x32 of the counter words (x0, x1) under `key` (two uint32)."""
  with np.errstate(over='ignore'):
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    x0 = np.asarray(x0, np.uint32) + ks[0]
    x1 = np.asarray(x1, np.uint32) + ks[1]
    for group in range(5):
      for r in _ROTATIONS[group % 2]:
        x0 = x0 + x1
        x1 = _rotl(x1, r) ^ x0
      x0 = x0 + ks[(group + 1) % 3]
      x1 = x1 + ks[(group + 2) % 3] + np.uint32(group + 1)
    return x0.astype(np.uint32), x1.astype(np.uint32)


def prng_key(seed: int) -> np.ndarray:
  if seed < 0:
    raise ValueError('seed must be non-negative')
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def _threefry_counts(key, counts: np.ndarray) -> np.ndarray:
  counts = np.asarray(counts, np.uint32).ravel()
  n = counts.size
  if n % 2:
    counts = np.concatenate([counts, np.zeros(1, np.uint32)])
  half = counts.size // 2
  a, b = threefry2x32(key, counts[:half], counts[half:])
  return np.concatenate([a, b])[:n]


def fold_in(key, data: int) -> np.ndarray:
  return _threefry_counts(key, np.array([0, data & 0xFFFFFFFF], np.uint32))


def split(key, num: int = 2) -> np.ndarray:
  return _threefry_counts(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits(key, n: int) -> np.ndarray:
  return _threefry_counts(key, np.arange(n, dtype=np.uint32))


_ERFINV_CENTRAL = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
                   -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_TAIL = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
                -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def erfinv_f32(x: np.ndarray) -> np.ndarray:
  """float32 inverse error function as XLA evaluates it (two polynomials in w = -log1p(-x^2))."""
  f32 = np.float32
  x = np.asarray(x, f32)
  w = (-np.log1p((-x * x).astype(f32))).astype(f32)
  wc = (w - f32(2.5)).astype(f32)
  wt = (np.sqrt(np.maximum(w, f32(0))) - f32(3)).astype(f32)
  pc = np.full_like(x, f32(_ERFINV_CENTRAL[0]))
  for c in _ERFINV_CENTRAL[1:]:
    pc = (f32(c) + pc * wc).astype(f32)
  pt = np.full_like(x, f32(_ERFINV_TAIL[0]))
  for c in _ERFINV_TAIL[1:]:
    pt = (f32(c) + pt * wt).astype(f32)
  return (np.where(w < f32(5), pc, pt) * x).astype(f32)


def normal(key, shape) -> np.ndarray:
  """jax.random.normal(key, shape, float32)."""
  f32 = np.float32
  n = int(np.prod(shape)) if len(tuple(shape)) else 1
  bits = random_bits(key, n)
  f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(f32) - f32(1.0)
  lo, hi = np.nextafter(f32(-1), f32(0)), f32(1)
  u = np.maximum(lo, (f * f32(hi - lo)).astype(f32) + lo).astype(f32)
  return (f32(np.sqrt(2)) * erfinv_f32(u)).astype(f32).reshape(tuple(shape))


def init_z(seed: int, shape) -> np.ndarray:
  """diffusion_utils.py:462 with rng = PRNGKey(seed)."""
  return normal(prng_key(seed), shape)


def step_noise(seed: int, step: int, shape) -> np.ndarray:
  """diffusion_utils.py:389-390 at scan index `step`."""
  return normal(fold_in(prng_key(seed), step), shape)


def step_keys(seed: int, num_steps: int) -> np.ndarray:
  """[num_steps + 1, 2] uint32: row 0 = PRNGKey(seed), row i + 1 = fold_in(key, i)."""
  key = prng_key(seed)
  out = np.zeros((num_steps + 1, 2), np.uint32)
  out[0] = key
  zeros = np.zeros(num_steps, np.uint32)
  a, b = threefry2x32(key, zeros, np.arange(num_steps, dtype=np.uint32))
  out[1:, 0], out[1:, 1] = a, b
  return out
