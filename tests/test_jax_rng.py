"""jax.random (threefry2x32) restated in numpy: Random123 known answers for the block function
and the values the JAX documentation prints for PRNGKey(0)."""
import numpy as np
import pytest

from music_spectrogram_diffusion_b200 import jax_rng as J


def test_threefry2x32_known_answers():
  cases = [((0x0, 0x0), (0x0, 0x0), (0x6b200159, 0x99ba4efe)),
           ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
           ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
  for key, ctr, want in cases:
    a, b = J.threefry2x32(key, [ctr[0]], [ctr[1]])
    assert (int(a[0]), int(b[0])) == want


def test_documented_values_for_key_zero():
  key = J.prng_key(0)
  np.testing.assert_array_equal(key, [0, 0])
  np.testing.assert_array_equal(J.normal(key, (3,)), np.array([1.8160863, -0.48262316, 0.33988908], np.float32))
  assert J.normal(key, ()) == np.float32(-0.20584226)
  ks = J.split(key)
  np.testing.assert_array_equal(ks, [[4146024105, 967050713], [2718843009, 1272950319]])
  assert J.normal(ks[1], (1,))[0] == np.float32(-1.2515389)
  np.testing.assert_array_equal(J.prng_key(42), [0, 42])
  np.testing.assert_array_equal(J.prng_key((7 << 32) | 9), [7, 9])
  with pytest.raises(ValueError):
    J.prng_key(-1)


def test_normal_statistics_and_layout():
  x = J.normal(J.prng_key(3), (4, 256, 128))
  assert x.dtype == np.float32 and x.shape == (4, 256, 128) and np.isfinite(x).all()
  assert abs(float(x.mean())) < 5e-3 and abs(float(x.std()) - 1.0) < 5e-3
  # odd sizes pad the counter array by one; element k of the stream does not depend on the padding
  odd = J.random_bits(J.prng_key(3), 5)
  a, b = J.threefry2x32(J.prng_key(3), [0, 1, 2], [3, 4, 0])
  np.testing.assert_array_equal(odd, np.concatenate([a, b])[:5])
  # a batch of B segments is NOT B independent single-segment streams (counter = flat index)
  assert not np.array_equal(J.normal(J.prng_key(3), (1, 256, 128))[0], x[0])


def test_step_keys_match_fold_in():
  keys = J.step_keys(11, 6)
  np.testing.assert_array_equal(keys[0], J.prng_key(11))
  for i in range(6):
    np.testing.assert_array_equal(keys[i + 1], J.fold_in(J.prng_key(11), i))
  np.testing.assert_array_equal(J.step_noise(11, 4, (2, 8)), J.normal(keys[5], (2, 8)))
  np.testing.assert_array_equal(J.init_z(11, (2, 8)), J.normal(keys[0], (2, 8)))


def test_erfinv_matches_double_precision():
  from scipy.special import erfinv
  u = np.linspace(-0.9999, 0.9999, 20001).astype(np.float32)
  np.testing.assert_allclose(J.erfinv_f32(u), erfinv(u.astype(np.float64)), rtol=5e-6, atol=2e-7)   # the polynomial itself is good to ~3e-6
