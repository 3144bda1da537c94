"""Consumes tests/golden/jax_tiny.npz -- vectors produced by running the REAL JAX reference
(tests/golden/make_jax_golden.py) -- whenever the file exists.  It cannot be generated in the
build image (no jax / flax / t5x, no network), so until a pod with those packages produces it
every test here skips with that reason and the network/sampler parity stays "unpinned"
(oracle/msd_oracle.py header).

CPU half: the oracle restatement and the restated jax.random stream against the reference's
outputs (tight: both are fp32 CPU).  GPU half: the CUDA path, through the C ABI, against the same
vectors with the tolerances of tests/test_gpu_model.py."""
import os

import numpy as np
import pytest
import torch

from music_spectrogram_diffusion_b200 import config, jax_rng as J, weights
from oracle import msd_oracle as O
from tests import helpers as H

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'jax_tiny.npz')
T = N = C = 128


@pytest.fixture(scope='module')
def g():
  if not os.path.exists(PATH):
    pytest.skip('tests/golden/jax_tiny.npz not generated (needs a JAX install: '
                'python tests/golden/make_jax_golden.py --reference <checkout>)')
  return np.load(PATH, allow_pickle=False)


def _setup(g):
  t5 = config.t5_tiny()
  params = weights.synthetic_params(t5, T, N, C, seed=int(g['weight_seed']))
  oc = H.oracle_config(t5, int(g['steps']), float(g['cond_weight']))
  batch = H.torch_batch(g['tokens'], g['ctx'], g['ctx_mask'])
  return t5, params, oc, batch


def test_generator_batch_is_the_seeded_one(g):
  toks, ctx, cmask = H.make_batch(2, T, C)
  cmask[1, :] = 1
  cmask[1, 40:] = 0
  np.testing.assert_array_equal(g['tokens'], toks)
  np.testing.assert_array_equal(g['ctx'], ctx)
  np.testing.assert_array_equal(g['ctx_mask'], cmask)


def test_jax_random_stream_including_fold_in(g):
  """init_z = normal(PRNGKey(seed)), noise_i = normal(fold_in(key, i)): bit-exact keys and bits,
  normals to an ulp of the erfinv polynomial."""
  seed, steps = int(g['seed']), int(g['steps'])
  keys = J.step_keys(seed, steps)
  np.testing.assert_allclose(J.init_z(seed, (2, N, 128)), g['init_z'], rtol=2e-6, atol=2e-7)
  for i in range(steps):
    np.testing.assert_array_equal(keys[i + 1], np.asarray(g[f'key_{i}'], np.uint32).reshape(-1)[-2:])
    np.testing.assert_array_equal(J.random_bits(keys[i + 1], 16), g[f'bits_{i}'])
    np.testing.assert_allclose(J.step_noise(seed, i, (2, N, 128)), g[f'noise_{i}'],
                               rtol=2e-6, atol=2e-7)


def test_oracle_primitives_match_reference(g):
  x, scale = torch.from_numpy(g['prim_x']), torch.from_numpy(g['prim_scale'])
  np.testing.assert_allclose(O.layer_norm(x, scale).numpy(), g['prim_layer_norm'], rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(O.gelu_tanh(x).numpy(), g['prim_gelu'], rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(O.swish(x).numpy(), g['prim_swish'], rtol=1e-5, atol=1e-6)
  film = O.film_layer(x, torch.from_numpy(g['prim_film_cond']), torch.from_numpy(g['prim_film_kernel']))
  np.testing.assert_allclose(film.numpy(), g['prim_film'], rtol=1e-4, atol=1e-5)


def test_oracle_network_and_sampler_match_reference(g):
  t5, params, oc, batch = _setup(g)
  P = O.params_to(params)
  encs = O.encode(P, oc, batch['encoder_input_tokens'],
                  O.scale_features(batch['encoder_continuous_inputs'], oc, clip=True),
                  batch['encoder_continuous_mask'])
  tok_valid = (batch['encoder_input_tokens'] > 0).unsqueeze(-1)
  ctx_valid = (batch['encoder_continuous_mask'] > 0).unsqueeze(-1)
  assert ((encs[0][0] - torch.from_numpy(g['enc_tokens'])).abs() * tok_valid).max() < 2e-4
  assert ((encs[1][0] - torch.from_numpy(g['enc_context'])).abs() * ctx_valid).max() < 2e-4
  init_z = torch.from_numpy(g['init_z'])
  steps = int(g['steps'])
  for i in (steps - 1, 5, 0):
    t = torch.full((2,), float(np.float32(i + 1.0) / np.float32(steps)))
    for flag, name in ((1.0, 'cond'), (0.0, 'uncond')):
      want = torch.from_numpy(g[f'eps_{name}_{i}'])
      got = O.decode(P, oc, [(e * flag, m * flag) for e, m in encs], init_z, t)
      assert (got - want).abs().max() < 5e-4 * want.abs().max(), (i, name)
  noise = torch.stack([torch.from_numpy(g[f'noise_{i}']) for i in range(steps)])

  def pred_fn(z, time_, cond):
    f = 1.0 if cond else 0.0
    return O.decode(P, oc, [(e * f, m * f) for e, m in encs], z, time_)
  z1 = O.eval_step(init_z, steps - 1, noise[steps - 1], pred_fn, oc)
  # first reverse step multiplies the network output by ~22026 before the clip
  assert (z1 - torch.from_numpy(g['z_after_first'])).abs().mean() < 1e-3
  mel, _ = O.predict_batch_with_aux(P, oc, batch, init_z, noise)
  span = oc.max_value - oc.min_value
  err = (mel - torch.from_numpy(g['mel'])).abs() / span * 2.0
  assert err.mean() < 1e-3 and (err > 0.1).float().mean() < 5e-3, (err.mean(), err.max())
  np.testing.assert_allclose(g['mel'], g['mel_scan_only'], atol=1e-5)


@pytest.mark.gpu
def test_cuda_path_matches_reference_vectors(g, cuda_device):
  """The product (libmsd_b200.so through the C ABI) against the reference's own outputs: encoder,
  decoder forward, full trajectory from the seed alone (rng='jax')."""
  t5, params, oc, batch = _setup(g)
  steps, weight, seed = int(g['steps']), float(g['cond_weight']), int(g['seed'])
  eng = H.build_engine(t5, T, N, C, 2, steps, weight, params)
  b = H.torch_batch(g['tokens'], g['ctx'], g['ctx_mask'], cuda_device)
  eng.encode(b['encoder_input_tokens'], b['encoder_continuous_inputs'], b['encoder_continuous_mask'])
  got = eng.encodings().cpu()
  want = torch.cat([torch.from_numpy(g['enc_tokens']), torch.from_numpy(g['enc_context'])], dim=1)
  valid = torch.cat([batch['encoder_input_tokens'] > 0, batch['encoder_continuous_mask'] > 0], dim=1)
  assert ((got - want).abs() * valid.unsqueeze(-1)).max() < 6e-2
  z = torch.from_numpy(g['init_z']).to(cuda_device)
  for i in (steps - 1, 5, 0):
    for cond, name in ((True, 'cond'), (False, 'uncond')):
      want = torch.from_numpy(g[f'eps_{name}_{i}'])
      eps = eng.decode_eps(z, i, cond).cpu()
      assert (eps - want).abs().max() < 3e-2 * want.abs().max(), (i, name)
  mel = eng.sample(seed=seed).cpu()
  span = oc.max_value - oc.min_value
  err = (mel - torch.from_numpy(g['mel'])).abs() / span * 2.0
  H.assert_trajectory_close(err, 'jax reference, tiny, seed-driven')
  eng.close()
