"""Dumps REFERENCE vectors for the hot path by running the real JAX reference.

This is the route from "parity unpinned" to pinned: on a machine that has `jax`, `flax` (and,
for the model-level call, `t5x` + `seqio` + `gin` + `clu` + `tensorflow`) and a checkout of
magenta/music-spectrogram-diffusion, run

    python tests/golden/make_jax_golden.py --reference /path/to/music-spectrogram-diffusion

It writes `tests/golden/jax_tiny.npz`; `tests/test_jax_golden.py` consumes the file whenever it is
present (CPU: oracle vs these vectors; GPU: the CUDA path vs these vectors) and skips otherwise.
None of the dependencies is installable in the build image (no network), so the file is NOT
committed yet; the script has been written against the reference sources, not executed.

What is dumped (tiny network = config.t5_tiny(), weights.synthetic_params(seed 0), the seeded
batch of tests/helpers.make_batch, all in fp32):
  enc_tokens / enc_context       module.encode outputs                 network.py:537-559
  eps_cond_<i> / eps_uncond_<i>  module.decode at steps i in STEPS_DEC  network.py:561-573
  z_after_first                  one eval_step body from init_z          diffusion_utils.py:398-453
  mel / mel_scan_only            predict_batch_with_aux(rng=PRNGKey(SEED)) models.py:340-400
                                 (model class when t5x is importable, else the same five calls
                                  made directly -- recorded in `via`)
  init_z, noise_<i>              jax.random.normal(key) / normal(fold_in(key, i))  -- pins the
                                 restated threefry stream incl. the fold_in composition
  bits_<i>                       jax.random.bits(fold_in(key, i), (16,)) raw uint32
  layer_norm / gelu / film       primitives on fixed inputs              layers.py:632-666
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

T = N = C = 128
STEPS, WEIGHT, B, SEED = 12, 2.0, 2, 7
STEPS_DEC = (STEPS - 1, 5, 0)


def nest(flat):
  tree = {}
  for k, v in flat.items():
    node = tree
    parts = k.split('/')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    node[parts[-1]] = v
  return tree


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--reference', default=os.environ.get('MSD_REFERENCE', '/root/reference'))
  ap.add_argument('--out', default=os.path.join(HERE, 'jax_tiny.npz'))
  args = ap.parse_args()
  sys.path.insert(0, args.reference)

  import jax
  import jax.numpy as jnp
  jax.config.update('jax_default_matmul_precision', 'highest')  # true fp32 dots on any backend
  from music_spectrogram_diffusion import layers as ref_layers
  from music_spectrogram_diffusion.models.diffusion import diffusion_utils as du
  from music_spectrogram_diffusion.models.diffusion import network as ref_network

  from music_spectrogram_diffusion_b200 import config, weights
  from tests import helpers as H

  t5 = config.t5_tiny()
  flat = weights.synthetic_params(t5, T, N, C, seed=0)
  params = jax.tree_util.tree_map(jnp.asarray, nest(flat))
  rcfg = ref_network.T5Config(
      vocab_size=t5.vocab_size, dtype=jnp.float32, emb_dim=t5.emb_dim, num_heads=t5.num_heads,
      num_encoder_layers=t5.num_encoder_layers, num_decoder_layers=t5.num_decoder_layers,
      head_dim=t5.head_dim, mlp_dim=t5.mlp_dim, mlp_activations=tuple(t5.mlp_activations),
      dropout_rate=t5.dropout_rate, max_decoder_noise_time=t5.max_decoder_noise_time,
      decoder_cross_attend_style=t5.decoder_cross_attend_style,
      position_encoding=t5.position_encoding, context_positions=t5.context_positions)
  module = ref_network.ContinuousContextTransformer(config=rcfg)
  dcfg = du.DiffusionConfig(
      classifier_free_guidance=du.ClassifierFreeGuidanceConfig(eval_condition_weight=WEIGHT),
      sampler=du.SamplerConfig(schedule=du.DiffusionSchedule(name='cosine', num_steps=STEPS)))

  toks, ctx, cmask = H.make_batch(B, T, C)
  cmask[1, :] = 1
  cmask[1, 40:] = 0  # partially filled context: terminal-relative roll by 40
  out = dict(tokens=toks, ctx=ctx, ctx_mask=cmask, steps=STEPS, cond_weight=WEIGHT, seed=SEED,
             weight_seed=0, jax_version=jax.__version__)

  codec_source = 'reference'
  try:
    from music_spectrogram_diffusion import audio_codecs as ref_codecs
    codec = ref_codecs.MelGAN()
  except Exception as e:  # tensorflow missing: same two affine maps, audio_codecs.py:166-183
    codec_source = f'local restatement ({type(e).__name__})'

    class _Codec:
      min_value, max_value, n_dims = float(np.log(1e-5)), 4.0, 128

      def scale_features(self, f, output_range=(-1.0, 1.0), clip=False):
        lo, hi = output_range
        if clip:
          f = jnp.clip(f, self.min_value, self.max_value)
        return (f - self.min_value) / (self.max_value - self.min_value) * (hi - lo) + lo

      def scale_to_features(self, o, input_range=(-1.0, 1.0), clip=False):
        lo, hi = input_range
        return (o - lo) / (hi - lo) * (self.max_value - self.min_value) + self.min_value
    codec = _Codec()
  out['codec_source'] = codec_source

  # ---- encode / decode ------------------------------------------------------------------------
  ctx_scaled = codec.scale_features(jnp.asarray(ctx), output_range=[-1.0, 1.0], clip=True)
  encs = module.apply({'params': params}, input_tokens=jnp.asarray(toks),
                      continuous_inputs=ctx_scaled, continuous_mask=jnp.asarray(cmask),
                      enable_dropout=False, method=module.encode)
  out['enc_tokens'] = np.asarray(encs[0][0])
  out['enc_context'] = np.asarray(encs[1][0])

  def pred_fn(z, time, include_conditioning):
    step_encs = jax.tree_util.tree_map(lambda x: x * include_conditioning, encs)
    return module.apply({'params': params}, encodings_and_masks=step_encs, input_tokens=z,
                        noise_time=time, enable_dropout=False, method=module.decode)

  key = jax.random.PRNGKey(SEED)
  init_z = jax.random.normal(key, shape=(B, N, 128), dtype=jnp.float32)
  out['init_z'] = np.asarray(init_z)
  for i in range(STEPS):
    k = jax.random.fold_in(key, i)
    out[f'noise_{i}'] = np.asarray(jax.random.normal(k, shape=(B, N, 128), dtype=jnp.float32))
    out[f'bits_{i}'] = np.asarray(jax.random.bits(k, (16,), dtype=jnp.uint32))
    out[f'key_{i}'] = np.asarray(jax.random.key_data(k) if hasattr(jax.random, 'key_data') else k)
  for i in STEPS_DEC:
    t = jnp.full((B,), (i + 1.0) / STEPS, dtype=jnp.float32)
    out[f'eps_cond_{i}'] = np.asarray(pred_fn(init_z, t, True))
    out[f'eps_uncond_{i}'] = np.asarray(pred_fn(init_z, t, False))
  body = du.eval_step(rng=key, diffusion_config=dcfg, batch_size=B, pred_fn=pred_fn)
  z1, _ = body(init_z, jnp.asarray(STEPS - 1))
  out['z_after_first'] = np.asarray(z1)
  x0 = du.eval_scan(key, (B, N, 128), pred_fn, dcfg)
  out['mel_scan_only'] = np.asarray(codec.scale_to_features(x0, input_range=[-1.0, 1.0]))

  # ---- the model-level call (needs t5x) -------------------------------------------------------
  via = 'module.apply + diffusion_utils.eval_scan (t5x not importable)'
  try:
    from music_spectrogram_diffusion.models.diffusion import models as ref_models
    model = ref_models.ContextDiffusionModel(
        module=module, input_vocabulary=None, output_vocabulary=None, optimizer_def=None,
        diffusion_config=dcfg, audio_codec=codec)
    batch = dict(encoder_input_tokens=jnp.asarray(toks), encoder_continuous_inputs=jnp.asarray(ctx),
                 encoder_continuous_mask=jnp.asarray(cmask),
                 decoder_target_tokens=jnp.zeros((B, N, 128), jnp.float32))
    mel, scores = model.predict_batch_with_aux(params, batch, rng=key)
    out['mel'] = np.asarray(mel)
    out['scores'] = np.asarray(scores)
    via = 'ContextDiffusionModel.predict_batch_with_aux'
  except Exception as e:  # pylint: disable=broad-except
    print(f'model-level call unavailable: {type(e).__name__}: {e}', file=sys.stderr)
    out['mel'] = out['mel_scan_only']
  out['via'] = via

  # ---- primitives the reference's own tests do not pin ----------------------------------------
  rng = np.random.default_rng(0)
  x = rng.standard_normal((4, 16, t5.emb_dim)).astype(np.float32) * 3
  scale = (1 + 0.1 * rng.standard_normal(t5.emb_dim)).astype(np.float32)
  ln = ref_layers.LayerNorm(dtype=jnp.float32)
  out['prim_x'] = x
  out['prim_scale'] = scale
  out['prim_layer_norm'] = np.asarray(ln.apply({'params': {'scale': jnp.asarray(scale)}}, jnp.asarray(x)))
  import flax.linen as nn
  out['prim_gelu'] = np.asarray(nn.gelu(jnp.asarray(x)))
  out['prim_swish'] = np.asarray(nn.swish(jnp.asarray(x)))
  cond = rng.standard_normal((4, 1, 4 * t5.emb_dim)).astype(np.float32)
  fk = (rng.standard_normal((4 * t5.emb_dim, 2 * t5.emb_dim)) * 0.05).astype(np.float32)
  film = ref_layers.FiLMLayer()
  out['prim_film_cond'] = cond
  out['prim_film_kernel'] = fk
  out['prim_film'] = np.asarray(film.apply(
      {'params': {'DenseGeneral_0': {'kernel': jnp.asarray(fk)}}}, jnp.asarray(x), jnp.asarray(cond)))

  np.savez_compressed(args.out, **out)
  print(f'wrote {args.out} ({via}; codec: {codec_source}; jax {jax.__version__})')


if __name__ == '__main__':
  main()
