"""Thin host wrapper over the C ABI: torch tensors are storage only.

`Engine` owns one `msd_ctx` (one GPU).  It mirrors the operator surface the
reference exposes one level below `InferenceModel.predict`:
  encode      <-> module.apply(..., method=module.encode)   models.py:365-371
  decode_eps  <-> module.apply(..., method=module.decode)   models.py:378-386
  sample      <-> diffusion_utils.eval_scan + scale_to_features  models.py:393-395
"""

from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import numpy as np
import torch

from music_spectrogram_diffusion_b200 import _native
from music_spectrogram_diffusion_b200.config import DiffusionConfig, T5Config


PRECISIONS = {'bf16': 0, 'fp32_accurate': 1}


def make_msd_config(t5: T5Config, diffusion: DiffusionConfig, inputs_length: int,
                    targets_length: int, context_length: int, max_batch: int,
                    n_dims: int = 128, feature_min: float = math.log(1e-5),
                    feature_max: float = 4.0, rng: str = 'jax',
                    precision: str = 'bf16') -> _native.MsdConfig:
  """Translate the reference's config objects into `struct msd_config`.  rng: 'jax' (the
  threefry stream of jax.random.PRNGKey(seed), as the reference draws its noise) or 'philox'.
  precision: 'bf16' (tensor-core operands in bf16, the fast path) or 'fp32_accurate' (3 x bf16
  split-precision dense layers + fp32 attention: what T5Config.dtype = float32 asks for)."""
  if rng not in ('jax', 'philox'):
    raise ValueError(f'unknown rng {rng!r}')
  if precision not in PRECISIONS:
    raise ValueError(f'unknown precision {precision!r} (expected one of {sorted(PRECISIONS)})')
  if tuple(t5.mlp_activations) != ('gelu', 'linear'):
    raise NotImplementedError(
        f'mlp_activations={t5.mlp_activations}: only the gated-GELU MLP of the '
        'diffusion configs (gin/models/diffusion/context/t5_base.gin:79) is built')
  styles = {'concat_encodings': 0, 'sum_cross_attends': 1}
  if t5.decoder_cross_attend_style not in styles:
    raise ValueError(f'Unknown decoder_cross_attend_style: {t5.decoder_cross_attend_style}')
  if diffusion.model_output == 'x0_and_eps':
    raise NotImplementedError(
        'model_output="x0_and_eps" needs a 2*n_dims output head; the context network emits n_dims '
        'channels (network.py:452-456), so the reference cannot run it on this path either')
  if diffusion.model_output not in ('eps', 'x0', 'v'):
    raise ValueError('Unknown model_output: %s' % diffusion.model_output)
  sched, tsched = diffusion.sampler.schedule, diffusion.train_schedule
  names = {'cosine': 0, 'linear': 1}
  for sc in (sched, tsched):
    if sc.name not in names:
      raise ValueError('Schedule %s not identified.' % sc.name)
    if sc.name == 'linear' and (sc.start is None or sc.stop is None or not sc.num_steps):
      raise ValueError('linear schedule needs start, stop and num_steps')
  if diffusion.sampler.name not in ('ddpm', 'ddim'):
    raise ValueError('Unknown sampler type: %s' % diffusion.sampler.name)
  sampler = {'ddpm': 0, 'ddim': 1}[diffusion.sampler.name]
  lv = diffusion.sampler.logvar_type
  logvar_frac = 0.0
  if lv.startswith('medium:'):
    logvar, logvar_frac = 2, float(lv.split(':')[1])
    if not 0 <= logvar_frac <= 1:
      raise ValueError(f'logvar_type {lv!r}: frac must be in [0, 1]')
  elif lv in ('large', 'small'):
    logvar = {'large': 0, 'small': 1}[lv]
  else:
    raise ValueError(f'unknown logvar_type {lv!r}')
  ctxpos = {'regular': 0, 'terminal_relative': 1}[t5.context_positions]
  return _native.MsdConfig(
      vocab_size=t5.vocab_size, emb_dim=t5.emb_dim, num_heads=t5.num_heads,
      head_dim=t5.head_dim, num_encoder_layers=t5.num_encoder_layers,
      num_decoder_layers=t5.num_decoder_layers, mlp_dim=t5.mlp_dim,
      inputs_length=inputs_length, targets_length=targets_length,
      context_length=context_length, n_dims=n_dims, num_steps=int(sched.num_steps),
      max_batch=max_batch, sampler=sampler, logvar_type=logvar,
      clip_x0=int(bool(diffusion.sampler.clip_x0)), context_positions=ctxpos,
      max_decoder_noise_time=float(t5.max_decoder_noise_time),
      eval_condition_weight=float(diffusion.classifier_free_guidance.eval_condition_weight),
      feature_min=float(feature_min), feature_max=float(feature_max),
      model_output={'eps': 0, 'x0': 1, 'v': 2}[diffusion.model_output],
      sampler_schedule=names[sched.name], train_schedule=names[tsched.name],
      train_num_steps=int(tsched.num_steps or 0), logvar_frac=logvar_frac,
      sampler_beta_start=float(sched.start or 0.0), sampler_beta_stop=float(sched.stop or 0.0),
      train_beta_start=float(tsched.start or 0.0), train_beta_stop=float(tsched.stop or 0.0),
      cross_attend_style=styles[t5.decoder_cross_attend_style],
      rng_kind={'philox': 0, 'jax': 1}[rng], precision=PRECISIONS[precision])


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
  return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device: torch.device) -> ctypes.c_void_p:
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
  """One msd_ctx on one GPU."""

  def __init__(self, cfg: _native.MsdConfig, device: int = 0):
    self.lib = _native.load()
    if not torch.cuda.is_available():
      raise _native.MsdError('no CUDA device: the sm_100a library cannot run (no CPU fallback)')
    self.cfg = cfg
    self.device = torch.device('cuda', device)
    torch.cuda.set_device(self.device)
    torch.zeros(1, device=self.device)  # make sure the primary context exists
    handle = ctypes.c_void_p()
    _native.check(self.lib.msd_create(ctypes.byref(cfg), device, ctypes.byref(handle)),
                  'msd_create')
    self._h = handle
    self._batch = 0

  def close(self) -> None:
    if getattr(self, '_h', None):
      self.lib.msd_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  # -- weights ---------------------------------------------------------------
  def load_weights(self, params: Dict[str, np.ndarray]) -> None:
    names = sorted(params)
    arr = (_native.MsdTensor * len(names))()
    keep = []
    for i, k in enumerate(names):
      a = np.ascontiguousarray(params[k], dtype=np.float32)
      keep.append(a)
      arr[i].name = k.encode()
      arr[i].data = a.ctypes.data
      arr[i].ndim = a.ndim
      for j, s in enumerate(a.shape):
        arr[i].shape[j] = s
    _native.check(self.lib.msd_load_weights(self._h, arr, len(names)), 'msd_load_weights')

  # -- operator surface --------------------------------------------------------
  def encode(self, tokens: torch.Tensor, ctx_features: torch.Tensor,
             ctx_mask: torch.Tensor) -> None:
    b = tokens.shape[0]
    assert tokens.dtype == torch.int32 and tokens.is_cuda and tokens.is_contiguous()
    assert ctx_features.dtype == torch.float32 and ctx_features.is_contiguous()
    assert ctx_mask.dtype == torch.int32 and ctx_mask.is_contiguous()
    assert tokens.shape == (b, self.cfg.inputs_length), tokens.shape
    assert ctx_features.shape == (b, self.cfg.context_length, self.cfg.n_dims)
    assert ctx_mask.shape == (b, self.cfg.context_length)
    _native.check(self.lib.msd_encode(self._h, _ptr(tokens), _ptr(ctx_features), _ptr(ctx_mask),
                                      b, _stream(self.device)), 'msd_encode')
    self._batch = b

  def encodings(self) -> torch.Tensor:
    out = torch.empty(self._batch, self.cfg.inputs_length + self.cfg.context_length,
                      self.cfg.emb_dim, dtype=torch.float32, device=self.device)
    _native.check(self.lib.msd_get_encodings(self._h, _ptr(out), _stream(self.device)),
                  'msd_get_encodings')
    return out

  def decode_eps(self, z: torch.Tensor, step_i: int, conditioned: bool) -> torch.Tensor:
    assert z.dtype == torch.float32 and z.is_cuda and z.is_contiguous()
    assert z.shape == (self._batch, self.cfg.targets_length, self.cfg.n_dims)
    out = torch.empty_like(z)
    _native.check(self.lib.msd_decode_eps(self._h, _ptr(z), step_i, int(conditioned), _ptr(out),
                                          _stream(self.device)), 'msd_decode_eps')
    return out

  def sample(self, init_z: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
             seed: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    shape = (self._batch, self.cfg.targets_length, self.cfg.n_dims)
    if init_z is not None:
      assert init_z.dtype == torch.float32 and init_z.is_contiguous() and init_z.shape == shape
    if noise is not None:
      assert noise.dtype == torch.float32 and noise.is_contiguous()
      assert noise.shape == (self.cfg.num_steps,) + shape, noise.shape
    if out is None:
      out = torch.empty(shape, dtype=torch.float32, device=self.device)
    _native.check(self.lib.msd_sample(self._h, _ptr(init_z), _ptr(noise), seed, _ptr(out),
                                      _stream(self.device)), 'msd_sample')
    return out

  # -- guidance split over two GPUs (msd_p2p_*, include/msd_b200.h) --------------------------
  def p2p_export(self) -> bytes:
    buf = ctypes.create_string_buffer(64)
    _native.check(self.lib.msd_p2p_export(self._h, buf), 'msd_p2p_export')
    return buf.raw

  def p2p_attach(self, peer_handle: bytes, role: str) -> None:
    """role: 'cond' (this GPU runs the conditional pass) or 'uncond'."""
    if len(peer_handle) != 64:
      raise ValueError('peer handle must be the 64 bytes msd_p2p_export returned on the other rank')
    self._peer_handle = ctypes.create_string_buffer(peer_handle, 64)
    _native.check(self.lib.msd_p2p_attach(self._h, self._peer_handle,
                                          {'cond': 1, 'uncond': 2}[role]), 'msd_p2p_attach')

  def p2p_detach(self) -> None:
    _native.check(self.lib.msd_p2p_detach(self._h), 'msd_p2p_detach')

  KERNEL_CLASSES = ('gemm', 'attention', 'rmsnorm_film', 'sampler', 'other')

  def profile_step(self, step_i: int = 500, reps: int = 3) -> Dict[str, Dict[str, float]]:
    """Per-kernel-class CUDA-event timing of one diffusion step (uncaptured)."""
    out = np.zeros((5, 4), dtype=np.float64)
    _native.check(self.lib.msd_profile_step(self._h, step_i, reps, out.ctypes.data),
                  'msd_profile_step')
    return {name: dict(ms=float(out[i, 0]), launches=float(out[i, 1]), flops=float(out[i, 2]),
                       bytes=float(out[i, 3]))
            for i, name in enumerate(self.KERNEL_CLASSES)}

  def step_table(self) -> np.ndarray:
    tab = np.zeros((self.cfg.num_steps, 16), dtype=np.float32)
    _native.check(self.lib.msd_get_step_table(self._h, tab.ctypes.data), 'msd_get_step_table')
    return tab


def launch_count() -> int:
  return int(_native.load().msd_launch_count())


# ---- operator-level hooks (unit parity with msd/layers.py) --------------------
def op_dense(a: torch.Tensor, w: torch.Tensor, variant: int = 0, block_n: int = 0) -> torch.Tensor:
  lib = _native.load()
  m, k = a.shape
  n = w.shape[1]
  out = torch.empty(m, n, dtype=torch.float32, device=a.device)
  _native.check(lib.msd_op_dense_variant(_ptr(a.contiguous()), _ptr(w.contiguous()), m, n, k,
                                         _ptr(out), variant, block_n, _stream(a.device)),
                'msd_op_dense_variant')
  return out


def op_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                 key_mask: Optional[torch.Tensor], heads: int) -> torch.Tensor:
  lib = _native.load()
  nb, lq, _ = q.shape
  lk = k.shape[1]
  out = torch.empty_like(q)
  _native.check(lib.msd_op_attention(_ptr(q.contiguous()), _ptr(k.contiguous()),
                                     _ptr(v.contiguous()), _ptr(key_mask), nb, heads, lq, lk,
                                     _ptr(out), _stream(q.device)), 'msd_op_attention')
  return out


def op_rmsnorm_film(x: torch.Tensor, gamma: torch.Tensor,
                    film: Optional[torch.Tensor]) -> torch.Tensor:
  lib = _native.load()
  rows, d = x.shape
  out = torch.empty_like(x)
  _native.check(lib.msd_op_rmsnorm_film(_ptr(x.contiguous()), _ptr(gamma), _ptr(film), rows, d,
                                        _ptr(out), _stream(x.device)), 'msd_op_rmsnorm_film')
  return out


EPILOGUES = {'bf16': 0, 'resid_f32': 2, 'gated_gelu': 3, 'pos_f32': 4, 'gated_gelu_split3': 5}


def op_dense_epilogue(a: torch.Tensor, w: torch.Tensor, epilogue: str, block_n: int = 0,
                      w1: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
                      pos: Optional[torch.Tensor] = None, pos_shift: Optional[torch.Tensor] = None,
                      dup_rows: int = 0) -> torch.Tensor:
  """The GEMM with one of its fused epilogues (kernels.h GemmEpilogue); see msd_op_dense_epilogue."""
  lib = _native.load()
  m, k = a.shape
  n = w.shape[1]
  out = torch.empty(m + dup_rows, n, dtype=torch.float32, device=a.device)
  pos_rows = 0 if pos is None else pos.shape[0]
  keep = [t.contiguous() if t is not None else None for t in (a, w, w1, resid, pos, pos_shift)]
  _native.check(lib.msd_op_dense_epilogue(_ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), m, n, k,
                                          EPILOGUES[epilogue], block_n, _ptr(keep[3]), _ptr(keep[4]),
                                          pos_rows, _ptr(keep[5]), dup_rows, _ptr(out),
                                          _stream(a.device)), 'msd_op_dense_epilogue')
  return out


def op_dense_deferred_norm(a: torch.Tensor, w_out: torch.Tensor, x: torch.Tensor, g_lo: torch.Tensor,
                           g_hi: torch.Tensor, split_row: int, w2: torch.Tensor,
                           w2b: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                           block_n1: int = 0, block_n2: int = 0):
  """Residual projection with the deferred-normalisation epilogue + the projection consuming it
  (msd_op_dense_deferred_norm): returns (x_out [M, d], y [M, N2])."""
  lib = _native.load()
  m, k = a.shape
  d = w_out.shape[1]
  n2 = w2.shape[1]
  x_out = torch.empty(m, d, dtype=torch.float32, device=a.device)
  y = torch.empty(m, n2, dtype=torch.float32, device=a.device)
  keep = [t.contiguous() if t is not None else None for t in (a, w_out, x, g_lo, g_hi, w2, w2b, bias)]
  _native.check(lib.msd_op_dense_deferred_norm(_ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), m, d, k,
                                               _ptr(keep[3]), _ptr(keep[4]), split_row, _ptr(keep[5]),
                                               _ptr(keep[6]), n2, _ptr(keep[7]), block_n1, block_n2,
                                               _ptr(x_out), _ptr(y), _stream(a.device)),
                'msd_op_dense_deferred_norm')
  return x_out, y


def op_attention_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                     key_mask: Optional[torch.Tensor], heads: int) -> torch.Tensor:
  lib = _native.load()
  nb, lq, _ = q.shape
  lk = k.shape[1]
  out = torch.empty_like(q)
  _native.check(lib.msd_op_attention_f32(_ptr(q.contiguous()), _ptr(k.contiguous()),
                                         _ptr(v.contiguous()), _ptr(key_mask), nb, heads, lq, lk,
                                         _ptr(out), _stream(q.device)), 'msd_op_attention_f32')
  return out


def op_jax_bits(seed: int, step: int, n: int, device: torch.device) -> torch.Tensor:
  """Raw uint32 words of the device jax.random stream (as int32 storage; view as uint32)."""
  out = torch.empty(n, dtype=torch.int32, device=device)
  _native.check(_native.load().msd_op_jax_bits(seed, step, n, _ptr(out), _stream(device)),
                'msd_op_jax_bits')
  return out


def op_jax_normal(seed: int, step: int, n: int, device: torch.device) -> torch.Tensor:
  """Device draw of the jax.random stream (step < 0: init_z; else the noise of scan index step)."""
  out = torch.empty(n, dtype=torch.float32, device=device)
  _native.check(_native.load().msd_op_jax_normal(seed, step, n, _ptr(out), _stream(device)),
                'msd_op_jax_normal')
  return out
