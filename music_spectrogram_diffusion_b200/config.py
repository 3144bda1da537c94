"""Config dataclasses that keep the reference's gin-configurable names.

Mirrors (names, fields, defaults):
  network.T5Config                       msd/models/diffusion/network.py:54-72
  diffusion_utils.DiffusionSchedule      msd/models/diffusion/diffusion_utils.py:25-30
  diffusion_utils.ClassifierFreeGuidanceConfig                          :33-36
  diffusion_utils.SamplerConfig                                         :39-46
  diffusion_utils.DiffusionConfig                                       :49-59
so that a training ``config.gin`` of the reference binds onto them unchanged
(see gin_lite.py).  Plain dataclasses: no flax/jax.
"""

from __future__ import annotations

import dataclasses
from typing import Any, Optional, Sequence


@dataclasses.dataclass
class T5Config:
  vocab_size: int = 1536
  dtype: Any = 'float32'
  emb_dim: int = 512
  num_heads: int = 8
  num_encoder_layers: int = 6
  num_decoder_layers: int = 6
  head_dim: int = 64
  mlp_dim: int = 2048
  mlp_activations: Sequence[str] = ('relu',)
  dropout_rate: float = 0.1
  max_decoder_noise_time: float = 2e4
  decoder_cross_attend_style: str = 'sum_cross_attends'
  position_encoding: str = 'fixed'
  context_positions: str = 'regular'


@dataclasses.dataclass
class DiffusionSchedule:
  name: str = 'cosine'
  start: Optional[float] = None
  stop: Optional[float] = None
  num_steps: Optional[int] = None


@dataclasses.dataclass
class ClassifierFreeGuidanceConfig:
  drop_condition_prob: float = 0.1
  eval_condition_weight: float = 5.0


@dataclasses.dataclass
class SamplerConfig:
  name: str = 'ddpm'
  schedule: DiffusionSchedule = dataclasses.field(
      default_factory=lambda: DiffusionSchedule(name='cosine', num_steps=1000))
  clip_x0: bool = True
  logvar_type: str = 'large'


@dataclasses.dataclass
class DiffusionConfig:
  time_continuous_or_discrete: str = 'continuous'
  train_schedule: DiffusionSchedule = dataclasses.field(
      default_factory=lambda: DiffusionSchedule(name='cosine'))
  loss_norm: str = 'l1'
  loss_type: str = 'eps'
  model_output: str = 'eps'
  classifier_free_guidance: ClassifierFreeGuidanceConfig = dataclasses.field(
      default_factory=ClassifierFreeGuidanceConfig)
  sampler: SamplerConfig = dataclasses.field(default_factory=SamplerConfig)


# ---------------------------------------------------------------------------
# Named model sizes (the gin files the benchmark names).
# ---------------------------------------------------------------------------
def t5_base() -> T5Config:
  """gin/models/diffusion/context/t5_base.gin:69-83."""
  return T5Config(
      vocab_size=1536, dtype='float32', emb_dim=768, num_heads=12,
      num_encoder_layers=12, num_decoder_layers=12, head_dim=64, mlp_dim=2048,
      mlp_activations=('gelu', 'linear'), dropout_rate=0.1,
      decoder_cross_attend_style='concat_encodings',
      position_encoding='fixed_permuted_offset',
      context_positions='terminal_relative')


def t5_small() -> T5Config:
  """gin/models/diffusion/context/t5_small.gin:5-11 on top of t5_base."""
  c = t5_base()
  c.emb_dim, c.num_heads, c.num_encoder_layers = 512, 6, 8
  c.num_decoder_layers, c.head_dim, c.mlp_dim = 8, 64, 1024
  return c


def t5_tiny(emb_dim=128, num_heads=2, layers=2, mlp_dim=256) -> T5Config:
  """Test-sized network with the base topology (not a reference config)."""
  c = t5_base()
  c.emb_dim, c.num_heads = emb_dim, num_heads
  c.num_encoder_layers = c.num_decoder_layers = layers
  c.mlp_dim = mlp_dim
  return c


TASK_FEATURE_LENGTHS_CONTEXT = {
    # gin/tasks/mt3/context_mega.gin:5
    'inputs': 2048, 'targets': 256, 'targets_context': 256,
}
