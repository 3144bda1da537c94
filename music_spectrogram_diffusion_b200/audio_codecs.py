"""Codec constants and feature scaling on the hot path's boundary.

Mirrors `AudioCodec.scale_features / scale_to_features` (msd/audio_codecs.py:166-183) and the
MelGAN constants (204-218).  The mel<->audio transforms (Audio2Mel STFT 43-143, the TF-Hub
SoundStream vocoder 249-264) are a separate GAN, out of this path's scope:
`encode`/`decode` raise.
"""

from __future__ import annotations

import numpy as np


class AudioCodec:
  name: str
  n_dims: int
  sample_rate: int
  hop_size: int
  min_value: float
  max_value: float
  pad_value: float
  additional_frames_for_encoding: int = 0

  @property
  def abbrev_str(self):
    return self.name

  @property
  def frame_rate(self):
    return int(self.sample_rate // self.hop_size)

  def scale_features(self, features, output_range=(-1.0, 1.0), clip=False):
    min_out, max_out = output_range
    if clip:
      features = np.clip(features, self.min_value, self.max_value)
    zero_one = (features - self.min_value) / (self.max_value - self.min_value)
    return zero_one * (max_out - min_out) + min_out

  def scale_to_features(self, outputs, input_range=(-1.0, 1.0), clip=False):
    min_out, max_out = input_range
    outputs = np.clip(outputs, min_out, max_out) if clip else outputs
    zero_one = (outputs - min_out) / (max_out - min_out)
    return zero_one * (self.max_value - self.min_value) + self.min_value

  def encode(self, audio):
    raise NotImplementedError('audio -> mel is outside the DDPM hot path (SURVEY §2)')

  def decode(self, features):
    raise NotImplementedError('mel -> audio vocoder is outside the DDPM hot path (SURVEY §2)')

  @property
  def context_codec(self):
    return self


class MelGAN(AudioCodec):
  """128-bin log-mel at 16 kHz, hop 320 -> 50 frames/s (msd/audio_codecs.py:204-218)."""
  name = 'melgan'
  n_dims = 128
  sample_rate = 16000
  hop_size = 320
  min_value = float(np.log(1e-5))
  max_value = 4.0
  pad_value = float(np.log(1e-5))
  additional_frames_for_encoding = 16

  def __init__(self, decode_dither_amount: float = 0.0):
    self._decode_dither_amount = decode_dither_amount
