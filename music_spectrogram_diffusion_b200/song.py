"""Full-song synthesis driver: notes (or a MIDI file) -> chained 5.12 s segments -> mel frames.

Library form of the loops the reference keeps in its colab ("Synthesize Audio" cell) and in
`beam/evaluation.py:156-276`: the first segment runs with a masked context, every later one
is conditioned on the previous segment's predicted mel; per-segment wall times are reported
with the reference's `model_timing` fields (first segment excluded, evaluation.py:217-220,
238-247).  The mel -> audio vocoder is outside this path.
"""

from __future__ import annotations

import time
from typing import Any, Dict, Optional, Union

import numpy as np

from music_spectrogram_diffusion_b200 import midi_file, midi_tokens


def event_vocabulary_of(model) -> midi_tokens.EventVocabulary:
  """The model's event vocabulary: `InferenceModel.codec` is the tokeniser's own object."""
  return model.codec


def load_notes(midi: Union[str, bytes], sustain: bool = True) -> np.ndarray:
  data = open(midi, 'rb').read() if isinstance(midi, str) else midi
  song = midi_file.read_midi(data)
  if sustain:
    song = midi_file.apply_sustain(song)
  return song.notes


def synthesize_song(model, notes: np.ndarray, seed: int = 0, always_mask_context: bool = False,
                    max_segments: Optional[int] = None) -> Dict[str, Any]:
  """model: an `InferenceModel` (anything with .predict, .sequence_length, .audio_codec, .codec).

  Returns {'full_pred_encoded': f32 [segments * targets_length, n_dims] in feature units,
  'num_frames': frames that belong to the song, 'tokens': the per-segment model inputs,
  'model_timing': {...}} -- the same keys `beam/evaluation.py` yields for this part."""
  ac = model.audio_codec
  lengths = model.sequence_length
  toks = midi_tokens.tokenize_song(
      notes, event_vocabulary_of(model), inputs_length=lengths['inputs'],
      frames_per_segment=lengths['targets'], frame_rate=ac.frame_rate, sample_rate=ac.sample_rate,
      hop_size=ac.hop_size)
  nseg = len(toks.tokens) if max_segments is None else min(max_segments, len(toks.tokens))
  ctx_len = lengths.get('targets_context') or 0
  pred = np.zeros((1, ctx_len, ac.n_dims), np.float32)
  full = np.zeros((1, 0, ac.n_dims), np.float32)
  seconds = []
  for i in range(nseg):
    batch = {
        'encoder_input_tokens': toks.tokens[i:i + 1],
        'encoder_continuous_inputs': pred[:1],
        # first segment: nothing to condition on; later ones: a full chunk of predicted context
        'encoder_continuous_mask': (np.zeros if (i == 0 or always_mask_context) else np.ones)(
            (1, ctx_len), np.int32),
        'decoder_target_tokens': np.zeros((1, lengths['targets'], ac.n_dims), np.float32),
    }
    tick = time.time()
    pred, _ = model.predict(batch, seed=seed)
    if i != 0:
      seconds.append(time.time() - tick)
    full = np.concatenate([full, pred[:1]], axis=1)
  seconds_per_chunk = lengths['targets'] * (ac.hop_size / ac.sample_rate)
  per_chunk = float(np.mean(seconds)) if seconds else float('nan')
  return {
      'full_pred_encoded': full[0],
      'num_frames': min(toks.num_frames, nseg * lengths['targets']),
      'tokens': toks.tokens[:nseg],
      'model_timing': {
          'prediction_seconds_per_chunk': per_chunk,
          'predictions_seconds_per_audio_second': per_chunk / seconds_per_chunk,
      },
  }
