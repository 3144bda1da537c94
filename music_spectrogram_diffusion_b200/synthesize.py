"""Command-line form of the colab notebook: MIDI file in, predicted log-mel frames out.

  python -m music_spectrogram_diffusion_b200.synthesize tune.mid out.npy \\
      --checkpoint base_with_context/checkpoint_500000 --gin base_with_context/config.gin

Without --checkpoint the weights are synthetic (`synthetic:0`), which exercises the whole pipeline
but produces noise-like mel frames.  The mel -> audio vocoder is not part of this package.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Optional, Sequence

import numpy as np

_DEFAULT_GIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests',
                            'golden', 'base_with_context.gin')


def build_parser() -> argparse.ArgumentParser:
  ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
  ap.add_argument('midi', help='standard MIDI file (format 0 or 1)')
  ap.add_argument('out', help='output .npy: float32 [frames, 128] log-mel (50 frames per second)')
  ap.add_argument('--checkpoint', default='synthetic:0',
                  help='T5X checkpoint directory, .npz, or synthetic:<seed>')
  ap.add_argument('--gin', default=_DEFAULT_GIN, help='training config.gin of the checkpoint')
  ap.add_argument('--gin-binding', action='append', default=[],
                  help='extra gin binding, e.g. "diffusion_utils.SamplerConfig.name = \'ddim\'"')
  ap.add_argument('--cond-weight', type=float, default=2.0,
                  help='classifier-free guidance weight (the colab uses 2.0)')
  ap.add_argument('--seed', type=int, default=0)
  ap.add_argument('--device', type=int, default=0)
  ap.add_argument('--no-sustain', action='store_true', help='ignore sustain-pedal controller events')
  ap.add_argument('--max-segments', type=int, default=None)
  return ap


def main(argv: Optional[Sequence[str]] = None) -> int:
  args = build_parser().parse_args(argv)
  from music_spectrogram_diffusion_b200 import inference, song
  bindings = [f'diffusion_utils.ClassifierFreeGuidanceConfig.eval_condition_weight = {args.cond_weight}']
  gin_config = inference.parse_training_gin_file(args.gin, bindings + list(args.gin_binding))
  model = inference.InferenceModel(args.checkpoint, gin_config, batch_size=1, device=args.device)
  notes = song.load_notes(args.midi, sustain=not args.no_sustain)
  result = song.synthesize_song(model, notes, seed=args.seed, max_segments=args.max_segments)
  mel = result['full_pred_encoded'][:result['num_frames']]
  np.save(args.out, mel.astype(np.float32))
  timing = result['model_timing']
  print(json.dumps({'notes': int(len(notes)), 'segments': int(len(result['tokens'])),
                    'frames': int(mel.shape[0]), 'seconds_of_audio': mel.shape[0] / 50.0,
                    'prediction_seconds_per_chunk': timing['prediction_seconds_per_chunk'],
                    'x_realtime': (1.0 / timing['predictions_seconds_per_audio_second']
                                   if timing['predictions_seconds_per_audio_second'] else None)}))
  return 0


if __name__ == '__main__':
  sys.exit(main())
