"""Standard MIDI file reader + sustain pedal (restated third-party behaviour, parity unpinned:
see midi_file.py) and the host part of the full-song driver."""
import numpy as np
import pytest

from music_spectrogram_diffusion_b200 import midi_file as F, midi_tokens as M, song


def test_hand_assembled_file_running_status_and_tempo_change():
  # format 0, 96 ticks per quarter; tempo 500000 us (120 bpm) then 250000 us from tick 96
  track = bytes([
      0x00, 0xFF, 0x51, 0x03, 0x07, 0xA1, 0x20,      # tempo 500000
      0x00, 0xC0, 0x28,                              # program 40 on channel 0
      0x00, 0x90, 0x3C, 0x64,                        # note on 60 vel 100 at tick 0
      0x60, 0x3E, 0x50,                              # running status: note on 62 vel 80 at tick 96
      0x00, 0xFF, 0x51, 0x03, 0x03, 0xD0, 0x90,      # tempo 250000 at tick 96
      0x60, 0x90, 0x3C, 0x00,                        # note on vel 0 == note off 60 at tick 192
      0x81, 0x40, 0x80, 0x3E, 0x00,                  # delta 192 (two-byte VLQ): note off 62 at 384
      0x00, 0xFF, 0x2F, 0x00])
  data = b'MThd' + (6).to_bytes(4, 'big') + bytes([0, 0, 0, 1, 0, 96]) + b'MTrk' + len(track).to_bytes(4, 'big') + track
  s = F.read_midi(data)
  assert s.ticks_per_quarter == 96 and len(s.notes) == 2
  n = s.notes[np.argsort(s.notes['pitch'])]
  # tick 96 = 0.5 s; afterwards a tick lasts 250000e-6 / 96 s
  np.testing.assert_allclose(n['start'], [0.0, 0.5])
  np.testing.assert_allclose(n['end'], [0.5 + 96 * 0.25 / 96, 0.5 + 288 * 0.25 / 96])
  assert list(n['velocity']) == [100, 80] and list(n['program']) == [40, 40] and not n['is_drum'].any()
  assert abs(s.total_time - 1.25) < 1e-12
  with pytest.raises(F.MidiError):
    F.read_midi(b'RIFF' + data[4:])


def test_write_read_round_trip_with_drums_and_programs():
  notes = M.make_notes([(0.0, 0.5, 60, 100, 0, False), (0.25, 1.0, 64, 90, 40, False),
                        (0.5, 0.6, 38, 127, 0, True), (1.0, 2.5, 60, 70, 0, False)])
  s = F.read_midi(F.write_midi(notes, ticks_per_quarter=480, bpm=100.0))
  key = lambda a: np.lexsort((a['pitch'], a['start']))
  got, want = s.notes[key(s.notes)], notes[key(notes)]
  np.testing.assert_allclose(got['start'], want['start'], atol=1e-3)
  np.testing.assert_allclose(got['end'], want['end'], atol=1e-3)
  for f in ('pitch', 'velocity', 'program', 'is_drum'):
    np.testing.assert_array_equal(got[f], want[f])


def test_one_note_off_closes_all_earlier_notes_of_the_key():
  notes = M.make_notes([(0.0, 1.0, 60, 100), (0.5, 1.0, 60, 90)])
  s = F.read_midi(F.write_midi(notes))
  assert len(s.notes) == 2
  np.testing.assert_allclose(sorted(s.notes['start']), [0.0, 0.5])
  np.testing.assert_allclose(s.notes['end'], [1.0, 1.0])


def test_sustain_pedal_extends_notes_until_release_or_restrike():
  notes = M.make_notes([(0.0, 0.5, 60, 100), (0.2, 0.6, 64, 100), (1.0, 1.2, 60, 100),
                        (3.0, 3.5, 67, 100)])
  data = F.write_midi(notes, sustain=[(0.1, 0, 127), (2.0, 0, 0), (3.2, 0, 100)])
  s = F.apply_sustain(F.read_midi(data))
  got = {(round(float(a['start']), 3), int(a['pitch'])): round(float(a['end']), 3) for a in s.notes}
  assert got[(0.0, 60)] == 1.0      # held by the pedal until the same pitch is struck again
  assert got[(0.2, 64)] == 2.0      # held until the pedal comes up
  assert got[(1.0, 60)] == 2.0
  assert got[(3.0, 67)] == 3.5      # pedal still down at the end: ends with the last event
  plain = F.read_midi(data)
  assert round(float(plain.notes['end'][0]), 3) == 0.5
  assert song.load_notes(data).shape == (4,)


class _FakeModel:
  """Records what the driver feeds `predict` (stands in for InferenceModel on the CPU)."""

  def __init__(self):
    from music_spectrogram_diffusion_b200 import audio_codecs, inference
    self.audio_codec = audio_codecs.MelGAN()
    self.codec = inference.build_codec(num_velocity_bins=1)
    self.sequence_length = {'inputs': 2048, 'targets': 256, 'targets_context': 256}
    self.calls = []

  def predict(self, batch, seed=0):
    self.calls.append({k: np.array(v) for k, v in batch.items()})
    k = len(self.calls)
    return np.full((1, 256, 128), float(k), np.float32), np.zeros(1, np.float32)


def test_song_driver_chains_context_like_the_colab_loop():
  notes = M.make_notes([(0.5, 11.0, 60, 100, 0, False), (6.0, 6.2, 40, 90, 33, False)])
  m = _FakeModel()
  out = song.synthesize_song(m, notes, seed=3)
  assert out['full_pred_encoded'].shape == (3 * 256, 128) and out['num_frames'] == M.num_song_frames(11.0)
  assert [c['encoder_continuous_mask'].sum() for c in m.calls] == [0, 256, 256]
  np.testing.assert_array_equal(m.calls[1]['encoder_continuous_inputs'], np.full((1, 256, 128), 1.0))
  np.testing.assert_array_equal(m.calls[2]['encoder_continuous_inputs'], np.full((1, 256, 128), 2.0))
  assert all(c['encoder_input_tokens'].shape == (1, 2048) and c['encoder_input_tokens'].dtype == np.int32
             for c in m.calls)
  np.testing.assert_array_equal(out['full_pred_encoded'][256:512], 2.0)
  t = out['model_timing']
  assert t['prediction_seconds_per_chunk'] >= 0
  assert abs(t['predictions_seconds_per_audio_second'] * 5.12 - t['prediction_seconds_per_chunk']) < 1e-9
  # second segment opens with the tie section of the long note
  v = song.event_vocabulary_of(m)
  assert list(m.calls[1]['encoder_input_tokens'][0, :3]) == [
      v.encode('program', 0) + 3, v.encode('pitch', 60) + 3, v.encode('tie', 0) + 3]
