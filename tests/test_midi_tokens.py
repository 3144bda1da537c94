"""Notes -> tokens (CPU).  The known answers are the ones the reference's own tests hold
(event_codec_test.py, vocabularies_test.py, run_length_encoding_test.py,
note_sequences_test.py:41-287), re-expressed without note_seq / TensorFlow; the rest are
cross-checks of the array formulation against a step-by-step simulation."""
import numpy as np
import pytest

from music_spectrogram_diffusion_b200 import midi_tokens as M

# the vocabulary the reference's tests use (note_sequences_test.py:25-37): other block order
# than the shipped one, so both layouts are exercised
TEST_VOCAB = M.EventVocabulary(100, 100, [('pitch', 0, 127), ('velocity', 0, 127), ('drum', 0, 127),
                                          ('program', 0, 127), ('tie', 0, 0)])


def test_event_vocabulary_known_answers():
  v = M.EventVocabulary(100, 100, [('pitch', 0, 127)])
  ids = [v.encode('pitch', 60), v.encode('shift', 5), v.encode('pitch', 62)]
  assert ids == [161, 5, 163]
  assert [v.decode(i) for i in ids] == [('pitch', 60), ('shift', 5), ('pitch', 62)]
  assert v.max_shift_steps == 100
  assert [v.is_shift(i) for i in (-1, 0, 100, 101)] == [False, True, True, False]
  with pytest.raises(ValueError, match='not within valid range'):
    v.encode('pitch', 128)
  with pytest.raises(ValueError, match='Unknown event type'):
    v.encode('drum', 1)
  with pytest.raises(ValueError, match='Unknown event index'):
    v.decode(229)


def test_shipped_vocabulary_layout():
  v = M.mt3_event_vocabulary(M.VocabularyConfig(num_velocity_bins=1))
  assert v.num_classes == 1388                      # 1001 + 128 + 2 + 1 + 128 + 128
  assert M.num_embeddings(v.num_classes) == 1536
  assert v.id_range('shift') == (0, 1000) and v.id_range('pitch') == (1001, 1128)
  assert v.id_range('velocity') == (1129, 1130) and v.id_range('tie') == (1131, 1131)
  assert v.id_range('program') == (1132, 1259) and v.id_range('drum') == (1260, 1387)
  assert M.num_embeddings(M.mt3_event_vocabulary().num_classes) == 1664


def test_velocity_quantisation_round_trips():
  for bins in (1, 127):
    assert M.velocity_to_bin(0, bins) == 0 and M.bin_to_velocity(0, bins) == 0
  assert M.velocity_to_bin(M.bin_to_velocity(1, 1), 1) == 1
  for b in range(1, 128):
    assert M.velocity_to_bin(M.bin_to_velocity(b, 127), 127) == b


def test_model_id_offsets_and_special_ids():
  np.testing.assert_array_equal(M.to_model_ids([1, 2, 3], 32), [4, 5, 6])
  np.testing.assert_array_equal(M.from_model_ids([4, 5, 6], 32), [1, 2, 3])
  # PAD / UNK / ids beyond the regular block decode to "invalid" (extra ids included)
  np.testing.assert_array_equal(M.from_model_ids([0, 2, 3, 4, 34, 35], 32), [-2, -2, 0, 1, 31, -2])
  enc = [0, 2, 3, 4, 1, 0, 1, 0]
  np.testing.assert_array_equal(M.from_model_ids(enc, 32), [-2, -2, 0, 1, -1])
  np.testing.assert_array_equal(M.from_model_ids(enc, 32, keep_length=True), [-2, -2, 0, 1, -1, -1, -1, -1])
  M.to_model_ids([0, 15, 31], 32)
  for bad in ([-1, 15, 31], [0, 15, 32]):
    with pytest.raises(ValueError, match='does not fall within valid range'):
      M.to_model_ids(bad, 32)


def test_run_length_encode_shifts_known_answers():
  r = lambda ev, kinds=(): list(M.run_length_encode_shifts(ev, TEST_VOCAB, kinds))
  assert r([1, 1, 1, 161, 1, 1, 1, 162, 1, 1, 1]) == [3, 161, 6, 162]
  assert r([1] * 202 + [161, 1, 1, 1]) == [100, 100, 2, 161]
  assert r([1, 1, 1, 161, 162, 1, 1, 1]) == [3, 161, 162]
  assert r([1, 1, 1, 525, 356, 161, 1, 1, 525, 356, 161, 355, 394], ('velocity', 'program')) == [
      3, 525, 356, 161, 5, 161, 355, 394]
  assert r([]) == [] and r([1, 1]) == []


def test_onsets_only_known_answer():
  notes = M.make_notes([(1.0, 1.1, 61, 100), (2.0, 2.1, 62, 100), (3.0, 3.1, 63, 100)])
  ft = np.arange(0, 4, step=.001)
  ix = M.encode_and_index_events(M.timed_note_events(notes, 'onsets'), TEST_VOCAB, ft)
  want = [1] * 100 + [162] + [1] * 100 + [163] + [1] * 100 + [164] + [1] * 100
  np.testing.assert_array_equal(ix.events, want)
  assert len(ix.event_start_indices) == len(ix.event_end_indices) == len(ft)
  assert (ix.event_start_indices[0], ix.event_end_indices[0]) == (0, 0)
  for frame, idx in ((1000, 100), (2000, 201), (3000, 302)):
    assert ix.event_start_indices[frame] == idx and ix.event_end_indices[frame] == idx
  assert ix.event_start_indices[-1] == 402 and ix.event_end_indices[-1] == len(want)


def test_onsets_offsets_velocities_known_answer():
  notes = M.make_notes([(1.0, 3.0, 61, 1), (2.0, 4.0, 62, 127)])
  ft = np.arange(0, 4, step=.001)
  ix = M.encode_and_index_events(M.timed_note_events(notes, 'onsets_offsets'), TEST_VOCAB, ft)
  want = ([1] * 100 + [230, 162] + [1] * 100 + [356, 163] + [1] * 100 + [229, 162] + [1] * 100 +
          [229, 163])
  np.testing.assert_array_equal(ix.events, want)
  for frame, idx in ((0, 0), (1000, 100), (2000, 202), (3000, 304)):
    assert ix.event_start_indices[frame] == idx and ix.event_end_indices[frame] == idx
  assert ix.event_start_indices[-1] == 405 and ix.event_end_indices[-1] == len(want)


def test_multitrack_with_tie_state_known_answer():
  notes = M.make_notes([(0.0, 1.0, 37, 127, 0, True), (1.0, 3.0, 61, 127, 0, False),
                        (2.0, 4.0, 62, 127, 40, False)])
  ft = np.arange(0, 4, step=.001)
  ix = M.encode_and_index_events(M.timed_note_events(notes), TEST_VOCAB, ft, with_tie_state=True)
  e = TEST_VOCAB.encode
  shift = [e('shift', 1)] * 100
  want = ([e('velocity', 127), e('drum', 37)] + shift +
          [e('program', 0), e('velocity', 127), e('pitch', 61)] + shift +
          [e('program', 40), e('velocity', 127), e('pitch', 62)] + shift +
          [e('program', 0), e('velocity', 0), e('pitch', 61)] + shift +
          [e('program', 40), e('velocity', 0), e('pitch', 62)])
  assert len(ix.events) == 414
  np.testing.assert_array_equal(ix.events, want)
  tie = e('tie', 0)
  want_state = [tie, tie,
                e('program', 0), e('pitch', 61), tie,
                e('program', 0), e('pitch', 61), e('program', 40), e('pitch', 62), tie,
                e('program', 40), e('pitch', 62), tie]
  np.testing.assert_array_equal(ix.state_events, want_state)
  for frame, ev_idx, st_idx in ((0, 0, 0), (1000, 102, 1), (2000, 205, 2), (3000, 308, 5)):
    assert ix.event_start_indices[frame] == ev_idx == ix.event_end_indices[frame]
    assert ix.state_event_indices[frame] == st_idx
  assert ix.event_start_indices[-1] == 410 and ix.event_end_indices[-1] == len(want)
  assert ix.state_event_indices[-1] == 10


def test_last_frame_is_covered_by_an_extra_shift():
  notes = M.make_notes([(0.0, 0.1, 60, 100)])
  ft = np.arange(0, 1.008, step=.008)
  ix = M.encode_and_index_events(M.timed_note_events(notes, 'onsets'), TEST_VOCAB, ft)
  np.testing.assert_array_equal(ix.events, [161] + [1] * 101)
  assert (ix.event_start_indices[0], ix.event_end_indices[0]) == (0, 0)
  assert (ix.event_start_indices[125], ix.event_end_indices[125]) == (101, 102)


def _simulate(ev, vocab, frame_times, with_state):
  """Step-by-step simulation of the same stream (one shift at a time), for cross-checking."""
  order = np.argsort(ev.time, kind='stable')
  sps = vocab.steps_per_second
  nb = M.num_velocity_bins_of(vocab)
  out, state_out, starts, sidx = [], [], [], []
  sounding = {}
  step = 0
  mark = smark = 0

  def flush_frames():
    while len(starts) < len(frame_times) and frame_times[len(starts)] < step / sps:
      starts.append(mark)
      sidx.append(smark)

  for i in order:
    target = round(float(ev.time[i]) * sps)
    while step < target:
      out.append(vocab.encode('shift', 1))
      step += 1
      flush_frames()
      mark, smark = len(out), len(state_out)
    if with_state:
      for (prog, pitch) in sorted(k for k, b in sounding.items() if b):
        state_out += [vocab.encode('program', prog), vocab.encode('pitch', pitch)]
      state_out.append(vocab.encode('tie', 0))
    vel, prog, pitch = int(ev.velocity[i]), int(ev.program[i]), int(ev.pitch[i])
    if vel < 0:
      out.append(vocab.encode('pitch', pitch))
      continue
    b = M.velocity_to_bin(vel, nb)
    if prog >= 0 and ev.is_drum[i]:
      out += [vocab.encode('velocity', b), vocab.encode('drum', pitch)]
      continue
    if prog >= 0:
      out.append(vocab.encode('program', prog))
    out += [vocab.encode('velocity', b), vocab.encode('pitch', pitch)]
    sounding[(max(prog, 0), pitch)] = b
  while step / sps <= frame_times[-1]:
    out.append(vocab.encode('shift', 1))
    step += 1
    flush_frames()
    mark = len(out)
  return out, starts, starts[1:] + [len(out)], state_out, sidx


@pytest.mark.parametrize('seed', range(6))
@pytest.mark.parametrize('mode', ['onsets', 'onsets_offsets', 'onsets_offsets_programs'])
def test_array_formulation_matches_step_simulation(seed, mode):
  rng = np.random.default_rng(seed)
  n = int(rng.integers(0, 60))
  start = np.round(rng.uniform(0, 6.0, n), int(rng.integers(1, 4)))
  dur = np.round(rng.uniform(0.0, 1.5, n), 2)
  rows = [(s, s + d, int(rng.integers(30, 40)), int(rng.integers(1, 128)),
           int(rng.choice([0, 0, 40, 41, 80])), bool(rng.random() < 0.2)) for s, d in zip(start, dur)]
  notes = M.make_notes(rows)
  vocab = M.mt3_event_vocabulary(M.VocabularyConfig(num_velocity_bins=int(rng.choice([1, 127]))))
  frame_rate = float(rng.choice([50.0, 125.0, 31.25]))
  ft = np.arange(int(rng.integers(1, 400))) / frame_rate
  ev = M.timed_note_events(notes, mode)
  with_state = mode == 'onsets_offsets_programs'
  ix = M.encode_and_index_events(ev, vocab, ft, with_tie_state=with_state)
  out, starts, ends, state_out, sidx = _simulate(ev, vocab, ft, with_state)
  np.testing.assert_array_equal(ix.events, out)
  np.testing.assert_array_equal(ix.event_start_indices, starts)
  np.testing.assert_array_equal(ix.event_end_indices, ends)
  if with_state:
    np.testing.assert_array_equal(ix.state_events, state_out)
    np.testing.assert_array_equal(ix.state_event_indices, sidx)


def test_program_granularities_and_slakh_mapping():
  v = M.mt3_event_vocabulary()
  ev = np.array([v.encode('program', 43), v.encode('velocity', 3), v.encode('pitch', 60)])
  np.testing.assert_array_equal(M.map_programs(ev, v, 'full'), ev)
  np.testing.assert_array_equal(M.map_programs(ev, v, 'midi_class'),
                                [v.encode('program', 40), ev[1], ev[2]])
  np.testing.assert_array_equal(M.map_programs(ev, v, 'flat'), ev[1:])
  assert [M.program_to_slakh_program(p) for p in (0, 3, 4, 25, 27, 44, 79, 80, 127)] == [
      0, 0, 4, 24, 26, 43, 73, 80, 88]


def test_song_frame_count_follows_the_reference_padding():
  assert M.num_song_frames(5.12) == 257            # exact multiple of the hop: a full extra frame
  assert M.num_song_frames(5.119) == 256
  assert M.num_song_frames(0.0) == 1


def test_tokenize_song_segments_ties_and_limits():
  # a long piano note crossing the first segment boundary (5.12 s), a drum hit, a later violin note
  notes = M.make_notes([(1.0, 7.0, 60, 100, 3, False), (2.0, 2.1, 38, 90, 0, True),
                        (6.0, 6.5, 72, 64, 41, False)])
  v = M.mt3_event_vocabulary(M.VocabularyConfig(num_velocity_bins=1))
  song = M.tokenize_song(notes, v)
  assert song.num_frames == M.num_song_frames(7.0) == 351 and song.tokens.shape == (2, 2048)
  e = lambda k, x: v.encode(k, x) + M.NUM_SPECIAL_IDS
  first = list(song.tokens[0, :song.lengths[0]])
  assert first == [e('tie', 0), e('shift', 100), e('program', 0), e('velocity', 1), e('pitch', 60),
                   e('shift', 200), e('drum', 38), M.EOS_ID]     # redundant velocity dropped
  second = list(song.tokens[1, :song.lengths[1]])
  # tie section lists the still-sounding piano note, then events relative to the segment start
  assert second == [e('program', 0), e('pitch', 60), e('tie', 0),
                    e('shift', 88), e('program', 41), e('velocity', 1), e('pitch', 72),
                    e('shift', 138), e('velocity', 0), e('pitch', 72),
                    e('shift', 188), e('program', 0), e('pitch', 60), M.EOS_ID]
  assert not song.tokens[0, song.lengths[0]:].any()
  assert song.tokens.max() < M.num_embeddings(v.num_classes)
  # without ties / with program classes dropped
  plain = M.tokenize_song(notes, v, include_ties=False, program_granularity='flat')
  assert list(plain.tokens[1, :3]) == [e('shift', 88), e('velocity', 1), e('pitch', 72)]
  # too many events for one segment is an error, as in handle_too_long(skip=False)
  dense = M.make_notes([(0.01 * i, 0.01 * i + 0.005, 40 + i % 40, 100, 0, False) for i in range(500)])
  with pytest.raises(ValueError, match='exceeds maximum length'):
    M.tokenize_song(dense, v)
  assert M.tokenize_song(M.make_notes([]), v).tokens.shape == (1, 2048)


# ---- tokens -> notes: known answers of note_sequences_test.py:289-504 ------------------------
def _decode(events, onsets_only, start=0.0, max_time=None):
  d = M.NoteDecoder(TEST_VOCAB, onsets_only=onsets_only)
  invalid, dropped = d.feed(events, start_time=start, max_time=max_time)
  notes = d.finish()
  rows = [(round(float(n['start']), 6), round(float(n['end']), 6), int(n['pitch']), int(n['velocity']),
           int(n['program']), bool(n['is_drum'])) for n in notes]
  return rows, invalid, dropped


def test_decode_known_answers():
  assert _decode([25, 161, 50, 162], True) == (
      [(0.25, 0.26, 60, 100, 0, False), (0.50, 0.51, 61, 100, 0, False)], 0, 0)
  assert _decode([5, 161, 25, 162], True) == (
      [(0.05, 0.06, 60, 100, 0, False), (0.25, 0.26, 61, 100, 0, False)], 0, 0)
  assert _decode([5, 356, 161, 25, 229, 161], False) == ([(0.05, 0.25, 60, 127, 0, False)], 0, 0)
  # a second onset without a note-off ends the first note where the second begins
  assert _decode([5, 356, 161, 10, 161, 25, 229, 161], False) == (
      [(0.05, 0.10, 60, 127, 0, False), (0.10, 0.25, 60, 127, 0, False)], 0, 0)
  rows, invalid, dropped = _decode([5, 525, 356, 161, 15, 356, 394, 25, 525, 229, 161], False)
  assert (invalid, dropped) == (0, 0)
  assert sorted(rows) == [(0.05, 0.25, 60, 127, 40, False), (0.15, 0.16, 37, 127, 0, True)]


def test_decode_invalid_and_dropped_events():
  assert _decode([5, -1, 161, -2, 25, 162, 9999], True) == (
      [(0.05, 0.06, 60, 100, 0, False), (0.25, 0.26, 61, 100, 0, False)], 3, 0)
  assert _decode([161, 25, 162], True, start=1.0, max_time=1.25) == (
      [(1.00, 1.01, 60, 100, 0, False), (1.25, 1.26, 61, 100, 0, False)], 0, 0)
  assert _decode([5, 161, 30, 162], True, start=1.0, max_time=1.25) == (
      [(1.05, 1.06, 60, 100, 0, False)], 0, 2)
  assert _decode([25, 230, 50, 161], True) == ([(0.50, 0.51, 60, 100, 0, False)], 1, 0)


@pytest.mark.parametrize('seed', range(5))
def test_tokenize_decode_round_trip(seed):
  """notes -> per-segment token rows -> notes gives back the song on the 10 ms grid."""
  rng = np.random.default_rng(100 + seed)
  rows, cursor = [], {}
  for _ in range(int(rng.integers(5, 80))):
    prog = int(rng.choice([0, 24, 40, 73]))
    pitch = int(rng.integers(40, 90))
    t0 = max(cursor.get((prog, pitch), 0.0), float(rng.uniform(0, 14))) + 0.02
    t0 = round(t0, 2)
    t1 = round(t0 + float(rng.uniform(0.03, 3.0)), 2)
    cursor[(prog, pitch)] = t1                       # no overlapping notes on one (program, pitch)
    rows.append((t0, t1, pitch, int(rng.integers(1, 128)), prog, False))
  for _ in range(int(rng.integers(0, 10))):
    t0 = round(float(rng.uniform(0, 14)), 2)
    rows.append((t0, t0 + 0.01, int(rng.integers(35, 50)), int(rng.integers(1, 128)), 0, True))
  notes = M.make_notes(rows)
  vocab = M.mt3_event_vocabulary()
  song = M.tokenize_song(notes, vocab, map_to_slakh_programs=False)
  back, invalid, dropped = M.decode_song(song.tokens, vocab)
  assert (invalid, dropped) == (0, 0) and len(back) == len(notes)
  key = lambda a: np.lexsort((a['pitch'], a['program'], a['is_drum'], np.round(a['start'], 2)))
  a, b = notes[key(notes)], back[key(back)]
  np.testing.assert_allclose(b['start'], a['start'], atol=1e-9)
  pitched = ~a['is_drum']
  np.testing.assert_allclose(b['end'][pitched], a['end'][pitched], atol=1e-9)
  for f in ('pitch', 'velocity', 'program', 'is_drum'):
    np.testing.assert_array_equal(a[f], b[f])


def test_run_length_encoding_preserves_event_times_property():
  """For any stream of single-step shifts and events: every kept event keeps its time since the
  segment start, runs never exceed max_shift_steps, and no trailing shift survives."""
  hypothesis = pytest.importorskip('hypothesis')
  from hypothesis import given, settings, strategies as st
  vocab = TEST_VOCAB
  shift1 = vocab.encode('shift', 1)
  lo, hi = vocab.id_range('pitch')

  @settings(max_examples=200, deadline=None)
  @given(st.lists(st.one_of(st.just(shift1), st.integers(lo, hi)), max_size=400))
  def check(stream):
    out = M.run_length_encode_shifts(stream, vocab)
    # times of the non-shift events in the input
    t, want = 0, []
    for tok in stream:
      if tok == shift1:
        t += 1
      else:
        want.append((t, tok))
    # times of the events in the output (absolute runs, re-armed by every event)
    got, run = [], 0
    last_time = 0
    for tok in out:
      if vocab.is_shift(int(tok)):
        assert 1 <= tok <= vocab.max_shift_steps
        run += int(tok)
      else:
        if run:
          last_time = run
        got.append((last_time, int(tok)))
        run = 0
    assert got == want
    assert not len(out) or not vocab.is_shift(int(out[-1]))

  check()
