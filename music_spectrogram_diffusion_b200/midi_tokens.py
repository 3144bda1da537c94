"""Notes -> encoder input tokens of the synthesis model, on the CPU, with numpy only.

This is the caller side of the hot path: what the colab does between
`note_seq.midi_to_note_sequence` and `InferenceModel.predict` (ipynb "Setup Synthesis Task"),
restated without note_seq / seqio / t5 / TensorFlow.  Behaviour follows, and is pinned in
tests/test_midi_tokens.py by the known answers of, the reference's own tests:

  event vocabulary layout, shift block first        msd/event_codec.py:35-112 (event_codec_test.py)
  mt3 vocabulary, velocity bins, id offset 3 + EOS   msd/vocabularies.py:56-144, 147-256
                                                     (vocabularies_test.py)
  timed note events and their tie-break order        msd/note_sequences.py:139-211
  note event -> tokens, tie-section state            msd/note_sequences.py:214-262
  single-step shifts + frame indexing                msd/run_length_encoding.py:62-166
                                                     (note_sequences_test.py:41-287)
  tie-section prefix of a segment                    msd/run_length_encoding.py:169-194
  shift run-length coding, redundant state changes   msd/run_length_encoding.py:197-271
                                                     (run_length_encoding_test.py)
  256-frame segment grid of a full song              msd/preprocessors.py:60-81, 863-921
  program -> Slakh class program                     msd/preprocessors.py:440-476, ipynb:701-707

The implementation is array based (prefix sums over step-sorted events instead of the
reference's token-by-token loops), so a song tokenises in milliseconds.
"""

from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MAX_MIDI_VELOCITY = 127
PAD_ID, EOS_ID, UNK_ID = 0, 1, 2
NUM_SPECIAL_IDS = 3          # GenericTokenVocabulary: regular id r is stored as r + 3
DECODED_EOS_ID, DECODED_INVALID_ID = -1, -2

NOTE_DTYPE = np.dtype([('start', 'f8'), ('end', 'f8'), ('pitch', 'i4'), ('velocity', 'i4'),
                       ('program', 'i4'), ('is_drum', '?')])


def make_notes(rows: Sequence[Sequence]) -> np.ndarray:
  """rows of (start, end, pitch, velocity[, program[, is_drum]]) -> structured note array."""
  out = np.zeros(len(rows), dtype=NOTE_DTYPE)
  for i, r in enumerate(rows):
    r = tuple(r)
    out[i] = r + (0, False)[len(r) - 4:] if len(r) < 6 else r
  return out


# ---------------------------------------------------------------------------------------------
# event vocabulary
# ---------------------------------------------------------------------------------------------
class EventVocabulary:
  """Contiguous id blocks per event kind; 'shift' always owns the first block, starting at 0."""

  def __init__(self, max_shift_steps: int, steps_per_second: float,
               ranges: Sequence[Tuple[str, int, int]]):
    blocks = [('shift', 0, int(max_shift_steps))] + [(k, int(lo), int(hi)) for k, lo, hi in ranges]
    kinds = [b[0] for b in blocks]
    if len(set(kinds)) != len(kinds):
      raise ValueError(f'event kinds must be unique: {kinds}')
    self.steps_per_second = steps_per_second
    self.kinds = kinds
    self._lo = np.array([b[1] for b in blocks], np.int64)
    self._hi = np.array([b[2] for b in blocks], np.int64)
    sizes = self._hi - self._lo + 1
    self._base = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    self._end = self._base + sizes            # one past the last id of each block
    self._slot = {k: i for i, k in enumerate(kinds)}

  @property
  def num_classes(self) -> int:
    return int(self._end[-1])

  @property
  def max_shift_steps(self) -> int:
    return int(self._hi[0])

  def is_shift(self, index) -> bool:
    return bool(self._lo[0] <= index <= self._hi[0])

  def _slot_of(self, kind: str) -> int:
    if kind not in self._slot:
      raise ValueError(f'Unknown event type: {kind}')
    return self._slot[kind]

  def encode(self, kind: str, value: int) -> int:
    s = self._slot_of(kind)
    if not self._lo[s] <= value <= self._hi[s]:
      raise ValueError(f'Event value {value} is not within valid range '
                       f'[{self._lo[s]}, {self._hi[s]}] for type {kind}')
    return int(self._base[s] + value - self._lo[s])

  def encode_array(self, kind: str, values) -> np.ndarray:
    s = self._slot_of(kind)
    values = np.asarray(values, np.int64)
    if values.size and (values.min() < self._lo[s] or values.max() > self._hi[s]):
      raise ValueError(f'Event value outside [{self._lo[s]}, {self._hi[s]}] for type {kind}')
    return self._base[s] + values - self._lo[s]

  def id_range(self, kind: str) -> Tuple[int, int]:
    """[first id, last id] of a kind."""
    s = self._slot_of(kind)
    return int(self._base[s]), int(self._end[s] - 1)

  def decode(self, index: int) -> Tuple[str, int]:
    s = int(np.searchsorted(self._end, index, side='right'))
    if index < 0 or s >= len(self.kinds):
      raise ValueError(f'Unknown event index: {index}')
    return self.kinds[s], int(self._lo[s] + index - self._base[s])

  # ---- the method names of the reference's event_codec.Codec (event_codec.py:64-112), so that
  # `InferenceModel.codec` can be used where callers expect that object (inference.py:110-111)
  def is_shift_event_index(self, index: int) -> bool:
    return self.is_shift(index)

  def encode_event(self, event) -> int:
    """event: anything with .type and .value (event_codec.Event) or a (type, value) pair."""
    kind, value = (event.type, event.value) if hasattr(event, 'type') else event
    return self.encode(kind, value)

  def event_type_range(self, event_type: str) -> Tuple[int, int]:
    return self.id_range(event_type)

  def decode_event_index(self, index: int) -> 'Event':
    return Event(*self.decode(index))

  @property
  def num_velocity_bins(self) -> int:
    lo, hi = self.id_range('velocity')
    return hi - lo

  @property
  def max_shift_seconds(self) -> int:
    return int(round(self.max_shift_steps / self.steps_per_second))


@dataclasses.dataclass(frozen=True)
class Event:
  """event_codec.Event (event_codec.py:28-31)."""
  type: str
  value: int


@dataclasses.dataclass
class VocabularyConfig:
  steps_per_second: int = 100
  max_shift_seconds: int = 10
  num_velocity_bins: int = 127


def mt3_event_vocabulary(cfg: VocabularyConfig = VocabularyConfig()) -> EventVocabulary:
  """The block order the checkpoints were trained with: shift, pitch, velocity (bin 0 = note
  off), tie, program, drum."""
  return EventVocabulary(
      cfg.steps_per_second * cfg.max_shift_seconds, cfg.steps_per_second,
      [('pitch', 0, 127), ('velocity', 0, cfg.num_velocity_bins), ('tie', 0, 0),
       ('program', 0, 127), ('drum', 0, 127)])


def num_velocity_bins_of(vocab: EventVocabulary) -> int:
  lo, hi = vocab.id_range('velocity')
  return hi - lo


def velocity_to_bin(velocity: int, num_velocity_bins: int) -> int:
  return 0 if velocity == 0 else math.ceil(num_velocity_bins * velocity / MAX_MIDI_VELOCITY)


def bin_to_velocity(velocity_bin: int, num_velocity_bins: int) -> int:
  return 0 if velocity_bin == 0 else int(MAX_MIDI_VELOCITY * velocity_bin / num_velocity_bins)


def num_embeddings(num_classes: int, extra_ids: int = 100) -> int:
  """Embedding rows: specials + classes + extra ids, rounded up to a multiple of 128."""
  return 128 * math.ceil((NUM_SPECIAL_IDS + num_classes + extra_ids) / 128)


def to_model_ids(event_ids: Sequence[int], num_classes: int) -> np.ndarray:
  """GenericTokenVocabulary.encode: shift regular ids past PAD/EOS/UNK (no EOS appended)."""
  ids = np.asarray(event_ids, np.int64)
  if ids.size and (ids.min() < 0 or ids.max() >= num_classes):
    bad = int(ids[(ids < 0) | (ids >= num_classes)][0])
    raise ValueError(f'token_id {bad} does not fall within valid range of [0, {num_classes})')
  return (ids + NUM_SPECIAL_IDS).astype(np.int32)


def from_model_ids(ids: Sequence[int], num_classes: int, keep_length: bool = False) -> np.ndarray:
  """GenericTokenVocabulary.decode (truncates after the first EOS) or, with keep_length, its
  TensorFlow twin (everything from the first EOS on becomes DECODED_EOS_ID)."""
  ids = np.asarray(ids, np.int64)
  out = np.where((ids >= NUM_SPECIAL_IDS) & (ids < NUM_SPECIAL_IDS + num_classes),
                 ids - NUM_SPECIAL_IDS, DECODED_INVALID_ID)
  eos = np.flatnonzero(ids == EOS_ID)
  if eos.size:
    out[eos[0]:] = DECODED_EOS_ID
    if not keep_length:
      out = out[:eos[0] + 1]
  return out


# ---------------------------------------------------------------------------------------------
# notes -> timed events -> single-step token stream with frame indices
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class TimedEvents:
  """Unsorted note on/off events; velocity < 0 / program < 0 mean "not modelled"."""
  time: np.ndarray
  pitch: np.ndarray
  velocity: np.ndarray
  program: np.ndarray
  is_drum: np.ndarray


def timed_note_events(notes: np.ndarray, mode: str = 'onsets_offsets_programs') -> TimedEvents:
  """mode: 'onsets' | 'onsets_offsets' | 'onsets_offsets_programs'.  The order inside the
  returned arrays is the tie-break for the later stable sort by time: offsets before onsets,
  and within each by pitch (or by (is_drum, program, pitch) when programs are modelled)."""
  n = len(notes)
  minus = np.full(n, -1, np.int64)
  if mode == 'onsets':
    o = np.argsort(notes['pitch'], kind='stable')
    v = notes[o]
    return TimedEvents(v['start'].astype(np.float64), v['pitch'].astype(np.int64), minus, minus,
                       np.zeros(n, bool))
  if mode == 'onsets_offsets':
    o = np.argsort(notes['pitch'], kind='stable')
    v = notes[o]
    return TimedEvents(np.concatenate([v['end'], v['start']]).astype(np.float64),
                       np.concatenate([v['pitch'], v['pitch']]).astype(np.int64),
                       np.concatenate([np.zeros(n, np.int64), v['velocity'].astype(np.int64)]),
                       np.concatenate([minus, minus]), np.zeros(2 * n, bool))
  if mode == 'onsets_offsets_programs':
    o = np.lexsort((notes['pitch'], notes['program'], notes['is_drum']))
    v = notes[o]
    pitched = v[~v['is_drum']]           # drums have no note-off
    m = len(pitched)
    return TimedEvents(np.concatenate([pitched['end'], v['start']]).astype(np.float64),
                       np.concatenate([pitched['pitch'], v['pitch']]).astype(np.int64),
                       np.concatenate([np.zeros(m, np.int64), v['velocity'].astype(np.int64)]),
                       np.concatenate([pitched['program'], v['program']]).astype(np.int64),
                       np.concatenate([np.zeros(m, bool), v['is_drum']]))
  raise ValueError(f'unknown mode {mode!r}')


@dataclasses.dataclass
class IndexedEvents:
  events: np.ndarray               # event ids, time as runs of single-step shifts
  event_start_indices: np.ndarray  # per frame: first event of the frame
  event_end_indices: np.ndarray    # per frame: == start index of the next frame
  state_events: np.ndarray         # active-note listings ("tie sections"), each ending in 'tie'
  state_event_indices: np.ndarray  # per frame: where its listing starts in state_events


def encode_and_index_events(ev: TimedEvents, vocab: EventVocabulary, frame_times: Sequence[float],
                            with_tie_state: bool = False) -> IndexedEvents:
  """Token stream of `ev` with one shift token per time step, and for every audio frame the
  slice of the stream it owns.  A frame at time t starts at the token right after shift number
  k - 1, k being the first step with t < k / steps_per_second; the stream is extended with
  shifts until the last frame is covered."""
  sps = vocab.steps_per_second
  frame_times = np.asarray(frame_times, np.float64)
  order = np.argsort(ev.time, kind='stable')
  steps = np.rint(ev.time[order] * sps).astype(np.int64)
  pitch, vel, prog, drum = ev.pitch[order], ev.velocity[order], ev.program[order], ev.is_drum[order]
  n = len(order)

  # tokens of each event: [program] [velocity bin] pitch|drum
  nbins = num_velocity_bins_of(vocab) if (vel >= 0).any() else 0
  vbin = np.where(vel > 0, np.ceil(nbins * vel / MAX_MIDI_VELOCITY).astype(np.int64), 0)
  has_vel = vel >= 0
  has_prog = has_vel & (prog >= 0) & ~drum
  count = 1 + has_vel.astype(np.int64) + has_prog.astype(np.int64)
  first = np.concatenate([[0], np.cumsum(count)[:-1]]) if n else np.zeros(0, np.int64)
  flat = np.zeros(int(count.sum()), np.int64)
  if has_prog.any():
    flat[first[has_prog]] = vocab.encode_array('program', prog[has_prog])
  if has_vel.any():
    flat[first[has_vel] + has_prog[has_vel]] = vocab.encode_array('velocity', vbin[has_vel])
  is_drum_tok = has_vel & (prog >= 0) & drum
  last = first + count - 1
  if is_drum_tok.any():
    flat[last[is_drum_tok]] = vocab.encode_array('drum', pitch[is_drum_tok])
  if (~is_drum_tok).any():
    flat[last[~is_drum_tok]] = vocab.encode_array('pitch', pitch[~is_drum_tok])

  # number of shifts: up to the last event, and until the last frame is covered
  cs = np.maximum(steps, 0)
  k_frame = np.floor(frame_times * sps).astype(np.int64)
  k_frame = np.where(frame_times < k_frame / sps, k_frame, k_frame + 1)
  k_frame = np.where(frame_times < k_frame / sps, k_frame, k_frame + 1)
  k_frame = np.where(frame_times < (k_frame - 1) / sps, k_frame - 1, k_frame)
  k_frame = np.maximum(k_frame, 1)
  last_step = int(cs[-1]) if n else 0
  total_shifts = max(last_step, int(k_frame[-1]))

  shift_id = vocab.encode('shift', 1)
  events = np.full(total_shifts + len(flat), shift_id, np.int64)
  tok_event = np.repeat(np.arange(n), count)                   # owning event of each flat token
  events[np.arange(len(flat)) + cs[tok_event]] = flat          # cs shifts precede event's tokens
  tokens_before_step = np.concatenate([[0], np.cumsum(count)])[np.searchsorted(cs, np.arange(total_shifts + 1), 'left')]
  pos_after_shift = np.arange(total_shifts + 1) + tokens_before_step
  starts = pos_after_shift[k_frame - 1]
  ends = np.concatenate([starts[1:], [len(events)]])

  state_tokens = np.zeros(0, np.int64)
  state_idx = np.zeros(len(frame_times), np.int64)
  if with_tie_state:
    tie_id = vocab.encode('tie', 0)
    prog_base = vocab.id_range('program')[0]
    pitch_base = vocab.id_range('pitch')[0]
    active: Dict[Tuple[int, int], int] = {}       # (program, pitch) -> velocity bin, 0 = released
    chunks: List[np.ndarray] = []
    dumped = np.zeros(n + 1, np.int64)            # state tokens emitted before event i
    cached: Optional[np.ndarray] = None
    for i in range(n):
      if cached is None:
        keys = sorted(k for k, b in active.items() if b)
        cached = np.empty(2 * len(keys) + 1, np.int64)
        if keys:
          ka = np.array(keys, np.int64)
          cached[0:-1:2] = prog_base + ka[:, 0]
          cached[1:-1:2] = pitch_base + ka[:, 1]
        cached[-1] = tie_id
      chunks.append(cached)
      dumped[i + 1] = dumped[i] + len(cached)
      if has_vel[i] and not is_drum_tok[i]:
        key = (int(prog[i]) if prog[i] >= 0 else 0, int(pitch[i]))
        if active.get(key, 0) != int(vbin[i]):
          cached = None
        active[key] = int(vbin[i])
    state_tokens = np.concatenate(chunks) if chunks else state_tokens
    # value after shift k: listings of all events before step k; the tail shifts after the last
    # event keep the value of the last shift that preceded an event
    k = np.minimum(np.arange(total_shifts + 1), last_step)
    after_shift = dumped[np.searchsorted(cs, k, 'left')]
    state_idx = after_shift[k_frame - 1]
  return IndexedEvents(events, starts, ends, state_tokens, state_idx)


# ---------------------------------------------------------------------------------------------
# per-segment token sequence
# ---------------------------------------------------------------------------------------------
def segment_events(ix: IndexedEvents, first_frame: int, last_frame: int,
                   tie_id: Optional[int]) -> np.ndarray:
  """Events owned by frames [first_frame, last_frame], preceded (if tie_id is given) by the
  listing of the notes sounding at the segment start, up to and including its 'tie' token."""
  body = ix.events[ix.event_start_indices[first_frame]:ix.event_end_indices[last_frame]]
  if tie_id is None:
    return body
  if not len(ix.state_events):
    # a song without a single note has no listing at all (the reference's pipeline indexes out
    # of range here); an empty tie section is the consistent answer
    return np.concatenate([[tie_id], body]).astype(np.int64)
  s = int(ix.state_event_indices[first_frame])
  rel = np.flatnonzero(ix.state_events[s:] == tie_id)
  if not rel.size:
    raise ValueError('state events carry no tie token after the segment start')
  return np.concatenate([ix.state_events[s:s + rel[0] + 1], body])


def map_programs(events: np.ndarray, vocab: EventVocabulary, granularity: str = 'full') -> np.ndarray:
  """'full' keeps program tokens, 'midi_class' maps each to the first program of its class of
  eight, 'flat' drops them (vocabularies.py:56-101)."""
  lo, hi = vocab.id_range('program')
  is_prog = (events >= lo) & (events <= hi)
  if granularity == 'full':
    return events
  if granularity == 'midi_class':
    return np.where(is_prog, lo + 8 * ((events - lo) // 8), events)
  if granularity == 'flat':
    return events[~is_prog]
  raise ValueError(f'unknown program granularity {granularity!r}')


def run_length_encode_shifts(events: Sequence[int], vocab: EventVocabulary,
                             state_change_kinds: Sequence[str] = ()) -> np.ndarray:
  """Single-step shifts -> one run per non-shift event group, counted from the segment start
  (split into pieces of at most max_shift_steps); trailing shifts are dropped.  A token of a
  state-change kind that repeats the value already in force is dropped."""
  events = np.asarray(events, np.int64)
  shift_lo, shift_hi = vocab.id_range('shift')
  is_shift = (events >= shift_lo) & (events <= shift_hi)
  keep = ~is_shift
  for kind in state_change_kinds:
    lo, hi = vocab.id_range(kind)
    idx = np.flatnonzero((events >= lo) & (events <= hi))
    if idx.size > 1:
      same = events[idx[1:]] == events[idx[:-1]]
      keep[idx[1:][same]] = False
  elapsed = np.cumsum(is_shift)                       # steps since the segment start
  kept = np.flatnonzero(keep)
  out: List[int] = []
  previous = 0
  max_run = vocab.max_shift_steps
  for i in kept:
    now = int(elapsed[i])
    if now > previous:
      full, rest = divmod(now, max_run)
      out.extend([shift_lo + max_run] * full)
      if rest:
        out.append(shift_lo + rest)
      previous = now
    out.append(int(events[i]))
  return np.asarray(out, np.int64)


# ---------------------------------------------------------------------------------------------
# full song -> model batches
# ---------------------------------------------------------------------------------------------
SLAKH_CLASS_PROGRAMS = (0, 4, 8, 16, 24, 26, 29, 32, 33, 40, 41, 42, 43, 46, 47, 48, 50, 52, 55, 56,
                        57, 58, 60, 61, 64, 66, 67, 68, 69, 70, 71, 73, 80, 88)


def program_to_slakh_program(program: int) -> int:
  """Largest Slakh class program not above `program` (the colab's mapping, ipynb:701-707)."""
  return max(p for p in SLAKH_CLASS_PROGRAMS if p <= program)


def num_song_frames(total_time: float, sample_rate: int = 16000, hop_size: int = 320) -> int:
  """Frames of a silent buffer of int(total_time * sample_rate) samples after the reference's
  padding, which always adds between 1 and hop_size samples (preprocessors.py:60-81)."""
  n = int(total_time * sample_rate)
  return (n + hop_size - n % hop_size) // hop_size


@dataclasses.dataclass
class SongTokens:
  tokens: np.ndarray          # int32 [segments, inputs_length], 0-padded, EOS-terminated
  lengths: np.ndarray         # tokens per segment including EOS
  num_frames: int             # mel frames of the whole song (last segment may be partial)
  frames_per_segment: int


def tokenize_song(notes: np.ndarray, vocab: Optional[EventVocabulary] = None, inputs_length: int = 2048,
                  frames_per_segment: int = 256, frame_rate: int = 50, sample_rate: int = 16000,
                  hop_size: int = 320, include_ties: bool = True, program_granularity: str = 'full',
                  map_to_slakh_programs: bool = True, total_time: Optional[float] = None) -> SongTokens:
  """The colab's "Setup Synthesis Task" cell as one function: notes (sustain already applied)
  -> one row of model input ids per 5.12 s segment."""
  vocab = vocab or mt3_event_vocabulary()
  notes = notes.copy()
  if map_to_slakh_programs and len(notes):
    pitched = ~notes['is_drum']
    notes['program'][pitched] = [program_to_slakh_program(int(p)) for p in notes['program'][pitched]]
  if total_time is None:
    total_time = float(notes['end'].max()) if len(notes) else 0.0
  nframes = num_song_frames(total_time, sample_rate, hop_size)
  frame_times = np.arange(nframes) / frame_rate
  ix = encode_and_index_events(timed_note_events(notes, 'onsets_offsets_programs'), vocab,
                               frame_times, with_tie_state=True)
  tie_id = vocab.encode('tie', 0) if include_ties else None
  nseg = -(-nframes // frames_per_segment)
  rows = np.zeros((nseg, inputs_length), np.int32)
  lengths = np.zeros(nseg, np.int64)
  for s in range(nseg):
    f0, f1 = s * frames_per_segment, min((s + 1) * frames_per_segment, nframes) - 1
    ev = segment_events(ix, f0, f1, tie_id)
    ev = map_programs(ev, vocab, program_granularity)
    ev = run_length_encode_shifts(ev, vocab, ('velocity', 'program'))
    if len(ev) > inputs_length - 1:                       # room for EOS (handle_too_long)
      raise ValueError(f'Value for "inputs" field exceeds maximum length: segment {s} has '
                       f'{len(ev)} tokens, limit {inputs_length - 1}')
    ids = to_model_ids(ev, vocab.num_classes)
    rows[s, :len(ids)] = ids
    rows[s, len(ids)] = EOS_ID
    lengths[s] = len(ids) + 1
  return SongTokens(rows, lengths, nframes, frames_per_segment)


# ---------------------------------------------------------------------------------------------
# tokens -> notes (the inverse direction; used for round-trip checks of the tokeniser and by
# callers that want to look at what a token row says)
# ---------------------------------------------------------------------------------------------
DEFAULT_VELOCITY = 100
DEFAULT_NOTE_DURATION = 0.01
MIN_NOTE_DURATION = 0.01


class NoteDecoder:
  """Streaming decoder of event ids back into notes (msd/note_sequences.py:264-407 and
  msd/run_length_encoding.py:274-326; known answers: note_sequences_test.py:289-504).

  Shift tokens give the time since the segment start (consecutive ones add up, any other event
  re-arms the counter).  With `onsets_only` every pitch token is a note of DEFAULT_NOTE_DURATION;
  otherwise velocity / program tokens set the state that the following pitch (or drum) tokens
  use, velocity 0 closing the sounding note of that (pitch, program).  A segment may open with a
  tie section (`begin_segment`): pitches listed before the 'tie' token stay sounding, every other
  sounding note is closed at the tie token's time."""

  def __init__(self, vocab: EventVocabulary, onsets_only: bool = False):
    self.vocab = vocab
    self.onsets_only = onsets_only
    self.time = 0.0
    self.velocity = DEFAULT_VELOCITY
    self.program = 0
    self.sounding: Dict[Tuple[int, int], Tuple[float, int]] = {}   # (pitch, program) -> (onset, vel)
    self.tied: set = set()
    self.in_tie_section = False
    self.rows: List[Tuple[float, float, int, int, int, bool]] = []

  # -- helpers -------------------------------------------------------------------------------
  def _emit(self, start: float, end: float, pitch: int, velocity: int, program: int = 0,
            is_drum: bool = False) -> None:
    self.rows.append((start, max(end, start + MIN_NOTE_DURATION), pitch, velocity, program, is_drum))

  def _close(self, key: Tuple[int, int], end: float) -> None:
    onset, velocity = self.sounding.pop(key)
    self._emit(onset, end, key[0], velocity, key[1])

  def begin_segment(self) -> None:
    self.tied = set()
    self.in_tie_section = True

  # -- one event -----------------------------------------------------------------------------
  def _event(self, time: float, kind: str, value: int) -> None:
    if self.onsets_only:
      if kind != 'pitch':
        raise ValueError(f'unexpected event type: {kind}')
      self._emit(time, time + DEFAULT_NOTE_DURATION, value, DEFAULT_VELOCITY)
      return
    if time < self.time:
      raise ValueError(f'event time < current time, {time} < {self.time}')
    self.time = time
    key = (value, self.program)
    if kind == 'pitch':
      if self.in_tie_section:
        if key not in self.sounding:
          raise ValueError(f'inactive pitch/program in tie section: {value}/{self.program}')
        if key in self.tied:
          raise ValueError(f'pitch/program is already tied: {value}/{self.program}')
        self.tied.add(key)
      elif self.velocity == 0:
        if key not in self.sounding:
          raise ValueError(f'note-off for inactive pitch/program: {value}/{self.program}')
        self._close(key, time)
      else:
        if key in self.sounding:        # re-struck without a note-off: end the earlier note here
          self._close(key, time)
        self.sounding[key] = (time, self.velocity)
    elif kind == 'drum':
      if self.velocity == 0:
        raise ValueError('velocity cannot be zero for drum event')
      self._emit(time, time + DEFAULT_NOTE_DURATION, value, self.velocity, 0, True)
    elif kind == 'velocity':
      self.velocity = bin_to_velocity(value, num_velocity_bins_of(self.vocab))
    elif kind == 'program':
      self.program = value
    elif kind == 'tie':
      if not self.in_tie_section:
        raise ValueError('tie section end event when not in tie section')
      for k in [k for k in self.sounding if k not in self.tied]:
        self._close(k, self.time)
      self.in_tie_section = False
    else:
      raise ValueError(f'unexpected event type: {kind}')

  # -- a run of event ids ------------------------------------------------------------------------
  def feed(self, event_ids: Sequence[int], start_time: float = 0.0,
           max_time: Optional[float] = None) -> Tuple[int, int]:
    """Returns (ids that could not be decoded or applied, ids dropped beyond max_time)."""
    invalid = dropped = 0
    steps = 0
    now = start_time
    event_ids = list(event_ids)
    for i, tok in enumerate(event_ids):
      try:
        kind, value = self.vocab.decode(int(tok))
      except ValueError:
        invalid += 1
        continue
      if kind == 'shift':
        steps += value
        now = start_time + steps / self.vocab.steps_per_second
        if max_time and now > max_time:
          dropped = len(event_ids) - i
          break
        continue
      steps = 0
      try:
        self._event(now, kind, value)
      except ValueError:
        invalid += 1
    return invalid, dropped

  def finish(self) -> np.ndarray:
    """Closes whatever still sounds (at the latest time seen, at least MIN_NOTE_DURATION after its
    onset) and returns the notes."""
    for onset, _ in self.sounding.values():
      self.time = max(self.time, onset + MIN_NOTE_DURATION)
    for k in list(self.sounding):
      self._close(k, self.time)
    return make_notes(self.rows)


def decode_song(tokens: np.ndarray, vocab: EventVocabulary, frames_per_segment: int = 256,
                frame_rate: int = 50, include_ties: bool = True) -> Tuple[np.ndarray, int, int]:
  """Rows of model ids as produced by `tokenize_song` -> (notes, invalid ids, dropped ids)."""
  dec = NoteDecoder(vocab)
  invalid = dropped = 0
  seconds = frames_per_segment / frame_rate
  for s, row in enumerate(np.asarray(tokens)):
    ids = from_model_ids(row, vocab.num_classes)
    ids = ids[ids >= 0]                                    # strip EOS marker and padding
    if include_ties:
      dec.begin_segment()
    a, b = dec.feed(ids, start_time=s * seconds, max_time=None)
    invalid += a
    dropped += b
  return dec.finish(), invalid, dropped
