"""Standard MIDI file -> note array (+ sustain pedal), without pretty_midi / note_seq.

The colab reads the upload with `note_seq.midi_to_note_sequence` and then calls
`note_seq.apply_sustain_control_changes` (ipynb "Upload MIDI File" / "Setup Synthesis Task").
Both live in third-party packages that are not part of /root/reference (note_seq==0.0.3 and the
pretty_midi it wraps), so this module restates their published behaviour:

  * tempo map in seconds from the set-tempo meta events of all tracks (default 120 bpm);
  * a note = note-on (velocity > 0) .. next note-off / note-on-velocity-0 of the same channel
    and pitch inside a track; one note-off closes every note of that key opened at an earlier
    tick; channel 10 (index 9) is percussion; a note carries the program in force on its
    channel when it is closed;
  * sustain (controller 64, value >= 64 = down): a note released while its channel's pedal is
    down keeps sounding until the pedal comes up or the same pitch is struck again, whichever
    is first; events at equal times are ordered pedal-down, pedal-up, note-on, note-off.

**Parity unpinned**: no golden vectors exist for these two steps in the reference's tests; the
tests here pin the reader against files written by `write_midi` and hand-assembled bytes.
"""

from __future__ import annotations

import dataclasses
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

from music_spectrogram_diffusion_b200.midi_tokens import NOTE_DTYPE

SUSTAIN_CONTROLLER = 64
CC_DTYPE = np.dtype([('time', 'f8'), ('number', 'i4'), ('value', 'i4'), ('channel', 'i4'),
                     ('track', 'i4')])


class MidiError(ValueError):
  pass


@dataclasses.dataclass
class MidiSong:
  notes: np.ndarray            # NOTE_DTYPE
  note_channel: np.ndarray     # int32 per note: track * 16 + channel ("instrument" for sustain)
  control_changes: np.ndarray  # CC_DTYPE
  ticks_per_quarter: int
  total_time: float


def _read_vlq(buf: bytes, pos: int) -> Tuple[int, int]:
  value = 0
  for _ in range(4):
    if pos >= len(buf):
      raise MidiError('truncated variable-length quantity')
    b = buf[pos]
    pos += 1
    value = (value << 7) | (b & 0x7F)
    if not b & 0x80:
      return value, pos
  raise MidiError('variable-length quantity longer than 4 bytes')


def _track_events(buf: bytes) -> List[Tuple[int, int, int, int, int]]:
  """(tick, kind, channel, a, b) with kind: 0 note-off, 1 note-on, 2 control, 3 program,
  4 tempo (a = microseconds per quarter)."""
  out = []
  pos, tick, status = 0, 0, 0
  while pos < len(buf):
    delta, pos = _read_vlq(buf, pos)
    tick += delta
    if pos >= len(buf):
      raise MidiError('event without status byte')
    b = buf[pos]
    if b & 0x80:
      status = b
      pos += 1
    elif not status:
      raise MidiError('running status without a previous status byte')
    if status == 0xFF:                                   # meta
      if pos >= len(buf):
        raise MidiError('truncated meta event')
      mtype = buf[pos]
      length, pos = _read_vlq(buf, pos + 1)
      data = buf[pos:pos + length]
      pos += length
      if mtype == 0x51 and length == 3:
        out.append((tick, 4, 0, int.from_bytes(data, 'big'), 0))
      if mtype == 0x2F:
        break
      status = 0                                         # meta / sysex cancel running status
      continue
    if status in (0xF0, 0xF7):                           # sysex
      length, pos = _read_vlq(buf, pos)
      pos += length
      status = 0
      continue
    hi, ch = status & 0xF0, status & 0x0F
    nbytes = 1 if hi in (0xC0, 0xD0) else 2
    if pos + nbytes > len(buf):
      raise MidiError('truncated channel event')
    a = buf[pos]
    c = buf[pos + 1] if nbytes == 2 else 0
    pos += nbytes
    if hi == 0x90 and c > 0:
      out.append((tick, 1, ch, a, c))
    elif hi == 0x80 or hi == 0x90:
      out.append((tick, 0, ch, a, 0))
    elif hi == 0xB0:
      out.append((tick, 2, ch, a, c))
    elif hi == 0xC0:
      out.append((tick, 3, ch, a, 0))
  return out


def read_midi(data: bytes) -> MidiSong:
  if data[:4] != b'MThd' or len(data) < 14:
    raise MidiError('not a standard MIDI file (no MThd header)')
  hlen, fmt, ntracks, division = struct.unpack('>IHHH', data[4:14])
  if division & 0x8000:
    raise MidiError('SMPTE time division is not supported')
  if fmt not in (0, 1):
    raise MidiError(f'MIDI format {fmt} is not supported (0 and 1 are)')
  pos = 8 + hlen
  tracks = []
  while pos + 8 <= len(data) and len(tracks) < ntracks:
    tag, length = data[pos:pos + 4], struct.unpack('>I', data[pos + 4:pos + 8])[0]
    body = data[pos + 8:pos + 8 + length]
    pos += 8 + length
    if tag == b'MTrk':
      tracks.append(_track_events(body))
  # tempo map: piecewise-linear tick -> seconds
  tempos = sorted((t, us) for tr in tracks for (t, kind, _, us, _) in tr if kind == 4)
  seg_tick, seg_time, seg_spt = [0], [0.0], [0.5 / division]      # 120 bpm until told otherwise
  for tick, us in tempos:
    spt = us * 1e-6 / division
    if tick == seg_tick[-1]:
      seg_spt[-1] = spt
      continue
    seg_time.append(seg_time[-1] + (tick - seg_tick[-1]) * seg_spt[-1])
    seg_tick.append(tick)
    seg_spt.append(spt)
  seg_tick_a, seg_time_a, seg_spt_a = np.array(seg_tick), np.array(seg_time), np.array(seg_spt)

  def seconds(tick: int) -> float:
    i = int(np.searchsorted(seg_tick_a, tick, side='right')) - 1
    return float(seg_time_a[i] + (tick - seg_tick_a[i]) * seg_spt_a[i])

  notes, chans, ccs = [], [], []
  for ti, tr in enumerate(tracks):
    program = [0] * 16
    open_notes: Dict[Tuple[int, int], List[Tuple[int, int]]] = {}
    for tick, kind, ch, a, b in tr:
      if kind == 3:
        program[ch] = a
      elif kind == 2:
        ccs.append((seconds(tick), a, b, ch, ti))
      elif kind == 1:
        open_notes.setdefault((ch, a), []).append((tick, b))
      elif kind == 0 and (ch, a) in open_notes:
        pending = open_notes[(ch, a)]
        closing = [p for p in pending if p[0] != tick]
        keeping = [p for p in pending if p[0] == tick]
        for start_tick, velocity in closing:
          notes.append((seconds(start_tick), seconds(tick), a, velocity, program[ch], ch == 9))
          chans.append(ti * 16 + ch)
        if closing and keeping:
          open_notes[(ch, a)] = keeping
        else:
          del open_notes[(ch, a)]
  note_arr = np.array(notes, dtype=NOTE_DTYPE) if notes else np.zeros(0, NOTE_DTYPE)
  cc_arr = np.array(ccs, dtype=CC_DTYPE) if ccs else np.zeros(0, CC_DTYPE)
  total = float(note_arr['end'].max()) if len(note_arr) else 0.0
  return MidiSong(note_arr, np.array(chans, np.int32), cc_arr, division, total)


def apply_sustain(song: MidiSong) -> MidiSong:
  """Extends note ends over pedal-down spans (see the module docstring for the rules)."""
  notes = song.notes.copy()
  n = len(notes)
  # (time, order, kind, index): kinds 0 pedal down, 1 pedal up, 2 note on, 3 note off
  events = []
  for i in range(n):
    events.append((float(notes['start'][i]), 2, i))
    events.append((float(notes['end'][i]), 3, i))
  for c in song.control_changes:
    if c['number'] == SUSTAIN_CONTROLLER and 0 <= c['value'] <= 127:
      events.append((float(c['time']), 0 if c['value'] >= 64 else 1, -(int(c['track']) * 16 + int(c['channel'])) - 1))
  events.sort(key=lambda e: (e[0], e[1]))
  pedal: Dict[int, bool] = {}
  sounding: Dict[int, List[int]] = {}
  alive = np.ones(n, bool)
  time = 0.0
  for time, kind, ref in events:
    if kind <= 1:
      inst = -ref - 1
      pedal[inst] = kind == 0
      if kind == 1:
        keep = []
        for i in sounding.get(inst, []):
          if notes['end'][i] < time:
            notes['end'][i] = time
          else:
            keep.append(i)
        sounding[inst] = keep
      continue
    inst = int(song.note_channel[ref])
    if kind == 2:
      if pedal.get(inst, False):
        keep = []
        for i in sounding.get(inst, []):
          if notes['pitch'][i] == notes['pitch'][ref]:
            notes['end'][i] = time
            if notes['start'][i] == notes['end'][i]:
              alive[i] = False
          else:
            keep.append(i)
        sounding[inst] = keep
      sounding.setdefault(inst, []).append(ref)
    elif not pedal.get(inst, False) and ref in sounding.get(inst, []):
      sounding[inst].remove(ref)
  for rest in sounding.values():            # pedal still down at the end: ring until the last event
    for i in rest:
      notes['end'][i] = max(notes['end'][i], time)
  total = float(notes['end'][alive].max()) if alive.any() else 0.0
  return MidiSong(notes[alive], song.note_channel[alive], song.control_changes,
                  song.ticks_per_quarter, total)


# ---------------------------------------------------------------------------------------------
# writer (fixtures and examples)
# ---------------------------------------------------------------------------------------------
def _vlq(value: int) -> bytes:
  out = [value & 0x7F]
  value >>= 7
  while value:
    out.append((value & 0x7F) | 0x80)
    value >>= 7
  return bytes(reversed(out))


def write_midi(notes: np.ndarray, ticks_per_quarter: int = 480, bpm: float = 120.0,
               sustain: Sequence[Tuple[float, int, int]] = ()) -> bytes:
  """Format-1 file: tempo track + one track per (program, is_drum); `sustain` = (time, channel,
  value) controller-64 events.  Times are rounded to ticks."""
  spt = 60.0 / bpm / ticks_per_quarter
  to_tick = lambda t: int(round(t / spt))
  tracks = [b'\x00\xff\x51\x03' + int(60e6 / bpm).to_bytes(3, 'big') + b'\x00\xff\x2f\x00']
  groups: Dict[Tuple[int, bool], List[int]] = {}
  for i in range(len(notes)):
    groups.setdefault((int(notes['program'][i]), bool(notes['is_drum'][i])), []).append(i)
  next_channel = 0
  for (program, is_drum), idx in sorted(groups.items()):
    if is_drum:
      ch = 9
    else:
      ch = next_channel if next_channel < 9 else next_channel + 1
      next_channel += 1
      if ch > 15:
        raise MidiError('more than 15 pitched programs do not fit one port')
    ev = [(0, 0, bytes([0xC0 | ch, program]))]
    for i in idx:
      ev.append((to_tick(notes['start'][i]), 2, bytes([0x90 | ch, int(notes['pitch'][i]), int(notes['velocity'][i])])))
      ev.append((to_tick(notes['end'][i]), 1, bytes([0x80 | ch, int(notes['pitch'][i]), 0])))
    for t, c, v in sustain:
      if c == ch:
        ev.append((to_tick(t), 0, bytes([0xB0 | ch, SUSTAIN_CONTROLLER, v])))
    ev.sort(key=lambda e: (e[0], e[1]))
    body, last = b'', 0
    for tick, _, msg in ev:
      body += _vlq(tick - last) + msg
      last = tick
    tracks.append(body + b'\x00\xff\x2f\x00')
  out = b'MThd' + struct.pack('>IHHH', 6, 1, len(tracks), ticks_per_quarter)
  for body in tracks:
    out += b'MTrk' + struct.pack('>I', len(body)) + body
  return out
