"""Per-key-block clock64 timeline of the softmax warpgroups of one attention CTA (GPU)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_spectrogram_diffusion_b200 import _native
lib = _native.load()
dev = torch.device('cuda', 0)
nb, H, Lq, Lk = int(os.environ.get('NB', '8')), 12, int(os.environ.get('LQ', '256')), int(os.environ.get('LK', '2304'))
w = H * 64
q = torch.randn(nb, Lq, w, device=dev) * 0.3
k = torch.randn(nb, Lk, w, device=dev) * 0.3
v = torch.randn(nb, Lk, w, device=dev)
mask = torch.ones(nb, Lk, dtype=torch.int32, device=dev) if os.environ.get('MASK', '1') == '1' else None
out = torch.empty_like(q)
trace = torch.zeros(2 * 64 * 8 + 8 + 3 * 1024, dtype=torch.int64, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
  rc = lib.msd_op_attention_trace(P(q), P(k), P(v), P(mask) if mask is not None else None, nb, H, Lq, Lk, P(out), P(trace), None)
  assert rc == 0, lib.msd_last_error()
torch.cuda.synchronize()
tt = trace.cpu()
k = tt[2 * 64 * 8:2 * 64 * 8 + 8]
cta = tt[2 * 64 * 8 + 8:].reshape(1024, 3)
cta = cta[cta[:, 1] > 0]
if len(cta):
  import collections
  t0g = int(cta[:, 1].min())
  per_sm = collections.Counter(int(x) for x in cta[:, 0])
  starts = sorted(int(x) - t0g for x in cta[:, 1]); ends = sorted(int(x) - t0g for x in cta[:, 2])
  durs = sorted(int(e) - int(s_) for s_, e in zip(cta[:, 1], cta[:, 2]))
  q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
  print(f'{len(cta)} CTAs on {len(per_sm)} SMs, CTAs per SM: {sorted(collections.Counter(per_sm.values()).items())}')
  print(f'start ns: min {starts[0]} p25 {q(starts, .25)} p50 {q(starts, .5)} p75 {q(starts, .75)} max {starts[-1]}')
  print(f'end   ns: min {ends[0]} p25 {q(ends, .25)} p50 {q(ends, .5)} p75 {q(ends, .75)} max {ends[-1]}')
  print(f'duration ns: min {durs[0]} p50 {q(durs, .5)} max {durs[-1]}')
  # overlapping residency: for each SM, the maximum number of CTAs alive at once
  by_sm = collections.defaultdict(list)
  for smid, a, b in cta.tolist():
    by_sm[smid].append((a, b))
  conc = collections.Counter()
  for smid, iv in by_sm.items():
    ev = sorted([(a, 1) for a, b in iv] + [(b, -1) for a, b in iv])
    cur = mx = 0
    for _, d in ev:
      cur += d; mx = max(mx, cur)
    conc[mx] += 1
  print(f'max concurrent CTAs per SM: {sorted(conc.items())}')
t = tt[:2 * 64 * 8].reshape(2, 64, 8)
nblk = Lk // 128
print('CTA stamps (cycles from entry): setup_done=%d first_block_start=%d loop_end=%d stores_done=%d after_sync=%d' % (int(k[1]-k[0]), int(t[0,0,0]-k[0]), int(k[2]-k[0]), int(k[3]-k[0]), int(k[4]-k[0])))
t0 = int(t[0, 0, 0])
names = ['start', 's_full', 'ldtm', 'max', 'turn', 'exp', 'pv_wait', 'arrive']
for tile in range(2):
  print('tile', tile, ' (cycles since first stamp; deltas between phases)')
  for it in range(min(18, Lk // 128)):
    row = [int(x) - t0 for x in t[tile, it]]
    d = [row[i + 1] - row[i] for i in range(7)]
    print(f'  blk {it:2d} start={row[0]:7d}  wait_s={d[0]:5d} ldtm={d[1]:5d} max={d[2]:5d} turn={d[3]:5d} exp={d[4]:5d} pvwait={d[5]:5d} store={d[6]:5d}  total={row[7]-row[0]:6d}')
