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
trace = torch.zeros(2 * 64 * 8 + 8, dtype=torch.int64, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
  rc = lib.msd_op_attention_trace(P(q), P(k), P(v), P(mask) if mask is not None else None, nb, H, Lq, Lk, P(out), P(trace), None)
  assert rc == 0, lib.msd_last_error()
torch.cuda.synchronize()
tt = trace.cpu()
k = tt[2 * 64 * 8:]
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
