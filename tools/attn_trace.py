"""Per-key-block clock64 timeline of the softmax warpgroups of one attention CTA (GPU)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_spectrogram_diffusion_b200 import _native
lib = _native.load()
dev = torch.device('cuda', 0)
nb, H, Lq, Lk = 8, 12, 256, 2304
w = H * 64
q = torch.randn(nb, Lq, w, device=dev) * 0.3
k = torch.randn(nb, Lk, w, device=dev) * 0.3
v = torch.randn(nb, Lk, w, device=dev)
mask = torch.ones(nb, Lk, dtype=torch.int32, device=dev)
out = torch.empty_like(q)
trace = torch.zeros(2, 64, 8, dtype=torch.int64, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
  rc = lib.msd_op_attention_trace(P(q), P(k), P(v), P(mask), nb, H, Lq, Lk, P(out), P(trace), None)
  assert rc == 0, lib.msd_last_error()
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[0, 0, 0])
names = ['start', 's_full', 'ldtm', 'max', 'turn', 'exp', 'pv_wait', 'arrive']
for tile in range(2):
  print('tile', tile, ' (cycles since first stamp; deltas between phases)')
  for it in range(18):
    row = [int(x) - t0 for x in t[tile, it]]
    d = [row[i + 1] - row[i] for i in range(7)]
    print(f'  blk {it:2d} start={row[0]:7d}  wait_s={d[0]:5d} ldtm={d[1]:5d} max={d[2]:5d} turn={d[3]:5d} exp={d[4]:5d} pvwait={d[5]:5d} store={d[6]:5d}  total={row[7]-row[0]:6d}')
