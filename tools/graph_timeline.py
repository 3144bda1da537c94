"""In-graph kernel timeline of one diffusion step (CUPTI through torch.profiler): start offset,
duration and gap of every kernel of a replayed step graph, plus per-class busy time.  Unlike the
CUDA-event classes of bench.py (which serialise the launches) this shows the overlap that
programmatic dependent launch actually achieves."""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from music_spectrogram_diffusion_b200 import inference

ap = argparse.ArgumentParser()
ap.add_argument('--model', default='base')
ap.add_argument('--segments', type=int, default=8)
ap.add_argument('--diffusion-steps', type=int, default=12)
ap.add_argument('--out', default='gpurun_out/graph_timeline.json')
args = ap.parse_args()
t5, diff, lengths = bench.model_configs(args)
model = inference.InferenceModel.from_config(t5, diff, lengths, 'synthetic:0', args.segments, 0)
eng = model.engine
dev = eng.device
b = bench.synthetic_batch(args.segments, lengths, 100)
tok = torch.from_numpy(b['encoder_input_tokens']).to(dev)
ctx = torch.from_numpy(b['encoder_continuous_inputs']).to(dev)
msk = torch.from_numpy(b['encoder_continuous_mask']).to(dev)
eng.encode(tok, ctx, msk)
for _ in range(2):
  eng.sample(seed=1)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
  eng.sample(seed=2)
  torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') == 'kernel']
ev.sort(key=lambda e: e['ts'])
names = [e['name'] for e in ev]
# a step = everything from one sampler_step kernel (exclusive) to the next (inclusive)
ends = [i for i, n in enumerate(names) if 'step_advance' in n]
assert len(ends) >= 4, f'found {len(ends)} steps in the trace'
lo, hi = ends[len(ends) // 2 - 1] + 1, ends[len(ends) // 2] + 1
step = ev[lo:hi]
t0 = step[0]['ts']


def cls(n):
  for k in ('gemm', 'attention_combine', 'attention', 'rmsnorm', 'sampler', 'step_advance'):
    if k in n:
      return k
  return 'other'


rows = []
prev_end = t0
busy = {}
for e in step:
  s, d = e['ts'] - t0, e['dur']
  rows.append({'kernel': cls(e['name']), 'start_us': round(s, 2), 'dur_us': round(d, 2),
               'gap_us': round(e['ts'] - prev_end, 2),
               'grid': e.get('args', {}).get('grid'), 'block': e.get('args', {}).get('block')})
  busy[cls(e['name'])] = busy.get(cls(e['name']), 0.0) + d
  prev_end = max(prev_end, e['ts'] + d)
total = step[-1]['ts'] + step[-1]['dur'] - t0
summary = {'kernels': len(step), 'step_us': round(total, 1),
           'sum_of_durations_us': round(sum(e['dur'] for e in step), 1),
           'busy_us_by_class': {k: round(v, 1) for k, v in busy.items()},
           'positive_gaps_us': round(sum(max(0.0, r['gap_us']) for r in rows), 1),
           'overlap_us': round(-sum(min(0.0, r['gap_us']) for r in rows), 1)}
os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
json.dump({'summary': summary, 'kernels': rows}, open(args.out, 'w'), indent=0)
print(json.dumps(summary))
for r in rows[:30]:
  print(r)
