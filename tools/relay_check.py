"""Run under torchrun with >= 2 ranks (NCCL): one chained song relayed over the ranks
(distributed.synthesize_song) must equal the same chain computed by rank 0 alone, bit for bit;
prints both wall times.  BASELINE config 5 in miniature (fewer diffusion steps)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from music_spectrogram_diffusion_b200 import distributed as D, inference, midi_tokens as M, song

ap = argparse.ArgumentParser()
ap.add_argument('--model', default='base')
ap.add_argument('--segments', type=int, default=1)
ap.add_argument('--diffusion-steps', type=int, default=50)
ap.add_argument('--song-segments', type=int, default=6)
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)
t5, diff, lengths = bench.model_configs(args)
model = inference.InferenceModel.from_config(t5, diff, lengths, 'synthetic:0', 1, local)
# a synthetic arrangement long enough for the requested number of 5.12 s segments
rng = np.random.default_rng(0)
dur = args.song_segments * 5.12 - 0.5
rows = []
for prog in (0, 33, 48):
  t = 0.0
  while t < dur - 0.5:
    d = float(rng.uniform(0.1, 0.8))
    rows.append((t, min(t + d, dur), int(rng.integers(40, 80)), int(rng.integers(40, 120)), prog, False))
    t += float(rng.uniform(0.1, 0.5))
toks = M.tokenize_song(M.make_notes(rows), song.event_vocabulary_of(model))
segs = [torch.from_numpy(toks.tokens[i]) for i in range(len(toks.tokens))]
C, nd = lengths['targets_context'], 128
torch.cuda.synchronize()
if world > 1:
  dist.barrier()
t0 = time.time()
relay = D.synthesize_song(model.predict_on_device, segs, C, nd, dev, seed=11)
torch.cuda.synchronize()
t_relay = time.time() - t0
if rank == 0:
  # the same chain on this rank alone (no process group involved)
  t0 = time.time()
  prev = torch.zeros(1, C, nd, device=dev)
  parts = []
  for k, s in enumerate(segs):
    mask = (torch.zeros if k == 0 else torch.ones)(1, C, dtype=torch.int32, device=dev)
    prev = model.predict_on_device(s.to(dev).reshape(1, -1), prev, mask, seed=11)[:1].clone()
    parts.append(prev)
  torch.cuda.synchronize()
  t_local = time.time() - t0
  local_song = torch.cat(parts, dim=1)
  same = bool(torch.equal(relay, local_song))
  print(json.dumps({'ranks': world, 'segments': len(segs), 'diffusion_steps': args.diffusion_steps,
                    'bit_identical': same, 'max_abs_diff': float((relay - local_song).abs().max()),
                    'relay_seconds': round(t_relay, 3), 'single_rank_seconds': round(t_local, 3),
                    'frames': int(relay.shape[1])}))
  assert same
if world > 1:
  dist.barrier()
  dist.destroy_process_group()
