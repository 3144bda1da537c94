"""CPU, world_size 2 over gloo: the segment sharding and the chained-song relay protocol of
music_spectrogram_diffusion_b200.distributed with a stand-in predict function."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from music_spectrogram_diffusion_b200 import distributed as D

FRAMES, DIMS, TOK = 8, 4, 16


def fake_predict(tokens, ctx, mask, seed):
  """Deterministic stand-in for InferenceModel.predict: depends on tokens, context, mask, seed."""
  b = tokens.shape[0]
  base = tokens.float().sum(dim=1).reshape(b, 1, 1) * 1e-3
  c = (ctx * mask.float().unsqueeze(-1)).mean(dim=(1, 2)).reshape(b, 1, 1)
  s = torch.arange(b, dtype=torch.float32).reshape(b, 1, 1) + seed
  return (base + 0.5 * c + 0.01 * s).expand(b, FRAMES, DIMS).contiguous()


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  g = torch.Generator().manual_seed(0)
  n = 5
  tokens = torch.randint(1, 100, (n, TOK), generator=g)
  ctx = torch.randn(n, FRAMES, DIMS, generator=g)
  mask = torch.ones(n, FRAMES, dtype=torch.int32)
  full = D.predict_sharded(fake_predict, tokens, ctx, mask, seed=3)
  segs = [torch.randint(1, 100, (TOK,), generator=g) for _ in range(5)]
  song = D.synthesize_song(fake_predict, segs, FRAMES, DIMS, torch.device('cpu'), seed=1)
  q.put((rank, full.numpy(), None if song is None else song.numpy()))
  dist.barrier()
  dist.destroy_process_group()


def test_shard_range_partitions():
  for n in (1, 5, 8, 64):
    for w in (1, 2, 3, 8):
      blocks = [D.shard_range(n, w, r) for r in range(w)]
      assert blocks[0][0] == 0 and blocks[-1][1] == n
      assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
      sizes = [b - a for a, b in blocks]
      assert max(sizes) - min(sizes) <= 1


def test_two_ranks_match_single_process():
  world = 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  results = {}
  for _ in range(world):
    r, full, song = q.get(timeout=120)
    results[r] = (full, song)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0

  # single-process expectations
  g = torch.Generator().manual_seed(0)
  n = 5
  tokens = torch.randint(1, 100, (n, TOK), generator=g)
  ctx_t = torch.randn(n, FRAMES, DIMS, generator=g)
  mask = torch.ones(n, FRAMES, dtype=torch.int32)
  want_full = torch.cat([fake_predict(tokens[a:b], ctx_t[a:b], mask[a:b], 3)
                         for a, b in (D.shard_range(n, world, r) for r in range(world))])
  segs = [torch.randint(1, 100, (TOK,), generator=g) for _ in range(5)]
  want_song = D.synthesize_song(fake_predict, segs, FRAMES, DIMS, torch.device('cpu'), seed=1)
  for r in range(world):
    np.testing.assert_allclose(results[r][0], want_full.numpy(), rtol=0, atol=0)
  np.testing.assert_allclose(results[0][1], want_song.numpy(), rtol=0, atol=0)
  assert results[1][1] is None
  # the relay is the single-process chain with ONE seed for every segment (beam/evaluation.py:209)
  prev = torch.zeros(1, FRAMES, DIMS)
  for k, seg in enumerate(segs):
    m = torch.zeros(1, FRAMES, dtype=torch.int32) if k == 0 else torch.ones(1, FRAMES, dtype=torch.int32)
    prev = fake_predict(seg.reshape(1, -1), prev, m, 1)
    np.testing.assert_array_equal(want_song[0, k * FRAMES:(k + 1) * FRAMES].numpy(), prev[0].numpy())
  # the chain really chains: segment 1 depends on segment 0's output
  assert not np.allclose(want_song[0, :FRAMES].numpy(), want_song[0, FRAMES:2 * FRAMES].numpy())


class _FakeEngine:
  """Stands in for engine.Engine in the guidance-split set-up protocol."""

  def __init__(self, rank):
    self.rank, self.attached, self.device = rank, None, torch.device('cpu')

  def p2p_export(self):
    return bytes([self.rank]) * 64

  def p2p_attach(self, handle, role):
    assert len(handle) == 64
    self.attached = (handle[0], role)

  def p2p_detach(self):
    self.attached = None


class _FakeModel:
  def __init__(self, rank):
    self.engine = _FakeEngine(rank)


def _pair_worker(rank, world, port, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  model = _FakeModel(rank)
  orig_sync = torch.cuda.synchronize
  torch.cuda.synchronize = lambda *a, **k: None   # CPU stand-in
  try:
    with D.CfgSplitPair(model, 0, 1) as pair:
      inside = (pair.active, model.engine.attached)
    q.put((rank, inside, model.engine.attached))
  finally:
    torch.cuda.synchronize = orig_sync
  dist.barrier()
  dist.destroy_process_group()


def test_guidance_split_pairing_protocol():
  """CfgSplitPair: ranks 0 / 1 swap their 64-byte handles and attach with roles cond / uncond,
  a third rank takes part in the barriers only; everybody detaches on exit."""
  world = 3
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_pair_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = {}
  for _ in range(world):
    r, inside, after = q.get(timeout=120)
    res[r] = (inside, after)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert res[0] == ((True, (1, 'cond')), None)      # rank 0 holds rank 1's handle, conditional pass
  assert res[1] == ((True, (0, 'uncond')), None)
  assert res[2] == ((False, None), None)
