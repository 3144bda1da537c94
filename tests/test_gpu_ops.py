"""-m gpu: operator-level parity of the sm_100a kernels against the oracle's restatement of
msd/layers.py (same shape of test as layers_test.py:375-387 / 285-330 / 450-484, at sizes the
tensor-core kernels accept)."""
import numpy as np
import pytest
import torch

from oracle import msd_oracle as O
from tests.helpers import bf16_round

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('variant,block_n', [(0, 0), (0, 64), (0, 96), (0, 128), (0, 192), (0, 256),
                                             (1, 0), (1, 64), (1, 128), (1, 256)])
@pytest.mark.parametrize('M,N,K', [(128, 768, 64), (256, 768, 128), (512, 2304, 768),
                                   (384, 768, 384), (2048 + 128, 1536, 2048)])
def test_dense_general(cuda_device, M, N, K, variant, block_n):
  from music_spectrogram_diffusion_b200 import engine
  if block_n and N % block_n:
    pytest.skip('tile width does not divide N')
  g = torch.Generator().manual_seed(M + N + K)
  a = bf16_round(torch.randn(M, K, generator=g))
  w = bf16_round(torch.randn(K, N, generator=g) / np.sqrt(K))
  got = engine.op_dense(a.to(cuda_device), w.to(cuda_device), variant, block_n).cpu()
  want = O.dense_general(a.double(), w.double()).float()
  err = (got - want).abs().max().item()
  assert err < 2e-4 * np.sqrt(K), f'max err {err}'


@pytest.mark.parametrize('nb,heads,Lq,Lk,tail,masked', [
    (2, 2, 128, 256, 1, False), (2, 3, 256, 384, 1, True), (2, 3, 256, 384, 2, True),
    (1, 2, 256, 2304, 5, True), (2, 2, 256, 768, 2, 'head'), (3, 2, 128, 256, 1, True),
    (8, 12, 256, 768, 0, False), (8, 12, 256, 2304, 0, True)])
def test_dot_product_attention_tail_split(cuda_device, monkeypatch, nb, heads, Lq, Lk, tail, masked):
  """Long/short CTA pairs with the in-kernel merge (tail 0 = the automatic choice, which is active
  for the 96-CTA grids of the last two cases); run twice to check the hand-shake words re-arm."""
  from music_spectrogram_diffusion_b200 import engine
  monkeypatch.setenv('MSD_ATTN_SPLITS', '1')
  monkeypatch.setenv('MSD_ATTN_TAIL', str(tail))
  g = torch.Generator().manual_seed(nb * 77 + Lk + tail)
  w = heads * 64
  q = bf16_round(torch.randn(nb, Lq, w, generator=g) * 0.5)
  k = bf16_round(torch.randn(nb, Lk, w, generator=g) * 0.5)
  v = bf16_round(torch.randn(nb, Lk, w, generator=g))
  mask = None
  bias = None
  if masked:
    mask = (torch.rand(nb, Lk, generator=g) > 0.3).to(torch.int32)
    if masked == 'head':
      mask[0, :Lk - 128] = 0                   # the long CTA of batch 0 has nothing to attend to
    else:
      mask[0, Lk // 2:] = 0                    # the short CTA of batch 0 has nothing to attend to
    if nb > 2:
      mask[2, :] = 0                           # neither has -> zeros
    qm = torch.ones(nb, Lq)
    m4 = O.make_attention_mask(qm, mask.float())
    bias = torch.where(m4 > 0, torch.zeros_like(m4), torch.full_like(m4, -1e10))
  want = O.dot_product_attention(q.view(nb, Lq, heads, 64), k.view(nb, Lk, heads, 64),
                                 v.view(nb, Lk, heads, 64), bias).reshape(nb, Lq, w)
  if masked:
    want = O.zero_activations_if_masked(want, m4)
  for _ in range(2):
    got = engine.op_attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device),
                              None if mask is None else mask.to(cuda_device), heads).cpu()
    err = (got - want).abs().max().item()
    assert err < 3e-2, f'max err {err}'


@pytest.mark.parametrize('splits', [0, 1, 3])
@pytest.mark.parametrize('nb,heads,Lq,Lk,masked', [
    (1, 1, 128, 128, False), (2, 2, 128, 256, False), (2, 3, 256, 384, True),
    (1, 2, 256, 2304, True), (3, 2, 128, 128, True), (2, 2, 256, 768, True)])
def test_dot_product_attention(cuda_device, monkeypatch, nb, heads, Lq, Lk, masked, splits):
  """splits: 0 = automatic split-KV choice, 1 = single pass, 3 = forced 3-way split + combine."""
  from music_spectrogram_diffusion_b200 import engine
  if splits == 3 and (Lk // 128) % 3:
    pytest.skip('key blocks not divisible by 3')
  if splits:
    monkeypatch.setenv('MSD_ATTN_SPLITS', str(splits))
  g = torch.Generator().manual_seed(nb * 1000 + Lk)
  w = heads * 64
  q = bf16_round(torch.randn(nb, Lq, w, generator=g) * 0.5)
  k = bf16_round(torch.randn(nb, Lk, w, generator=g) * 0.5)
  v = bf16_round(torch.randn(nb, Lk, w, generator=g))
  mask = None
  bias = None
  if masked:
    mask = (torch.rand(nb, Lk, generator=g) > 0.3).to(torch.int32)
    mask[0, Lk // 2:] = 0                      # a run of fully masked key blocks
    if nb > 2:
      mask[2, :] = 0                           # a row with nothing to attend to -> zeros
    qm = torch.ones(nb, Lq)
    m4 = O.make_attention_mask(qm, mask.float())
    bias = torch.where(m4 > 0, torch.zeros_like(m4), torch.full_like(m4, -1e10))
  want = O.dot_product_attention(q.view(nb, Lq, heads, 64), k.view(nb, Lk, heads, 64),
                                 v.view(nb, Lk, heads, 64), bias).reshape(nb, Lq, w)
  if masked:
    want = O.zero_activations_if_masked(want, m4)
  got = engine.op_attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device),
                            None if mask is None else mask.to(cuda_device), heads).cpu()
  err = (got - want).abs().max().item()
  assert torch.isfinite(got).all()
  assert err < 3e-2, f'max err {err}'


@pytest.mark.parametrize('rows,d,film', [(128, 128, False), (256, 768, True), (100, 512, True)])
def test_layer_norm_film(cuda_device, rows, d, film):
  from music_spectrogram_diffusion_b200 import engine
  g = torch.Generator().manual_seed(rows + d)
  x = torch.randn(rows, d, generator=g) * 3
  gamma = 1 + 0.1 * torch.randn(d, generator=g)
  fv = torch.randn(2 * d, generator=g) * 0.2 if film else None
  want = O.layer_norm(x, gamma)
  if film:
    want = want * (fv[:d] + 1.0) + fv[d:]
  got = engine.op_rmsnorm_film(x.to(cuda_device), gamma.to(cuda_device),
                               None if fv is None else fv.to(cuda_device)).cpu()
  err = (got - want).abs().max().item()
  assert err < 4e-2, f'max err {err}'   # bf16 output rounding of O(4) values
  assert (got - bf16_round(want)).abs().max().item() < 2e-2
