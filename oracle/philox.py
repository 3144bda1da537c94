"""numpy restatement of the library's in-kernel noise generator (TEST INFRASTRUCTURE).

csrc/elementwise.cu `philox_normal4`: Philox4x32-10 keyed by the 64-bit seed, counter =
(element_index / 4, stream, 0x6d7364), four uniforms -> two Box-Muller pairs.  stream 0 draws
init_z, stream i + 1 draws the noise of reverse step i.  This is NOT jax.random (threefry is a
third-party detail the reference does not pin); it exists so that seed-driven runs of the CUDA
path can be compared with the oracle on identical noise without shipping the noise tensors.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
  k0, k1 = np.uint32(k0), np.uint32(k1)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = M0 * c0.astype(np.uint64)
      p1 = M1 * c2.astype(np.uint64)
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
  return c0, c1, c2, c3


def normal(seed: int, stream: int, n: int) -> np.ndarray:
  """n (multiple of 4) standard normals, element order identical to the kernel's."""
  assert n % 4 == 0
  idx4 = np.arange(n // 4, dtype=np.uint64)
  c0 = (idx4 & MASK).astype(np.uint32)
  c1 = (idx4 >> np.uint64(32)).astype(np.uint32)
  c2 = np.full_like(c0, np.uint32(stream))
  c3 = np.full_like(c0, np.uint32(0x6d7364))
  r = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
  f32 = np.float32
  s = f32(2.3283064365386963e-10)
  u = [(ri.astype(np.float32) + f32(0.5)) * s for ri in r]
  ra = np.sqrt(f32(-2.0) * np.log(np.clip(u[0], f32(1e-12), f32(1.0))))
  rb = np.sqrt(f32(-2.0) * np.log(np.clip(u[2], f32(1e-12), f32(1.0))))
  a1 = (f32(2.0) * u[1]).astype(np.float64) * np.pi
  a3 = (f32(2.0) * u[3]).astype(np.float64) * np.pi
  out = np.stack([ra * np.cos(a1).astype(np.float32), ra * np.sin(a1).astype(np.float32),
                  rb * np.cos(a3).astype(np.float32), rb * np.sin(a3).astype(np.float32)], axis=1)
  return out.reshape(-1).astype(np.float32)


def init_z(seed: int, shape) -> np.ndarray:
  return normal(seed, 0, int(np.prod(shape))).reshape(shape)


def step_noise(seed: int, step: int, shape) -> np.ndarray:
  return normal(seed, step + 1, int(np.prod(shape))).reshape(shape)
