// Device functions of the reverse-diffusion update shared by the stand-alone sampler kernel
// (elementwise.cu) and the sampler epilogue of the final-projection GEMM (gemm_tcgen05.cu):
// noise generators (Philox4x32-10, jax.random threefry2x32 + XLA erfinv), split-precision stores
// and the update itself (msd/models/diffusion/diffusion_utils.py:398-453, 382-395, 120-163,
// 215-222; audio_codecs.py:176-183).
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace msd {
namespace {

__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (perf-mode noise; parity runs inject noise instead)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, uint32_t stream,
                                                 unsigned long long idx4) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(idx4), static_cast<uint32_t>(idx4 >> 32), stream, 0x6d7364u,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float s = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = (static_cast<float>(r[0]) + 0.5f) * s, u1 = (static_cast<float>(r[1]) + 0.5f) * s;
  const float u2 = (static_cast<float>(r[2]) + 0.5f) * s, u3 = (static_cast<float>(r[3]) + 0.5f) * s;
  const float ra = sqrtf(-2.0f * logf(fminf(fmaxf(u0, 1e-12f), 1.0f)));
  const float rb = sqrtf(-2.0f * logf(fminf(fmaxf(u2, 1e-12f), 1.0f)));
  float sa, ca, sb, cb;
  sincospif(2.0f * u1, &sa, &ca);
  sincospif(2.0f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__device__ __forceinline__ void store_split4(bf16* zs, long long idx, int n_dims, float4 v) {
  const long long row = idx / n_dims;
  const int col = static_cast<int>(idx - row * n_dims);
  bf16* o = zs + row * (3LL * n_dims) + col;
  bf16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1);
  split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
  __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
  __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
  uint2 uh, ul;
  uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b);
  ul.x = *reinterpret_cast<uint32_t*>(&c); ul.y = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint2*>(o) = uh;
  *reinterpret_cast<uint2*>(o + n_dims) = ul;
  *reinterpret_cast<uint2*>(o + 2 * n_dims) = uh;
}

// ---------------------------------------------------------------------------
// jax.random (threefry2x32) noise, restated from the published algorithm (jax 0.3.16 defaults;
// CPU twin and derivation: music_spectrogram_diffusion_b200/jax_rng.py).  Element e of an
// n-element draw is word e of threefry_2x32(key, arange(n)): the counters are split into halves,
// so e < n/2 is the first output word of the pair (e, e + n/2) and e >= n/2 the second word of
// (e - n/2, e).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    const int r0 = (g & 1) ? 17 : 13, r1 = (g & 1) ? 29 : 15, r2 = (g & 1) ? 16 : 26, r3 = (g & 1) ? 24 : 6;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r0) ^ x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r1) ^ x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r2) ^ x0;
    x0 += x1; x1 = __funnelshift_l(x1, x1, r3) ^ x0;
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + static_cast<uint32_t>(g + 1);
  }
  return make_uint2(x0, x1);
}

// XLA's float32 erfinv (Giles' two single-precision polynomials in w = -log1p(-x^2))
__device__ __forceinline__ float erfinv_xla(float x) {
  const float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    const float v = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * v; p = -3.5233877e-06f + p * v; p = -4.39150654e-06f + p * v;
    p = 0.00021858087f + p * v; p = -0.00125372503f + p * v; p = -0.00417768164f + p * v;
    p = 0.246640727f + p * v; p = 1.50140941f + p * v;
  } else {
    const float v = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * v; p = 0.00134934322f + p * v; p = -0.00367342844f + p * v;
    p = 0.00573950773f + p * v; p = -0.0076224613f + p * v; p = 0.00943887047f + p * v;
    p = 1.00167406f + p * v; p = 2.83297682f + p * v;
  }
  return p * x;
}

__device__ __forceinline__ float jax_normal_from_bits(uint32_t bits) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;                       // nextafter(-1, 0); (1 - lo) rounds to 2
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
  return 1.41421354f * erfinv_xla(u);
}

// random words for elements [4*i4, 4*i4 + 4) of an n-element draw (n a multiple of 8)
__device__ __forceinline__ uint4 jax_bits4(const uint32_t* key, long long n, long long i4) {
  const uint32_t k0 = key[0], k1 = key[1];
  const long long half = n >> 1, e = i4 * 4;
  uint32_t r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long ej = e + j;
    const bool second = ej >= half;
    const uint32_t c0 = static_cast<uint32_t>(second ? ej - half : ej);
    const uint2 o = threefry2x32(k0, k1, c0, static_cast<uint32_t>(c0 + half));
    r[j] = second ? o.y : o.x;
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}
// normals for the same elements
__device__ __forceinline__ float4 jax_normal4(const uint32_t* key, long long n, long long i4) {
  const uint4 b = jax_bits4(key, n, i4);
  return make_float4(jax_normal_from_bits(b.x), jax_normal_from_bits(b.y),
                     jax_normal_from_bits(b.z), jax_normal_from_bits(b.w));
}

// ---------------------------------------------------------------------------
// One reverse-diffusion update (CFG combine + x0 + clip + DDPM/DDIM mean + noise)
// ---------------------------------------------------------------------------
// One reverse-diffusion update of elements [4*i4, 4*i4 + 4): `mo` is the model output of the
// conditional pass for them, `mu` of the unconditional pass (ignored when a.passes == 1).
__device__ __forceinline__ void sampler_update4(const SamplerArgs& a, int step,
                                                const float* noise_base, float* mel_base,
                                                unsigned long long seed, long long i4, float4 mo,
                                                float4 mu) {
  const long long idx = i4 * 4;
  const float* cf = a.coef + static_cast<size_t>(step) * MSD_STEP_COLS;
  const float x0_scale = cf[0], eps_scale = cf[1], c_z = cf[2], c_x0 = cf[3], sigma = cf[4];
  const bool last = cf[5] != 0.f;
  const float p0 = cf[8], p1 = cf[9], q0 = cf[10], q1 = cf[11], e1 = cf[12], e2 = cf[13];
  const float4 z = *reinterpret_cast<const float4*>(a.z + idx);
  // _get_x0_and_eps_from_model_output (diffusion_utils.py:288-321): eps = p0 z + p1 out and
  // x0 = q0 z + q1 out (for model_output == 'eps': p0 = 0, p1 = 1, i.e. eps = out exactly)
  float4 e, x0;
  e.x = fmaf(p1, mo.x, p0 * z.x); e.y = fmaf(p1, mo.y, p0 * z.y);
  e.z = fmaf(p1, mo.z, p0 * z.z); e.w = fmaf(p1, mo.w, p0 * z.w);
  if (p0 == 0.f && p1 == 1.f) e = mo;
  if (a.passes == 2) {
    // classifier-free guidance on eps, then x0 from the combined eps at logsnr_t (424-433)
    float4 eu;
    eu.x = fmaf(p1, mu.x, p0 * z.x); eu.y = fmaf(p1, mu.y, p0 * z.y);
    eu.z = fmaf(p1, mu.z, p0 * z.z); eu.w = fmaf(p1, mu.w, p0 * z.w);
    if (p0 == 0.f && p1 == 1.f) eu = mu;
    const float w = a.cond_weight, wu = 1.0f - a.cond_weight;
    e.x = w * e.x + wu * eu.x; e.y = w * e.y + wu * eu.y;
    e.z = w * e.z + wu * eu.z; e.w = w * e.w + wu * eu.w;
    x0.x = x0_scale * (z.x - e.x * eps_scale); x0.y = x0_scale * (z.y - e.y * eps_scale);
    x0.z = x0_scale * (z.z - e.z * eps_scale); x0.w = x0_scale * (z.w - e.w * eps_scale);
  } else if (q0 == 0.f && q1 == 1.f) {
    x0 = mo;
  } else if (p0 == 0.f && p1 == 1.f) {
    // predict_x0_from_eps at the train schedule's logsnr: q0 = A, q1 = -A * B
    const float A = q0, Bc = -q1 / q0;
    x0.x = A * (z.x - mo.x * Bc); x0.y = A * (z.y - mo.y * Bc);
    x0.z = A * (z.z - mo.z * Bc); x0.w = A * (z.w - mo.w * Bc);
  } else {
    x0.x = fmaf(q1, mo.x, q0 * z.x); x0.y = fmaf(q1, mo.y, q0 * z.y);
    x0.z = fmaf(q1, mo.z, q0 * z.z); x0.w = fmaf(q1, mo.w, q0 * z.w);
  }
  if (a.clip_x0) {
    x0.x = fminf(fmaxf(x0.x, -1.f), 1.f); x0.y = fminf(fmaxf(x0.y, -1.f), 1.f);
    x0.z = fminf(fmaxf(x0.z, -1.f), 1.f); x0.w = fminf(fmaxf(x0.w, -1.f), 1.f);
    if (a.ddim) {  // pred_eps = predict_eps_from_x0(z, clipped x0, logsnr_t) (437-439)
      e.x = e1 * (z.x - x0.x * e2); e.y = e1 * (z.y - x0.y * e2);
      e.z = e1 * (z.z - x0.z * e2); e.w = e1 * (z.w - x0.w * e2);
    }
  }
  float4 zn;
  if (last) {
    zn = x0;
  } else if (a.ddim) {
    // ddim_step (369-379): z_s = alpha_s x0 + stdv_s eps; table columns 3 / 2
    zn.x = c_x0 * x0.x + c_z * e.x; zn.y = c_x0 * x0.y + c_z * e.y;
    zn.z = c_x0 * x0.z + c_z * e.z; zn.w = c_x0 * x0.w + c_z * e.w;
  } else {
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sigma != 0.f) {
      if (noise_base != nullptr) {
        nz = *reinterpret_cast<const float4*>(noise_base + static_cast<size_t>(step) * a.n + idx);
      } else {
        nz = a.rng_kind == 1
                 ? jax_normal4(a.rng_keys + 2 * (step + 1), a.n, i4)
                 : philox_normal4(seed, static_cast<uint32_t>(step) + 1u,
                                  static_cast<unsigned long long>(i4));
      }
    }
    zn.x = c_z * z.x + c_x0 * x0.x + sigma * nz.x; zn.y = c_z * z.y + c_x0 * x0.y + sigma * nz.y;
    zn.z = c_z * z.z + c_x0 * x0.z + sigma * nz.z; zn.w = c_z * z.w + c_x0 * x0.w + sigma * nz.w;
  }
  *reinterpret_cast<float4*>(a.z + idx) = zn;
  store_split4(a.z_split, idx, a.n_dims, zn);
  if (last && mel_base != nullptr) {
    // scale_to_features, msd/audio_codecs.py:176-183 with input_range (-1, 1)
    const float span = a.feat_max - a.feat_min;
    float4 f;
    f.x = (zn.x + 1.f) * 0.5f * span + a.feat_min; f.y = (zn.y + 1.f) * 0.5f * span + a.feat_min;
    f.z = (zn.z + 1.f) * 0.5f * span + a.feat_min; f.w = (zn.w + 1.f) * 0.5f * span + a.feat_min;
    *reinterpret_cast<float4*>(mel_base + idx) = f;
  }
}

__device__ __forceinline__ void prefetch_next_film(const SamplerArgs& a, int step, long long gid) {
  if (a.film == nullptr || step < 1) return;
  const long long off = gid * 32;  // one 128-byte line per thread
  if (off < a.film_step_floats) {
    const float* ptr = a.film + static_cast<long long>(step - 1) * a.film_step_floats + off;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
  }
}

// the same, with the model outputs read from memory
__device__ __forceinline__ void sampler_step_body(const SamplerArgs& a, int step,
                                                  const float* noise_base, float* mel_base,
                                                  unsigned long long seed, long long i4,
                                                  const float* eps_cond = nullptr,
                                                  const float* eps_uncond = nullptr) {
  if (eps_cond == nullptr) {
    eps_cond = a.eps;
    eps_uncond = a.eps + a.n;
  }
  const float4 mo = *reinterpret_cast<const float4*>(eps_cond + i4 * 4);
  const float4 mu = a.passes == 2 ? *reinterpret_cast<const float4*>(eps_uncond + i4 * 4)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  sampler_update4(a, step, noise_base, mel_base, seed, i4, mo, mu);
}

}  // namespace
}  // namespace msd
