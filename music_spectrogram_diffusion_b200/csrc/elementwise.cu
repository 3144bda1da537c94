// Row-wise / element-wise kernels of the DDPM hot path (HBM-bound; vectorised, coalesced):
//   rmsnorm(+FiLM)   msd/layers.py:632-649 (T5 RMS norm), 652-666 (FiLM x*(1+s)+b)
//   sampler step     msd/models/diffusion/diffusion_utils.py:398-453, 382-395, 120-163, 215-222
//   token embedding  msd/layers.py:556-559 + network.py:278-287
//   feature scaling  msd/audio_codecs.py:166-183
//   masks            msd/models/diffusion/network.py:28-51, 546; msd/layers.py:341-348
#include "common.cuh"
#include "kernels.h"
#include "sampler.cuh"

#define MSD_TRY_RC(expr)      \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

namespace msd {

std::atomic<unsigned long long> g_launch_count{0};

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// RMSNorm (+FiLM) : one warp per row, row kept in registers, bf16 output
// ---------------------------------------------------------------------------
struct NormDev {
  const float* x;
  const float* gamma;
  bf16* out;
  const float* film;
  const int* step;
  long long film_step_stride, film_offset;
  int rows, d, ldo, split3;
  int src_len, dst_len, dst_off;  // row remap when src_len > 0
};

constexpr int NORM_MAX_ITERS = 8;  // d <= 1024

// ITERS = d / 128 (float4 per lane), a template parameter so that the per-row constants (gamma
// and the FiLM scale | bias rows) can be requested together with the row itself: their L2 round
// trip then overlaps the row's instead of following the warp reduction (the kernel is a pure
// latency chain: ~5 us for 19 MB of traffic).
template <int ITERS>
__global__ void __launch_bounds__(256) rmsnorm_film_kernel(const NormDev p) {
  griddep_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.rows) return;
  constexpr int D = ITERS * 128;
  // per-segment constants: safe to read ahead of the dependency wait (written at load time)
  float4 g[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i)
    g[i] = __ldg(reinterpret_cast<const float4*>(p.gamma + (i * 32 + lane) * 4));
  griddep_wait();
  const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<size_t>(warp) * D);
  float4 v[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) v[i] = xr[i * 32 + lane];
  float4 fsv[ITERS], fbv[ITERS];
  const bool has_film = p.film != nullptr;
  if (has_film) {
    const long long base = static_cast<long long>(*p.step) * p.film_step_stride + p.film_offset;
    const float* fs = p.film + base;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int c = (i * 32 + lane) * 4;
      fsv[i] = __ldg(reinterpret_cast<const float4*>(fs + c));
      fbv[i] = __ldg(reinterpret_cast<const float4*>(fs + D + c));
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i)
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / static_cast<float>(D) + 1e-6f);
  int orow = warp;
  if (p.src_len > 0) {
    const int b = warp / p.src_len;
    orow = b * p.dst_len + p.dst_off + (warp - b * p.src_len);
  }
  bf16* o = p.out + static_cast<size_t>(orow) * p.ldo;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = (i * 32 + lane) * 4;
    float y0 = v[i].x * inv * g[i].x, y1 = v[i].y * inv * g[i].y;
    float y2 = v[i].z * inv * g[i].z, y3 = v[i].w * inv * g[i].w;
    if (has_film) {
      y0 = y0 * (fsv[i].x + 1.0f) + fbv[i].x; y1 = y1 * (fsv[i].y + 1.0f) + fbv[i].y;
      y2 = y2 * (fsv[i].z + 1.0f) + fbv[i].z; y3 = y3 * (fsv[i].w + 1.0f) + fbv[i].w;
    }
    if (!p.split3) {
      uint2 u;
      u.x = pack_bf16(y0, y1);
      u.y = pack_bf16(y2, y3);
      *reinterpret_cast<uint2*>(o + c) = u;
    } else {
      bf16 h0, h1, h2, h3, l0, l1, l2, l3;
      split_bf16(y0, h0, l0); split_bf16(y1, h1, l1);
      split_bf16(y2, h2, l2); split_bf16(y3, h3, l3);
      __nv_bfloat162 a = __halves2bfloat162(h0, h1), b2 = __halves2bfloat162(h2, h3);
      __nv_bfloat162 c0 = __halves2bfloat162(l0, l1), c1 = __halves2bfloat162(l2, l3);
      uint2 uh, ul;
      uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b2);
      ul.x = *reinterpret_cast<uint32_t*>(&c0); ul.y = *reinterpret_cast<uint32_t*>(&c1);
      *reinterpret_cast<uint2*>(o + c) = uh;            // hi
      *reinterpret_cast<uint2*>(o + D + c) = ul;        // lo
      *reinterpret_cast<uint2*>(o + 2 * D + c) = uh;    // hi
    }
  }
}

template <int ITERS>
int launch_norm_iters(const NormDev& p, cudaStream_t stream) {
  static const int configured = [] {
    return cudaFuncSetAttribute(rmsnorm_film_kernel<ITERS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                100) == cudaSuccess ? 0 : -2;
  }();
  MSD_REQUIRE(configured == 0, "rmsnorm: cudaFuncSetAttribute failed");
  MSD_CUDA_CHECK(launch_kernel(rmsnorm_film_kernel<ITERS>, dim3((p.rows + 7) / 8), dim3(256), 0, stream, p));
  return 0;
}
int launch_norm(const NormDev& p, cudaStream_t stream) {
  switch (p.d >> 7) {
    case 1: return launch_norm_iters<1>(p, stream);
    case 2: return launch_norm_iters<2>(p, stream);
    case 3: return launch_norm_iters<3>(p, stream);
    case 4: return launch_norm_iters<4>(p, stream);
    case 5: return launch_norm_iters<5>(p, stream);
    case 6: return launch_norm_iters<6>(p, stream);
    case 7: return launch_norm_iters<7>(p, stream);
    default: return launch_norm_iters<8>(p, stream);
  }
}

__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerArgs a) {
  griddep_launch_dependents();
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  griddep_wait();
  if (a.run == nullptr) {
    const int step = *a.step;
    prefetch_next_film(a, step, i4);
    if (i4 * 4 < a.n) sampler_step_body(a, step, a.noise, a.mel_out, a.seed, i4);
    return;
  }
  // Per-call arguments and the step index live in device memory (RunArgs).  The step advance is
  // folded in: every block counts itself done once all its threads hold `step` in a register, and
  // the last one to arrive decrements it for the next graph launch.
  __shared__ int s_step;
  if (threadIdx.x == 0) {
    // thread 0 alone reads the step index (its store to shared memory needs the loaded value, so
    // the load has completed before the atomic below is issued) and hands it to the block
    const int st = *reinterpret_cast<volatile int*>(&a.run->step);
    s_step = st;
    __threadfence();
    const unsigned int prev = atomicAdd(&a.run->done, 1u);
    if (prev == gridDim.x - 1) {
      a.run->done = 0u;
      a.run->step = st - 1;
    }
  }
  __syncthreads();
  const int step = s_step;
  prefetch_next_film(a, step, i4);
  const float* noise_base = a.run->noise;
  float* mel_base = a.run->mel_out;
  const unsigned long long seed = a.run->seed;
  if (a.xrole != 0) {
    // ---- guidance split: send my pass's eps to the peer, receive the peer's
    __shared__ unsigned int s_seq;
    if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(&a.run->xseq);
    __syncthreads();
    const unsigned int seq = s_seq;
    const long long par = static_cast<long long>(seq & 1u) * a.xparity_floats;
    if (i4 * 4 < a.n)
      *reinterpret_cast<float4*>(a.xpeer + par + i4 * 4) = *reinterpret_cast<const float4*>(a.eps + i4 * 4);
    __threadfence_system();   // my stores are visible to the peer before the flag is
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(&a.run->xsent, 1u);
      if (prev == gridDim.x - 1) {   // every block's share is on its way: raise the peer's flag
        a.run->xsent = 0u;
        a.run->xseq = seq + 1u;
        unsigned int* pflag = reinterpret_cast<unsigned int*>(a.xpeer + a.xflags_off) + (seq & 1u);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pflag), "r"(seq) : "memory");
      }
      // wait for the peer's values of this step
      const unsigned int* lflag = reinterpret_cast<const unsigned int*>(a.xlocal + a.xflags_off) + (seq & 1u);
      unsigned int v;
      unsigned long long spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(lflag) : "memory");
        if (v != seq) {
          __nanosleep(200);
          if (++spins > (1ull << 26)) __trap();   // ~15 s: the peer is gone, fail loudly
        }
      } while (v != seq);
    }
    __syncthreads();
    const float* other = a.xlocal + par;
    const float* ec = a.xrole == 1 ? a.eps : other;
    const float* eu = a.xrole == 1 ? other : a.eps;
    if (i4 * 4 < a.n) sampler_step_body(a, step, noise_base, mel_base, seed, i4, ec, eu);
    return;
  }
  if (i4 * 4 < a.n) sampler_step_body(a, step, noise_base, mel_base, seed, i4);
}

__global__ void __launch_bounds__(256)
init_z_kernel(const float* init_z, float* z, bf16* zs, long long n, int n_dims,
              unsigned long long seed, int rng_kind, const uint32_t* rng_keys) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long idx = i4 * 4;
  if (idx >= n) return;
  float4 v;
  if (init_z != nullptr) v = *reinterpret_cast<const float4*>(init_z + idx);
  else if (rng_kind == 1) v = jax_normal4(rng_keys, n, i4);
  else v = philox_normal4(seed, 0u, static_cast<unsigned long long>(i4));
  *reinterpret_cast<float4*>(z + idx) = v;
  store_split4(zs, idx, n_dims, v);
}

// ---------------------------------------------------------------------------
// Encoder front ends
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const int* tokens, const float* emb, const float* pos, float* x, int rows,
                    int T, int d, int vocab) {
  const int d4 = d >> 2;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(rows) * d4) return;
  const int row = static_cast<int>(gid / d4), c = static_cast<int>(gid - static_cast<long long>(row) * d4);
  int tok = tokens[row];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const int t = row % T;
  const float4 e = __ldg(reinterpret_cast<const float4*>(emb + static_cast<size_t>(tok) * d) + c);
  const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * d) + c);
  reinterpret_cast<float4*>(x + static_cast<size_t>(row) * d)[c] =
      make_float4(e.x + pp.x, e.y + pp.y, e.z + pp.z, e.w + pp.w);
}

__global__ void __launch_bounds__(256)
scale_split_kernel(const float* feat, bf16* out, long long n, int n_dims, float fmin, float fmax) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long idx = i4 * 4;
  if (idx >= n) return;
  float4 f = *reinterpret_cast<const float4*>(feat + idx);
  // scale_features(clip=True), msd/audio_codecs.py:166-174 with output_range (-1, 1)
  const float inv = 1.0f / (fmax - fmin);
  f.x = (fminf(fmaxf(f.x, fmin), fmax) - fmin) * inv * 2.0f - 1.0f;
  f.y = (fminf(fmaxf(f.y, fmin), fmax) - fmin) * inv * 2.0f - 1.0f;
  f.z = (fminf(fmaxf(f.z, fmin), fmax) - fmin) * inv * 2.0f - 1.0f;
  f.w = (fminf(fmaxf(f.w, fmin), fmax) - fmin) * inv * 2.0f - 1.0f;
  store_split4(out, idx, n_dims, f);
}

// rows of fp32 -> [hi | lo | hi] bf16 rows (A operand of a split-precision GEMM)
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* src, bf16* out, long long n, int cols) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long idx = i4 * 4;
  if (idx >= n) return;
  store_split4(out, idx, cols, *reinterpret_cast<const float4*>(src + idx));
}

// One block per batch row: key-mask bit words for [tokens | context] and the
// terminal-relative roll amount (= get_sequence_length of the context mask).
__global__ void __launch_bounds__(256)
build_masks_kernel(const int* tokens, const int* ctx_mask, int T, int C, uint32_t* bits,
                   int* ctx_seq_len, int terminal_relative) {
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int words_t = T >> 5, words_c = C >> 5;
  uint32_t* brow = bits + static_cast<size_t>(b) * (words_t + words_c);
  __shared__ int first_zero;
  if (threadIdx.x == 0) first_zero = C;
  __syncthreads();
  for (int w = warp; w < words_t; w += nw) {
    const uint32_t m = __ballot_sync(0xffffffffu, tokens[static_cast<size_t>(b) * T + w * 32 + lane] > 0);
    if (lane == 0) brow[w] = m;
  }
  for (int w = warp; w < words_c; w += nw) {
    const int v = ctx_mask[static_cast<size_t>(b) * C + w * 32 + lane];
    const uint32_t m = __ballot_sync(0xffffffffu, v > 0);
    const uint32_t z = __ballot_sync(0xffffffffu, v == 0);
    if (lane == 0) {
      brow[words_t + w] = m;
      if (z) atomicMin(&first_zero, w * 32 + __ffs(z) - 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // network.py:28-39: first zero index, or the full length when there is no zero.
    int len = first_zero;  // == C when no zero was found
    ctx_seq_len[b] = terminal_relative ? (len % C) : 0;  // roll by C == roll by 0
  }
}

// ---------------------------------------------------------------------------
// Load-time: weight packing and fp32 SIMT GEMM
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* W, int K, int N, bf16* dst, int ldd, int n_off, int k_off,
                   int part) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(K) * N) return;
  const int n = static_cast<int>(gid / K), k = static_cast<int>(gid - static_cast<long long>(n) * K);
  const float w = W[static_cast<size_t>(k) * N + n];
  bf16 hi, lo;
  split_bf16(w, hi, lo);
  dst[static_cast<size_t>(n_off + n) * ldd + k_off + k] = part ? lo : hi;
}

__global__ void __launch_bounds__(256)
pack_gated_kernel(const float* W0, const float* W1, int K, int F, bf16* dst, int ldd, int k_off,
                  int part) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= 2LL * F * K) return;
  const int r = static_cast<int>(gid / K), k = static_cast<int>(gid - static_cast<long long>(r) * K);
  const int g = r >> 6, j = r & 63;
  const float* W = (j < 32) ? W0 : W1;
  const int col = g * 32 + (j & 31);
  bf16 hi, lo;
  split_bf16(W[static_cast<size_t>(k) * F + col], hi, lo);
  dst[static_cast<size_t>(r) * ldd + k_off + k] = part ? lo : hi;
}

constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                 int ldc, int M, int N, int K, int act) {
  __shared__ float sA[SG_K][SG_T + 1];
  __shared__ float sB[SG_K][SG_T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    for (int e = threadIdx.x; e < SG_T * SG_K; e += 256) {
      const int am = e / SG_K, ak = e % SG_K;
      const int gm = m0 + am, gk = k0 + ak;
      sA[ak][am] = (gm < M && gk < K) ? A[static_cast<size_t>(gm) * K + gk] : 0.f;
      const int bk = e / SG_T, bn = e % SG_T;
      const int gn = n0 + bn, gk2 = k0 + bk;
      sB[bk][bn] = (gn < N && gk2 < K) ? B[static_cast<size_t>(gk2) * N + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float v = acc[i][j];
        if (act == 1) v = v / (1.0f + expf(-v));  // swish = x * sigmoid(x)
        C[static_cast<size_t>(gm) * ldc + gn] = v;
      }
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* s, bf16* d, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const bf16* s, float* d, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __bfloat162float(s[i]);
}
// dst[r][c] = src[r * ld + c] (+ src[r * ld + lo_off + c] when lo_off > 0: hi + lo of a split row)
__global__ void __launch_bounds__(256)
bf16_rows_to_f32_kernel(const bf16* s, int ld, int lo_off, float* d, long long rows, int cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = static_cast<int>(i - r * cols);
  float v = __bfloat162float(s[r * ld + c]);
  if (lo_off > 0) v += __bfloat162float(s[r * ld + lo_off + c]);
  d[i] = v;
}
__global__ void __launch_bounds__(256)
mask_bits_kernel(const int* mask, long long words, uint32_t* bits) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= words) return;
  const uint32_t m = __ballot_sync(0xffffffffu, mask[w * 32 + lane] > 0);
  if (lane == 0) bits[w] = m;
}

inline int blocks_for(long long n, int per_block) {
  return static_cast<int>((n + per_block - 1) / per_block);
}

}  // namespace

// Every kernel of the per-step graph asks for the maximum shared-memory carve-out, including the
// ones that use no shared memory: alternating carve-outs between consecutive kernels forces an SM
// reconfiguration (the SM must drain first), which also defeats programmatic dependent launch.
int elementwise_configure() {
  MSD_CUDA_CHECK(cudaFuncSetAttribute(sampler_step_kernel,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(init_z_kernel,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return 0;
}

int launch_rmsnorm(const float* x, const float* gamma, int rows, int d, bf16* out, int ldo,
                   const float* film, const int* step, long long film_step_stride,
                   long long film_offset, int split3, cudaStream_t stream) {
  MSD_REQUIRE(d % 128 == 0 && d <= 128 * NORM_MAX_ITERS, "rmsnorm: d=%d must be k*128 <= 1024", d);
  NormDev p;
  p.x = x; p.gamma = gamma; p.out = out; p.film = film; p.step = step;
  p.film_step_stride = film_step_stride; p.film_offset = film_offset;
  p.rows = rows; p.d = d; p.ldo = ldo; p.split3 = split3;
  p.src_len = 0; p.dst_len = 0; p.dst_off = 0;
  ProfScope prof(KC_NORM, 0.0, static_cast<double>(rows) * d * (4.0 + (split3 ? 6.0 : 2.0)), stream);
  MSD_TRY_RC(launch_norm(p, stream));
  ++g_launch_count;
  return 0;
}

int launch_rmsnorm_rows_remap(const float* x, const float* gamma, int B, int src_len, int d,
                              bf16* out, int dst_len, int dst_off, cudaStream_t stream,
                              int split3) {
  MSD_REQUIRE(d % 128 == 0 && d <= 128 * NORM_MAX_ITERS, "rmsnorm: d=%d must be k*128 <= 1024", d);
  NormDev p;
  p.x = x; p.gamma = gamma; p.out = out; p.film = nullptr; p.step = nullptr;
  p.film_step_stride = 0; p.film_offset = 0;
  p.rows = B * src_len; p.d = d; p.ldo = split3 ? 3 * d : d; p.split3 = split3;
  p.src_len = src_len; p.dst_len = dst_len; p.dst_off = dst_off;
  MSD_TRY_RC(launch_norm(p, stream));
  ++g_launch_count;
  return 0;
}

int launch_sampler_step(const SamplerArgs& a, cudaStream_t stream) {
  MSD_REQUIRE(a.n % 4 == 0 && a.n_dims % 4 == 0, "sampler: sizes must be multiples of 4");
  ProfScope prof(KC_SAMPLER, 0.0, static_cast<double>(a.n) * (4.0 * (a.passes + 3) + 6.0), stream);
  MSD_CUDA_CHECK(launch_kernel(sampler_step_kernel, dim3(blocks_for(a.n / 4, 256)), dim3(256), 0, stream, a));
  ++g_launch_count;
  return 0;
}

__global__ void __launch_bounds__(256)
jax_normal_kernel(uint32_t k0, uint32_t k1, long long n, float* out) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  const uint32_t key[2] = {k0, k1};
  *reinterpret_cast<float4*>(out + i4 * 4) = jax_normal4(key, n, i4);
}

__global__ void __launch_bounds__(256)
jax_bits_kernel(uint32_t k0, uint32_t k1, long long n, uint32_t* out) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  const uint32_t key[2] = {k0, k1};
  *reinterpret_cast<uint4*>(out + i4 * 4) = jax_bits4(key, n, i4);
}

int launch_jax_bits(uint32_t k0, uint32_t k1, long long n, uint32_t* out, cudaStream_t stream) {
  MSD_REQUIRE(n > 0 && n % 8 == 0 && n < (1ll << 32), "jax_bits: n must be k*8 < 2^32");
  jax_bits_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(k0, k1, n, out);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_jax_normal(uint32_t k0, uint32_t k1, long long n, float* out, cudaStream_t stream) {
  MSD_REQUIRE(n > 0 && n % 8 == 0 && n < (1ll << 32), "jax_normal: n must be k*8 < 2^32");
  jax_normal_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(k0, k1, n, out);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_init_z(const float* init_z, float* z, bf16* z_split, long long n, int n_dims,
                  unsigned long long seed, cudaStream_t stream, int rng_kind,
                  const uint32_t* rng_keys) {
  MSD_REQUIRE(rng_kind == 0 || (rng_keys != nullptr && n % 8 == 0 && n < (1ll << 32)),
              "init_z: the jax stream needs its key table and a draw of k*8 < 2^32 elements");
  init_z_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(init_z, z, z_split, n, n_dims, seed,
                                                            rng_kind, rng_keys);
  MSD_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_embed_tokens(const int* tokens, const float* emb, const float* pos, float* x, int B,
                        int T, int d, int vocab, cudaStream_t stream) {
  const long long n = static_cast<long long>(B) * T * (d / 4);
  embed_tokens_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(tokens, emb, pos, x, B * T, T, d,
                                                              vocab);
  MSD_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_scale_split(const float* feat, bf16* out_split, long long rows, int n_dims, float fmin,
                       float fmax, cudaStream_t stream) {
  const long long n = rows * n_dims;
  scale_split_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(feat, out_split, n, n_dims, fmin,
                                                                 fmax);
  MSD_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_split3_rows(const float* src, bf16* out_split, long long rows, int cols,
                       cudaStream_t stream) {
  MSD_REQUIRE(cols % 4 == 0, "split3_rows: cols must be a multiple of 4");
  const long long n = rows * cols;
  split3_rows_kernel<<<blocks_for(n / 4, 256), 256, 0, stream>>>(src, out_split, n, cols);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_build_masks(const int* tokens, const int* ctx_mask, int B, int T, int C, uint32_t* bits,
                       int* ctx_seq_len, int terminal_relative, cudaStream_t stream) {
  MSD_REQUIRE(T % 128 == 0 && C % 128 == 0, "masks: lengths must be multiples of 128");
  build_masks_kernel<<<B, 256, 0, stream>>>(tokens, ctx_mask, T, C, bits, ctx_seq_len,
                                            terminal_relative);
  MSD_CUDA_CHECK(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int launch_pack_weight(const float* W, int K, int N, bf16* dst, int ldd, int n_off, int k_off,
                       int part, cudaStream_t stream) {
  pack_weight_kernel<<<blocks_for(static_cast<long long>(K) * N, 256), 256, 0, stream>>>(
      W, K, N, dst, ldd, n_off, k_off, part);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_pack_gated(const float* W0, const float* W1, int K, int F, bf16* dst, int ldd,
                      cudaStream_t stream, int k_off, int part) {
  MSD_REQUIRE(F % 32 == 0, "pack_gated: F must be a multiple of 32");
  pack_gated_kernel<<<blocks_for(2LL * F * K, 256), 256, 0, stream>>>(W0, W1, K, F, dst, ldd, k_off,
                                                                      part);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_f32_to_bf16(const float* src, bf16* dst, long long n, cudaStream_t stream) {
  f32_to_bf16_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(src, dst, n);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_bf16_to_f32(const bf16* src, float* dst, long long n, cudaStream_t stream) {
  bf16_to_f32_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(src, dst, n);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_bf16_rows_to_f32(const bf16* src, int ld, int lo_off, float* dst, long long rows, int cols,
                            cudaStream_t stream) {
  bf16_rows_to_f32_kernel<<<blocks_for(rows * cols, 256), 256, 0, stream>>>(src, ld, lo_off, dst,
                                                                            rows, cols);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int launch_mask_bits(const int* mask, int nb, int L, uint32_t* bits, cudaStream_t stream) {
  MSD_REQUIRE(L % 128 == 0, "mask_bits: L must be a multiple of 128");
  const long long words = static_cast<long long>(nb) * (L / 32);
  mask_bits_kernel<<<blocks_for(words * 32, 256), 256, 0, stream>>>(mask, words, bits);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_sgemm_f32(const float* A, const float* B, float* C, int ldc, int M, int N, int K,
                     int act, cudaStream_t stream) {
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T);
  sgemm_f32_kernel<<<grid, 256, 0, stream>>>(A, B, C, ldc, M, N, K, act);
  MSD_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace msd
