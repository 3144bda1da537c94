// bf16 GEMM on the 5th-gen tensor cores: D[M,N] = A[M,K] * B[N,K]^T.
//
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor 2D tiles (SWIZZLE_128B) into a
//                      STAGES-deep smem ring, completion on `full` mbarriers
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BN x K=16,
//                      fp32 accumulator in TMEM; tcgen05.commit frees ring slots (`empty`)
//                      and finally signals `tmem_full`
//   warps 2..5         epilogue: tcgen05.ld (thread = accumulator row), fused residual /
//                      gated-GELU / position-add, vectorised global stores
//
// Replaces the XLA dot_general lowering of DenseGeneral (msd/layers.py:397-442) for every
// projection on the hot path.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace msd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

__host__ __device__ inline bool epi_is_bf16_out(int e) {
  return e == EPI_BF16 || e == EPI_GATED_GELU || e == EPI_GATED_GELU_SPLIT3;
}

// exact-tanh GELU of the fp32-accurate mode (flax.linen.gelu(approximate=True))
__device__ __forceinline__ float gelu_tanh_exact(float x) {
  const float k0 = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(k0 * (x + 0.044715f * x * x * x)));
}

struct GemmDev {
  int M, N, K;
  int epilogue;
  void* out;
  int ldo;
  const float* resid;
  const float* pos;
  int pos_rows;
  const int* pos_shift;
  int dup_rows;
  // debugging (MSD_GEMM_TRACE with msd_bench_gemm): per CTA 8 int64 -- smid, globaltimer at entry /
  // exit, then clock64 at: entry, set-up done (after the cluster sync), dependency wait returned
  // (epilogue warp), first accumulator ready, last chunk stored, exit
  long long* trace;
  GemmPrep prep;       // EPI_RESID_PREP (DN instances only)
  GemmRowScale rs;     // row scale + bias on EPI_BF16 / EPI_GATED_GELU (DN instances only)
  const int* step;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ void store_bf16x32(bf16* dst, const float* v) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    d4[i] = u;
  }
}

template <int BN>
__global__ void __launch_bounds__(192, (BN <= 128 ? 2 : 1))
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BLOCK_M;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  griddep_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      griddep_wait();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tma_load_2d(sA + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BLOCK_K, m0);
        tma_load_2d(sB + s * Cfg::B_STAGE_BYTES, &tmap_b, &full_bar[s], kb * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, 0, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * Cfg::B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2, 1024, 16);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2, 1024, 16);
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // slot reusable once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue: 4 warps, TMEM lane group = warp % 4 ----------------
    const int lg = warp & 3;
    const int row = m0 + lg * 32 + lane;
    griddep_wait();
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const bool row_ok = row < p.M;
    uint32_t r[32];
    if (p.epilogue == EPI_GATED_GELU) {
      uint32_t g[32];
      bf16* out = reinterpret_cast<bf16*>(p.out);
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        tmem_ld_32x32b_x32(t_row + c, r);
        tmem_ld_32x32b_x32(t_row + c + 32, g);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v[i] = gelu_tanh(__uint_as_float(r[i])) * __uint_as_float(g[i]);
        if (row_ok) store_bf16x32(out + static_cast<size_t>(row) * p.ldo + (n0 + c) / 2, v);
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        tmem_ld_32x32b_x32(t_row + c, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const int col = n0 + c;
        if (!row_ok) continue;
        if (p.epilogue == EPI_BF16) {
          store_bf16x32(reinterpret_cast<bf16*>(p.out) + static_cast<size_t>(row) * p.ldo + col,
                        v);
        } else {
          float* out = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col;
          if (p.epilogue == EPI_RESID_F32) {
            const float4* rs = reinterpret_cast<const float4*>(
                p.resid + static_cast<size_t>(row) * p.ldo + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 q = rs[i];
              v[4 * i + 0] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          } else if (p.epilogue == EPI_POS_F32) {
            const int seq = row / p.pos_rows;
            int pr = row - seq * p.pos_rows;
            if (p.pos_shift != nullptr) {
              pr -= p.pos_shift[seq];
              if (pr < 0) pr += p.pos_rows;
            }
            const float4* ps =
                reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(pr) * p.N + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 q = __ldg(ps + i);
              v[4 * i + 0] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o4[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (p.epilogue == EPI_POS_F32 && p.dup_rows > 0) {
            float4* o4b = reinterpret_cast<float4*>(out + static_cast<size_t>(p.dup_rows) * p.ldo);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o4b[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair persistent GEMM: a 2-CTA cluster owns a 256 x BN output tile (each CTA 128 rows and
// half of the B tile in its shared memory); the even CTA issues tcgen05.mma.cta_group::2 for
// both.  Per k-block each CTA pulls 16 KB of A and BN/2 x 128 B of B: half the L2 traffic per
// FLOP of the 1-CTA kernel above, which ncu showed pinned at the L2 bandwidth cap with the
// tensor pipe ~30 % active.  Clusters are persistent over tiles; the TMEM accumulator is
// double-buffered so tile i's epilogue overlaps tile i+1's main loop.
//   warp 0 (one lane, both CTAs)   TMA producer (own A rows, own half of B), bytes credited to
//                                  the leader's `full` barrier
//   warp 1 (one lane, leader)      MMA issuer; commits multicast to both CTAs' barriers
//   warps 2..5 (both CTAs)         epilogue of the CTA's own 128 accumulator rows
// ---------------------------------------------------------------------------------------------
template <int BN, int DN = 0>
struct PairCfg {
  static constexpr int HALF_N = BN / 2;
  static constexpr int B_STAGE_BYTES = HALF_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // epilogue staging: O ring 4 x [128 rows x 128 B] (source of the TMA stores / TMA reduce-adds;
  // the bf16 path uses the first slot as 4 per-warp transpose tiles).  The deferred-normalisation
  // producer (DN == 2) adds an operand ring 4 x [128 rows x 64 B] behind it and pays with stages.
  static constexpr int O_RING_BYTES = 4 * 16384;
  // (+ 2 x BN floats: the tile's two column-gain rows, staged once per tile)
  static constexpr int A_RING_BYTES = DN == 2 ? 4 * 8192 + 2 * BN * 4 : 0;
  static constexpr int EPI_BYTES = O_RING_BYTES + A_RING_BYTES;
  static constexpr int STAGE_BUDGET = DN == 2 ? (227 * 1024 - 1024 - 512 - EPI_BYTES) : 160 * 1024;
  static constexpr int STAGES = STAGE_BUDGET / STAGE_BYTES > 8 ? 8 : STAGE_BUDGET / STAGE_BYTES;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + EPI_BYTES + 512 /*barriers*/ + 1024 /*align*/;
  static constexpr uint32_t ACC_COLS = BN;  // columns per accumulator buffer
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
};

// ---- coalesced epilogue ---------------------------------------------------------------------
// tcgen05.ld hands every thread ONE accumulator row; storing rows straight from registers makes
// each warp-level store touch 32 cache lines.  Instead each epilogue warp transposes its
// 32 rows x (64 or 128) bytes through a private, XOR-swizzled smem tile so that global loads and
// stores cover whole 64/128-byte row segments (8 or 4 lanes per row).
//   stage_write<K16>: thread `lane` (row) writes its K16 16-byte chunks
//   stage_read <K16>: lane reads chunk (lane % K16) of row (it * (32 / K16) + lane / K16)
template <int K16>
__device__ __forceinline__ void stage_write(uint8_t* tile, int lane, const uint4* chunks) {
#pragma unroll
  for (int c = 0; c < K16; ++c)
    *reinterpret_cast<uint4*>(tile + lane * 128 + ((c ^ (lane & 7)) * 16)) = chunks[c];
}
template <int K16>
__device__ __forceinline__ uint4 stage_read(const uint8_t* tile, int lane, int it, int* row_out,
                                            int* chunk_out) {
  constexpr int RPI = 32 / K16;  // rows per iteration
  const int row = it * RPI + lane / K16;
  const int chunk = lane % K16;
  *row_out = row;
  *chunk_out = chunk;
  return *reinterpret_cast<const uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) * 16));
}

// fp32 outputs: 32 accumulator columns [col, col+32) of rows [row0, row0+32)
__device__ __forceinline__ void epilogue_f32_chunk(const GemmDev& p, uint8_t* tile, int lane,
                                                   int row0, int col, const uint32_t (&r)[32]) {
  uint4 ch[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) ch[c] = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
  stage_write<8>(tile, lane, ch);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int rr, cc;
    const uint4 u = stage_read<8>(tile, lane, it, &rr, &cc);
    const int row = row0 + rr;
    if (row >= p.M) continue;
    float4 v = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z),
                           __uint_as_float(u.w));
    const size_t off = static_cast<size_t>(row) * p.ldo + col + cc * 4;
    if (p.epilogue == EPI_RESID_F32) {
      const float4 q = *reinterpret_cast<const float4*>(p.resid + off);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    } else if (p.epilogue == EPI_POS_F32) {
      const int seq = row / p.pos_rows;
      int pr = row - seq * p.pos_rows;
      if (p.pos_shift != nullptr) {
        pr -= p.pos_shift[seq];
        if (pr < 0) pr += p.pos_rows;
      }
      const float4 q = __ldg(reinterpret_cast<const float4*>(
          p.pos + static_cast<size_t>(pr) * p.N + col + cc * 4));
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    float* out = reinterpret_cast<float*>(p.out);
    *reinterpret_cast<float4*>(out + off) = v;
    if (p.epilogue == EPI_POS_F32 && p.dup_rows > 0)
      *reinterpret_cast<float4*>(out + off + static_cast<size_t>(p.dup_rows) * p.ldo) = v;
  }
  __syncwarp();
}

// bf16 outputs: K16 16-byte chunks per row (64 or 32 output columns) starting at out column `col`
template <int K16>
__device__ __forceinline__ void epilogue_bf16_rows(const GemmDev& p, uint8_t* tile, int lane,
                                                   int row0, int col, const uint4* chunks) {
  stage_write<K16>(tile, lane, chunks);
  __syncwarp();
  bf16* out = reinterpret_cast<bf16*>(p.out);
#pragma unroll
  for (int it = 0; it < K16; ++it) {
    int rr, cc;
    const uint4 u = stage_read<K16>(tile, lane, it, &rr, &cc);
    const int row = row0 + rr;
    if (row < p.M)
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * p.ldo + col + cc * 8) = u;
  }
  __syncwarp();
}

// acc[0:32] = r, acc[32:64] = g:  acc = acc * inv_r + bias[0:64]
__device__ __forceinline__ void scale_bias_64(uint32_t (&r)[32], uint32_t (&g)[32], float inv_r,
                                              const float* bias) {
  if (bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b0 = b4[q], b1 = b4[8 + q];  // staged in shared memory: broadcast reads
      r[4 * q + 0] = __float_as_uint(fmaf(__uint_as_float(r[4 * q + 0]), inv_r, b0.x));
      r[4 * q + 1] = __float_as_uint(fmaf(__uint_as_float(r[4 * q + 1]), inv_r, b0.y));
      r[4 * q + 2] = __float_as_uint(fmaf(__uint_as_float(r[4 * q + 2]), inv_r, b0.z));
      r[4 * q + 3] = __float_as_uint(fmaf(__uint_as_float(r[4 * q + 3]), inv_r, b0.w));
      g[4 * q + 0] = __float_as_uint(fmaf(__uint_as_float(g[4 * q + 0]), inv_r, b1.x));
      g[4 * q + 1] = __float_as_uint(fmaf(__uint_as_float(g[4 * q + 1]), inv_r, b1.y));
      g[4 * q + 2] = __float_as_uint(fmaf(__uint_as_float(g[4 * q + 2]), inv_r, b1.z));
      g[4 * q + 3] = __float_as_uint(fmaf(__uint_as_float(g[4 * q + 3]), inv_r, b1.w));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      r[i] = __float_as_uint(__uint_as_float(r[i]) * inv_r);
      g[i] = __float_as_uint(__uint_as_float(g[i]) * inv_r);
    }
  }
}

// DN: instances carrying the deferred-normalisation epilogues (kernels.h): 1 = consumer (row
// scale + bias row, GemmRowScale), 2 = producer (EPI_RESID_PREP, GemmPrep; tmap_aux = the bf16
// operand it writes).  The plain instances (0) stay free of that code.
template <int BN, int DN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                              const __grid_constant__ CUtensorMap tmap_b,
                              const __grid_constant__ CUtensorMap tmap_out,
                              const __grid_constant__ CUtensorMap tmap_aux, const GemmDev p) {
  using Cfg = PairCfg<BN, DN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sEpi = sB + STAGES * Cfg::B_STAGE_BYTES;  // O ring [4][16 KB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2] (leader's copy is the one used)
  uint64_t* xload_bar = tmem_empty_bar + 2;       // [4] residual chunk landed in its O-ring slot (DN 2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xload_bar + 4);
  uint8_t* sAring = sEpi + Cfg::O_RING_BYTES;     // [4][8 KB] (DN 2)

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  long long* trc = (p.trace != nullptr && threadIdx.x == 64) ? p.trace + blockIdx.x * 8 : nullptr;
  if (trc) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trc[0] = smid; trc[1] = t; trc[3] = clock64();
  }
  const int num_kb = p.K / BLOCK_K;
  const int m_pairs = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = p.N / BN;
  const int num_tiles = m_pairs * n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);  // 4 epilogue warps x 2 CTAs
    }
    for (int a = 0; a < 4; ++a) mbar_init(&xload_bar[a], 1);
    fence_barrier_init();
  }
  if (warp == 2 && lane == 0 && !epi_is_bf16_out(p.epilogue)) tma_prefetch_desc(&tmap_out);
  if (DN == 2 && warp == 3 && lane == 0) tma_prefetch_desc(&tmap_aux);
  if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (trc) trc[4] = clock64();

  griddep_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      // The weight (B) tiles do not depend on the previous kernel: the first ring-full of them
      // is requested BEFORE griddepcontrol.wait so the fetch overlaps the predecessor's tail.
      int prefetched = 0;
      if (cluster_id < num_tiles) {
        const int n0 = (cluster_id / m_pairs) * BN + static_cast<int>(rank) * Cfg::HALF_N;
        prefetched = num_kb < STAGES ? num_kb : STAGES;
        for (int kb = 0; kb < prefetched; ++kb) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[kb], 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2sm(sB + kb * Cfg::B_STAGE_BYTES, &tmap_b, &full_bar[kb], kb * BLOCK_K, n0);
        }
      }
      griddep_wait();
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile % m_pairs) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
        const int n0 = (tile / m_pairs) * BN + static_cast<int>(rank) * Cfg::HALF_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (it < prefetched) {  // slot was armed and its B tile requested above
            tma_load_2d_2sm(sA + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BLOCK_K, m0);
            continue;
          }
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2sm(sA + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BLOCK_K, m0);
          tma_load_2d_2sm(sB + s * Cfg::B_STAGE_BYTES, &tmap_b, &full_bar[s], kb * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BN, 0, 0);
      const uint32_t a_lo0 = desc_lo_sw128(smem_u32(sA)), b_lo0 = desc_lo_sw128(smem_u32(sB));
      int it = 0, tcount = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
        const int acc = tcount & 1;
        const uint32_t acc_ph = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);  // both epilogues drained this buffer
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_lo = a_lo0 + s * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + s * (Cfg::B_STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_2sm_lo(d_tmem, a_lo + k * 2, b_lo + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[s]);  // frees the slot in BOTH CTAs
        }
        umma_commit_2sm(&tmem_full_bar[acc]);  // accumulator ready in BOTH CTAs
      }
    }
  } else {
    // ---------------- epilogue: 4 warps, TMEM lane group = warp % 4 ----------------
    const int lg = warp & 3;
    uint8_t* tile_smem = sEpi + lg * 4096;
    uint8_t* sO = sEpi;               // [4][16 KB]
    const bool has_res = p.epilogue == EPI_RESID_F32;
    const bool epi_leader = (warp == 2 && lane == 0);
    constexpr int NCH = BN / 32;
    uint32_t gc = 0;  // running fp32 chunk counter (ring slot = gc & 3)
    int tcount = 0;
    griddep_wait();  // residual reads / output writes come after the predecessor is complete
    if (trc) trc[5] = clock64();
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
      const int acc = tcount & 1;
      const uint32_t acc_ph = (tcount >> 1) & 1;
      const int tile_row = (tile % m_pairs) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
      const int row0 = tile_row + lg * 32;
      const int n0 = (tile / m_pairs) * BN;
      // deferred normalisation, consumer side: this thread's row scale and the bias row
      // (requested before the accumulator is awaited)
      float inv_r = 1.0f;
      const float* bias = nullptr;
      if constexpr (DN == 1) {
        if (p.rs.ss_lo != nullptr) {
          const int grow = row0 + lane;
          const bool lo = grow < p.rs.split_row;
          const float* ssp = (lo ? p.rs.ss_lo : p.rs.ss_hi) + grow;
          const int parts = lo ? p.rs.parts_lo : p.rs.parts_hi;
          float ss = 0.f;
          if (grow < p.M)
            for (int t = 0; t < parts; ++t) ss += ssp[static_cast<size_t>(t) * p.rs.ss_stride];
          inv_r = rsqrtf(ss * p.rs.inv_d + 1e-6f);
          if (p.rs.col_bias != nullptr) {
            // the tile's bias row goes to shared memory once (slots 1.. of the O ring are unused by
            // the bf16 epilogues), double-buffered over tiles: a warp that runs ahead by one tile
            // must not overwrite what a slower warp still reads
            float* sb = reinterpret_cast<float*>(sEpi + 16384) + (tcount & 1) * BN;
            const int e = static_cast<int>(threadIdx.x) - 64;
            if (e < BN / 4)
              reinterpret_cast<float4*>(sb)[e] = __ldg(reinterpret_cast<const float4*>(
                  p.rs.col_bias + static_cast<long long>(*p.step) * p.rs.bias_step_stride + n0) + e);
            named_barrier_sync_c<1>(128);
            bias = sb;
          }
        }
      }
      // deferred normalisation, producer side: the residual chunks are TMA-loaded into the O-ring
      // slots they will be stored from (the first four before the accumulator is awaited)
      bool prep = false;
      const float* gvec = nullptr;
      if constexpr (DN == 2) {
        prep = p.epilogue == EPI_RESID_PREP && tile_row < p.M;
        if (prep) {
          if (epi_leader) {
            if (tcount > 0) tma_store_wait_read<0>();  // the previous tile's stores have left the ring
            for (int c = 0; c < (NCH < 4 ? NCH : 4); ++c) {
              const uint32_t slot = (gc + c) & 3;
              mbar_arrive_expect_tx(&xload_bar[slot], 16384);
              tma_load_2d(sO + slot * 16384, &tmap_out, &xload_bar[slot], n0 + c * 32, tile_row);
            }
          }
          // the tile's two column-gain rows (rows below / from split_row) go to shared memory once
          float* sg = reinterpret_cast<float*>(sAring + 4 * 8192);
          {
            const long long st = *p.step;
            const int e = static_cast<int>(threadIdx.x) - 64;
            const int which = e / (BN / 4), idx = e - which * (BN / 4);
            if (which < 2) {
              const float* src = which == 0 ? p.prep.g_lo + st * p.prep.g_lo_step_stride
                                            : p.prep.g_hi + st * p.prep.g_hi_step_stride;
              reinterpret_cast<float4*>(sg + which * BN)[idx] =
                  __ldg(reinterpret_cast<const float4*>(src + n0) + idx);
            }
          }
          gvec = sg + (row0 + lane < p.prep.split_row ? 0 : BN);
          named_barrier_sync_c<1>(128);  // gains staged (their last readers passed the previous tile's barriers)
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after_sync();
      if (trc && tcount == 0) trc[6] = clock64();
      const uint32_t t_row = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(lg * 32) << 16);
      uint32_t r[32];
      if (p.epilogue == EPI_GATED_GELU) {
        uint32_t g[32];
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          tmem_ld_32x32b_x32(t_row + c, r);
          tmem_ld_32x32b_x32(t_row + c + 32, g);
          tmem_ld_wait();
          if constexpr (DN == 1) {
            if (p.rs.ss_lo != nullptr) scale_bias_64(r, g, inv_r, bias ? bias + c : nullptr);
          }
          uint4 ch[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[i] = gelu_tanh(__uint_as_float(r[8 * q + i])) * __uint_as_float(g[8 * q + i]);
            ch[q] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                               pack_bf16(v[6], v[7]));
          }
          epilogue_bf16_rows<4>(p, tile_smem, lane, row0, (n0 + c) / 2, ch);
        }
      } else if (p.epilogue == EPI_GATED_GELU_SPLIT3) {
        // fp32-accurate mode: exact tanh, result kept to ~16 mantissa bits as [hi | lo | hi]
        uint32_t g[32];
        const int F = p.N / 2;
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          tmem_ld_32x32b_x32(t_row + c, r);
          tmem_ld_32x32b_x32(t_row + c + 32, g);
          tmem_ld_wait();
          uint4 chh[4], chl[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i] = gelu_tanh_exact(__uint_as_float(r[8 * q + i])) * __uint_as_float(g[8 * q + i]);
              lo[i] = v[i] - __bfloat162float(__float2bfloat16_rn(v[i]));
            }
            chh[q] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                pack_bf16(v[6], v[7]));
            chl[q] = make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]),
                                pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
          }
          const int oc = (n0 + c) / 2;
          epilogue_bf16_rows<4>(p, tile_smem, lane, row0, oc, chh);
          epilogue_bf16_rows<4>(p, tile_smem, lane, row0, F + oc, chl);
          epilogue_bf16_rows<4>(p, tile_smem, lane, row0, 2 * F + oc, chh);
        }
      } else if (p.epilogue == EPI_BF16) {
        uint32_t g[32];
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          tmem_ld_32x32b_x32(t_row + c, r);
          tmem_ld_32x32b_x32(t_row + c + 32, g);
          tmem_ld_wait();
          if constexpr (DN == 1) {
            if (p.rs.ss_lo != nullptr) scale_bias_64(r, g, inv_r, bias ? bias + c : nullptr);
          }
          uint4 ch[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ch[q] = make_uint4(
                pack_bf16(__uint_as_float(r[8 * q + 0]), __uint_as_float(r[8 * q + 1])),
                pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
            ch[4 + q] = make_uint4(
                pack_bf16(__uint_as_float(g[8 * q + 0]), __uint_as_float(g[8 * q + 1])),
                pack_bf16(__uint_as_float(g[8 * q + 2]), __uint_as_float(g[8 * q + 3])),
                pack_bf16(__uint_as_float(g[8 * q + 4]), __uint_as_float(g[8 * q + 5])),
                pack_bf16(__uint_as_float(g[8 * q + 6]), __uint_as_float(g[8 * q + 7])));
          }
          epilogue_bf16_rows<8>(p, tile_smem, lane, row0, n0 + c, ch);
        }
      } else {
        // fp32 outputs: registers (+ position rows) -> swizzled smem tile -> one TMA store per
        // 128 x 32 chunk; the residual form (out == resid: x += acc, the only way the engine uses
        // it) is a TMA REDUCE-ADD straight into the residual stream, so the residual is never
        // loaded: no load latency in the epilogue and half its memory traffic.  Four smem slots:
        // the store of chunk c - 4 must have read its slot before chunk c overwrites it.
        const int trow = lg * 32 + lane;   // row inside the CTA's 128-row tile
        const int grow = tile_row + trow;  // global row
        float ssum = 0.f;
        // (DN 2: a CTA whose rows are padding loads and stores nothing: skip its chunks, so that
        // the slot / barrier phase counter gc only counts chunks that were really loaded)
        const int nch_run = (DN == 2 && !prep) ? 0 : NCH;
#pragma unroll 1
        for (int c = 0; c < nch_run; ++c, ++gc) {
          const uint32_t slot = gc & 3;
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.epilogue == EPI_POS_F32 && grow < p.M) {
            const int seq = grow / p.pos_rows;
            int pr = grow - seq * p.pos_rows;
            if (p.pos_shift != nullptr) {
              pr -= p.pos_shift[seq];
              if (pr < 0) pr += p.pos_rows;
            }
            const float4* ps = reinterpret_cast<const float4*>(
                p.pos + static_cast<size_t>(pr) * p.N + n0 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 f = __ldg(ps + q);
              v[4 * q + 0] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
            }
          }
          if constexpr (DN == 2) {
            if (prep) {
              // v = acc + residual chunk (landed in this chunk's slot); operand chunk = bf16(v * g)
              mbar_wait(&xload_bar[slot], (gc >> 2) & 1u);
              const uint8_t* xr = sO + slot * 16384 + trow * 128;
              const float4* g4 = reinterpret_cast<const float4*>(gvec + c * 32);
              uint8_t* arow = sAring + slot * 8192 + trow * 64;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 xa = *reinterpret_cast<const float4*>(xr + (((2 * q) ^ (trow & 7)) * 16));
                const float4 xb = *reinterpret_cast<const float4*>(xr + (((2 * q + 1) ^ (trow & 7)) * 16));
                const float4 ga = g4[2 * q], gb = g4[2 * q + 1];
                float* w = v + 8 * q;
                w[0] += xa.x; w[1] += xa.y; w[2] += xa.z; w[3] += xa.w;
                w[4] += xb.x; w[5] += xb.y; w[6] += xb.z; w[7] += xb.w;
#pragma unroll
                for (int i = 0; i < 8; ++i) ssum = fmaf(w[i], w[i], ssum);
                // SWIZZLE_64B: 16-byte unit index ^= bits 7..8 of the tile offset = (row >> 1) & 3
                *reinterpret_cast<uint4*>(arow + ((q ^ ((trow >> 1) & 3)) * 16)) =
                    make_uint4(pack_bf16(w[0] * ga.x, w[1] * ga.y), pack_bf16(w[2] * ga.z, w[3] * ga.w),
                               pack_bf16(w[4] * gb.x, w[5] * gb.y), pack_bf16(w[6] * gb.z, w[7] * gb.w));
              }
            }
          }
          uint8_t* orow = sO + slot * 16384 + trow * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(orow + ((q ^ (trow & 7)) * 16)) =
                make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          fence_proxy_async_smem();
          // after this barrier everyone may write the NEXT chunk's slot: its previous user is the
          // store issued three chunks ago, so at most the two newest groups may still be reading
          // (DN 2: a slot is rewritten only by a residual load issued after wait_read<0> below)
          if (DN != 2 && epi_leader) tma_store_wait_read<2>();
          named_barrier_sync_c<1>(128);
          if (epi_leader) {
            if (tile_row < p.M) {  // M % 128 == 0: a CTA's rows are all valid or all padding
              if (DN == 2 && prep) {
                tma_store_2d(&tmap_out, sO + slot * 16384, n0 + c * 32, tile_row);
                tma_store_2d(&tmap_aux, sAring + slot * 8192, n0 + c * 32, tile_row);
              } else if (has_res) {
                tma_reduce_add_2d(&tmap_out, sO + slot * 16384, n0 + c * 32, tile_row);
              } else {
                tma_store_2d(&tmap_out, sO + slot * 16384, n0 + c * 32, tile_row);
                if (p.epilogue == EPI_POS_F32 && p.dup_rows > 0)
                  tma_store_2d(&tmap_out, sO + slot * 16384, n0 + c * 32, tile_row + p.dup_rows);
              }
            }
            tma_store_commit();
            if constexpr (DN == 2) {
              if (prep && c + 4 < NCH) {  // refill this slot with the residual chunk four ahead
                tma_store_wait_read<0>();
                mbar_arrive_expect_tx(&xload_bar[slot], 16384);
                tma_load_2d(sO + slot * 16384, &tmap_out, &xload_bar[slot], n0 + (c + 4) * 32, tile_row);
              }
            }
          }
        }
        if constexpr (DN == 2) {
          if (prep) p.prep.ss[static_cast<size_t>(n0 / BN) * p.prep.ss_stride + grow] = ssum;
        }
      }
      // accumulator drained: let the leader's MMA warp reuse it
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
  }
  if (trc) trc[7] = clock64();
  if (warp == 2 && lane == 0) tma_store_wait_all();
  tc_fence_before_sync();
  cluster_sync_all();
  if (trc) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trc[2] = t;
    trc[3] = clock64() - trc[3];   // total cycles in the CTA; the stamps below become offsets
  }
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

// SMs of the current device (148 on B200): the persistent grid is one CTA pair per two SMs.
static int gemm_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

template <int BN, int DN = 0>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                const CUtensorMap& taux, const GemmDev& d, cudaStream_t st) {
  using Cfg = PairCfg<BN, DN>;
  const int m_pairs = (d.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_tiles = m_pairs * (d.N / BN);
  const int sms = gemm_sm_count();
  const int clusters = num_tiles < sms / 2 ? num_tiles : sms / 2;
  ProfScope prof(KC_GEMM, 2.0 * d.M * d.N * d.K,
                 2.0 * (static_cast<double>(d.M) * d.K + static_cast<double>(d.N) * d.K) +
                     4.0 * d.M * d.N, st);
  MSD_CUDA_CHECK(launch_kernel(gemm_bf16_tcgen05_pair_kernel<BN, DN>, dim3(2 * clusters), dim3(192),
                               Cfg::SMEM_BYTES, st, ta, tb, tout, taux, d));
  ++g_launch_count;
  return 0;
}

template <int BN>
int configure_pair() {
  MSD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<BN, 0>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairCfg<BN, 0>::SMEM_BYTES));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<BN, 1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairCfg<BN, 1>::SMEM_BYTES));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<BN, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairCfg<BN, 2>::SMEM_BYTES));
  static_assert(PairCfg<BN, 2>::SMEM_BYTES <= 227 * 1024 && PairCfg<BN, 2>::STAGES >= 3, "smem budget");
  return 0;
}

template <int BN>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& d, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  dim3 grid(d.N / BN, (d.M + BLOCK_M - 1) / BLOCK_M);
  ProfScope prof(KC_GEMM, 2.0 * d.M * d.N * d.K,
                 2.0 * (static_cast<double>(d.M) * d.K + static_cast<double>(d.N) * d.K) +
                     4.0 * d.M * d.N, st);
  MSD_CUDA_CHECK(launch_kernel(gemm_bf16_tcgen05_kernel<BN>, grid, dim3(192), Cfg::SMEM_BYTES, st, ta,
                               tb, d));
  ++g_launch_count;
  return 0;
}

template <int BN>
int configure_bn() {
  MSD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<BN>::SMEM_BYTES));
  return 0;
}

}  // namespace

int gemm_configure() {
  if (int rc = configure_bn<64>()) return rc;
  if (int rc = configure_bn<128>()) return rc;
  if (int rc = configure_bn<256>()) return rc;
  if (int rc = configure_pair<64>()) return rc;
  if (int rc = configure_pair<96>()) return rc;
  if (int rc = configure_pair<128>()) return rc;
  if (int rc = configure_pair<192>()) return rc;
  return configure_pair<256>();
}

// Tile width of the CTA-pair kernel: minimise (waves x per-tile MMA time) over the widths that
// divide N.  74 clusters run concurrently; a tile costs ~BN cycles per k-step.
int gemm_pick_pair_bn(int M, int N) {
  // Measured on B200 (tools/gemm_bench.py): the widest tile wins (least L2->SM traffic per FLOP)
  // unless it leaves clusters idle that a narrower tiling would use: among the widths whose tile
  // count fits the 74 concurrently running clusters, take the one with the most tiles.
  // (A 96-wide instance exists for the fp32 epilogues, block_n = 96; as a candidate here it
  // gained nothing, and two overlapped rounds of 96 instead of one of 192 lost 5 %.)
  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int widths[4] = {256, 192, 128, 64};
  const int conc = gemm_sm_count() / 2;  // concurrently running CTA pairs
  int best = 0, best_tiles = 0;
  for (int i = 0; i < 4; ++i) {
    const int bn = widths[i];
    if (N % bn != 0) continue;
    const int tiles = m_pairs * (N / bn);
    if (best == 0) { best = bn; best_tiles = tiles; continue; }   // widest dividing width
    if (best_tiles < conc && tiles <= conc && tiles > best_tiles) { best = bn; best_tiles = tiles; }
  }
  return best;
}

int gemm_pick_block_n(int M, int N) {
  const int mt = (M + BLOCK_M - 1) / BLOCK_M;
  if (N % 256 == 0 && mt * (N / 256) >= 296) return 256;
  if (N % 128 == 0 && mt * (N / 128) >= 148) return 128;
  if (N % 64 == 0) return 64;
  if (N % 128 == 0) return 128;
  return 0;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  static int configured = gemm_configure();
  if (configured != 0) return configured;
  MSD_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  MSD_REQUIRE(a.K % BLOCK_K == 0, "gemm: K=%d must be a multiple of %d", a.K, BLOCK_K);
  MSD_REQUIRE(a.M % BLOCK_M == 0, "gemm: M=%d must be a multiple of %d", a.M, BLOCK_M);
  static const int forced_variant = [] {
    const char* e = getenv("MSD_GEMM_VARIANT");  // debugging aid: 1 forces the single-CTA kernel
    return e ? atoi(e) : 0;
  }();
  const bool pair = (forced_variant ? forced_variant : a.variant) != 1;
  int bn = a.block_n ? a.block_n
                     : (pair ? gemm_pick_pair_bn(a.M, a.N) : gemm_pick_block_n(a.M, a.N));
  MSD_REQUIRE(pair || a.epilogue != EPI_GATED_GELU_SPLIT3,
              "gemm: the split-precision gated epilogue exists in the CTA-pair kernel only");
  MSD_REQUIRE(bn == 64 || bn == 128 || bn == 256 || (pair && bn == 192) ||
                  (pair && bn == 96 && !epi_is_bf16_out(a.epilogue)),
              "gemm: N=%d has no valid tile width (block_n %d)", a.N, bn);
  MSD_REQUIRE(a.N % bn == 0, "gemm: N=%d not a multiple of block_n=%d", a.N, bn);
  MSD_REQUIRE(a.ldo % 8 == 0, "gemm: ldo=%d must be a multiple of 8", a.ldo);

  CUtensorMap ta, tb;
  if (a.tmap_a) {
    ta = *a.tmap_a;
  } else if (int rc = make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M)) {
    return rc;
  }
  if (a.tmap_b) {
    tb = *a.tmap_b;
  } else if (int rc = make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, pair ? bn / 2 : bn)) {
    return rc;
  }
  GemmDev d;
  d.M = a.M; d.N = a.N; d.K = a.K;
  d.epilogue = a.epilogue;
  d.out = a.out; d.ldo = a.ldo;
  d.resid = a.resid; d.pos = a.pos; d.pos_rows = a.pos_rows > 0 ? a.pos_rows : 1;
  d.pos_shift = a.pos_shift; d.dup_rows = a.dup_rows;
  d.trace = a.trace;
  d.prep = a.prep; d.rs = a.rs; d.step = a.step;
  const bool dn = a.epilogue == EPI_RESID_PREP || a.rs.ss_lo != nullptr;
  if (dn) {
    MSD_REQUIRE(pair, "gemm: deferred normalisation exists in the CTA-pair kernel only");
    if (a.epilogue == EPI_RESID_PREP) {
      MSD_REQUIRE(a.resid != nullptr && a.resid == a.out, "gemm: EPI_RESID_PREP works in place (out == resid)");
      MSD_REQUIRE(a.prep.a && a.prep.ss && a.prep.g_lo && a.prep.g_hi && a.prep.lda % 8 == 0 &&
                      a.prep.ss_stride >= a.M,
                  "gemm: EPI_RESID_PREP needs prep.a / ss / g_lo / g_hi (lda %% 8 == 0, ss_stride >= M)");
      MSD_REQUIRE(a.step != nullptr || (a.prep.g_lo_step_stride == 0 && a.prep.g_hi_step_stride == 0),
                  "gemm: step-dependent column scales need the device step index");
    }
    if (a.rs.ss_lo != nullptr) {
      MSD_REQUIRE(a.epilogue == EPI_BF16 || a.epilogue == EPI_GATED_GELU,
                  "gemm: the row scale applies to the bf16 and gated epilogues only");
      MSD_REQUIRE(a.rs.ss_hi != nullptr && a.rs.parts_lo > 0 && a.rs.parts_hi > 0 && a.rs.inv_d > 0.f,
                  "gemm: incomplete row-scale description");
      MSD_REQUIRE(a.step != nullptr || a.rs.col_bias == nullptr || a.rs.bias_step_stride == 0,
                  "gemm: a step-dependent bias row needs the device step index");
    }
  }
  if (pair) {
    CUtensorMap tout = ta;  // placeholder unless the epilogue is an fp32 one
    if (!epi_is_bf16_out(a.epilogue)) {
      const int rows = a.M + (a.epilogue == EPI_POS_F32 ? a.dup_rows : 0);
      if (int rc = make_tmap_f32_2d(&tout, a.out, rows, a.N, a.ldo, BLOCK_M)) return rc;
      // the residual epilogue adds into `out` with a TMA reduction: out must already hold resid
      if (a.epilogue == EPI_RESID_F32 && a.resid != a.out)
        MSD_CUDA_CHECK(cudaMemcpy2DAsync(a.out, static_cast<size_t>(a.ldo) * 4, a.resid,
                                         static_cast<size_t>(a.ldo) * 4, static_cast<size_t>(a.N) * 4,
                                         a.M, cudaMemcpyDeviceToDevice, stream));
    }
    if (a.epilogue == EPI_RESID_PREP) {
      CUtensorMap taux;
      if (int rc = make_tmap_bf16_2d_half(&taux, a.prep.a, a.M, a.N, a.prep.lda, BLOCK_M)) return rc;
      switch (bn) {
        case 64: return launch_pair<64, 2>(ta, tb, tout, taux, d, stream);
        case 96: return launch_pair<96, 2>(ta, tb, tout, taux, d, stream);
        case 128: return launch_pair<128, 2>(ta, tb, tout, taux, d, stream);
        case 192: return launch_pair<192, 2>(ta, tb, tout, taux, d, stream);
        default: return launch_pair<256, 2>(ta, tb, tout, taux, d, stream);
      }
    }
    if (dn) {
      switch (bn) {
        case 64: return launch_pair<64, 1>(ta, tb, tout, ta, d, stream);
        case 96: return launch_pair<96, 1>(ta, tb, tout, ta, d, stream);
        case 128: return launch_pair<128, 1>(ta, tb, tout, ta, d, stream);
        case 192: return launch_pair<192, 1>(ta, tb, tout, ta, d, stream);
        default: return launch_pair<256, 1>(ta, tb, tout, ta, d, stream);
      }
    }
    switch (bn) {
      case 64: return launch_pair<64>(ta, tb, tout, ta, d, stream);
      case 96: return launch_pair<96>(ta, tb, tout, ta, d, stream);
      case 128: return launch_pair<128>(ta, tb, tout, ta, d, stream);
      case 192: return launch_pair<192>(ta, tb, tout, ta, d, stream);
      default: return launch_pair<256>(ta, tb, tout, ta, d, stream);
    }
  }
  switch (bn) {
    case 64: return launch_bn<64>(ta, tb, d, stream);
    case 128: return launch_bn<128>(ta, tb, d, stream);
    default: return launch_bn<256>(ta, tb, d, stream);
  }
}

}  // namespace msd
