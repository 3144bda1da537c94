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
#include "common.cuh"
#include "kernels.h"

namespace msd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

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

  if (warp == 0) {
    if (lane == 0) {
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

template <int BN>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& d, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  dim3 grid(d.N / BN, (d.M + BLOCK_M - 1) / BLOCK_M);
  ProfScope prof(KC_GEMM, 2.0 * d.M * d.N * d.K,
                 2.0 * (static_cast<double>(d.M) * d.K + static_cast<double>(d.N) * d.K) +
                     4.0 * d.M * d.N, st);
  gemm_bf16_tcgen05_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, st>>>(ta, tb, d);
  MSD_CUDA_CHECK(cudaGetLastError());
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
  return configure_bn<256>();
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
  int bn = a.block_n ? a.block_n : gemm_pick_block_n(a.M, a.N);
  MSD_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: N=%d has no valid tile width", a.N);
  MSD_REQUIRE(a.N % bn == 0, "gemm: N=%d not a multiple of block_n=%d", a.N, bn);
  if (a.epilogue == EPI_GATED_GELU)
    MSD_REQUIRE(bn >= 64, "gemm: gated epilogue needs block_n >= 64");
  MSD_REQUIRE(a.ldo % 8 == 0, "gemm: ldo=%d must be a multiple of 8", a.ldo);

  CUtensorMap ta, tb;
  if (a.tmap_a) {
    ta = *a.tmap_a;
  } else if (int rc = make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M)) {
    return rc;
  }
  if (a.tmap_b) {
    tb = *a.tmap_b;
  } else if (int rc = make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, bn)) {
    return rc;
  }
  GemmDev d;
  d.M = a.M; d.N = a.N; d.K = a.K;
  d.epilogue = a.epilogue;
  d.out = a.out; d.ldo = a.ldo;
  d.resid = a.resid; d.pos = a.pos; d.pos_rows = a.pos_rows > 0 ? a.pos_rows : 1;
  d.pos_shift = a.pos_shift; d.dup_rows = a.dup_rows;
  switch (bn) {
    case 64: return launch_bn<64>(ta, tb, d, stream);
    case 128: return launch_bn<128>(ta, tb, d, stream);
    default: return launch_bn<256>(ta, tb, d, stream);
  }
}

}  // namespace msd
