// Fused attention O = softmax(Q K^T + keymask) V on tcgen05 tensor cores, head_dim 64,
// NO 1/sqrt(d) scaling (msd/layers.py:158-181, note at 254-258), key-padding mask as in
// msd/layers.py:341-348 (0 / -1e10 bias == masked keys get exactly zero weight in fp32), and
// rows with no attendable key produce 0 (msd/layers.py:882-902 zero_activations_if_masked).
//
// One CTA per (256-query group, head, batch row): TWO 128-query tiles ping-pong so that the
// tensor core works on one tile while the other tile's softmax runs (the exp throughput of the
// SFU, not the tensor core, bounds head_dim-64 attention on this part).
//   warp 8 (one lane)   TMA producer: both Q tiles once, then K/V 128-key tiles into a ring
//   warp 9 (one lane)   MMA issuer:   S_t = Q_t K^T  (M128 x N128 x K64, both operands K-major)
//                                     PV_t = P_t V   (M128 x N64 x K128, P K-major from smem,
//                                                     V MN-major straight from its [key,64] tile)
//   warps 0..3 / 4..7   softmax group of tile 0 / tile 1 (setmaxnreg gives them the registers): thread = query row; the 128 logits of a
//                       key block are read ONCE from TMEM into registers, max / exp2 / sum in
//                       fp32, P written to smem as bf16 in the 128B-swizzled K-major layout.
//                       O accumulates in TMEM across key blocks (MMA accumulate); the running
//                       max is only raised when it grows by > 2^8 (lazy rescale: then O is
//                       rescaled in TMEM with tcgen05.ld/st), exact because the final division
//                       uses the same reference max for numerator and denominator
// Key blocks whose 128 mask bits are all zero are skipped by every role.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace msd {

namespace {

constexpr int BQ = 128;   // queries per tile
constexpr int HD = 64;    // head dim
constexpr int Q_BYTES = BQ * HD * 2;         // 16 KB per tile
constexpr int ATTN_THREADS = 384;  // warps 0-3 / 4-7 softmax tile 0 / 1, 8 TMA, 9 MMA, 10-11 idle
constexpr float LOG2E = 1.4426950408889634f;

// Two instances, by keys per block:
//   BKV 128: one CTA per SM (193 KB of shared memory, 384 TMEM columns); two query tiles in flight
//            per SM.  The per-tile chain of a key block (S wait, TMEM load, max, exp, P store: ~3150
//            cycles, clock64 trace in profiles/) runs strictly in sequence, so with two tiles the SFU
//            is busy ~65 % and the issue slots ~26 %: latency-bound.
//   BKV 64:  half-size K/V stages and P tiles, 256 TMEM columns, 104 registers per softmax thread:
//            TWO CTAs per SM, i.e. four query tiles in flight per SM, whose phases interleave on
//            the sub-partitions; grids of up to 2 x SMs CTAs run as a single wave (B = 8
//            self-attention: 192 CTAs).
template <int BKV>
struct ACfg {
  static constexpr int KV_TILE_BYTES = BKV * HD * 2;   // 16 / 8 KB
  static constexpr int KV_STAGES = 3;
  static constexpr int P_BYTES = BQ * BKV * 2;         // 32 / 16 KB per tile ([128 x 64] sub-tiles)
  static constexpr int NCH = BKV / 32;                 // 32-column chunks of a logits row
  static constexpr int WPB = BKV / 32;                 // mask words per key block
  static constexpr uint32_t TMEM_COLS = BKV == 128 ? 512 : 256;  // S0 S1 (BKV each) PV0 PV1 (64 each)
  static constexpr int SMEM = 2 * Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES + 2 * P_BYTES + 256;
  static constexpr int MIN_CTAS = BKV == 128 ? 1 : 2;
};

struct AttnDev {
  bf16* O;
  int ldo;
  int heads, Lq, Lk;
  const uint32_t* mask_bits;
  int mask_stride_words;
  long long* trace;  // optional [2 tiles][64 blocks][8] clock64 stamps of CTA (0,0,0), else null
  // split-KV: blockIdx.z = batch * splits + split; each split covers nkb / splits key blocks.
  //   merge == 0: every split writes an unnormalised partial (O fp32, reference max m, sum l) that
  //               the combine kernel merges;
  //   merge == 1: splits 1.. publish their partial per softmax warp and raise a flag, split 0 (the
  //               owner) merges them into its own result before the final store -- no second kernel.
  //               All CTAs of the grid must be co-resident (the launcher checks).
  int splits, merge;
  float* part_o;   // merge 0: [rows * heads * splits][64]; merge 1: per-warp slots, see below
  float* part_ml;  // merge 0: [rows * heads * splits][2]
  // tail mode (BKV 128, splits == 1, tail > 0): blockIdx.z = role * nbatch + batch.  Role 1
  // ("short") CTAs cover the last `tail` key blocks and publish an unnormalised partial per softmax
  // warp; role 0 ("long") CTAs cover the rest, are scheduled first (lower block index), and merge
  // the partial of their short partner before the final store.
  int tail, nbatch;
  int kv_static;    // K, V and mask_bits are not produced by the preceding kernels (see TMA warp)
  int kv_batch_rows, kv_row0;  // K/V row of (batch b, key block j) = b*kv_batch_rows + kv_row0 + j*BKV
  uint32_t* flags;  // one per (softmax warp, partner), 0 outside a launch
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Bit j = key block j has at least one attendable key.  One coalesced pass by a whole warp
// (a serial loop of dependent global loads cost ~3000 cycles per CTA for 18 blocks).
template <int WPB>
__device__ __forceinline__ uint64_t active_blocks(const uint32_t* mrow, int nkb_all, int lane) {
  if (mrow == nullptr) return ~0ull;
  uint32_t word[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = h * 32 + lane;
    bool a = false;
    if (j < nkb_all) {
      if (WPB == 4) {
        const uint4 w = *reinterpret_cast<const uint4*>(mrow + j * 4);
        a = (w.x | w.y | w.z | w.w) != 0u;
      } else {
        const uint2 w = *reinterpret_cast<const uint2*>(mrow + j * 2);
        a = (w.x | w.y) != 0u;
      }
    }
    word[h] = __ballot_sync(0xffffffffu, a);
  }
  return static_cast<uint64_t>(word[0]) | (static_cast<uint64_t>(word[1]) << 32);
}
__device__ __forceinline__ bool block_active(uint64_t act, int blk) { return (act >> blk) & 1ull; }
__device__ __forceinline__ int next_active(uint64_t act, int from, int nkb) {
  for (int j = from; j < nkb; ++j)
    if (block_active(act, j)) return j;
  return -1;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32x2 arithmetic (sm_100): halves the FMA-pipe instruction count of the softmax.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Per-warp partial slot of the in-kernel merges: [16 float4][32 lanes] of O (512-byte coalesced
// stores / loads) followed by [32 lanes] float2 (m, l).
constexpr int SLOT_FLOATS = 32 * HD + 64;

template <int BKV>
__global__ void __launch_bounds__(ATTN_THREADS, ACfg<BKV>::MIN_CTAS)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q,
                         const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const AttnDev p) {
  using Cfg = ACfg<BKV>;
  constexpr int KV_STAGES = Cfg::KV_STAGES;
  constexpr int KV_TILE_BYTES = Cfg::KV_TILE_BYTES;
  constexpr int P_BYTES = Cfg::P_BYTES;
  constexpr int NCH = Cfg::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  const bool ktr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
                   threadIdx.x == 0;
  long long* ktrp = p.trace + 2 * 64 * 8;  // CTA-level stamps: entry, setup done, loop end, stores done
  if (ktr) ktrp[0] = clock64();
  // debugging: per-CTA (smid, start, end) in nanoseconds of the global timer, after the per-block
  // stamps and the 8 CTA-level stamps of the trace buffer
  long long* ctr = nullptr;
  if (p.trace != nullptr && threadIdx.x == 0) {
    const int lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (lin < 1024) {
      ctr = p.trace + 2 * 64 * 8 + 8 + lin * 3;
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      ctr[0] = smid;
      ctr[1] = t;
    }
  }
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // the swizzled tiles assume a 1024-byte aligned base
  uint8_t* sQ = smem;                                   // [2][16 KB]
  uint8_t* sK = sQ + 2 * Q_BYTES;                       // [KV_STAGES][KV tile]
  uint8_t* sV = sK + KV_STAGES * KV_TILE_BYTES;         // [KV_STAGES][KV tile]
  uint8_t* sP = sV + KV_STAGES * KV_TILE_BYTES;         // [2][P tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES; // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;  // [2]
  uint64_t* p_full = s_full + 2;            // [2] (128 arrivals each)
  uint64_t* pv_full = p_full + 2;           // [2]
  uint64_t* s_free = pv_full + 2;           // [2] (128 arrivals: logits copied to registers)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int qgrp = blockIdx.x, head = blockIdx.y;
  const int role = p.tail > 0 ? static_cast<int>(blockIdx.z) / p.nbatch : 0;
  const int b = p.tail > 0 ? static_cast<int>(blockIdx.z) - role * p.nbatch
                           : static_cast<int>(blockIdx.z) / p.splits;
  const int split = p.tail > 0 ? 0 : static_cast<int>(blockIdx.z) - b * p.splits;
  const int q0 = qgrp * 2 * BQ;                       // first query row of this CTA
  const int nq = (p.Lq - q0 >= 2 * BQ) ? 2 : 1;       // query tiles handled here
  const int nkb_all = p.Lk / BKV;
  // this CTA's key blocks [kb0, nkb)
  const int kb0 = p.tail > 0 ? (role ? nkb_all - p.tail : 0) : split * nkb_all / p.splits;
  const int nkb = p.tail > 0 ? (role ? nkb_all : nkb_all - p.tail) : (split + 1) * nkb_all / p.splits;
  const uint32_t* mrow =
      p.mask_bits ? p.mask_bits + static_cast<size_t>(b) * p.mask_stride_words : nullptr;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_full[t], 1);
      mbar_init(&s_free[t], 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (ktr) ktrp[1] = clock64();

  griddep_launch_dependents();
  if (warp >= 8) {
  if (BKV == 128) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");  // 384 x 80 = 256 x 104 + 128 x 32
  if (warp == 8) {
    // kv_static: K, V and the key mask were written long before the preceding kernel (the
    // cross-attention cache of a diffusion step), so the first ring-full of K/V tiles is
    // requested ahead of the dependency wait; only Q comes from the preceding kernel.
    if (!p.kv_static) griddep_wait();
    const uint64_t act = active_blocks<Cfg::WPB>(mrow, nkb_all, lane);
    if (lane == 0) {
      int it = 0, j = kb0;
      auto load_kv = [&](int jb) {
        const int s = it % KV_STAGES;
        const uint32_t ph = (it / KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], 2 * KV_TILE_BYTES);
        const int krow = b * p.kv_batch_rows + p.kv_row0 + jb * BKV;
        tma_load_2d(sK + s * KV_TILE_BYTES, &tmap_k, &kv_full[s], head * HD, krow);
        tma_load_2d(sV + s * KV_TILE_BYTES, &tmap_v, &kv_full[s], head * HD, krow);
        ++it;
      };
      if (p.kv_static) {
        for (; j < nkb && it < KV_STAGES; ++j)
          if (block_active(act, j)) load_kv(j);
        griddep_wait();
      }
      mbar_arrive_expect_tx(q_full, nq * Q_BYTES);
      for (int t = 0; t < nq; ++t)
        tma_load_2d(sQ + t * Q_BYTES, &tmap_q, q_full, head * HD, b * p.Lq + q0 + t * BQ);
      for (; j < nkb; ++j)
        if (block_active(act, j)) load_kv(j);
    }
  } else if (warp == 9) {
    if (!p.kv_static) griddep_wait();  // mask words may come from the previous kernel
    const uint64_t act = active_blocks<Cfg::WPB>(mrow, nkb_all, lane);
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);  // P K-major, V MN-major
      // descriptor low words (address >> 4 | LBO); stepping = adding (bytes >> 4)
      const uint32_t q_lo = desc_lo_sw128(smem_u32(sQ));
      const uint32_t k_lo = desc_lo_sw128(smem_u32(sK));
      const uint32_t v_lo = desc_lo_sw128(smem_u32(sV));
      const uint32_t p_lo = desc_lo_sw128(smem_u32(sP));
      // S_t = Q_t K^T : 4 k-steps of 16 along head_dim (32 bytes each inside the swizzle atom)
      auto issue_s = [&](int t, int stage) {
        const uint32_t qa = q_lo + t * (Q_BYTES >> 4);
        const uint32_t kb = k_lo + stage * (KV_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_lo(tmem_base + t * BKV, qa + k * 2, kb + k * 2, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      // PV_t = P_t V : BKV/16 k-steps of 16 keys.  P: sub-tile (k/4) of 16 KB, +32 B per step
      // inside.  V (MN-major): 16 keys = 16 rows of 128 B -> +2048 B per step; 8-key groups 1024 B
      // apart.
      auto issue_pv = [&](int t, int stage, uint32_t acc_first) {
        const uint32_t pa = p_lo + t * (P_BYTES >> 4);
        const uint32_t vb = v_lo + stage * (KV_TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_bf16_lo(tmem_base + 2 * BKV + t * 64, pa + (k >> 2) * ((BQ * 128) >> 4) + (k & 3) * 2,
                       vb + k * (2048 >> 4), idesc_pv, k != 0 ? 1u : acc_first);
        umma_commit(&pv_full[t]);
      };
      int jn = next_active(act, kb0, nkb);
      int it = 0;
      if (jn >= 0) {
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after_sync();
        for (int t = 0; t < nq; ++t) issue_s(t, 0);
      }
      while (jn >= 0) {
        jn = next_active(act, jn + 1, nkb);
        const int stage = it % KV_STAGES;
        const int nstage = (it + 1) % KV_STAGES;
        const uint32_t par = it & 1;
        // Next block's logits first: they only need the S buffer back (s_free), not P, so the
        // softmax groups never wait for the tensor core in steady state.
        if (jn >= 0) {
          mbar_wait(&kv_full[nstage], ((it + 1) / KV_STAGES) & 1);
          for (int t = 0; t < nq; ++t) {
            mbar_wait(&s_free[t], par);
            tc_fence_after_sync();
            issue_s(t, nstage);
          }
        }
        for (int t = 0; t < nq; ++t) {
          mbar_wait(&p_full[t], par);  // P_t published; previous PV_t consumed / rescaled
          tc_fence_after_sync();
          issue_pv(t, stage, it == 0 ? 0u : 1u);
        }
        umma_commit(&kv_empty[stage]);  // K/V of this block are dead once the MMAs above retire
        ++it;
      }
    }
  }
  } else {
    // ------------------------- softmax / output warp groups -------------------------
    if (BKV == 128) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int tile = warp >> 2;  // 0: warps 0..3, 1: warps 4..7
    if (!p.kv_static) griddep_wait();  // mask words may come from the previous kernel
    const uint64_t act = active_blocks<Cfg::WPB>(mrow, nkb_all, lane);
    griddep_wait();  // O is written by these warps
    if (tile < nq) {
      const int lg = warp & 3;
      const int r = lg * 32 + lane;  // query row inside the tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
      const uint32_t tmem_s = tmem_base + tile * BKV + lane_off;
      const uint32_t tmem_pv = tmem_base + 2 * BKV + tile * 64 + lane_off;
      uint8_t* sPt = sP + tile * P_BYTES;
      float m = -INFINITY, l = 0.f;  // reference max (natural units) and running sum
      uint32_t s[NCH][32];
      // Ping-pong of the SFU-bound exp phase between the two tiles' warpgroups (named barriers
      // 2 + tile: "tile may run its exps"): without it both groups run in lockstep and collide on
      // the SFU while it idles during their max / store / wait phases (ncu: XU 41 % busy).
      int nact = 0;
      for (int j = kb0; j < nkb; ++j) nact += block_active(act, j) ? 1 : 0;
      const bool pingpong = (nq == 2);
      // barrier ids as immediates (2: tile 0 may run its exps, 3: tile 1 may)
      auto turn_sync = [&]() {
        if (tile == 0) named_barrier_sync_c<2>(256);
        else named_barrier_sync_c<3>(256);
      };
      auto turn_give = [&]() {   // let the OTHER tile run
        if (tile == 0) named_barrier_arrive_c<3>(256);
        else named_barrier_arrive_c<2>(256);
      };
      if (pingpong && tile == 1 && nact > 0) named_barrier_arrive_c<2>(256);  // tile 0 goes first
      constexpr float RESCALE_THRESHOLD = 5.545177444f;  // 8 * ln 2: P stays below 2^8
      int it = 0;
      for (int j = kb0; j < nkb; ++j) {
        if (!block_active(act, j)) continue;
        uint32_t mw[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) mw[c] = 0xffffffffu;
        if (mrow != nullptr) {
          if (NCH == 4) {
            const uint4 w = *reinterpret_cast<const uint4*>(mrow + j * 4);
            mw[0] = w.x; mw[1] = w.y; mw[NCH - 2] = w.z; mw[NCH - 1] = w.w;
          } else {
            const uint2 w = *reinterpret_cast<const uint2*>(mrow + j * 2);
            mw[0] = w.x; mw[1] = w.y;
          }
        }
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 &&
                        blockIdx.z == 0 && lg == 0 && lane == 0 && it < 64;
        long long* trp = p.trace + (tile * 64 + it) * 8;
        if (tr) trp[0] = clock64();
        mbar_wait(&s_full[tile], it & 1);
        tc_fence_after_sync();
        if (tr) trp[1] = clock64();
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tmem_s + c * 32, s[c]);
        tmem_ld_wait();
        if (tr) trp[2] = clock64();
        tc_fence_before_sync();
        mbar_arrive(&s_free[tile]);  // the S buffer may be overwritten by the next block's QK^T
        // row max over attendable keys (finite: an active block has >= 1 attendable key);
        // 8 independent partial maxima: a single 128-deep fmax chain cost ~720 cycles per block
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
        bool all_on = true;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint32_t bits = mw[c];
          all_on = all_on && (bits == 0xffffffffu);
          if (bits == 0xffffffffu) {
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              mx[(i >> 1) & 7] = fmax3(mx[(i >> 1) & 7], __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((bits >> i) & 1u) mx[i & 7] = fmaxf(mx[i & 7], __uint_as_float(s[c][i]));
          }
        }
        const float bmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                                 fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        bool waited_pv = false;
        if (it == 0) {
          m = bmax;
        } else if (__any_sync(0xffffffffu, bmax > m + RESCALE_THRESHOLD)) {
          // rare: raise the reference max of the rows that need it and rescale O in TMEM
          const float m_new = (bmax > m + RESCALE_THRESHOLD) ? bmax : m;
          const float alpha = ex2_approx((m - m_new) * LOG2E);
          mbar_wait(&pv_full[tile], (it - 1) & 1);  // previous PV has landed in O
          tc_fence_after_sync();
          waited_pv = true;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t ob[32];
            tmem_ld_32x32b_x32(tmem_pv + c * 32, ob);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ob[i] = __float_as_uint(__uint_as_float(ob[i]) * alpha);
            tmem_st_32x32b_x32(tmem_pv + c * 32, ob);
          }
          tmem_st_wait();
          l *= alpha;
          m = m_new;
        }
        const float mb = m * LOG2E;
        if (tr) trp[3] = clock64();
        if (pingpong) turn_sync();  // my turn on the SFU
        if (tr) trp[4] = clock64();
        // p = exp(s - m) as bf16 pairs (packed in place into s[c][0..15]), row sum in fp32.
        const uint64_t l2e2 = pack2(LOG2E, LOG2E), nmb2 = pack2(-mb, -mb);
        uint64_t acc2 = pack2(0.f, 0.f), acc2b = pack2(0.f, 0.f);
        // the last MUFU of this block has been issued: hand the SFU to the other tile, its exps
        // overlap the remaining adds / packs and the P store below (the very last hand-over has no
        // taker and is skipped)
        auto handover = [&]() {
          if (pingpong && !(tile == 1 && it + 1 == nact)) turn_give();
        };
        // Two copies of the phase behind a warp-uniform branch: with every key of the block
        // attendable (the common case) no select instructions are issued.
        auto exp_phase = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
          auto exp_chunk = [&](int c, float (&dst)[32]) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const uint64_t t2 =
                  ffma2(pack2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), l2e2, nmb2);
              float t0, t1;
              unpack2(t2, t0, t1);
              dst[i] = ex2_approx(t0);
              dst[i + 1] = ex2_approx(t1);
            }
          };
          auto finish_chunk = [&](int c, float (&src)[32]) {
            if (MASKED) {
              const uint32_t bits = mw[c];
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (!((bits >> i) & 1u)) src[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              acc2 = fadd2(acc2, pack2(src[i], src[i + 1]));
              acc2b = fadd2(acc2b, pack2(src[i + 2], src[i + 3]));
            }
#pragma unroll
            for (int i = 0; i < 32; i += 2) s[c][i >> 1] = pack_bf16(src[i], src[i + 1]);
          };
          if constexpr (NCH == 4) {
            // Software-pipelined by one 32-column chunk: the SFU exps of chunk c are independent of
            // the adds / packs of chunk c-1 issued next to them, so the MUFUs go out back to back.
            float e[2][32];
            exp_chunk(0, e[0]);
            exp_chunk(1, e[1]);
            finish_chunk(0, e[0]);
            exp_chunk(2, e[0]);
            finish_chunk(1, e[1]);
            exp_chunk(3, e[1]);
            handover();
            finish_chunk(2, e[0]);
            finish_chunk(3, e[1]);
          } else {
            // 104 registers per thread: one 32-wide staging buffer; the second chunk's exps are
            // issued while the first chunk is summed and packed
            float e0[32], e1[32];
            exp_chunk(0, e0);
            exp_chunk(1, e1);
            handover();
            finish_chunk(0, e0);
            finish_chunk(1, e1);
          }
        };
        if (all_on) exp_phase(std::false_type{});
        else exp_phase(std::true_type{});
        if (tr) trp[5] = clock64();
        float lsum, lsum_hi, lsum2, lsum2_hi;
        unpack2(acc2, lsum, lsum_hi);
        unpack2(acc2b, lsum2, lsum2_hi);
        lsum = (lsum + lsum_hi) + (lsum2 + lsum2_hi);
        l += lsum;
        // the P buffer is free once the previous PV MMA (which read it) has completed
        if (it > 0 && !waited_pv) mbar_wait(&pv_full[tile], (it - 1) & 1);
        if (tr) trp[6] = clock64();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          // columns [c*32, c*32+32) -> sub-tile c/2, 16-byte chunks (c&1)*4 .. +3, XOR row&7
          uint8_t* prow = sPt + (c >> 1) * (BQ * 128) + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = make_uint4(s[c][4 * q], s[c][4 * q + 1], s[c][4 * q + 2], s[c][4 * q + 3]);
            const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
          }
        }
        fence_proxy_async_smem();  // st.shared -> visible to the tensor core (async proxy)
        tc_fence_before_sync();    // order our tcgen05.ld/st before the next MMAs
        mbar_arrive(&p_full[tile]);
        if (tr) trp[7] = clock64();
        ++it;
      }
      if (ktr) ktrp[2] = clock64();
      if (it > 0) {
        mbar_wait(&pv_full[tile], (it - 1) & 1);
        tc_fence_after_sync();
      }
      // ---- output.  The O row (64 fp32) is handled in two halves of 32 columns to keep the
      // register footprint of this section at ~80.
      auto load_half = [&](int c, float (&o)[32]) {
        if (it > 0) {
          uint32_t ob[32];
          tmem_ld_32x32b_x32(tmem_pv + c * 32, ob);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(ob[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0.f;
        }
      };
      const int qtiles = (p.Lq + BQ - 1) / BQ;
      // per-warp slot index of the in-kernel merges (tail mode / merge mode)
      const size_t wbase =
          ((static_cast<size_t>(b) * p.heads + head) * qtiles + qgrp * 2 + tile) * 4 + lg;
      const bool publisher = (p.tail > 0 && role == 1) || (p.merge && p.splits > 1 && split > 0);
      const int npartners = p.tail > 0 ? (role == 0 ? 1 : 0)
                                       : ((p.merge && p.splits > 1 && split == 0) ? p.splits - 1 : 0);
      if (publisher) {
        // slot of partner index (split - 1) (tail mode: 0)
        const size_t wslot = wbase * (p.tail > 0 ? 1 : (p.splits - 1)) + (p.tail > 0 ? 0 : split - 1);
        float* slot = p.part_o + wslot * SLOT_FLOATS;
        float4* po = reinterpret_cast<float4*>(slot);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float o[32];
          load_half(c, o);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            po[(c * 8 + q) * 32 + lane] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        reinterpret_cast<float2*>(slot + 32 * HD)[lane] = make_float2(m, l);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(p.flags + wslot, 1u);
      } else if (p.splits > 1 && !p.merge) {
        const size_t prow =
            (static_cast<size_t>(b * p.Lq + q0 + tile * BQ + r) * p.heads + head) * p.splits + split;
        float4* po = reinterpret_cast<float4*>(p.part_o + prow * HD);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float o[32];
          load_half(c, o);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            po[c * 8 + q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        *reinterpret_cast<float2*>(p.part_ml + prow * 2) = make_float2(m, l);
      } else {
        // owner: merge the partners' partials (O = sum_i w_i O_i / sum_i w_i l_i, w_i = exp(m_i -
        // max m)).  Partners are handled in groups of kGroup whose loads of one L2 round trip are
        // all in flight together; any number of partners (the launcher only requires that the whole
        // grid is resident, so that waiting for them cannot deadlock).
        constexpr int kGroup = 3;
        const size_t wslot0 = wbase * (p.tail > 0 ? 1 : (p.splits > 1 ? p.splits - 1 : 1));
        auto slot_ml = [&](int pi) {
          return __ldcg(reinterpret_cast<const float2*>(p.part_o + (wslot0 + pi) * SLOT_FLOATS + 32 * HD) + lane);
        };
        for (int pi = 0; pi < npartners; ++pi) {
          for (uint32_t spins = 0; ld_acquire_gpu(p.flags + wslot0 + pi) == 0u; ++spins) {
            __nanosleep(64);
            if (spins > (1u << 24)) __trap();  // partner never published: fail loudly, do not hang
          }
        }
        float mm = m;
        for (int g = 0; g < npartners; g += kGroup) {
          float pm[kGroup];
#pragma unroll
          for (int u = 0; u < kGroup; ++u) pm[u] = (g + u < npartners) ? slot_ml(g + u).x : -INFINITY;
#pragma unroll
          for (int u = 0; u < kGroup; ++u) mm = fmaxf(mm, pm[u]);
        }
        const float wl = (m == -INFINITY) ? 0.f : ex2_approx((m - mm) * LOG2E);
        float lt = l * wl, inv = 0.f;
        // Coalesced store: each warp transposes its 32 rows x 128 B through its own 4 KB slice of
        // the (now idle) P tile, then writes whole 128-byte row segments (8 lanes per row); the
        // thread-per-row store cost ~3000 cycles per CTA (32 cache lines per instruction).
        uint8_t* stg = sPt + lg * 4096;
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float o[32];
          load_half(c, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= wl;
#pragma unroll 1
          for (int g = 0; g < npartners; g += kGroup) {
            float pw[kGroup];
#pragma unroll
            for (int u = 0; u < kGroup; ++u) {
              pw[u] = 0.f;
              if (g + u < npartners) {
                const float2 ml = slot_ml(g + u);
                pw[u] = (ml.x == -INFINITY) ? 0.f : ex2_approx((ml.x - mm) * LOG2E);
                if (c == 0) lt = fmaf(ml.y, pw[u], lt);
              }
            }
            // 16 columns at a time: 4 float4 of every partner of the group in flight together
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float4 v[kGroup][4];
#pragma unroll
              for (int u = 0; u < kGroup; ++u)
                if (g + u < npartners) {
                  const float4* po = reinterpret_cast<const float4*>(p.part_o + (wslot0 + g + u) * SLOT_FLOATS);
#pragma unroll
                  for (int q = 0; q < 4; ++q) v[u][q] = __ldcg(po + (c * 8 + h * 4 + q) * 32 + lane);
                }
#pragma unroll
              for (int u = 0; u < kGroup; ++u)
                if (g + u < npartners) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const int i0 = (h * 4 + q) * 4;
                    o[i0 + 0] = fmaf(v[u][q].x, pw[u], o[i0 + 0]);
                    o[i0 + 1] = fmaf(v[u][q].y, pw[u], o[i0 + 1]);
                    o[i0 + 2] = fmaf(v[u][q].z, pw[u], o[i0 + 2]);
                    o[i0 + 3] = fmaf(v[u][q].w, pw[u], o[i0 + 3]);
                  }
                }
            }
          }
          if (c == 0) inv = lt > 0.f ? 1.0f / lt : 0.f;   // every partner's l has been added by now
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(o[8 * q + 0] * inv, o[8 * q + 1] * inv);
            u.y = pack_bf16(o[8 * q + 2] * inv, o[8 * q + 3] * inv);
            u.z = pack_bf16(o[8 * q + 4] * inv, o[8 * q + 5] * inv);
            u.w = pack_bf16(o[8 * q + 6] * inv, o[8 * q + 7] * inv);
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((c * 4 + q) ^ (lane & 7)) * 16)) = u;
          }
        }
        __syncwarp();
        if (lane == 0)
          for (int pi = 0; pi < npartners; ++pi)
            p.flags[wslot0 + pi] = 0u;  // re-armed for the next launch (stream ordered)
        bf16* obase = p.O + static_cast<size_t>(b * p.Lq + q0 + tile * BQ + lg * 32) * p.ldo + head * HD;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), ch = lane & 7;
          const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((ch ^ (rr & 7)) * 16));
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * p.ldo + ch * 8) = u;
        }
      }
      if (ktr) ktrp[3] = clock64();
      tc_fence_before_sync();
    }
  }
  __syncthreads();
  if (ktr) ktrp[4] = clock64();
  if (ctr != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ctr[2] = t;
  }
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// Merge the split-KV partials of every (row, head): out = sum_s w_s O_s / sum_s w_s l_s with
// w_s = exp(m_s - max_s m_s); a split with no attendable key has m = -inf, l = 0 (weight 0).
__global__ void __launch_bounds__(256)
attention_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                         bf16* __restrict__ O, int ldo, int heads, int splits, long long n_rh) {
  griddep_launch_dependents();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rh = gid >> 4;  // (row, head) pair; 16 threads x 4 columns each
  const int c4 = static_cast<int>(gid & 15);
  if (rh >= n_rh) return;
  griddep_wait();
  float mmax = -INFINITY;
  for (int s = 0; s < splits; ++s) mmax = fmaxf(mmax, part_ml[(rh * splits + s) * 2]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float lt = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 ml = *reinterpret_cast<const float2*>(part_ml + (rh * splits + s) * 2);
    const float w = (ml.x == -INFINITY) ? 0.f : __expf(ml.x - mmax);
    const float4 v = *reinterpret_cast<const float4*>(part_o + (rh * splits + s) * HD + c4 * 4);
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    lt += w * ml.y;
  }
  const float inv = lt > 0.f ? 1.0f / lt : 0.f;
  const long long row = rh / heads;
  const int head = static_cast<int>(rh - row * heads);
  uint2 u;
  u.x = pack_bf16(acc.x * inv, acc.y * inv);
  u.y = pack_bf16(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(O + row * ldo + head * HD + c4 * 4) = u;
}

}  // namespace

// SMs of the current device (148 on B200).  The long/short split below relies on every short CTA
// finding an SM that no long CTA occupies, so the real count is used, not the nominal one.
static int device_sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
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

// CTAs of the 64-key instance that can be resident at the same time.  The runtime's occupancy
// calculator answers 1 for every kernel that contains tcgen05.alloc (tools/ubench/occ_probe.cu), so
// the count is derived from the resources themselves: shared memory (+ the per-block reserve),
// registers and tensor memory (256 of the SM's 512 columns per CTA) -> 2 per SM on B200.
static int slots_bkv64() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int smem_sm = 0, smem_resv = 0, regs_sm = 0;
    cudaFuncAttributes fa;
    int per_sm = 0;
    if (cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&smem_resv, cudaDevAttrReservedSharedMemoryPerBlock, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev) == cudaSuccess &&
        cudaFuncGetAttributes(&fa, attention_tcgen05_kernel<64>) == cudaSuccess) {
      const int by_smem = smem_sm / (ACfg<64>::SMEM + smem_resv);
      const int by_regs = regs_sm / (fa.numRegs * ATTN_THREADS);
      const int by_tmem = 512 / static_cast<int>(ACfg<64>::TMEM_COLS);
      per_sm = by_smem < by_regs ? by_smem : by_regs;
      per_sm = per_sm < by_tmem ? per_sm : by_tmem;
    }
    cached = per_sm * device_sm_count();
    cached_dev = dev;
  }
  return cached;
}

// Which instance runs: 64-key blocks (2 CTAs per SM) unless MSD_ATTN_BKV=128 asks for the
// one-CTA-per-SM kernel (kept for comparison and as the fallback when the device grants only one
// CTA of the small instance per SM).
static int attention_bkv(int Lk, int ctas) {
  const char* e = getenv("MSD_ATTN_BKV");  // read per launch: the tests switch instances
  const int forced = e ? atoi(e) : 0;
  if (forced == 128 || forced == 64) return (forced == 64 && Lk % 64 == 0) ? 64 : 128;
  // Measured on B200 (tools/attn_bench.py, profiles/README.md): the two-CTAs-per-SM instance wins
  // when the grid has at least one CTA per SM (B = 8 self-attention, 192 CTAs in ONE wave: 13.9 vs
  // 17.3 us; token encoder, 768 CTAs: 193 vs 210 us).  Two co-resident CTAs do not interleave
  // evenly (the warp scheduler serves the higher warp slots first: one CTA of an SM runs at full
  // speed, the other mostly afterwards), so a grid that fits one CTA per SM is still better off
  // with the 128-key blocks and their long/short tail split (B = 8 cross-attention, 96 CTAs: 34.6
  // vs 37.4 us), and small grids (one segment: 12-24 CTAs) are latency chains where fewer, larger
  // blocks win (16.5 vs 24.2 us).
  return (Lk % 64 == 0 && slots_bkv64() >= 2 * device_sm_count() && ctas >= device_sm_count()) ? 64
                                                                                                : 128;
}

int attention_pick_splits(int nbatch, int heads, int Lq, int Lk) {
  const int ctas = ((Lq + 2 * BQ - 1) / (2 * BQ)) * heads * nbatch;
  if (attention_bkv(Lk, ctas) == 64) {
    // every CTA of the launch should be resident at once (one wave of 2 CTAs per SM): the largest
    // split count that fits, with at least four 64-key blocks per CTA
    const int nkb = Lk / 64, slots = slots_bkv64();
    int best = 1;
    for (int s = 2; s <= 4; ++s)
      if (nkb % s == 0 && nkb / s >= 4 && ctas * s <= slots) best = s;
    return best;
  }
  const int nkb = Lk / 128;
  // measured on B200: splitting pays only when fewer than half the SMs would be busy (B = 8
  // cross-attention, 96 CTAs, is no faster split 3-way: per-CTA fixed costs eat the gain)
  const int sms = device_sm_count();
  if (ctas >= sms / 2 || nkb < 6) return 1;
  int best = 1;
  for (int s = 2; s <= 8; ++s) {
    if (nkb % s != 0 || nkb / s < 3) continue;  // >= 3 key blocks per CTA keeps the prologue small
    if (ctas * s <= 2 * sms) best = s;
  }
  return best;
}

// Tail split (128-key instance) for grids that fill between half and all of the SMs (one CTA per
// SM): long CTAs take nkb - t key blocks, short CTAs t, shorts run in the SMs the longs leave free
// (several rounds).  Cost model in key-block units with a fixed per-CTA cost F (setup + epilogue,
// measured ~4.5 on B200: 18 blocks / 96 CTAs -> tail 4).
int attention_pick_tail(int nbatch, int heads, int Lq, int Lk) {
  const int ctas = ((Lq + 2 * BQ - 1) / (2 * BQ)) * heads * nbatch;
  if (attention_bkv(Lk, ctas) == 64) return 0;
  const int nkb = Lk / 128;
  const int sms = device_sm_count();
  if (ctas < sms / 2 || ctas >= sms || nkb < 6) return 0;
  const int free_sms = sms - ctas;
  const int rounds = (ctas + free_sms - 1) / free_sms;
  const float F = 4.5f;
  int best = 0;
  float best_t = (nkb + F) * 0.92f;  // must beat the unsplit kernel by a margin
  for (int t = 1; t <= nkb / 2; ++t) {
    const float tl = (nkb - t) + F + 0.3f, ts = rounds * (t + F);
    const float tt = tl > ts ? tl : ts;
    if (tt < best_t) { best_t = tt; best = t; }
  }
  return best;
}

int attention_configure() {
  MSD_CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel<128>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<128>::SMEM));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel<64>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<64>::SMEM));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel<64>,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  MSD_CUDA_CHECK(cudaFuncSetAttribute(attention_combine_kernel,
                                      cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  return 0;
}

// floats of the part_o workspace / flag words a launch may need (engine and op hooks size theirs
// with these): the larger of the combine-kernel layout and the per-warp slot layout
size_t attention_workspace_floats(int nbatch, int heads, int Lq, int max_splits) {
  const size_t rows = static_cast<size_t>(nbatch) * Lq;
  const size_t combine = rows * heads * max_splits * HD;
  const size_t warps = static_cast<size_t>(nbatch) * heads * ((Lq + BQ - 1) / BQ) * 4;
  const size_t slots = warps * (max_splits > 1 ? max_splits - 1 : 1) * SLOT_FLOATS;
  return combine > slots ? combine : slots;
}
size_t attention_flag_words(int nbatch, int heads, int Lq, int max_splits) {
  return static_cast<size_t>(nbatch) * heads * ((Lq + BQ - 1) / BQ) * 4 * (max_splits > 1 ? max_splits - 1 : 1);
}

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  static int configured = attention_configure();
  if (configured != 0) return configured;
  const int bkv = attention_bkv(a.Lk, ((a.Lq + 2 * BQ - 1) / (2 * BQ)) * a.heads * a.nbatch);
  MSD_REQUIRE(a.Lq % BQ == 0 && a.Lk % 128 == 0,
              "attention: Lq=%d and Lk=%d must be multiples of 128", a.Lq, a.Lk);
  MSD_REQUIRE(a.Lk / bkv <= 64, "attention: Lk=%d exceeds 64 key blocks", a.Lk);
  MSD_REQUIRE(a.nbatch > 0 && a.heads > 0, "attention: empty problem");
  MSD_REQUIRE(a.ldo % 8 == 0, "attention: ldo must be a multiple of 8");
  if (a.mask_bits)
    MSD_REQUIRE(a.mask_stride_words % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(a.mask_bits) & 15) == 0,
                "attention: mask words must be 16-byte aligned per row");
  CUtensorMap tq, tk, tv;
  const int width = a.heads * HD;
  if (a.tmap_q) tq = *a.tmap_q;
  else if (int rc = make_tmap_bf16_2d(&tq, a.Q, (uint64_t)a.nbatch * a.Lq, width, a.ldq, BQ)) return rc;
  const int kv_batch_rows = a.kv_batch_rows > 0 ? a.kv_batch_rows : a.Lk;
  MSD_REQUIRE(a.kv_row0 >= 0 && a.kv_row0 + a.Lk <= kv_batch_rows,
              "attention: key rows [%d, %d) exceed the %d rows per batch", a.kv_row0, a.kv_row0 + a.Lk,
              kv_batch_rows);
  const uint64_t kv_rows = (uint64_t)a.nbatch * kv_batch_rows;
  if (a.tmap_k) tk = *a.tmap_k;
  else if (int rc = make_tmap_bf16_2d(&tk, a.K, kv_rows, width, a.ldk, bkv)) return rc;
  if (a.tmap_v) tv = *a.tmap_v;
  else if (int rc = make_tmap_bf16_2d(&tv, a.V, kv_rows, width, a.ldv, bkv)) return rc;
  AttnDev d;
  d.O = a.O; d.ldo = a.ldo; d.heads = a.heads; d.Lq = a.Lq; d.Lk = a.Lk;
  d.mask_bits = a.mask_bits; d.mask_stride_words = a.mask_stride_words;
  d.trace = a.trace;
  const int nkb = a.Lk / bkv;
  const int ctas = ((a.Lq + 2 * BQ - 1) / (2 * BQ)) * a.heads * a.nbatch;
  int splits = (a.part_o != nullptr && a.part_ml != nullptr)
                   ? (a.splits > 0 ? a.splits : attention_pick_splits(a.nbatch, a.heads, a.Lq, a.Lk))
                   : 1;
  if (splits > a.max_splits) splits = a.max_splits > 0 ? a.max_splits : 1;
  MSD_REQUIRE(nkb % splits == 0, "attention: %d key blocks not divisible by %d splits", nkb, splits);
  d.splits = splits; d.part_o = a.part_o; d.part_ml = a.part_ml;
  // In-kernel merge of the splits (no combine kernel): the owner CTAs wait for their partners, so
  // the whole grid must be resident at once.  MSD_ATTN_MERGE=0 forces the combine kernel.
  const char* merge_env = getenv("MSD_ATTN_MERGE");
  const bool merge_allowed = !(merge_env && merge_env[0] == '0');
  // Default: only with the 64-key instance (few partners, all CTAs resident).  For the small
  // grids of one segment (128-key instance, 6 splits) the owner's wait + serial merge measured
  // slower than the combine kernel (batch-1 step 1035 vs 977 us); MSD_ATTN_MERGE=2 forces it.
  const bool merge_forced = merge_env && merge_env[0] == '2';
  d.merge = (splits > 1 && a.flags != nullptr && merge_allowed && (bkv == 64 || merge_forced) &&
             ctas * splits <= (bkv == 64 ? slots_bkv64() : device_sm_count())) ? 1 : 0;
  int tail = 0;
  if (bkv == 128 && splits == 1 && a.flags != nullptr && a.part_o != nullptr && a.part_ml != nullptr &&
      a.tail >= 0)
    tail = a.tail > 0 ? a.tail : attention_pick_tail(a.nbatch, a.heads, a.Lq, a.Lk);
  MSD_REQUIRE(tail < nkb, "attention: tail %d must be below %d key blocks", tail, nkb);
  if (tail > 0) {
    // long CTAs wait for their short partners: every long CTA must be resident together with at
    // least one SM left for the short ones, or the wait could never end
    MSD_REQUIRE(ctas < device_sm_count(), "attention: tail split needs fewer long CTAs (%d) than SMs (%d)",
                ctas, device_sm_count());
  }
  d.tail = tail; d.nbatch = a.nbatch; d.flags = a.flags; d.kv_static = a.kv_static;
  d.kv_batch_rows = kv_batch_rows; d.kv_row0 = a.kv_row0;
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.heads, a.nbatch * (tail > 0 ? 2 : splits));
  if (getenv("MSD_ATTN_DEBUG"))
    fprintf(stderr, "[attn] nb=%d Lq=%d Lk=%d bkv=%d ctas=%d splits=%d merge=%d tail=%d\n", a.nbatch, a.Lq,
            a.Lk, bkv, ctas, splits, d.merge, tail);
  ProfScope prof(KC_ATTENTION, 4.0 * a.nbatch * a.heads * static_cast<double>(a.Lq) * a.Lk * HD,
                 2.0 * a.nbatch * a.heads * HD * (2.0 * a.Lq + 2.0 * a.Lk), stream);
  if (bkv == 64)
    MSD_CUDA_CHECK(launch_kernel(attention_tcgen05_kernel<64>, grid, dim3(ATTN_THREADS), ACfg<64>::SMEM,
                                 stream, tq, tk, tv, d));
  else
    MSD_CUDA_CHECK(launch_kernel(attention_tcgen05_kernel<128>, grid, dim3(ATTN_THREADS),
                                 ACfg<128>::SMEM, stream, tq, tk, tv, d));
  ++g_launch_count;
  if (splits > 1 && !d.merge) {
    const long long n_rh = static_cast<long long>(a.nbatch) * a.Lq * a.heads;
    const long long threads = n_rh * 16;
    MSD_CUDA_CHECK(launch_kernel(attention_combine_kernel, dim3(static_cast<unsigned>((threads + 255) / 256)),
                                 dim3(256), 0, stream, static_cast<const float*>(a.part_o),
                                 static_cast<const float*>(a.part_ml), a.O, a.ldo, a.heads, splits,
                                 n_rh));
    ++g_launch_count;
  }
  return 0;
}

}  // namespace msd
