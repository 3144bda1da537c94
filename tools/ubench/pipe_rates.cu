// Issue-rate micro-benchmark for the instructions of the attention softmax (one warp per SM
// sub-partition, 8 independent chains, clock64 around an unrolled loop).  Prints cycles per
// warp-instruction.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 256

template <int OP>
__global__ void bench(float* out, long long* cyc, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1) * 1e-3f;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 77u + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) {  // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (OP == 1) {  // F2FP bf16x2 pack
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        a[i] = __uint_as_float(u[i]);
      } else if (OP == 2) {  // FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(seed));
      } else if (OP == 3) {  // FFMA2
        uint64_t v, s2;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(s2) : "f"(seed));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(v) : "l"(s2));
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
        a[i] = lo + 0.f * hi;
      } else if (OP == 4) {  // FMNMX3
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
      } else if (OP == 5) {  // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      } else if (OP == 6) {  // IADD3-ish
        asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      } else if (OP == 7) {  // LEA (shift-add)
        asm volatile("{ .reg .u32 t; shl.b32 t, %1, 23; add.u32 %0, %0, t; }" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      } else if (OP == 8) {  // FSEL
        asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, 0f00000000; selp.f32 %0, %0, %1, p; }" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      } else if (OP == 9) {  // cvt f32 -> f16x2 pack
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        a[i] = __uint_as_float(u[i]);
      } else if (OP == 10) {  // ex2 f16x2
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      } else if (OP == 11) {  // ex2 bf16x2
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      } else if (OP == 12) {  // FADD2
        uint64_t v;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(v));
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
        a[i] = lo + 0.f * hi;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; x ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x & 0xff);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, 8 * sizeof(long long));
  bench<OP><<<1, threads>>>(out, cyc, 0.999f);
  bench<OP><<<1, threads>>>(out, cyc, 0.999f);
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-22s warps/SMSP=%d  cycles per warp-instr per SMSP = %.2f  (%s)\n", name, threads / 128,
         double(h) / (ITERS * 8.0 * (threads / 128)), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256}) {
    run<0>("MUFU.EX2 f32", threads);
    run<1>("F2FP bf16x2 <- f32", threads);
    run<9>("F2FP f16x2 <- f32", threads);
    run<10>("MUFU.EX2 f16x2", threads);
    run<11>("MUFU.EX2 bf16x2", threads);
    run<2>("FFMA", threads);
    run<3>("FFMA2 (+movs, fadd)", threads);
    run<12>("FADD2 (+movs, fadd)", threads);
    run<4>("FMNMX3", threads);
    run<5>("PRMT", threads);
    run<6>("IADD", threads);
    run<7>("SHL+IADD", threads);
    run<8>("FSETP+FSEL", threads);
  }
  return 0;
}
