// Which feature of the attention kernel pins the occupancy calculator to one CTA per SM?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o occ_probe occ_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 2) k_plain(float* o) { o[threadIdx.x] = 1.f; }

__global__ void __launch_bounds__(384, 2) k_setmaxnreg(float* o) {
  if (threadIdx.x >= 256) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  o[threadIdx.x] = 1.f;
}

__global__ void __launch_bounds__(384, 2) k_tmem(float* o) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
  o[threadIdx.x] = 1.f;
}

__global__ void __launch_bounds__(384, 2) k_pdl(float* o) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  o[threadIdx.x] = 1.f;
}

__global__ void __launch_bounds__(384, 2) k_bar(float* o) {
  asm volatile("bar.sync 2, 256;" ::: "memory");
  o[threadIdx.x] = 1.f;
}

template <typename K>
void probe(const char* name, K k) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  for (int kb : {0, 64, 112}) {
    int occ = -1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 384, (size_t)kb * 1024);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k);
    printf("%-14s dyn smem %3d KB: occupancy %d (%s) regs %d\n", name, kb, occ, cudaGetErrorString(e), fa.numRegs);
  }
}

int main() {
  probe("plain", k_plain);
  probe("setmaxnreg", k_setmaxnreg);
  probe("tcgen05.alloc", k_tmem);
  probe("griddepcontrol", k_pdl);
  probe("named barrier", k_bar);
  return 0;
}
