// Micro-benchmark: throughput of narrow tcgen05.ld (32x32b.xN) "column gathers" out of tensor memory,
// i.e. the access pattern of an aggregation that keeps P^T (lane = channel, column = node) in TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_gather ldtm_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ float ld_cols(uint32_t taddr);
template <>
__device__ __forceinline__ float ld_cols<1>(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}

template <int BATCH>
__global__ void __launch_bounds__(1024, 1) k_ldtm(int iters, unsigned long long* cycles, float* sink, int mode) {
  __shared__ uint32_t tbase_s;
  __shared__ float s_tile[256 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256 * 33; i += blockDim.x) s_tile[i] = (float)i;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t col = (warp * 37 + 11) & 511;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
      float v[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        v[j] = ld_cols<1>(tbase + col);
        col = (col * 5 + 17) & 511;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < BATCH; ++j) acc += v[j];
    }
  } else {   // the same gather out of shared memory: one row of 32 floats per "column"
    for (int it = 0; it < iters; ++it) {
      float v[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        v[j] = s_tile[(col & 255) * 33 + lane];
        col = (col * 5 + 17) & 511;
      }
#pragma unroll
      for (int j = 0; j < BATCH; ++j) acc += v[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "r"(512u) : "memory");
}

int main() {
  unsigned long long* d_cyc; float* d_sink;
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 148 * 1024 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        k_ldtm<8><<<148, threads>>>(iters, d_cyc, d_sink, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long c[148];
      cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
      const double bytes = (double)(threads / 32) * iters * 8 * 128.0;
      printf("%s warps=%2d cycles(sm0)=%llu  -> %.1f B/clk/SM  (%.2f clk per warp-load)\n", mode ? "LDS " : "LDTM", threads / 32,
             c[0], bytes / (double)c[0], (double)c[0] / ((double)(threads / 32) * iters * 8) * (threads / 32));
    }
  return 0;
}
