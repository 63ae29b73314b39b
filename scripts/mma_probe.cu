// Micro-probe: cycles per tcgen05.mma (kind::tf32, SS) for different descriptor / accumulator patterns.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xva-trainer_b200/csrc -o /tmp/mma_probe scripts/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace xva;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n, int amn, int bmn) {
  uint32_t d = 0;
  d |= 1u << 4; d |= 2u << 7; d |= 2u << 10;
  d |= (uint32_t)amn << 15; d |= (uint32_t)bmn << 16;
  d |= (uint32_t)(n >> 3) << 17; d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// mode bits: 0 = K-major A & B; 1 = MN-major A; 2 = MN-major B; 4 = alternate two accumulators; 8 = advance k within tile (4 slices) ; 16 = rotate over 4 smem stages
__global__ void __launch_bounds__(384, 1) probe(int mode, int M, int N, int iters, long long* out, int spin) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) {
    float v = 1.0f;
    if (spin & 8) {  // pseudo-random normal-ish data (what real activations look like to the datapath)
      uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = ((int)(h & 0xFFFF) - 32768) * (1.0f / 16384.0f) * (1.0f + (h >> 28) * 0.1f);
    }
    ((float*)smem)[i] = v;
  }
  spin &= 7;
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 1) {
    const bool amn = mode & 1, bmn = mode & 2;
    const uint32_t id = idesc(M, N, amn, bmn);
    const uint64_t da_hi = desc(0, amn ? 4096 : 16, amn ? 512 : 1024, amn ? 1 : 2);
    const uint64_t db_hi = desc(0, bmn ? 4096 : 16, bmn ? 512 : 1024, bmn ? 1 : 2);
    const uint32_t base = ptx::smem_u32(smem) >> 4;
    const uint32_t stage_u = (48 * 1024) >> 4;
    const uint32_t ak = (amn ? 1024 : 32) >> 4, bk = (bmn ? 1024 : 32) >> 4;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (ptx::elect_one()) {
        const uint32_t st = (mode & 16) ? (i & 3) * stage_u : 0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint32_t kk = (mode & 8) ? k4 : 0;
          const uint64_t da = da_hi | (uint64_t)(base + st + kk * ak);
          const uint64_t db = db_hi | (uint64_t)(base + st + (16384 >> 4) + kk * bk);
          const uint32_t acc = (mode & 4) ? (k4 & 1) * 256 : 0;
          ptx::mma_tf32(tm + acc, da, db, id, 1u);
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = (t1 - t0);
    if (ptx::elect_one()) ptx::mbar_arrive(&bar2);
    __syncwarp();
  } else if (warp >= 4 && spin) {
    // the waiting pattern of the GEMM's epilogue warps
    if (spin == 1) ptx::mbar_wait(&bar2, 0);
    else if (spin == 2) { while (!ptx::mbar_try_wait(&bar2, 0)) __nanosleep(200); }
    else if (spin == 3) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ptx::smem_u32(&bar2)), "r"(0), "r"(1000000) : "memory");
      }
    }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  struct Cfg { int mode, M, N; const char* name; } cfgs[] = {
    {0, 128, 256, "K-major A,B  same slice, 1 acc"}, {8, 128, 256, "K-major A,B  4 k-slices"}, {8 | 16, 128, 256, "K-major A,B  4 k-slices, 4 stages"},
    {4, 128, 256, "K-major, 2 accumulators"}, {0, 128, 128, "K-major N=128"}, {0, 128, 64, "K-major N=64"}, {0, 64, 256, "K-major M=64"},
    {1 | 8, 128, 256, "MN-major A, K-major B"}, {2 | 8, 128, 256, "K-major A, MN-major B"}, {3 | 8, 128, 256, "MN-major A,B"},
    {3 | 8, 128, 128, "MN-major A,B N=128"},
  };
  for (int spin : {0, 8})
  for (auto& c : cfgs) {
    for (int grid : {148}) {
      printf("spin=%d ", spin);
      probe<<<grid, 384, 200 * 1024>>>(c.mode, c.M, c.N, iters, d, spin);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("%-40s grid %3d: %7.1f cyc/MMA  (%s)\n", c.name, grid, (double)cyc / (iters * 4), cudaGetErrorString(e));
    }
  }
  return 0;
}
