// Correctness probe: can a K-major SWIZZLE_128B A operand be read starting at an arbitrary ROW of a shared-memory tile
// (descriptor start address advanced by shift * 128 B), and does the descriptor's base-offset field have to carry
// (address >> 7) & 7 for that?  This is what a k-tap convolution needs to take all its taps from ONE activation tile
// loaded with a halo, instead of re-fetching the tile once per tap.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I xva-trainer_b200/csrc -o scripts/_probe_rowshift scripts/rowshift_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace xva;

constexpr int ROWS = 192;  // rows in the staged A tile (128 + halo)
constexpr int N = 32;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4; d |= 2u << 7; d |= 2u << 10;
  d |= (uint32_t)(n >> 3) << 17; d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// a [ROWS, 32], b [N, 32] (row-major fp32, tf32-exact values) -> out[128, N] = a[shift : shift + 128] @ b^T
__global__ void __launch_bounds__(128, 1) probe(const float* a, const float* b, int shift, int use_base_off, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // ROWS x 128 B
  uint8_t* sb = smem + 32 * 1024;     // N x 128 B
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 32); ptx::tmem_relinquish(); }
  // software SWIZZLE_128B (what TMA writes): 16-byte chunk c of row r lands at r * 128 + ((c ^ (r & 7)) * 16)
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    *(float*)(sa + r * 128 + (((c / 4) ^ (r & 7)) * 16) + (c % 4) * 4) = a[i];
  }
  for (int i = threadIdx.x; i < N * 32; i += blockDim.x) {
    const int r = i / 32, c = i % 32;
    *(float*)(sb + r * 128 + (((c / 4) ^ (r & 7)) * 16) + (c % 4) * 4) = b[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t a_addr = ptx::smem_u32(sa) + shift * 128;
      const uint32_t bo = use_base_off ? ((a_addr >> 7) & 7) : 0;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t da = desc(a_addr + k4 * 32, 16, 1024, 2, bo);
        const uint64_t db = desc(ptx::smem_u32(sb) + k4 * 32, 16, 1024, 2, 0);
        ptx::mma_tf32(tm, da, db, idesc(128, N), k4 ? 1u : 0u);
      }
      ptx::mma_commit(&bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  uint32_t v[32];
  ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  ptx::tmem_wait_ld();
  for (int n = 0; n < N; ++n) out[(warp * 32 + (threadIdx.x & 31)) * N + n] = __uint_as_float(v[n]);
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 32);
}

int main() {
  std::vector<float> ha(ROWS * 32), hb(N * 32);
  for (size_t i = 0; i < ha.size(); ++i) ha[i] = (float)((int)((i * 2654435761u) >> 24) - 128) / 64.0f;   // tf32-exact
  for (size_t i = 0; i < hb.size(); ++i) hb[i] = (float)((int)((i * 40503u + 7) % 255) - 127) / 128.0f;
  float *da, *db, *dout;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(128 * N);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 8, 1, 3, 5, 13, 25, 50}) {
      cudaMemset(dout, 0, 128 * N * 4);
      probe<<<1, 128, 64 * 1024>>>(da, db, shift, bo, dout);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0, ref2 = 0;
      int bad_rows = 0;
      for (int r = 0; r < 128; ++r) {
        double row_err = 0;
        for (int n = 0; n < N; ++n) {
          double want = 0;
          for (int k = 0; k < 32; ++k) want += (double)ha[(r + shift) * 32 + k] * hb[n * 32 + k];
          row_err += (ho[r * N + n] - want) * (ho[r * N + n] - want);
          ref2 += want * want;
        }
        err += row_err;
        bad_rows += row_err > 1e-6;
      }
      printf("base_offset=%s shift=%2d: rel err %.3e, wrong rows %3d / 128  (%s)\n", bo ? "(addr>>7)&7" : "0", shift,
             std::sqrt(err / ref2), bad_rows, cudaGetErrorString(e));
    }
  return 0;
}
