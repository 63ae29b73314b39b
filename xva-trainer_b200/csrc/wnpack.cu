// Weight-norm reparametrisation + re-packing of every convolution of a HiFi-GAN model in ONE launch, and its backward
// in one more (hifigan/models.py wraps each conv in torch.nn.utils.weight_norm: w = g * v / ||v||, the norm taken over
// everything but dim 0). The packed layout is what the tap-GEMM reads: one [rows, ld] matrix per kernel tap, groups
// with fewer than 32 input channels laid out as block-diagonal super-groups (see xva_gemm_args.groups), transposed
// convolutions with the roles of the two channel dimensions swapped. PyTorch eager spends ~20 launches per convolution
// on this (norm, div, mul, permute, index_select, pad, copy + the autograd mirror of each): ~9 000 launches per training
// step for the generator + 8 discriminators, about a quarter of the step's device time before this kernel existed.
//
// One block per normalised row (dim-0 slice of v). The row is staged in shared memory so that both the read of v
// (k fastest) and the write of the packed matrices (channel fastest) are coalesced.
#include "common.cuh"
#include "ops.cuh"
#include "../../include/xva_b200.h"

namespace xva {

namespace {

constexpr int kThreadsWn = 256;

__device__ __forceinline__ float block_sum_f(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < kThreadsWn / 32; ++w) t += sh[w];
  return t;  // valid in every thread
}

__device__ __forceinline__ const xva_wn_desc* find_desc(const xva_wn_desc* table, int n_desc, int row) {
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {  // last descriptor whose row_start <= row
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  return table + lo;
}

// offset of element (r, c, j) of v in the packed arena
__device__ __forceinline__ long dst_index(const xva_wn_desc& d, int r, int c, int j) {
  if (d.flags & XVA_WN_TRANSPOSED) return d.tap_off[j] + static_cast<long>(c) * d.ld + r;
  return d.tap_off[j] + static_cast<long>(r) * d.ld + ((r / d.og) % d.f) * d.cg + c;
}

__global__ void __launch_bounds__(kThreadsWn)
wn_pack_fwd_kernel(const xva_wn_desc* __restrict__ table, int n_desc) {
  extern __shared__ float row_s[];
  __shared__ float red[kThreadsWn / 32];
  const xva_wn_desc& d = *find_desc(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  const float* v = d.v + static_cast<long>(r) * inner;
  float ss = 0.0f;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const float x = v[i];
    row_s[i] = x;
    ss += x * x;
  }
  ss = block_sum_f(ss, red);  // (also orders the row_s writes before the reads below)
  const float scale = (d.flags & XVA_WN_PLAIN) ? 1.0f : d.g[r] / sqrtf(ss);
  const bool rnd = !(d.flags & XVA_WN_NO_ROUND);
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {  // i = j * c2 + c: channel fastest in the packed matrices
    const int j = i / c2, c = i - j * c2;
    const float w = row_s[c * k + j] * scale;
    d.dst[dst_index(d, r, c, j)] = rnd ? tf32_rn(w) : w;
  }
}

// dL/dv = (g / ||v||) * (dW - v * (v . dW) / ||v||^2),  dL/dg = (v . dW) / ||v||,  dW gathered from the packed layout
__global__ void __launch_bounds__(kThreadsWn)
wn_pack_bwd_kernel(const xva_wn_desc* __restrict__ table, int n_desc) {
  extern __shared__ float smem_f[];
  __shared__ float red[kThreadsWn / 32];
  const xva_wn_desc& d = *find_desc(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  float* row_v = smem_f;
  float* row_d = smem_f + inner;
  const float* v = d.v + static_cast<long>(r) * inner;
  float ss = 0.0f;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const float x = v[i];
    row_v[i] = x;
    ss += x * x;
  }
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const int j = i / c2, c = i - j * c2;
    row_d[c * k + j] = d.ddst[dst_index(d, r, c, j)];
  }
  ss = block_sum_f(ss, red);
  float* dv = d.dv + static_cast<long>(r) * inner;
  if (d.flags & XVA_WN_PLAIN) {  // w = v: the packed gradient is the parameter's (block-uniform branch)
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) dv[i] += row_d[i];
    return;
  }
  float dot = 0.0f;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) dot += row_v[i] * row_d[i];
  dot = block_sum_f(dot, red);
  const float inv_norm = rsqrtf(ss);
  const float scale = d.g[r] * inv_norm;
  const float coef = scale * dot / ss;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) dv[i] += scale * row_d[i] - coef * row_v[i];
  if (threadIdx.x == 0) d.dg[r] += dot * inv_norm;
}

}  // namespace

int wn_pack(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, int backward, cudaStream_t stream) {
  XVA_CHECK_ARG(table_dev && n_desc >= 1 && total_rows >= 1, "wn_pack: empty table");
  const size_t smem = static_cast<size_t>(max_inner) * sizeof(float) * (backward ? 2 : 1);
  XVA_CHECK_ARG(max_inner >= 1 && smem <= 96 * 1024, "wn_pack: max_inner=%d does not fit shared memory", max_inner);
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(wn_pack_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    XVA_CHECK_CUDA(cudaFuncSetAttribute(wn_pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  if (backward) wn_pack_bwd_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc);
  else wn_pack_fwd_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(wnpack)

}  // namespace xva
