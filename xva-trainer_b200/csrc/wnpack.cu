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

// Vector path (float4 global loads / stores): the kernel is a pure HBM stream, and with one 4-byte access in flight per
// thread it sat at ~1.2 TB/s (Little's law: 148 SMs x 2048 threads x 4 B = 1.2 MB in flight against the ~5 MB the
// memory system needs). Taken when every tap's channel run is a whole number of aligned float4s.
// (Parameter rows may start at any 4-byte offset -- the optimizers keep all parameters in one flat arena and a
// 1-element bias shifts everything behind it -- so their alignment is checked per row; the packed arena is aligned by
// construction.)
__device__ __forceinline__ bool vec_ok(const xva_wn_desc& d, int c2) {
  return !(d.flags & XVA_WN_TRANSPOSED) && (c2 & 3) == 0 && (d.ld & 3) == 0 && (d.cg & 3) == 0 && (d.inner & 3) == 0;
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(kThreadsWn)
wn_pack_fwd_kernel(const xva_wn_desc* __restrict__ table, int n_desc, int allow_vec) {
  extern __shared__ float row_s[];
  __shared__ float red[kThreadsWn / 32];
  const xva_wn_desc& d = *find_desc(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  const float* v = d.v + static_cast<long>(r) * inner;
  const bool vec = allow_vec && vec_ok(d, c2);
  float ss = 0.0f;
  if (vec && aligned16(v)) {
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (int i = threadIdx.x; i < (inner >> 2); i += kThreadsWn) {
      const float4 x = v4[i];
      *reinterpret_cast<float4*>(row_s + 4 * i) = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  } else {
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
      const float x = v[i];
      row_s[i] = x;
      ss += x * x;
    }
  }
  ss = block_sum_f(ss, red);  // (also orders the row_s writes before the reads below)
  const float scale = (d.flags & XVA_WN_PLAIN) ? 1.0f : d.g[r] / sqrtf(ss);
  const bool rnd = !(d.flags & XVA_WN_NO_ROUND);
  if (vec) {
    for (int i4 = threadIdx.x; i4 < (inner >> 2); i4 += kThreadsWn) {  // 4 consecutive channels of one tap
      const int i = i4 << 2, j = i / c2, c = i - j * c2;
      float4 w = make_float4(row_s[c * k + j] * scale, row_s[(c + 1) * k + j] * scale, row_s[(c + 2) * k + j] * scale,
                             row_s[(c + 3) * k + j] * scale);
      if (rnd) w = tf32_rn4(w);
      *reinterpret_cast<float4*>(d.dst + dst_index(d, r, c, j)) = w;
    }
  } else {
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) {  // i = j * c2 + c: channel fastest in the packed matrices
      const int j = i / c2, c = i - j * c2;
      const float w = row_s[c * k + j] * scale;
      d.dst[dst_index(d, r, c, j)] = rnd ? tf32_rn(w) : w;
    }
  }
}

// dL/dv = (g / ||v||) * (dW - v * (v . dW) / ||v||^2),  dL/dg = (v . dW) / ||v||,  dW gathered from the packed layout
__global__ void __launch_bounds__(kThreadsWn)
wn_pack_bwd_kernel(const xva_wn_desc* __restrict__ table, int n_desc, int allow_vec) {
  extern __shared__ float smem_f[];
  __shared__ float red[kThreadsWn / 32];
  const xva_wn_desc& d = *find_desc(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  float* row_v = smem_f;
  float* row_d = smem_f + inner;
  const float* v = d.v + static_cast<long>(r) * inner;
  const bool vec = allow_vec && vec_ok(d, c2);
  float ss = 0.0f;
  if (vec && aligned16(v)) {
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (int i = threadIdx.x; i < (inner >> 2); i += kThreadsWn) {
      const float4 x = v4[i];
      *reinterpret_cast<float4*>(row_v + 4 * i) = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  } else {
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
      const float x = v[i];
      row_v[i] = x;
      ss += x * x;
    }
  }
  if (vec) {
    for (int i4 = threadIdx.x; i4 < (inner >> 2); i4 += kThreadsWn) {
      const int i = i4 << 2, j = i / c2, c = i - j * c2;
      const float4 g4 = *reinterpret_cast<const float4*>(d.ddst + dst_index(d, r, c, j));
      row_d[c * k + j] = g4.x;
      row_d[(c + 1) * k + j] = g4.y;
      row_d[(c + 2) * k + j] = g4.z;
      row_d[(c + 3) * k + j] = g4.w;
    }
  } else {
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
      const int j = i / c2, c = i - j * c2;
      row_d[c * k + j] = d.ddst[dst_index(d, r, c, j)];
    }
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
  if (allow_vec && (inner & 3) == 0 && aligned16(dv)) {
    float4* dv4 = reinterpret_cast<float4*>(dv);
    for (int i = threadIdx.x; i < (inner >> 2); i += kThreadsWn) {
      float4 o = dv4[i];
      const float4 gd = *reinterpret_cast<const float4*>(row_d + 4 * i), gv = *reinterpret_cast<const float4*>(row_v + 4 * i);
      o.x += scale * gd.x - coef * gv.x;
      o.y += scale * gd.y - coef * gv.y;
      o.z += scale * gd.z - coef * gv.z;
      o.w += scale * gd.w - coef * gv.w;
      dv4[i] = o;
    }
  } else {
    for (int i = threadIdx.x; i < inner; i += kThreadsWn) dv[i] += scale * row_d[i] - coef * row_v[i];
  }
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
  static const int allow_vec = [] {  // XVA_WNPACK_VEC=0: the scalar path (A/B and debugging)
    const char* e = getenv("XVA_WNPACK_VEC");
    return !(e && e[0] == '0') ? 1 : 0;
  }();
  if (backward) wn_pack_bwd_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc, allow_vec);
  else wn_pack_fwd_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc, allow_vec);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(wnpack)

}  // namespace xva
