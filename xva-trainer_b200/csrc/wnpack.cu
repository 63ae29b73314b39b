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
template <class Desc>
__device__ __forceinline__ long dst_index(const Desc& d, int r, int c, int j) {
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


// ================================================================================================ spectral norm
// torch.nn.utils.spectral_norm (dim 0, one power iteration per training forward, eps 1e-12) + the same re-packing, for
// the spectral-normed scale discriminator (hifigan/models.py:207-215 with use_spectral_norm): with W = weight_orig
// reshaped [rows, inner], u [rows], v [inner]
//     v <- normalize(W^T u),  u <- normalize(W v),  sigma = u . (W v),  w_eff = W / sigma
// and, with u, v constants as in torch, dL/dW = dW_eff / sigma - (<dW_eff, W> / sigma^2) u v^T.
// PyTorch eager spends ~14 launches per convolution and call on this (reshape, 3 mv, 2 normalize, clones, dot, div,
// index_select, permute, mask, contiguous, round) plus their autograd mirror -- and the reference calls the
// discriminator four times per step: ~450 launches, 3.2 ms of a 29.7 ms step (measured by replacing the spectral norm
// with weight norm, profiles/r02_spectral_norm_ab.txt). Here a call is 5 launches for all 8 convolutions (3 in eval
// mode), its backward 3; deterministic (two-stage sums, no floating-point atomics).
constexpr int kSnRowChunk = 64;

__device__ __forceinline__ const xva_sn_desc* find_sn_row(const xva_sn_desc* table, int n_desc, int row) {
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  return table + lo;
}
__device__ __forceinline__ const xva_sn_desc* find_sn_blk(const xva_sn_desc* table, int n_desc, int blk) {
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].blk_start <= blk) lo = mid;
    else hi = mid - 1;
  }
  return table + lo;
}
// scratch layout of one descriptor: [chunks * inner] partial column sums | [rows] row products | sigma, coef
__device__ __forceinline__ int sn_chunks(const xva_sn_desc& d) { return (d.rows + kSnRowChunk - 1) / kSnRowChunk; }
__device__ __forceinline__ float* sn_rowbuf(const xva_sn_desc& d) { return d.work + static_cast<long>(sn_chunks(d)) * d.inner; }
__device__ __forceinline__ float* sn_scalars(const xva_sn_desc& d) { return sn_rowbuf(d) + d.rows; }

// K1: partial[rc][c] = sum over the 64 rows of chunk rc of W[r, c] * u[r]
__global__ void __launch_bounds__(kThreadsWn)
sn_colsum_kernel(const xva_sn_desc* __restrict__ table, int n_desc) {
  const xva_sn_desc& d = *find_sn_blk(table, n_desc, blockIdx.x);
  const int local = blockIdx.x - d.blk_start;
  const int ncc = (d.inner + kThreadsWn - 1) / kThreadsWn;
  const int cc = local % ncc, rc = local / ncc;
  const int c = cc * kThreadsWn + threadIdx.x;
  if (c >= d.inner) return;
  const int r0 = rc * kSnRowChunk, r1 = min(d.rows, r0 + kSnRowChunk);
  const float* w = d.w + static_cast<long>(r0) * d.inner + c;
  float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
  int r = r0;
  for (; r + 4 <= r1; r += 4, w += 4L * d.inner) {  // four independent loads in flight per thread
    acc0 += w[0] * d.u[r];
    acc1 += w[d.inner] * d.u[r + 1];
    acc2 += w[2L * d.inner] * d.u[r + 2];
    acc3 += w[3L * d.inner] * d.u[r + 3];
  }
  for (; r < r1; ++r, w += d.inner) acc0 += w[0] * d.u[r];
  d.work[static_cast<long>(rc) * d.inner + c] = (acc0 + acc1) + (acc2 + acc3);
}

// K2 (one block per descriptor): t = sum of the partials; v = t / max(||t||, eps) -> module buffer and the call's copy
__global__ void __launch_bounds__(kThreadsWn)
sn_normalize_v_kernel(const xva_sn_desc* __restrict__ table) {
  __shared__ float red[kThreadsWn / 32];
  const xva_sn_desc& d = table[blockIdx.x];
  const int chunks = sn_chunks(d);
  float ss = 0.0f;
  for (int c = threadIdx.x; c < d.inner; c += kThreadsWn) {
    float t = 0.0f;
    for (int rc = 0; rc < chunks; ++rc) t += d.work[static_cast<long>(rc) * d.inner + c];
    d.work[c] = t;  // (chunk 0's slot: read by this thread only)
    ss += t * t;
  }
  ss = block_sum_f(ss, red);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = threadIdx.x; c < d.inner; c += kThreadsWn) {
    const float x = d.work[c] * inv;
    d.v[c] = x;
    d.v_sav[c] = x;
  }
}

// K3 (one block per row): s[r] = W[r, :] . v
__global__ void __launch_bounds__(kThreadsWn)
sn_rowdot_kernel(const xva_sn_desc* __restrict__ table, int n_desc, int training) {
  __shared__ float red[kThreadsWn / 32];
  const xva_sn_desc& d = *find_sn_row(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const float* w = d.w + static_cast<long>(r) * d.inner;
  const float* v = training ? d.v_sav : d.v;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < d.inner; i += kThreadsWn) acc += w[i] * v[i];
  acc = block_sum_f(acc, red);
  if (threadIdx.x == 0) sn_rowbuf(d)[r] = acc;
}

// K4 (one block per descriptor): training: u = s / max(||s||, eps); sigma = u . s.  eval: sigma = u . s with the stored u.
__global__ void __launch_bounds__(kThreadsWn)
sn_sigma_kernel(const xva_sn_desc* __restrict__ table, int training) {
  __shared__ float red[kThreadsWn / 32];
  const xva_sn_desc& d = table[blockIdx.x];
  const float* s = sn_rowbuf(d);
  float sigma;
  if (training) {
    float ss = 0.0f;
    for (int r = threadIdx.x; r < d.rows; r += kThreadsWn) ss += s[r] * s[r];
    ss = block_sum_f(ss, red);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    float dot = 0.0f;
    for (int r = threadIdx.x; r < d.rows; r += kThreadsWn) {
      const float x = s[r] * inv;
      d.u[r] = x;
      d.u_sav[r] = x;
      dot += x * s[r];
    }
    sigma = block_sum_f(dot, red);
  } else {
    float dot = 0.0f;
    for (int r = threadIdx.x; r < d.rows; r += kThreadsWn) {
      d.u_sav[r] = d.u[r];
      dot += d.u[r] * s[r];
    }
    sigma = block_sum_f(dot, red);
    for (int c = threadIdx.x; c < d.inner; c += kThreadsWn) d.v_sav[c] = d.v[c];
  }
  if (threadIdx.x == 0) sn_scalars(d)[0] = sigma;
}

// K5 (one block per row): dst = W[r, :] / sigma in the packed layout, tf32-rounded
__global__ void __launch_bounds__(kThreadsWn)
sn_pack_kernel(const xva_sn_desc* __restrict__ table, int n_desc) {
  extern __shared__ float row_s[];
  const xva_sn_desc& d = *find_sn_row(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  const float* w = d.w + static_cast<long>(r) * inner;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) row_s[i] = w[i];
  __syncthreads();
  const float scale = 1.0f / sn_scalars(d)[0];
  const bool rnd = !(d.flags & XVA_WN_NO_ROUND);
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const int j = i / c2, c = i - j * c2;
    const float x = row_s[c * k + j] * scale;
    d.dst[dst_index(d, r, c, j)] = rnd ? tf32_rn(x) : x;
  }
}

// B1 (one block per row): rowdot[r] = <dW_eff[r, :], W[r, :]>
__global__ void __launch_bounds__(kThreadsWn)
sn_bwd_rowdot_kernel(const xva_sn_desc* __restrict__ table, int n_desc) {
  __shared__ float red[kThreadsWn / 32];
  const xva_sn_desc& d = *find_sn_row(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  const float* w = d.w + static_cast<long>(r) * inner;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const int j = i / c2, c = i - j * c2;
    acc += d.ddst[dst_index(d, r, c, j)] * w[c * k + j];
  }
  acc = block_sum_f(acc, red);
  if (threadIdx.x == 0) sn_rowbuf(d)[r] = acc;
}

// B2 (one block per descriptor): coef = <dW_eff, W> / sigma^2
__global__ void __launch_bounds__(kThreadsWn)
sn_bwd_coef_kernel(const xva_sn_desc* __restrict__ table) {
  __shared__ float red[kThreadsWn / 32];
  const xva_sn_desc& d = table[blockIdx.x];
  const float* s = sn_rowbuf(d);
  float acc = 0.0f;
  for (int r = threadIdx.x; r < d.rows; r += kThreadsWn) acc += s[r];
  acc = block_sum_f(acc, red);
  if (threadIdx.x == 0) {
    const float sigma = sn_scalars(d)[0];
    sn_scalars(d)[1] = acc / (sigma * sigma);
  }
}

// B3 (one block per row): dw[r, i] += dW_eff[r, i] / sigma - coef * u[r] * v[i]
__global__ void __launch_bounds__(kThreadsWn)
sn_bwd_apply_kernel(const xva_sn_desc* __restrict__ table, int n_desc) {
  extern __shared__ float row_d[];
  const xva_sn_desc& d = *find_sn_row(table, n_desc, blockIdx.x);
  const int r = blockIdx.x - d.row_start;
  const int inner = d.inner, k = d.k, c2 = inner / k;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) {
    const int j = i / c2, c = i - j * c2;
    row_d[c * k + j] = d.ddst[dst_index(d, r, c, j)];
  }
  __syncthreads();
  const float inv_sigma = 1.0f / sn_scalars(d)[0];
  const float cu = sn_scalars(d)[1] * d.u_sav[r];
  float* dw = d.dw + static_cast<long>(r) * inner;
  for (int i = threadIdx.x; i < inner; i += kThreadsWn) dw[i] += row_d[i] * inv_sigma - cu * d.v_sav[i];
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

int sn_pack(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, int training,
            int backward, cudaStream_t stream) {
  XVA_CHECK_ARG(table_dev && n_desc >= 1 && total_rows >= 1 && total_blocks >= 1, "sn_pack: empty table");
  const size_t smem = static_cast<size_t>(max_inner) * sizeof(float);
  XVA_CHECK_ARG(max_inner >= 1 && smem <= 96 * 1024, "sn_pack: max_inner=%d does not fit shared memory", max_inner);
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(sn_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    XVA_CHECK_CUDA(cudaFuncSetAttribute(sn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  if (!backward) {
    if (training) {
      sn_colsum_kernel<<<total_blocks, kThreadsWn, 0, stream>>>(table_dev, n_desc);
      sn_normalize_v_kernel<<<n_desc, kThreadsWn, 0, stream>>>(table_dev);
    }
    sn_rowdot_kernel<<<total_rows, kThreadsWn, 0, stream>>>(table_dev, n_desc, training);
    sn_sigma_kernel<<<n_desc, kThreadsWn, 0, stream>>>(table_dev, training);
    sn_pack_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc);
  } else {
    sn_bwd_rowdot_kernel<<<total_rows, kThreadsWn, 0, stream>>>(table_dev, n_desc);
    sn_bwd_coef_kernel<<<n_desc, kThreadsWn, 0, stream>>>(table_dev);
    sn_bwd_apply_kernel<<<total_rows, kThreadsWn, smem, stream>>>(table_dev, n_desc);
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(wnpack)

}  // namespace xva
