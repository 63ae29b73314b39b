// Row-wise HBM-bound kernels around the tap-GEMM: masked softmax (+dropout) forward/backward, LayerNorm backward,
// column sums (bias gradients), token embedding + sinusoidal positions, the 1->C k3 "scalar" conv used by
// pitch_emb / energy_emb, and the C->1 projection of the temporal predictors.
//
// Layout everywhere: [rows, C] fp32 with C contiguous; one warp per row, lanes stride over columns so every
// global access is a coalesced 128-byte line; reductions are warp shuffles; cross-row reductions (dgamma, dbeta,
// dbias, dW of the tiny convs) are accumulated per thread over a grid-stride loop of rows, combined through shared
// memory once per block, then one fp32 atomicAdd per column per block.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxColsPerLane = 32;  // rows up to 1024 columns are held in registers (softmax)
constexpr uint64_t kSeedStep = 0xA24BAED4963EE407ull;  // seed += *seed_dev * kSeedStep (same constant as the tap-GEMM)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}


// Sum per-thread column accumulators (thread owns columns lane + 32*i) over the warps of the block and add the
// block total to dst[0..C) with one atomic per column. `red` is reused across calls (leading __syncthreads()).
template <int kColsPerLane>
__device__ __forceinline__ void block_colsum_atomic(const float (&acc)[kColsPerLane],
                                                    float (*red)[32 * kColsPerLane + 1], int C, float* dst, int stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kColsPerLane; ++i) red[warp][lane + 32 * i] = acc[i];
  __syncthreads();
  if (dst == nullptr) return;
  for (int n = threadIdx.x; n < C; n += blockDim.x) {
    float a = 0.0f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) a += red[w][n];
    atomicAdd(dst + static_cast<long>(n) * stride, a);
  }
}

// ------------------------------------------------------------------------------------------------ softmax
// transformer.py:120-127: masked_fill(key >= len, -inf) -> softmax(dim=2) -> dropout.
// s [Z,R,N] holds alpha*q.k (written by the tap-GEMM); p_out gets softmax, pd_out (optional) softmax*dropout.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_fwd_kernel(const float* __restrict__ s, const int* __restrict__ lens, int R, int N, int ld, long rows,
                   float* __restrict__ p_out, float* __restrict__ pd_out, uint64_t seed,
                   const uint64_t* __restrict__ seed_dev, uint32_t thresh, float inv_keep) {
  const int lane = threadIdx.x & 31;
  seed += seed_dev ? __ldg(seed_dev) * kSeedStep : 0ull;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int z = static_cast<int>(row / R);
  const int nk = lens ? min(lens[z], N) : N;
  const float* src = s + row * ld;
  float v[kMaxColsPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxColsPerLane; ++i) {
    const int n = lane + 32 * i;
    v[i] = (n < nk) ? src[n] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxColsPerLane; ++i) {
    v[i] = (lane + 32 * i < nk) ? expf(v[i] - mx) : 0.0f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kMaxColsPerLane; ++i) {
    const int n = lane + 32 * i;
    if (n < ld) {  // pad columns [N, ld) are written as zero: the MN-major consumers multiply them
      const float p = tf32_rn(v[i] * inv);  // both outputs are GEMM operands (P*V, dV = P^T dO)
      p_out[row * ld + n] = p;
      if (pd_out) pd_out[row * ld + n] = tf32_rn(p * dropout_scale(seed, static_cast<uint64_t>(row) * ld + n, thresh, inv_keep));
    }
  }
}

// ds = p * (g - sum_n g*p) with g = dpd * dropout_scale   (in place on dpd)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dpd, int N, int ld, long rows, float alpha, uint64_t seed,
                   const uint64_t* __restrict__ seed_dev, uint32_t thresh, float inv_keep) {
  const int lane = threadIdx.x & 31;
  seed += seed_dev ? __ldg(seed_dev) * kSeedStep : 0ull;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float pv[kMaxColsPerLane], gv[kMaxColsPerLane];
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxColsPerLane; ++i) {
    const int n = lane + 32 * i;
    if (n < N) {
      pv[i] = p[row * ld + n];
      gv[i] = dpd[row * ld + n] * dropout_scale(seed, static_cast<uint64_t>(row) * ld + n, thresh, inv_keep);
    } else {
      pv[i] = 0.0f;
      gv[i] = 0.0f;
    }
    dot += pv[i] * gv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < kMaxColsPerLane; ++i) {
    const int n = lane + 32 * i;
    if (n < ld) dpd[row * ld + n] = (n < N) ? tf32_rn(alpha * pv[i] * (gv[i] - dot)) : 0.0f;  // operand of dQ, dK
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// y = (LN(x)*gamma + beta) [* post-dropout] * rowmask ; given dy and the saved pre-LN x, mean, rstd:
//   dx = rstd * (g - mean_n(g) - xhat * mean_n(g*xhat)),  g = dy_eff * gamma
//   dgamma += sum_rows dy_eff * xhat ; dbeta += sum_rows dy_eff
// dx_drop (optional) = dx * pre-dropout scale: the gradient of the GEMM branch when dropout sat between the GEMM
// and the residual add (transformer.py:51,139); dbias (optional) += column sums of that branch gradient.
template <int kColsPerLane>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const int* __restrict__ lens,
                     int R, int C, long rows, float* __restrict__ dx, float* __restrict__ dx_drop,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                     uint64_t seed_post, uint32_t thresh_post, float inv_keep_post, uint64_t seed_pre,
                     uint32_t thresh_pre, float inv_keep_pre, const uint64_t* __restrict__ seed_dev, int relu_gate) {
  __shared__ float red[kWarpsPerBlock][32 * kColsPerLane + 1];
  if (seed_dev) {
    const uint64_t add = __ldg(seed_dev) * kSeedStep;
    seed_post += add;
    seed_pre += add;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[kColsPerLane], acc_g[kColsPerLane], acc_b[kColsPerLane], acc_bias[kColsPerLane];
#pragma unroll
  for (int i = 0; i < kColsPerLane; ++i) {
    const int n = lane + 32 * i;
    gm[i] = n < C ? gamma[n] : 0.0f;
    acc_g[i] = acc_b[i] = acc_bias[i] = 0.0f;
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  for (long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long>(gridDim.x) * kWarpsPerBlock) {
    const int z = static_cast<int>(row / R), r = static_cast<int>(row - static_cast<long>(z) * R);
    const bool live = (lens == nullptr) || (r < lens[z]);
    const float mu = mean[row], rs = rstd[row];
    float xh[kColsPerLane], g[kColsPerLane];
    float s1 = 0.0f, s2 = 0.0f;
    uint32_t pos = 0u;
#pragma unroll
    for (int i = 0; i < kColsPerLane; ++i) {
      const int n = lane + 32 * i;
      float d = 0.0f;
      xh[i] = 0.0f;
      if (n < C && live) {
        d = dy[row * C + n] * dropout_scale(seed_post, static_cast<uint64_t>(row) * C + n, thresh_post, inv_keep_post);
        const float xv = x[row * C + n];
        xh[i] = (xv - mu) * rs;
        pos |= (xv > 0.0f ? 1u : 0u) << i;
      }
      acc_g[i] += d * xh[i];
      acc_b[i] += d;
      g[i] = d * gm[i];
      s1 += g[i];
      s2 += g[i] * xh[i];
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < kColsPerLane; ++i) {
      const int n = lane + 32 * i;
      if (n < C) {
        float v = rs * (g[i] - s1 - xh[i] * s2);
        // ConvReLUNorm (common/layers.py:94-97): x is the ReLU output, so the gradient passes only where x > 0
        if (relu_gate && !((pos >> i) & 1u)) v = 0.0f;
        // the tensor that feeds the dgrad / wgrad GEMMs (dx_drop if there is one, else dx) is stored tf32-rounded
        if (dx_drop) {
          dx[row * C + n] = v;
          const float vd = tf32_rn(v * dropout_scale(seed_pre, static_cast<uint64_t>(row) * C + n, thresh_pre, inv_keep_pre));
          dx_drop[row * C + n] = vd;
          acc_bias[i] += vd;
        } else {
          v = tf32_rn(v);
          dx[row * C + n] = v;
          acc_bias[i] += v;
        }
      }
    }
  }
  block_colsum_atomic<kColsPerLane>(acc_g, red, C, dgamma, 1);
  block_colsum_atomic<kColsPerLane>(acc_b, red, C, dbeta, 1);
  block_colsum_atomic<kColsPerLane>(acc_bias, red, C, dbias, 1);
}

// ------------------------------------------------------------------------------------------------ float4 variants
// Same math as the scalar kernels above with 16-byte accesses (lane l owns float4 l + 32 i of its row) and one
// dropout hash per float4. Used whenever the row stride is a multiple of 4 and the pointers are 16-byte aligned.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int kV4>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_fwd_v4_kernel(const float* __restrict__ s, const int* __restrict__ lens, int R, int N, int ld, long rows,
                      float* __restrict__ p_out, float* __restrict__ pd_out, uint64_t seed,
                      const uint64_t* __restrict__ seed_dev, uint32_t thresh, float inv_keep) {
  const int lane = threadIdx.x & 31;
  seed += seed_dev ? __ldg(seed_dev) * kSeedStep : 0ull;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int z = static_cast<int>(row / R);
  const int nk = lens ? min(lens[z], N) : N;
  const float* src = s + row * ld;
  float4 v[kV4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (n < nk) {  // columns >= nk (masked keys, and the never-written pad columns) are not used
      const float4 t = ld4(src + n);
      v[i].x = t.x;
      if (n + 1 < nk) v[i].y = t.y;
      if (n + 2 < nk) v[i].z = t.z;
      if (n + 3 < nk) v[i].w = t.w;
    }
    mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {  // expf(-inf) = 0 for the masked columns
    v[i] = make_float4(expf(v[i].x - mx), expf(v[i].y - mx), expf(v[i].z - mx), expf(v[i].w - mx));
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    if (n < ld) {  // pad columns [N, ld) are written as zero: the MN-major consumers multiply them
      const float4 pr = tf32_rn4(make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
      st4(p_out + row * ld + n, pr);
      if (pd_out) {
        const float4 d = dropout_scale4(seed, static_cast<uint64_t>(row) * ld + n, thresh, inv_keep);
        st4(pd_out + row * ld + n, tf32_rn4(make_float4(pr.x * d.x, pr.y * d.y, pr.z * d.z, pr.w * d.w)));
      }
    }
  }
}

template <int kV4>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
softmax_bwd_v4_kernel(const float* __restrict__ p, float* __restrict__ dpd, int N, int ld, long rows, float alpha,
                      uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t thresh, float inv_keep) {
  const int lane = threadIdx.x & 31;
  seed += seed_dev ? __ldg(seed_dev) * kSeedStep : 0ull;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 pv[kV4], gv[kV4];
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    pv[i] = gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {  // p is exactly zero in the pad columns; dpd's pad columns hold nothing
      pv[i] = ld4(p + row * ld + n);
      const float4 g = ld4(dpd + row * ld + n);
      const float4 d = dropout_scale4(seed, static_cast<uint64_t>(row) * ld + n, thresh, inv_keep);
      gv[i].x = g.x * d.x;
      if (n + 1 < N) gv[i].y = g.y * d.y;
      if (n + 2 < N) gv[i].z = g.z * d.z;
      if (n + 3 < N) gv[i].w = g.w * d.w;
    }
    dot += (pv[i].x * gv[i].x + pv[i].y * gv[i].y) + (pv[i].z * gv[i].z + pv[i].w * gv[i].w);
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    if (n < ld)  // operand of dQ, dK; pv = 0 in the pad columns makes them zero
      st4(dpd + row * ld + n, tf32_rn4(make_float4(alpha * pv[i].x * (gv[i].x - dot), alpha * pv[i].y * (gv[i].y - dot),
                                                   alpha * pv[i].z * (gv[i].z - dot), alpha * pv[i].w * (gv[i].w - dot))));
  }
}

// column = 4 * (lane + 32 i) + e  <->  acc[4 i + e]
template <int kV4>
__device__ __forceinline__ void block_colsum_atomic_v4(const float (&acc)[4 * kV4], float (*red)[128 * kV4 + 4], int C,
                                                       float* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kV4; ++i)
    *reinterpret_cast<float4*>(&red[warp][4 * (lane + 32 * i)]) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
  __syncthreads();
  if (dst == nullptr) return;
  for (int n = threadIdx.x; n < C; n += blockDim.x) {
    float a = 0.0f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) a += red[w][n];
    atomicAdd(dst + n, a);
  }
}

template <int kV4>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_bwd_v4_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma, const int* __restrict__ lens,
                        int R, int C, long rows, float* __restrict__ dx, float* __restrict__ dx_drop,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                        uint64_t seed_post, uint32_t thresh_post, float inv_keep_post, uint64_t seed_pre,
                        uint32_t thresh_pre, float inv_keep_pre, const uint64_t* __restrict__ seed_dev, int relu_gate) {
  __shared__ __align__(16) float red[kWarpsPerBlock][128 * kV4 + 4];
  if (seed_dev) {
    const uint64_t add = __ldg(seed_dev) * kSeedStep;
    seed_post += add;
    seed_pre += add;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 gm[kV4];
  float acc_g[4 * kV4], acc_b[4 * kV4], acc_bias[4 * kV4];
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    gm[i] = n < C ? ld4(gamma + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 4; ++e) acc_g[4 * i + e] = acc_b[4 * i + e] = acc_bias[4 * i + e] = 0.0f;
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  for (long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long>(gridDim.x) * kWarpsPerBlock) {
    const int z = static_cast<int>(row / R), r = static_cast<int>(row - static_cast<long>(z) * R);
    const bool live = (lens == nullptr) || (r < lens[z]);
    const float mu = mean[row], rs = rstd[row];
    float4 xh[kV4], g[kV4];
    float s1 = 0.0f, s2 = 0.0f;
    uint32_t pos = 0u;
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      const int n = 4 * (lane + 32 * i);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      xh[i] = d;
      if (n < C && live) {
        d = ld4(dy + row * C + n);
        const float4 sc = dropout_scale4(seed_post, static_cast<uint64_t>(row) * C + n, thresh_post, inv_keep_post);
        d = make_float4(d.x * sc.x, d.y * sc.y, d.z * sc.z, d.w * sc.w);
        const float4 xv = ld4(x + row * C + n);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        pos |= ((xv.x > 0.f ? 1u : 0u) | (xv.y > 0.f ? 2u : 0u) | (xv.z > 0.f ? 4u : 0u) | (xv.w > 0.f ? 8u : 0u)) << (4 * i);
      }
      acc_g[4 * i] += d.x * xh[i].x; acc_g[4 * i + 1] += d.y * xh[i].y; acc_g[4 * i + 2] += d.z * xh[i].z; acc_g[4 * i + 3] += d.w * xh[i].w;
      acc_b[4 * i] += d.x; acc_b[4 * i + 1] += d.y; acc_b[4 * i + 2] += d.z; acc_b[4 * i + 3] += d.w;
      g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      const int n = 4 * (lane + 32 * i);
      if (n < C) {
        float4 v = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                               rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
        if (relu_gate) {  // ConvReLUNorm (common/layers.py:94-97): the gradient passes only where the ReLU output > 0
          const uint32_t m = pos >> (4 * i);
          if (!(m & 1u)) v.x = 0.f;
          if (!(m & 2u)) v.y = 0.f;
          if (!(m & 4u)) v.z = 0.f;
          if (!(m & 8u)) v.w = 0.f;
        }
        // the tensor that feeds the dgrad / wgrad GEMMs (dx_drop if there is one, else dx) is stored tf32-rounded
        if (dx_drop) {
          st4(dx + row * C + n, v);
          const float4 sc = dropout_scale4(seed_pre, static_cast<uint64_t>(row) * C + n, thresh_pre, inv_keep_pre);
          v = tf32_rn4(make_float4(v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w));
          st4(dx_drop + row * C + n, v);
        } else {
          v = tf32_rn4(v);
          st4(dx + row * C + n, v);
        }
        acc_bias[4 * i] += v.x; acc_bias[4 * i + 1] += v.y; acc_bias[4 * i + 2] += v.z; acc_bias[4 * i + 3] += v.w;
      }
    }
  }
  block_colsum_atomic_v4<kV4>(acc_g, red, C, dgamma);
  block_colsum_atomic_v4<kV4>(acc_b, red, C, dbeta);
  block_colsum_atomic_v4<kV4>(acc_bias, red, C, dbias);
}

// LayerNorm forward on a stored pre-LN tensor (the un-fused path of the FFT blocks: the producing GEMM keeps narrow,
// double-buffered n tiles and this pass costs one read + one write at HBM speed).
//   y = ((x - mean) * rstd * gamma + beta) * (r < lens[z]), tf32-rounded (next GEMM operand); mean / rstd saved.
template <int kV4>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_fwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        const int* __restrict__ lens, int R, int C, long rows, float eps, float* __restrict__ y,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int z = static_cast<int>(row / R), r = static_cast<int>(row - static_cast<long>(z) * R);
  const bool live = (lens == nullptr) || (r < lens[z]);
  float4 v[kV4];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    v[i] = n < C ? ld4(x + row * C + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) / static_cast<float>(C);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    if (n < C) {
      const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mu;
    if (rstd_out) rstd_out[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const int n = 4 * (lane + 32 * i);
    if (n < C) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const float4 gmv = ld4(gamma + n), bt = ld4(beta + n);
        o = tf32_rn4(make_float4((v[i].x - mu) * rs * gmv.x + bt.x, (v[i].y - mu) * rs * gmv.y + bt.y,
                                 (v[i].z - mu) * rs * gmv.z + bt.z, (v[i].w - mu) * rs * gmv.w + bt.w));
      }
      st4(y + row * C + n, o);
    }
  }
}

// out[n] += sum_rows x[row, n], 16-byte loads: a block owns a 128-column strip, its 8 warps take rows 8 apart, four
// rows in flight per lane.
__global__ void __launch_bounds__(256)
colsum_v4_kernel(const float* __restrict__ x, long rows, int C, long ld, float* __restrict__ out) {
  __shared__ __align__(16) float red[8][132];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + 4 * lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < C) {
    const long step = static_cast<long>(gridDim.y) * 8;
    long r = static_cast<long>(blockIdx.y) * 8 + warp;
    for (; r + 3 * step < rows; r += 4 * step) {
      const float4 a = ld4(x + r * ld + n), b = ld4(x + (r + step) * ld + n), c = ld4(x + (r + 2 * step) * ld + n),
                   d = ld4(x + (r + 3 * step) * ld + n);
      acc.x += (a.x + b.x) + (c.x + d.x);
      acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z);
      acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < rows; r += step) {
      const float4 a = ld4(x + r * ld + n);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  *reinterpret_cast<float4*>(&red[warp][4 * lane]) = acc;
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < C) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    atomicAdd(out + blockIdx.x * 128 + threadIdx.x, sum);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] += sum_rows x[row, n]   (bias gradients). Thread per column within a 32-column strip, rows strided over
// blockIdx.y; coalesced 128-byte reads.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long rows, int C, long ld, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float acc = 0.0f;
  if (n < C)
    for (long r = static_cast<long>(blockIdx.y) * 8 + warp; r < rows; r += static_cast<long>(gridDim.y) * 8)
      acc += x[r * ld + n];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < C) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][lane];
    atomicAdd(out + n, s);
  }
}

// ------------------------------------------------------------------------------------------------ embedding
// out[b,t,:] = (tokens ? emb[tok] : in[b,t,:]) + pos(t) * live     transformer.py:212-227 (conditioning = 0)
//   pos(t)[c] = sin(t*f[c]) for c < C/2, cos(t*f[c-C/2]) otherwise (transformer.py:28-35)
//   live = tok != 0 (encoder) or t < lens[b] (decoder)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_pos_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ in,
                 const int* __restrict__ lens, const float* __restrict__ inv_freq, int T, int C, long rows,
                 float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long>(b) * T);
  const float* src;
  bool live;
  if (tokens) {
    const long long tok = tokens[row];
    src = emb + tok * C;
    live = tok != 0;
  } else {
    src = in + row * C;
    live = t < lens[b];
  }
  const int half = C >> 1;
  const float tf = static_cast<float>(t);
  for (int c = lane; c < C; c += 32) {
    float v = src[c];
    if (live && inv_freq != nullptr) {
      const float ang = tf * inv_freq[c < half ? c : c - half];
      v += (c < half) ? sinf(ang) : cosf(ang);
    }
    out[row * C + c] = tf32_rn(v);  // first GEMM operand of the FFT stack
  }
}

// d_emb[tok] += dout[row]  (tok != padding 0)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_bwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ dout, int C, long rows,
                 float* __restrict__ demb) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long tok = tokens[row];
  if (tok == 0) return;
  for (int c = lane; c < C; c += 32) atomicAdd(demb + tok * C + c, dout[row * C + c]);
}

// ------------------------------------------------------------------------------------------------ scalar conv
// io[b,t,c] += bias[c] + sum_j w[c,j] * x[b, t+j-1]     pitch_emb / energy_emb: Conv1d(1, C, 3, padding=1)
// (model.py:403-404,417-418). x [B,T]; w [C,3].
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scalar_conv_add_kernel(float* __restrict__ io, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, const int* __restrict__ lens, int T, int C, long rows) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int t = static_cast<int>(row % T);
  if (lens && t >= lens[row / T]) return;  // padded token rows stay untouched (zero): nothing downstream reads them
  const float x0 = t > 0 ? x[row - 1] : 0.0f, x1 = x[row], x2 = t + 1 < T ? x[row + 1] : 0.0f;
  for (int c = lane; c < C; c += 32)  // the sum feeds the predictor convs and the length regulator -> decoder GEMMs
    io[row * C + c] = tf32_rn(io[row * C + c] + bias[c] + w[c * 3] * x0 + w[c * 3 + 1] * x1 + w[c * 3 + 2] * x2);
}

// dw[c,j] += sum_rows dout[row,c] * x[row+j-1] ; dbias[c] += sum_rows dout[row,c]
template <int kColsPerLane>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scalar_conv_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, int T, int C, long rows,
                       float* __restrict__ dw, float* __restrict__ dbias) {
  __shared__ float red[kWarpsPerBlock][32 * kColsPerLane + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a[4][kColsPerLane];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < kColsPerLane; ++i) a[q][i] = 0.0f;
  for (long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long>(gridDim.x) * kWarpsPerBlock) {
    const int t = static_cast<int>(row % T);
    const float x0 = t > 0 ? x[row - 1] : 0.0f, x1 = x[row], x2 = t + 1 < T ? x[row + 1] : 0.0f;
#pragma unroll
    for (int i = 0; i < kColsPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float d = dout[row * C + c];
        a[0][i] += d * x0;
        a[1][i] += d * x1;
        a[2][i] += d * x2;
        a[3][i] += d;
      }
    }
  }
  block_colsum_atomic<kColsPerLane>(a[0], red, C, dw, 3);
  block_colsum_atomic<kColsPerLane>(a[1], red, C, dw + 1, 3);
  block_colsum_atomic<kColsPerLane>(a[2], red, C, dw + 2, 3);
  block_colsum_atomic<kColsPerLane>(a[3], red, C, dbias, 1);
}

// ------------------------------------------------------------------------------------------------ C -> 1 projection
// out[row] = (dot(x[row,:], w) + b) * live      TemporalPredictor.fc + mask, model.py:121
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rowdot_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  const int* __restrict__ lens, int R, int C, long rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.0f;
  for (int c = lane; c < C; c += 32) acc += x[row * C + c] * w[c];
  acc = warp_sum(acc);
  if (lane == 0) {
    const int z = static_cast<int>(row / R), r = static_cast<int>(row - static_cast<long>(z) * R);
    out[row] = (lens == nullptr || r < lens[z]) ? acc + bias[0] : 0.0f;
  }
}

// dx[row,:] = g*w ; dw += sum g*x[row,:] ; db += sum g,   g = dout[row]*live
template <int kColsPerLane>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rowdot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w,
                  const int* __restrict__ lens, int R, int C, long rows, float* __restrict__ dx,
                  float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[kWarpsPerBlock][32 * kColsPerLane + 1];
  __shared__ float redb[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a[kColsPerLane];
  float ab = 0.0f;
#pragma unroll
  for (int i = 0; i < kColsPerLane; ++i) a[i] = 0.0f;
  for (long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + warp; row < rows;
       row += static_cast<long>(gridDim.x) * kWarpsPerBlock) {
    const int z = static_cast<int>(row / R), r = static_cast<int>(row - static_cast<long>(z) * R);
    const float g = (lens == nullptr || r < lens[z]) ? dout[row] : 0.0f;
    ab += g;
#pragma unroll
    for (int i = 0; i < kColsPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        a[i] += g * x[row * C + c];
        dx[row * C + c] = g * w[c];
      }
    }
  }
  if (lane == 0) redb[warp] = ab;
  block_colsum_atomic<kColsPerLane>(a, red, C, dw, 1);
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int wv = 0; wv < kWarpsPerBlock; ++wv) s += redb[wv];
    atomicAdd(db, s);
  }
}

// out[row] = dot(a[row, :], b[row, :])   (the softmax-backward row term dO . O, see XVA_GEMM_SOFTMAX_BWD)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rowdot2_kernel(const float* __restrict__ a, const float* __restrict__ b, long rows, int C, long a_ld, long b_ld,
               float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.0f;
  for (int c = lane; c < C; c += 32) acc = fmaf(a[row * a_ld + c], b[row * b_ld + c], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

__global__ void __launch_bounds__(256)
round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long n) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    dst[i] = tf32_rn(src[i]);
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }

inline void drop_consts(float p, uint32_t* thresh, float* inv_keep) {
  if (p > 0.0f) {
    *thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
    *inv_keep = 1.0f / (1.0f - p);
  } else {
    *thresh = 0;
    *inv_keep = 1.0f;
  }
}

inline int row_blocks(long rows) { return static_cast<int>(ceil_div_l(rows, kWarpsPerBlock)); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }  // nullptr counts as aligned

}  // namespace

int softmax_fwd(const float* s, const int* lens, int Z, int R, int N, int ld, float* p_out, float* pd_out, float drop_p,
                uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream) {
  if (ld <= 0) ld = N;
  XVA_CHECK_ARG(N >= 1 && ld >= N && ld <= 32 * kMaxColsPerLane, "softmax: N=%d ld=%d out of range (max %d)", N, ld,
                32 * kMaxColsPerLane);
  XVA_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "softmax: dropout p=%f", drop_p);
  uint32_t th;
  float ik;
  drop_consts(pd_out ? drop_p : 0.0f, &th, &ik);
  const long rows = static_cast<long>(Z) * R;
  if (ld % 4 == 0 && aligned16(s) && aligned16(p_out) && aligned16(pd_out)) {
    if (ld <= 512) softmax_fwd_v4_kernel<4><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(s, lens, R, N, ld, rows, p_out, pd_out, seed, seed_dev, th, ik);
    else softmax_fwd_v4_kernel<8><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(s, lens, R, N, ld, rows, p_out, pd_out, seed, seed_dev, th, ik);
  } else {
    softmax_fwd_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(s, lens, R, N, ld, rows, p_out, pd_out, seed, seed_dev, th, ik);
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int softmax_bwd(const float* p, float* dpd, int Z, int R, int N, int ld, float alpha, float drop_p, uint64_t seed,
                const uint64_t* seed_dev, cudaStream_t stream) {
  if (ld <= 0) ld = N;
  XVA_CHECK_ARG(N >= 1 && ld >= N && ld <= 32 * kMaxColsPerLane, "softmax bwd: N=%d ld=%d out of range", N, ld);
  uint32_t th;
  float ik;
  drop_consts(drop_p, &th, &ik);
  const long rows = static_cast<long>(Z) * R;
  if (ld % 4 == 0 && aligned16(p) && aligned16(dpd)) {
    if (ld <= 512) softmax_bwd_v4_kernel<4><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(p, dpd, N, ld, rows, alpha, seed, seed_dev, th, ik);
    else softmax_bwd_v4_kernel<8><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(p, dpd, N, ld, rows, alpha, seed, seed_dev, th, ik);
  } else {
    softmax_bwd_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(p, dpd, N, ld, rows, alpha, seed, seed_dev, th, ik);
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const int* lens, int Z, int R, int C, float* dx, float* dx_drop, float* dgamma, float* dbeta,
                  float* dbias, float drop_post_p, uint64_t seed_post, float drop_pre_p, uint64_t seed_pre,
                  const uint64_t* seed_dev, int relu_gate, cudaStream_t stream) {
  XVA_CHECK_ARG(C >= 1 && C <= 1024, "layernorm bwd: C=%d (max 1024)", C);
  uint32_t th_post, th_pre;
  float ik_post, ik_pre;
  drop_consts(drop_post_p, &th_post, &ik_post);
  drop_consts(dx_drop ? drop_pre_p : 0.0f, &th_pre, &ik_pre);
  const long rows = static_cast<long>(Z) * R;
  int grid = static_cast<int>(ceil_div_l(rows, kWarpsPerBlock * 4));
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  if (grid < 1) grid = 1;
#define XVA_LN_BWD(CPL)                                                                                           \
  layernorm_bwd_kernel<CPL><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dy, x, mean, rstd, gamma, lens, R, C, rows, \
                                                                      dx, dx_drop, dgamma, dbeta, dbias, seed_post, \
                                                                      th_post, ik_post, seed_pre, th_pre, ik_pre, \
                                                                      seed_dev, relu_gate)
#define XVA_LN_BWD4(V4)                                                                                             \
  layernorm_bwd_v4_kernel<V4><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dy, x, mean, rstd, gamma, lens, R, C, rows, \
                                                                        dx, dx_drop, dgamma, dbeta, dbias, seed_post, \
                                                                        th_post, ik_post, seed_pre, th_pre, ik_pre, \
                                                                        seed_dev, relu_gate)
  if (C % 4 == 0 && aligned16(dy) && aligned16(x) && aligned16(gamma) && aligned16(dx) && aligned16(dx_drop)) {
    if (C <= 256) XVA_LN_BWD4(2);
    else if (C <= 384) XVA_LN_BWD4(3);
    else if (C <= 512) XVA_LN_BWD4(4);
    else XVA_LN_BWD4(8);  // 513..1024 channels: the xVAPitch pitch predictor's 708 / 780 (python/xvapitch/model.py:154-168)
  } else if (C <= 256) XVA_LN_BWD(8);
  else if (C <= 384) XVA_LN_BWD(12);
  else if (C <= 512) XVA_LN_BWD(16);
  else XVA_LN_BWD(32);
#undef XVA_LN_BWD4
#undef XVA_LN_BWD
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int layernorm_fwd(const float* x, const float* gamma, const float* beta, const int* lens, int Z, int R, int C, float eps,
                  float* y, float* mean, float* rstd, cudaStream_t stream) {
  XVA_CHECK_ARG(C >= 4 && C <= 1024 && C % 4 == 0, "layernorm fwd: C=%d (multiple of 4, max 1024)", C);
  XVA_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "layernorm fwd: pointers must be 16-byte aligned");
  const long rows = static_cast<long>(Z) * R;
  if (rows == 0) return XVA_OK;
  if (C <= 256) layernorm_fwd_v4_kernel<2><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(x, gamma, beta, lens, R, C, rows, eps, y, mean, rstd);
  else if (C <= 384) layernorm_fwd_v4_kernel<3><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(x, gamma, beta, lens, R, C, rows, eps, y, mean, rstd);
  else if (C <= 512) layernorm_fwd_v4_kernel<4><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(x, gamma, beta, lens, R, C, rows, eps, y, mean, rstd);
  else layernorm_fwd_v4_kernel<8><<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(x, gamma, beta, lens, R, C, rows, eps, y, mean, rstd);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int colsum(const float* x, long rows, int C, long ld, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(rows >= 0 && C >= 1, "colsum: rows=%ld C=%d", rows, C);
  if (rows == 0) return XVA_OK;
  if (C % 4 == 0 && ld % 4 == 0 && aligned16(x)) {
    const int strips4 = ceil_div(C, 128);
    int gy4 = static_cast<int>(ceil_div_l(rows, 8 * 16));
    const int cap4 = ceil_div(6 * num_sms(), strips4);
    if (gy4 > cap4) gy4 = cap4;
    if (gy4 < 1) gy4 = 1;
    colsum_v4_kernel<<<dim3(strips4, gy4), 256, 0, stream>>>(x, rows, C, ld, out);
    XVA_CHECK_LAUNCH();
    return XVA_OK;
  }
  int gy = static_cast<int>(ceil_div_l(rows, 8 * 16));
  const int strips = ceil_div(C, 32);
  const int cap = ceil_div(8 * num_sms(), strips);
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3(strips, gy), 256, 0, stream>>>(x, rows, C, ld, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int embed_pos(const long long* tokens, const float* emb, const float* in, const int* lens, const float* inv_freq,
              int B, int T, int C, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG((tokens && emb) || (in && lens), "embed_pos: need tokens+emb or in+lens");
  XVA_CHECK_ARG(C % 2 == 0, "embed_pos: C=%d must be even", C);
  const long rows = static_cast<long>(B) * T;
  embed_pos_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(tokens, emb, in, lens, inv_freq, T, C, rows, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int embed_bwd(const long long* tokens, const float* dout, int B, int T, int C, float* demb, cudaStream_t stream) {
  const long rows = static_cast<long>(B) * T;
  embed_bwd_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(tokens, dout, C, rows, demb);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int scalar_conv_add(float* io, const float* x, const float* w, const float* bias, const int* lens, int B, int T, int C,
                    cudaStream_t stream) {
  const long rows = static_cast<long>(B) * T;
  scalar_conv_add_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(io, x, w, bias, lens, T, C, rows);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int scalar_conv_bwd(const float* dout, const float* x, int B, int T, int C, float* dw, float* dbias,
                    cudaStream_t stream) {
  XVA_CHECK_ARG(C <= 512, "scalar_conv bwd: C=%d (max 512)", C);
  const long rows = static_cast<long>(B) * T;
  int grid = static_cast<int>(ceil_div_l(rows, kWarpsPerBlock * 8));
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  if (grid < 1) grid = 1;
  if (C <= 384) scalar_conv_bwd_kernel<12><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dout, x, T, C, rows, dw, dbias);
  else scalar_conv_bwd_kernel<16><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dout, x, T, C, rows, dw, dbias);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int rowdot_fwd(const float* x, const float* w, const float* bias, const int* lens, int Z, int R, int C, float* out,
               cudaStream_t stream) {
  const long rows = static_cast<long>(Z) * R;
  rowdot_fwd_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(x, w, bias, lens, R, C, rows, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int rowdot_bwd(const float* dout, const float* x, const float* w, const int* lens, int Z, int R, int C, float* dx,
               float* dw, float* db, cudaStream_t stream) {
  XVA_CHECK_ARG(C <= 512, "rowdot bwd: C=%d (max 512)", C);
  const long rows = static_cast<long>(Z) * R;
  int grid = static_cast<int>(ceil_div_l(rows, kWarpsPerBlock * 8));
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  if (grid < 1) grid = 1;
  if (C <= 256) rowdot_bwd_kernel<8><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dout, x, w, lens, R, C, rows, dx, dw, db);
  else rowdot_bwd_kernel<16><<<grid, kWarpsPerBlock * 32, 0, stream>>>(dout, x, w, lens, R, C, rows, dx, dw, db);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int rowdot2(const float* a, const float* b, long rows, int C, long a_ld, long b_ld, float* out, cudaStream_t stream) {
  if (rows <= 0) return XVA_OK;
  rowdot2_kernel<<<row_blocks(rows), kWarpsPerBlock * 32, 0, stream>>>(a, b, rows, C, a_ld, b_ld, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int round_tf32(const float* src, float* dst, long n, cudaStream_t stream) {
  if (n <= 0) return XVA_OK;
  long b = ceil_div_l(n, 256 * 8);
  if (b > 16L * num_sms()) b = 16L * num_sms();
  round_tf32_kernel<<<static_cast<int>(b), 256, 0, stream>>>(src, dst, n);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int counter_add(unsigned long long* counter, unsigned long long inc, cudaStream_t stream) {
  counter_add_kernel<<<1, 1, 0, stream>>>(counter, inc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(rowops)

}  // namespace xva
