// Length regulator and per-token averaging: the integer index path of FastPitch.
//
//   regulate_len   reference: python/fastpitch1_1/fastpitch/model.py:59-79
//   average_pitch  reference: python/fastpitch1_1/fastpitch/model.py:82-100
//
// The reference expands encoder rows to frames with a dense one-hot [B,Tm,Tt] matmul. Here the same map is an
// int32 prefix sum over the rounded durations plus a coalesced row gather (forward) / contiguous-segment sum
// (backward). Both are HBM-bound: algorithmic traffic = B*Tm*C*(read+write)*4 bytes.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

constexpr int kScanThreads = 256;

// One block per utterance: reps[j] = trunc(d[j]*pace + 0.5), cum[b, 0..Tt] = exclusive prefix sum, dec_len = total.
__global__ void __launch_bounds__(kScanThreads)
duration_scan_kernel(const float* __restrict__ durs, int Tt, float pace, int mel_max_len,
                     int* __restrict__ cum, int* __restrict__ dec_lens) {
  const int b = blockIdx.x;
  const float* d = durs + static_cast<long>(b) * Tt;
  int* c = cum + static_cast<long>(b) * (Tt + 1);
  const int per = (Tt + kScanThreads - 1) / kScanThreads;
  const int beg = threadIdx.x * per;
  const int end = min(beg + per, Tt);

  int local = 0;
  for (int j = beg; j < end; ++j) {
    // two separately rounded fp32 ops, exactly like `(durations.float() * pace + 0.5).long()`; no FMA contraction
    local += __float2int_rz(__fadd_rn(__fmul_rn(d[j], pace), 0.5f));
  }
  // block-wide exclusive scan of the per-thread totals
  __shared__ int warp_tot[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kScanThreads / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    if (lane < kScanThreads / 32) warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? warp_tot[warp - 1] : 0);
  for (int j = beg; j < end; ++j) {
    c[j] = run;
    run += __float2int_rz(__fadd_rn(__fmul_rn(d[j], pace), 0.5f));
  }
  if (threadIdx.x == 0) {
    const int total = warp_tot[kScanThreads / 32 - 1];
    c[Tt] = total;
    dec_lens[b] = (mel_max_len >= 0 && total > mel_max_len) ? mel_max_len : total;
  }
}

// Forward gather. One warp per output frame; the utterance's prefix sums sit in shared memory and the
// token index is found by binary search (== searchsorted(cum[1:], t, right=True)).
// grid = (ceil(T_out / frames_per_block), B), block = 32 * frames_per_block.
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
regulate_gather_kernel(const float* __restrict__ enc, const int* __restrict__ cum, int Tt, int C, int T_out,
                       float* __restrict__ out, int* __restrict__ idx_out) {
  extern __shared__ int s_cum[];
  const int b = blockIdx.y;
  const int* c = cum + static_cast<long>(b) * (Tt + 1);
  for (int i = threadIdx.x; i <= Tt; i += blockDim.x) s_cum[i] = c[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kWarps + warp;
  if (t >= T_out) return;
  const int total = s_cum[Tt];
  int j = -1;
  if (t < total) {
    // largest j with cum[j] <= t  (cum is non-decreasing; zero-length tokens are skipped automatically
    // because the search takes the LAST j whose start is <= t and cum[j+1] > t)
    int lo = 0, hi = Tt;  // invariant: cum[lo] <= t < cum[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_cum[mid] <= t) lo = mid;
      else hi = mid;
    }
    j = lo;
  }
  float4* dst = reinterpret_cast<float4*>(out + (static_cast<long>(b) * T_out + t) * C);
  const int c4 = C >> 2;
  if (j >= 0) {
    const float4* src = reinterpret_cast<const float4*>(enc + (static_cast<long>(b) * Tt + j) * C);
    for (int i = lane; i < c4; i += 32) dst[i] = __ldg(src + i);
  } else {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < c4; i += 32) dst[i] = z;
  }
  if (idx_out && lane == 0) idx_out[static_cast<long>(b) * T_out + t] = j;
}

// Backward: d_enc[b,j,:] = sum of d_out[b,t,:] over the token's frames (contiguous segment, clipped to T_out).
// One block per (token, utterance); threads stride over channels (float4), loop over the segment.
__global__ void __launch_bounds__(128)
regulate_scatter_kernel(const float* __restrict__ dout, const int* __restrict__ cum, int Tt, int C, int T_out,
                        float* __restrict__ denc, int accumulate) {
  const int j = blockIdx.x, b = blockIdx.y;
  const int* c = cum + static_cast<long>(b) * (Tt + 1);
  const int t0 = min(c[j], T_out), t1 = min(c[j + 1], T_out);
  const int c4 = C >> 2;
  float4* dst = reinterpret_cast<float4*>(denc + (static_cast<long>(b) * Tt + j) * C);
  for (int i = threadIdx.x; i < c4; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = t0; t < t1; ++t) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dout + (static_cast<long>(b) * T_out + t) * C) + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    if (accumulate) {
      float4 o = dst[i];
      acc.x += o.x;
      acc.y += o.y;
      acc.z += o.z;
      acc.w += o.w;
    }
    dst[i] = acc;
  }
}

// average_pitch: one warp per (utterance, formant, token). Mean of the non-zero frame values, 0 if none.
__global__ void __launch_bounds__(128)
average_pitch_kernel(const float* __restrict__ pitch, const float* __restrict__ durs, int F, int Tm, int Tt,
                     float* __restrict__ out, int log1p_out) {
  extern __shared__ int s_cum[];
  const int b = blockIdx.y;
  // cumsum(durs).long(): the running sum is accumulated in fp32 and truncated, like torch.cumsum(...).long()
  if (threadIdx.x == 0) {
    float run = 0.0f;
    s_cum[0] = 0;
    for (int j = 0; j < Tt; ++j) {
      run += durs[static_cast<long>(b) * Tt + j];
      s_cum[j + 1] = static_cast<int>(run);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;  // (f, j) flattened
  if (item >= F * Tt) return;
  const int f = item / Tt, j = item - f * Tt;
  const int t0 = min(s_cum[j], Tm), t1 = min(s_cum[j + 1], Tm);
  const float* p = pitch + (static_cast<long>(b) * F + f) * Tm;
  float sum = 0.0f;
  int cnt = 0;
  for (int t = t0 + lane; t < t1; t += 32) {
    const float v = __ldg(p + t);
    sum += v;
    cnt += (v != 0.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    const float mean = cnt > 0 ? sum / static_cast<float>(cnt) : 0.0f;
    // energy target: torch.log(1.0 + average_pitch(...)), model.py:415
    out[(static_cast<long>(b) * F + f) * Tt + j] = log1p_out ? logf(1.0f + mean) : mean;
  }
}

}  // namespace

int duration_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int* cum, int* dec_lens,
                  cudaStream_t stream) {
  XVA_CHECK_ARG(B >= 1 && Tt >= 1, "duration_scan: B=%d Tt=%d", B, Tt);
  duration_scan_kernel<<<B, kScanThreads, 0, stream>>>(durs, Tt, pace, mel_max_len, cum, dec_lens);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int regulate_gather(const float* enc, const int* cum, int B, int Tt, int C, int T_out, float* out, int* idx_out,
                    cudaStream_t stream) {
  XVA_CHECK_ARG(C % 4 == 0, "regulate_len: C=%d must be a multiple of 4", C);
  XVA_CHECK_ARG((Tt + 1) * 4 <= 200 * 1024, "regulate_len: Tt=%d too long", Tt);
  if (T_out == 0) return XVA_OK;
  constexpr int kWarps = 8;
  dim3 grid(ceil_div(T_out, kWarps), B);
  const int smem = (Tt + 1) * 4;
  if (smem > 48 * 1024)
    XVA_CHECK_CUDA(cudaFuncSetAttribute(regulate_gather_kernel<kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  regulate_gather_kernel<kWarps><<<grid, kWarps * 32, smem, stream>>>(enc, cum, Tt, C, T_out, out, idx_out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int regulate_scatter(const float* dout, const int* cum, int B, int Tt, int C, int T_out, float* denc, int accumulate,
                     cudaStream_t stream) {
  XVA_CHECK_ARG(C % 4 == 0, "regulate_len bwd: C=%d must be a multiple of 4", C);
  dim3 grid(Tt, B);
  regulate_scatter_kernel<<<grid, 128, 0, stream>>>(dout, cum, Tt, C, T_out, denc, accumulate);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, int log1p_out,
                  cudaStream_t stream) {
  XVA_CHECK_ARG((Tt + 1) * 4 <= 48 * 1024, "average_pitch: Tt=%d too long", Tt);
  dim3 grid(ceil_div(F * Tt, 4), B);
  average_pitch_kernel<<<grid, 128, (Tt + 1) * 4, stream>>>(pitch, durs, F, Tm, Tt, out, log1p_out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

}  // namespace xva
