// HBM-bound pieces of the mel-spectrogram extractor (hifigan/meldataset.py:217-240) around its two GEMMs.
//
// The STFT itself is a 4-tap tap-GEMM: with hop 256 and n_fft 1024 the frame f of the reflect-padded signal is rows
// f..f+3 of its [len/256, 256] view, so  spec[f, :] = sum_j view[f+j, :] @ Bw[j]^T  with Bw[j] = window * DFT basis
// columns 256j..256j+255 (no framing copy: the 4x overlap is four shifted TMA loads). The mel projection is a plain
// GEMM. What is left for this file: reflect padding, |.|, log-clamp, the L1 loss, and their gradients.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

inline int grid_for(long n) {
  long b = ceil_div_l(n, 256 * 4);
  const long cap = 16L * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// torch 'reflect' padding (edge sample not repeated): out[i] = y[reflect(i - pad)]
__device__ __forceinline__ long reflect_index(long t, long n) {
  if (t < 0) t = -t;
  if (t >= n) t = 2 * (n - 1) - t;
  return t;
}

__global__ void __launch_bounds__(256)
reflect_pad_fwd_kernel(const float* __restrict__ y, long n, int pad, long total, float* __restrict__ out) {
  const long np = n + 2L * pad;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / np, t = i - b * np;
    out[i] = tf32_rn(y[b * n + reflect_index(t - pad, n)]);  // operand of the STFT GEMM
  }
}

// dy[t] = dyp[t + pad] + mirrored contributions of the two padded margins
__global__ void __launch_bounds__(256)
reflect_pad_bwd_kernel(const float* __restrict__ dyp, long n, int pad, long total, float* __restrict__ dy) {
  const long np = n + 2L * pad;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / n, t = i - b * n;
    const float* row = dyp + b * np;
    float g = row[t + pad];
    if (t >= 1 && t <= pad) g += row[pad - t];                      // left margin: out[pad - t] = y[t]
    if (t <= n - 2 && t >= n - 1 - pad) g += row[pad + 2 * (n - 1) - t];  // right margin: out[pad + 2(n-1) - t] = y[t]
    dy[i] = g;
  }
}

// spec [rows, ld_s]: re in columns [0, nb), im in [nb, 2 nb).  mag [rows, ld_m], columns >= nb zero.
__global__ void __launch_bounds__(256)
spec_mag_fwd_kernel(const float* __restrict__ spec, long rows, int nb, int ld_s, int ld_m, float eps, float* __restrict__ mag) {
  const long total = rows * ld_m;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / ld_m;
    const int c = static_cast<int>(i - r * ld_m);
    float v = 0.0f;
    if (c < nb) {
      const float re = spec[r * ld_s + c], im = spec[r * ld_s + nb + c];
      // eps >= 0: sqrt(p + eps); eps < 0: sqrt(max(p, -eps)) (xVAPitch's TorchSTFT clamps instead of adding)
      const float p = re * re + im * im;
      v = tf32_rn(sqrtf(eps >= 0.0f ? p + eps : fmaxf(p, -eps)));  // operand of the mel projection
    }
    mag[i] = v;
  }
}

// dspec = dmag * spec / mag  (re and im parts; pad columns of dspec zeroed)
__global__ void __launch_bounds__(256)
spec_mag_bwd_kernel(const float* __restrict__ dmag, const float* __restrict__ spec, long rows, int nb, int ld_s, int ld_m,
                    float eps, float* __restrict__ dspec) {
  const long total = rows * ld_s;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / ld_s;
    const int c = static_cast<int>(i - r * ld_s);
    float v = 0.0f;
    if (c < 2 * nb) {
      const int bin = c < nb ? c : c - nb;
      const float re = spec[r * ld_s + bin], im = spec[r * ld_s + nb + bin];
      const float p = re * re + im * im;
      if (eps >= 0.0f || p >= -eps) {  // the clamped variant passes no gradient below its floor
        const float m = sqrtf(eps >= 0.0f ? p + eps : p);  // exact magnitude, not the tf32-rounded copy
        v = tf32_rn(dmag[r * ld_m + bin] * spec[i] / m);
      }
    }
    dspec[i] = v;
  }
}

__global__ void __launch_bounds__(256)
log_clamp_fwd_kernel(const float* __restrict__ x, long n, float lo, float* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = logf(fmaxf(x[i], lo));
}

// d/dx log(clamp(x, lo)) = 1/x where x >= lo, else 0   (torch.clamp passes the gradient on the boundary)
__global__ void __launch_bounds__(256)
log_clamp_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, long n, float lo, float* __restrict__ dx) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    dx[i] = v >= lo ? tf32_rn(dy[i] / v) : 0.0f;
  }
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

// acc[0] += sum |a - b| ;  kind 1: acc[0] += sum (c - a)^2 (b unused)
__global__ void __launch_bounds__(256)
reduce_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, int kind, float c,
                   double* __restrict__ acc) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    if (kind == 0) s += static_cast<double>(fabsf(a[i] - b[i]));
    else {
      const float d = c - a[i];
      s += static_cast<double>(d * d);
    }
  }
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// kind 0: out = scale * sign(b - a)  (gradient of scale * sum|a - b| wrt b), optionally accumulated into out
// kind 1: out = scale * 2 (a - c)    (gradient of scale * sum (c - a)^2 wrt a)
__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, int kind, float c, float scale,
                 float gate_slope, int accumulate, float* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float g;
    if (kind == 0) {
      const float d = b[i] - a[i];
      g = d > 0.0f ? scale : (d < 0.0f ? -scale : 0.0f);
      if (!(b[i] > 0.0f)) g *= gate_slope;  // b is a leaky-ReLU output: gradient wrt its pre-activation
    } else {
      g = scale * 2.0f * (a[i] - c);
    }
    out[i] = accumulate ? out[i] + g : g;
  }
}

// feature_loss term + its gradient in one pass (models.py:263-269): acc += sum |a - b| ; out = scale * sign(b - a),
// times gate_slope where b <= 0 (b is a leaky-ReLU output: gradient wrt its pre-activation).
__global__ void __launch_bounds__(256)
l1_loss_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, float scale, float gate_slope,
                    double* __restrict__ acc, float* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float bv = b[i], d = bv - a[i];
    s += static_cast<double>(fabsf(d));
    float g = d > 0.0f ? scale : (d < 0.0f ? -scale : 0.0f);
    if (!(bv > 0.0f)) g *= gate_slope;
    out[i] = g;
  }
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

}  // namespace

int l1_loss_grad(const float* a, const float* b, long n, float scale, float gate_slope, double* acc, float* out,
                 cudaStream_t stream) {
  if (n == 0) return XVA_OK;
  l1_loss_grad_kernel<<<grid_for(n), 256, 0, stream>>>(a, b, n, scale, gate_slope, acc, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int reflect_pad_fwd(const float* y, int B, long n, int pad, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(pad >= 0 && pad < n, "reflect_pad: pad=%d must be < n=%ld", pad, n);
  const long total = static_cast<long>(B) * (n + 2L * pad);
  reflect_pad_fwd_kernel<<<grid_for(total), 256, 0, stream>>>(y, n, pad, total, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int reflect_pad_bwd(const float* dyp, int B, long n, int pad, float* dy, cudaStream_t stream) {
  XVA_CHECK_ARG(pad >= 0 && pad < n, "reflect_pad bwd: pad=%d must be < n=%ld", pad, n);
  const long total = static_cast<long>(B) * n;
  reflect_pad_bwd_kernel<<<grid_for(total), 256, 0, stream>>>(dyp, n, pad, total, dy);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int spec_mag_fwd(const float* spec, long rows, int nb, int ld_s, int ld_m, float eps, float* mag, cudaStream_t stream) {
  XVA_CHECK_ARG(ld_s >= 2 * nb && ld_m >= nb, "spec_mag: nb=%d ld_s=%d ld_m=%d", nb, ld_s, ld_m);
  spec_mag_fwd_kernel<<<grid_for(rows * ld_m), 256, 0, stream>>>(spec, rows, nb, ld_s, ld_m, eps, mag);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int spec_mag_bwd(const float* dmag, const float* spec, long rows, int nb, int ld_s, int ld_m, float eps, float* dspec,
                 cudaStream_t stream) {
  XVA_CHECK_ARG(ld_s >= 2 * nb && ld_m >= nb, "spec_mag bwd: nb=%d ld_s=%d ld_m=%d", nb, ld_s, ld_m);
  spec_mag_bwd_kernel<<<grid_for(rows * ld_s), 256, 0, stream>>>(dmag, spec, rows, nb, ld_s, ld_m, eps, dspec);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int log_clamp_fwd(const float* x, long n, float lo, float* out, cudaStream_t stream) {
  log_clamp_fwd_kernel<<<grid_for(n), 256, 0, stream>>>(x, n, lo, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int log_clamp_bwd(const float* dy, const float* x, long n, float lo, float* dx, cudaStream_t stream) {
  log_clamp_bwd_kernel<<<grid_for(n), 256, 0, stream>>>(dy, x, n, lo, dx);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int reduce_loss(const float* a, const float* b, long n, int kind, float c, double* acc, cudaStream_t stream) {
  XVA_CHECK_ARG(kind == 0 || kind == 1, "reduce_loss: kind=%d", kind);
  if (n == 0) return XVA_OK;
  reduce_loss_kernel<<<grid_for(n), 256, 0, stream>>>(a, b, n, kind, c, acc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int loss_grad(const float* a, const float* b, long n, int kind, float c, float scale, float gate_slope, int accumulate,
              float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(kind == 0 || kind == 1, "loss_grad: kind=%d", kind);
  if (n == 0) return XVA_OK;
  loss_grad_kernel<<<grid_for(n), 256, 0, stream>>>(a, b, n, kind, c, scale, gate_slope, accumulate, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(melspec)

}  // namespace xva
