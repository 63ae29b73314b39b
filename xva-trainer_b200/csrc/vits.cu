// Element-wise kernels of the xVAPitch posterior encoder, the parts of its WaveNet stack
// that are not tap-GEMMs:
//   gated activation   acts = tanh(a) * sigmoid(b) with x_in = [a | b]     python/xvapitch/wavenet.py:6-13
//   posterior sample   z = (mean + eps * exp(log_scale)) * mask            python/xvapitch/model.py:1472-1475
// and their backward. Both are HBM-bound streaming kernels (one read of each input, one write of each output, float4
// accesses, grid sized to the data); the conditioning add of wavenet.py:9 is done by the producing GEMM's epilogue.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// x_in [rows, ld_in] (a in columns [0, H), b in [H, 2H)) -> acts [rows, H], tf32-rounded (operand of the 1x1 conv)
__global__ void __launch_bounds__(256)
gated_act_fwd_kernel(const float* __restrict__ x_in, long rows, int H, long ld_in, float* __restrict__ acts) {
  const int h4 = H >> 2;
  const long total = rows * h4;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / h4;
    const int c = static_cast<int>(i - r * h4) << 2;
    const float4 a = *reinterpret_cast<const float4*>(x_in + r * ld_in + c);
    const float4 b = *reinterpret_cast<const float4*>(x_in + r * ld_in + H + c);
    float4 o;
    o.x = tf32_rn(tanhf(a.x) * sigmoidf_(b.x));
    o.y = tf32_rn(tanhf(a.y) * sigmoidf_(b.y));
    o.z = tf32_rn(tanhf(a.z) * sigmoidf_(b.z));
    o.w = tf32_rn(tanhf(a.w) * sigmoidf_(b.w));
    *reinterpret_cast<float4*>(acts + r * H + c) = o;
  }
}

// d(a) = d * s * (1 - t^2), d(b) = d * t * s * (1 - s); t and s recomputed from x_in. dx_in [rows, 2H], tf32-rounded
// (operand of the input-gradient and weight-gradient GEMMs of the dilated convolution)
__global__ void __launch_bounds__(256)
gated_act_bwd_kernel(const float* __restrict__ dacts, const float* __restrict__ x_in, long rows, int H, long ld_in,
                     float* __restrict__ dx_in) {
  const int h4 = H >> 2;
  const long total = rows * h4;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / h4;
    const int c = static_cast<int>(i - r * h4) << 2;
    const float4 a = *reinterpret_cast<const float4*>(x_in + r * ld_in + c);
    const float4 b = *reinterpret_cast<const float4*>(x_in + r * ld_in + H + c);
    const float4 d = *reinterpret_cast<const float4*>(dacts + r * H + c);
    float4 da, db;
#define XVA_GATE_BWD(f)                          \
  {                                              \
    const float t = tanhf(a.f), s = sigmoidf_(b.f); \
    da.f = tf32_rn(d.f * s * (1.0f - t * t));    \
    db.f = tf32_rn(d.f * t * s * (1.0f - s));    \
  }
    XVA_GATE_BWD(x) XVA_GATE_BWD(y) XVA_GATE_BWD(z) XVA_GATE_BWD(w)
#undef XVA_GATE_BWD
    *reinterpret_cast<float4*>(dx_in + r * 2 * H + c) = da;
    *reinterpret_cast<float4*>(dx_in + r * 2 * H + H + c) = db;
  }
}

// stats [B, T, 2C] = [mean | log_scale], eps [B, T, C] -> z [B, T, C] (fp32: the decoder rounds its own operand copy),
// rows t >= lens[b] zero
__global__ void __launch_bounds__(256)
vits_sample_fwd_kernel(const float* __restrict__ stats, const float* __restrict__ eps, const int* __restrict__ lens, int B,
                       int T, int C, float* __restrict__ z) {
  const long total = static_cast<long>(B) * T * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / C;
    const int c = static_cast<int>(i - row * C);
    const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long>(b) * T);
    float v = 0.0f;
    if (t < lens[b]) v = stats[row * 2 * C + c] + eps[i] * expf(stats[row * 2 * C + C + c]);
    z[i] = v;
  }
}

// dstats = [dz | dz * eps * exp(log_scale)] on valid rows, zero elsewhere (tf32-rounded: operand of proj's gradients)
__global__ void __launch_bounds__(256)
vits_sample_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ eps, const float* __restrict__ stats,
                       const int* __restrict__ lens, int B, int T, int C, float* __restrict__ dstats) {
  const long total = static_cast<long>(B) * T * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / C;
    const int c = static_cast<int>(i - row * C);
    const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long>(b) * T);
    float dm = 0.0f, dl = 0.0f;
    if (t < lens[b]) {
      dm = dz[i];
      dl = dm * eps[i] * expf(stats[row * 2 * C + C + c]);
    }
    dstats[row * 2 * C + c] = tf32_rn(dm);
    dstats[row * 2 * C + C + c] = tf32_rn(dl);
  }
}

// out[z, c] += sum over the rows of item z of x[z, t, c]: the gradient of a per-utterance vector that the forward pass
// broadcast over the frames (the conditioning slices of the WaveNet layers, the decoder's cond_layer output).
// Block = 32 columns x 8 row lanes of one item; rows are strided over the lanes, partial sums meet in shared memory.
__global__ void __launch_bounds__(256)
colsum_items_kernel(const float* __restrict__ x, int rows, int C, long ld, long zs, float* __restrict__ out, long out_ld) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int z = blockIdx.y;
  float s = 0.0f;
  if (c < C) {
    const float* xp = x + static_cast<long>(z) * zs + c;
    for (int t = threadIdx.y; t < rows; t += 8) s += xp[static_cast<long>(t) * ld];
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float tot = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x];
    out[static_cast<long>(z) * out_ld + c] += tot;
  }
}


// ---- alignment of the latent frames to the text tokens (xVAPitch.train_step, python/xvapitch/model.py:766-771) as ONE
// batched product: with o = exp(-2 logs_p),
//   logp[b, j, i] = sum_c o[i,c] * (-z[j,c]^2 / 2) + sum_c (m o)[i,c] * z[j,c] + r[i],
//   r[i] = sum_c (-log(2 pi) / 2 - logs_p[i,c]) + sum_c (-m[i,c]^2 o[i,c] / 2)
// = <T[i, :], F[j, :]> with token rows T = [o | m o | r | 0..] and frame rows F = [-z^2 / 2 | z | 1 | 0..], K = 2C + 32.
// One block per token row (the row sums r are block reductions); one thread per element for the frame rows.
__global__ void __launch_bounds__(256)
vits_prior_operand_kernel(const float* __restrict__ m, const float* __restrict__ logs, int C, int K, float* __restrict__ out) {
  __shared__ double red[8];
  const long row = blockIdx.x;
  const float* mr = m + row * C;
  const float* lr = logs + row * C;
  float* o = out + row * K;
  double acc = 0.0;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float l = lr[c], mm = mr[c], os = expf(-2.0f * l);
    o[c] = os;
    o[C + c] = mm * os;
    acc += static_cast<double>(-0.91893853320467274178f - l) + static_cast<double>(-0.5f * (mm * mm) * os);
  }
  for (int c = 2 * C + 1 + threadIdx.x; c < K; c += 256) o[c] = 0.0f;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    o[2 * C] = static_cast<float>(t);
  }
}

__global__ void __launch_bounds__(256)
vits_latent_operand_kernel(const float* __restrict__ z, long rows, int C, int K, float* __restrict__ out) {
  const long total = rows * K;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / K;
    const int c = static_cast<int>(i - r * K);
    float v = 0.0f;
    if (c < C) {
      const float x = z[r * C + c];
      v = -0.5f * (x * x);
    } else if (c < 2 * C) {
      v = z[r * C + (c - C)];
    } else if (c == 2 * C) {
      v = 1.0f;
    }
    out[i] = v;
  }
}

// KL(q || p) of the posterior at the flow output against the expanded prior (VitsGeneratorLoss.kl_loss,
// python/xvapitch/losses.py:86-103) and its four gradients in one pass over [B, T, C]:
//   kl = logs_p - logs_q - 1/2 + (z - m)^2 exp(-2 logs_p) / 2 on frames t < lens[b];  acc += sum kl  (the caller divides
//   by the number of valid frames = sum(lens));  gradients times scale / sum(lens).
__global__ void __launch_bounds__(256)
vits_kl_kernel(const float* __restrict__ z, const float* __restrict__ lq, const float* __restrict__ m, const float* __restrict__ lp,
               const int* __restrict__ lens, int B, int T, int C, float scale, double* __restrict__ acc, float* __restrict__ dz,
               float* __restrict__ dlq, float* __restrict__ dm, float* __restrict__ dlp) {
  __shared__ double red[8];
  long count = 0;
  for (int b = 0; b < B; ++b) count += lens[b];
  const float gs = scale / static_cast<float>(count > 0 ? count : 1);
  const long total = static_cast<long>(B) * T * C;
  double s = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / C;
    const int b = static_cast<int>(row / T), t = static_cast<int>(row - static_cast<long>(b) * T);
    float g_z = 0.0f, g_lq = 0.0f, g_m = 0.0f, g_lp = 0.0f;
    if (t < lens[b]) {
      const float d = z[i] - m[i], e = expf(-2.0f * lp[i]);
      s += static_cast<double>(lp[i] - lq[i] - 0.5f + 0.5f * (d * d) * e);
      g_z = gs * d * e;
      g_m = -g_z;
      g_lq = -gs;
      g_lp = gs * (1.0f - (d * d) * e);
    }
    dz[i] = g_z;
    dlq[i] = g_lq;
    dm[i] = g_m;
    dlp[i] = g_lp;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(acc, t);
  }
}

inline int grid_for_n(long n) {
  const long blocks = (n + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

int gated_act_fwd(const float* x_in, long rows, int H, long ld_in, float* acts, cudaStream_t stream) {
  XVA_CHECK_ARG(x_in && acts && rows >= 0 && H >= 4 && H % 4 == 0 && ld_in >= 2 * H && ld_in % 4 == 0,
                "gated_act: rows=%ld H=%d ld_in=%ld (H and ld_in multiples of 4, ld_in >= 2H)", rows, H, ld_in);
  if (rows == 0) return XVA_OK;
  gated_act_fwd_kernel<<<grid_for_n(rows * (H / 4)), 256, 0, stream>>>(x_in, rows, H, ld_in, acts);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int gated_act_bwd(const float* dacts, const float* x_in, long rows, int H, long ld_in, float* dx_in, cudaStream_t stream) {
  XVA_CHECK_ARG(dacts && x_in && dx_in && rows >= 0 && H >= 4 && H % 4 == 0 && ld_in >= 2 * H && ld_in % 4 == 0,
                "gated_act bwd: rows=%ld H=%d ld_in=%ld", rows, H, ld_in);
  if (rows == 0) return XVA_OK;
  gated_act_bwd_kernel<<<grid_for_n(rows * (H / 4)), 256, 0, stream>>>(dacts, x_in, rows, H, ld_in, dx_in);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int vits_sample_fwd(const float* stats, const float* eps, const int* lens, int B, int T, int C, float* z, cudaStream_t stream) {
  XVA_CHECK_ARG(stats && eps && lens && z && B >= 1 && T >= 1 && C >= 1, "vits_sample: B=%d T=%d C=%d", B, T, C);
  vits_sample_fwd_kernel<<<grid_for_n(static_cast<long>(B) * T * C), 256, 0, stream>>>(stats, eps, lens, B, T, C, z);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int vits_sample_bwd(const float* dz, const float* eps, const float* stats, const int* lens, int B, int T, int C, float* dstats,
                    cudaStream_t stream) {
  XVA_CHECK_ARG(dz && eps && stats && lens && dstats && B >= 1 && T >= 1 && C >= 1, "vits_sample bwd: B=%d T=%d C=%d", B, T, C);
  vits_sample_bwd_kernel<<<grid_for_n(static_cast<long>(B) * T * C), 256, 0, stream>>>(dz, eps, stats, lens, B, T, C, dstats);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int colsum_items(const float* x, int Z, int rows, int C, long ld, long zs, float* out, long out_ld, cudaStream_t stream) {
  XVA_CHECK_ARG(x && out && Z >= 1 && Z <= 65535 && rows >= 0 && C >= 1 && ld >= C && out_ld >= C,
                "colsum_items: Z=%d rows=%d C=%d ld=%ld out_ld=%ld", Z, rows, C, ld, out_ld);
  colsum_items_kernel<<<dim3((C + 31) / 32, Z), dim3(32, 8), 0, stream>>>(x, rows, C, ld, zs, out, out_ld);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int vits_logp_operands(const float* m, const float* logs, const float* z, int B, int Tt, int Ts, int C, float* tok, float* frm,
                       cudaStream_t stream) {
  XVA_CHECK_ARG(m && logs && z && tok && frm && B >= 1 && Tt >= 1 && Ts >= 1 && C >= 1, "vits_logp_operands: B=%d Tt=%d Ts=%d C=%d",
                B, Tt, Ts, C);
  const int K = 2 * C + 32;
  vits_prior_operand_kernel<<<B * Tt, 256, 0, stream>>>(m, logs, C, K, tok);
  const long rows = static_cast<long>(B) * Ts;
  vits_latent_operand_kernel<<<grid_for_n(rows * K), 256, 0, stream>>>(z, rows, C, K, frm);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int vits_kl(const float* z, const float* lq, const float* m, const float* lp, const int* lens, int B, int T, int C, float scale,
            double* acc, float* dz, float* dlq, float* dm, float* dlp, cudaStream_t stream) {
  XVA_CHECK_ARG(z && lq && m && lp && lens && acc && dz && dlq && dm && dlp && B >= 1 && T >= 1 && C >= 1,
                "vits_kl: B=%d T=%d C=%d", B, T, C);
  vits_kl_kernel<<<grid_for_n(static_cast<long>(B) * T * C), 256, 0, stream>>>(z, lq, m, lp, lens, B, T, C, scale, acc, dz, dlq, dm, dlp);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(vits)

}  // namespace xva
