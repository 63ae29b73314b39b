// HiFi-GAN discriminator pieces that are not GEMM-shaped (hifigan/models.py:140-260):
//   conv_c1_*      the first convolution of every discriminator (one input channel: Conv2d(1, 32, (5,1), (3,1)) per
//                  period column in DiscriminatorP :144, Conv1d(1, 128, 15) in DiscriminatorS :207), including the
//                  period reshape and its reflect padding (:159-163) as index arithmetic on the raw waveform;
//   avgpool4_*     AvgPool1d(4, 2, padding=2) between the scales of MultiScaleDiscriminator (:240-243);
//   zero_tail_rows clears the alignment rows of an activation buffer so they can be read as conv zero padding.
// All HBM-bound: one pass over the waveform / the [rows, Cout] output.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

struct C1 {
  long xs_b;      // waveform stride between batch items
  int xs_q, xs_c; // sample index of element q of sequence (b, c) = q * xs_q + c * xs_c, reflected when >= Lsrc
  int P;          // sequences per batch item (period columns)
  int Lsrc;       // valid samples per batch item
  int L;          // logical sequence length (after the reflect padding)
  int k, s, pad;  // kernel size, stride, zero padding
  int Lout, Lout_p, Cout;
  float slope;
};

__device__ __forceinline__ long c1_src(const C1& p, int z, int q) {
  const int b = z / p.P, c = z - b * p.P;
  int idx = q * p.xs_q + c * p.xs_c;
  if (idx >= p.Lsrc) idx = 2 * (p.Lsrc - 1) - idx;
  return static_cast<long>(b) * p.xs_b + idx;
}

constexpr int kMaxK = 16;

constexpr int kRowsFwd = 64;    // output rows per block (forward)
constexpr int kRowsBwdW = 256;  // output rows per block (weight gradient)
constexpr int kMaxWin = (kRowsBwdW - 1) * 4 + kMaxK;  // input window of a row tile (stride <= 4)

// A block owns kRowsFwd consecutive output rows of one sequence: the input window ((rows-1)*s + k samples, gathered
// through the period reshape / reflect padding once) and the transposed filter bank [k][Cout] sit in shared memory;
// thread = (row, channel) with the channel fastest, so filter reads are conflict-free, the window read is a broadcast
// and the [rows, Cout] store is fully coalesced.
__global__ void __launch_bounds__(256)
conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, C1 p,
                   int tiles_per_seq, float* __restrict__ out) {
  __shared__ float xs[(kRowsFwd - 1) * 4 + kMaxK];
  __shared__ float ws[kMaxK * 128 + 128];  // [k][Cout] then bias[Cout]
  const int z = blockIdx.x / tiles_per_seq, t0 = (blockIdx.x - z * tiles_per_seq) * kRowsFwd;
  const int win = (kRowsFwd - 1) * p.s + p.k;
  for (int i = threadIdx.x; i < win; i += blockDim.x) {
    const int q = t0 * p.s - p.pad + i;
    xs[i] = (q >= 0 && q < p.L) ? x[c1_src(p, z, q)] : 0.0f;
  }
  for (int i = threadIdx.x; i < p.Cout * p.k; i += blockDim.x) {
    const int co = i / p.k, j = i - co * p.k;
    ws[j * p.Cout + co] = w[i];
  }
  float* bs = ws + p.k * p.Cout;
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) bs[i] = bias[i];
  __syncthreads();
  const int rows = min(kRowsFwd, p.Lout_p - t0);
  float* o = out + (static_cast<long>(z) * p.Lout_p + t0) * p.Cout;
  for (int i = threadIdx.x; i < rows * p.Cout; i += blockDim.x) {
    const int r = i / p.Cout, co = i - r * p.Cout;
    float acc = 0.0f;
    if (t0 + r < p.Lout) {  // alignment rows [Lout, Lout_p) stay zero
      acc = bs[co];
      const float* xw = xs + r * p.s;
#pragma unroll
      for (int j = 0; j < kMaxK; ++j)
        if (j < p.k) acc = fmaf(ws[j * p.Cout + co], xw[j], acc);
      acc = tf32_rn(acc > 0.0f ? acc : p.slope * acc);  // operand of the next (tensor-core) convolution
    }
    o[i] = acc;
  }
}


// Register-tiled variant for the two filter shapes the discriminators use (k15 s1: scale discriminators; k5 s3: period
// discriminators). The kernel above issues two shared-memory loads per FMA and runs at ~0.6 TB/s of output; here a
// thread keeps its channel's K filter taps in registers and slides a register window of the input over 8 consecutive
// output rows: (7 S + K) broadcast loads per 8 K FMAs. thread = (channel, row group), channel fastest: the [rows, Cout]
// store stays fully coalesced.
template <int K, int S>
__global__ void __launch_bounds__(256)
conv_c1_fwd_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, C1 p,
                         int tiles_per_seq, float* __restrict__ out) {
  constexpr int R = 8, WIN = (R - 1) * S + K;
  __shared__ float xs[(kRowsFwd - 1) * S + K + WIN];  // (+ slack: a partial last chunk reads a full window)
  const int z = blockIdx.x / tiles_per_seq, t0 = (blockIdx.x - z * tiles_per_seq) * kRowsFwd;
  constexpr int win_tile = (kRowsFwd - 1) * S + K;
  for (int i = threadIdx.x; i < win_tile + WIN; i += blockDim.x) {
    const int q = t0 * S - p.pad + i;
    xs[i] = (i < win_tile && q >= 0 && q < p.L) ? x[c1_src(p, z, q)] : 0.0f;
  }
  const int co = threadIdx.x % p.Cout, rg = threadIdx.x / p.Cout, groups = blockDim.x / p.Cout;
  float wr[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wr[j] = w[co * K + j];
  const float b = bias[co];
  __syncthreads();
  if (rg >= groups) return;
  const int rows = min(kRowsFwd, p.Lout_p - t0);
  float* o = out + (static_cast<long>(z) * p.Lout_p + t0) * p.Cout + co;
  for (int r0 = rg * R; r0 < rows; r0 += groups * R) {
    float xw[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) xw[i] = xs[r0 * S + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r0 + r >= rows) break;
      float acc = 0.0f;
      if (t0 + r0 + r < p.Lout) {  // alignment rows [Lout, Lout_p) stay zero
        acc = b;
#pragma unroll
        for (int j = 0; j < K; ++j) acc = fmaf(wr[j], xw[r * S + j], acc);
        acc = tf32_rn(acc > 0.0f ? acc : p.slope * acc);
      }
      o[static_cast<long>(r0 + r) * p.Cout] = acc;
    }
  }
}

// dw[co, j] += sum_rows dpre[row, co] * x(row, j) ; db[co] += sum_rows dpre[row, co].
// A block owns kRowsBwdW rows of one sequence (input window in shared memory); thread = (channel, row slice): the
// Cout channels are spread over the lanes (coalesced dpre rows), 256 / Cout row slices run in parallel and are summed
// through shared memory before one atomic per (co, j) per block.
__global__ void __launch_bounds__(256)
conv_c1_bwd_w_kernel(const float* __restrict__ dpre, const float* __restrict__ x, C1 p, int tiles_per_seq,
                     float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float xs[kMaxWin];
  __shared__ float red[256 * (kMaxK + 1)];
  const int z = blockIdx.x / tiles_per_seq, t0 = (blockIdx.x - z * tiles_per_seq) * kRowsBwdW;
  const int rows = min(kRowsBwdW, p.Lout - t0);
  const int win = (kRowsBwdW - 1) * p.s + p.k;
  for (int i = threadIdx.x; i < win; i += blockDim.x) {
    const int q = t0 * p.s - p.pad + i;
    xs[i] = (q >= 0 && q < p.L) ? x[c1_src(p, z, q)] : 0.0f;
  }
  __syncthreads();
  const int co = threadIdx.x % p.Cout, slice = threadIdx.x / p.Cout, n_slices = blockDim.x / p.Cout;
  float acc[kMaxK + 1];
#pragma unroll
  for (int j = 0; j <= kMaxK; ++j) acc[j] = 0.0f;
  if (slice < n_slices) {
    const float* d = dpre + (static_cast<long>(z) * p.Lout_p + t0) * p.Cout + co;
    for (int r = slice; r < rows; r += n_slices) {
      const float g = d[static_cast<long>(r) * p.Cout];
      const float* xw = xs + r * p.s;
#pragma unroll
      for (int j = 0; j < kMaxK; ++j)
        if (j < p.k) acc[j] = fmaf(g, xw[j], acc[j]);
      acc[kMaxK] += g;
    }
  }
#pragma unroll
  for (int j = 0; j <= kMaxK; ++j) red[j * 256 + threadIdx.x] = acc[j];
  __syncthreads();
  // (co, j) sums over the slices: thread i handles pairs i, i + 256, ...
  for (int i = threadIdx.x; i < p.Cout * (p.k + 1); i += blockDim.x) {
    const int c = i % p.Cout, j = i / p.Cout;  // j == k: the bias column
    const int jj = j < p.k ? j : kMaxK;
    float sum = 0.0f;
    for (int sl = 0; sl < n_slices; ++sl) sum += red[jj * 256 + sl * p.Cout + c];
    if (j < p.k) atomicAdd(dw + c * p.k + j, sum);
    else atomicAdd(db + c, sum);
  }
}

// dx[src(z, q)] += sum_{t, j : t*s + j - pad = q} dot(dpre[z, t, :], w[:, j]) ; one warp per (z, q)
__global__ void __launch_bounds__(256)
conv_c1_bwd_x_kernel(const float* __restrict__ dpre, const float* __restrict__ w, C1 p, long items, float scale,
                     float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long item = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (item >= items) return;
  const int z = static_cast<int>(item / p.L), q = static_cast<int>(item - static_cast<long>(z) * p.L);
  float acc = 0.0f;
  for (int j = 0; j < p.k; ++j) {
    const int num = q + p.pad - j;
    if (num < 0 || num % p.s) continue;
    const int t = num / p.s;
    if (t >= p.Lout) continue;
    const float* d = dpre + (static_cast<long>(z) * p.Lout_p + t) * p.Cout;
    for (int co = lane; co < p.Cout; co += 32) acc = fmaf(d[co], w[co * p.k + j], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) atomicAdd(dx + c1_src(p, z, q), scale * acc);
}


// Tiled variant: a block owns 64 consecutive input samples of one sequence; the dpre rows they touch and the filter
// bank [k][Cout] sit in shared memory (the kernel above re-reads every dpre row k / s times from L2 with one dependent
// load chain per warp: latency-bound at ~100 us for the 16 x 8192 x 128 scale discriminator). One warp per sample,
// lanes over the channels.
constexpr int kTileQ = 64;
__global__ void __launch_bounds__(256)
conv_c1_bwd_x_tiled_kernel(const float* __restrict__ dpre, const float* __restrict__ w, C1 p, int tiles_per_seq, int t_rows,
                           float scale, float* __restrict__ dx) {
  extern __shared__ float sm[];
  float* ws = sm;                      // [k][Cout]
  float* ds = sm + p.k * p.Cout;       // [t_rows][Cout]
  const int z = blockIdx.x / tiles_per_seq, q0 = (blockIdx.x - z * tiles_per_seq) * kTileQ;
  // first dpre row any sample of the tile can touch: t = ceil((q0 + pad - (k - 1)) / s), clamped at 0
  const int lo_num = q0 + p.pad - (p.k - 1);
  const int t_lo = lo_num <= 0 ? 0 : (lo_num + p.s - 1) / p.s;
  for (int i = threadIdx.x; i < p.Cout * p.k; i += blockDim.x) {
    const int co = i / p.k, j = i - co * p.k;
    ws[j * p.Cout + co] = w[i];
  }
  const float* d = dpre + (static_cast<long>(z) * p.Lout_p + t_lo) * p.Cout;
  const int valid = max(0, min(t_rows, p.Lout - t_lo));
  for (int i = threadIdx.x; i < t_rows * p.Cout; i += blockDim.x) ds[i] = (i < valid * p.Cout) ? d[i] : 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int qq = warp; qq < kTileQ; qq += 8) {
    const int q = q0 + qq;
    if (q >= p.L) break;
    float acc = 0.0f;
    for (int j = 0; j < p.k; ++j) {
      const int num = q + p.pad - j;
      if (num < 0 || num % p.s) continue;
      const int t = num / p.s - t_lo;
      if (t < 0 || t >= valid) continue;
      const float* dr = ds + t * p.Cout;
      const float* wj = ws + j * p.Cout;
      for (int co = lane; co < p.Cout; co += 32) acc = fmaf(dr[co], wj[co], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(dx + c1_src(p, z, q), scale * acc);
  }
}

// AvgPool1d(4, 2, padding=2), count_include_pad: out[i] = (x[2i-2] + x[2i-1] + x[2i] + x[2i+1]) / 4
__global__ void __launch_bounds__(256)
avgpool4_fwd_kernel(const float* __restrict__ x, int L, int Lout, long total, float* __restrict__ out) {
  for (long n = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; n < total; n += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = n / Lout;
    const int i = static_cast<int>(n - b * Lout);
    const float* r = x + b * L;
    float s = 0.0f;
#pragma unroll
    for (int d = -2; d <= 1; ++d) {
      const int m = 2 * i + d;
      if (m >= 0 && m < L) s += r[m];
    }
    out[n] = 0.25f * s;
  }
}

__global__ void __launch_bounds__(256)
avgpool4_bwd_kernel(const float* __restrict__ dout, int L, int Lout, long total, float* __restrict__ dx) {
  for (long n = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; n < total; n += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = n / L;
    const int m = static_cast<int>(n - b * L);
    const float* r = dout + b * Lout;
    float s = 0.0f;
    // 2i + d = m, d in {-2,-1,0,1}  ->  i in {(m+2)/2, (m+1)/2, m/2, (m-1)/2} where the division is exact
#pragma unroll
    for (int d = -2; d <= 1; ++d) {
      const int num = m - d;
      if (num >= 0 && (num & 1) == 0 && (num >> 1) < Lout) s += r[num >> 1];
    }
    dx[n] = 0.25f * s;
  }
}

__global__ void __launch_bounds__(256)
zero_tail_rows_kernel(float* __restrict__ x, int Lp, int Lvalid, int C, long total) {
  const int tail = (Lp - Lvalid) * C;
  for (long n = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; n < total; n += static_cast<long>(gridDim.x) * blockDim.x) {
    const long z = n / tail;
    x[(z * Lp + Lvalid) * C + (n - z * tail)] = 0.0f;
  }
}

inline int grid_for(long n) {
  long b = ceil_div_l(n, 256 * 4);
  const long cap = 16L * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

inline int fill(C1* p, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k, int s, int pad, int Lout, int Lout_p,
                int Cout, float slope) {
  XVA_CHECK_ARG(k >= 1 && k <= kMaxK, "conv_c1: kernel size %d (max %d)", k, kMaxK);
  XVA_CHECK_ARG(s >= 1 && Lout <= Lout_p && Cout >= 1, "conv_c1: stride %d Lout %d/%d Cout %d", s, Lout, Lout_p, Cout);
  XVA_CHECK_ARG(L - Lsrc < Lsrc, "conv_c1: reflect padding %d longer than the signal %d", L - Lsrc, Lsrc);
  p->xs_b = xs_b; p->xs_q = xs_q; p->xs_c = xs_c; p->P = P; p->Lsrc = Lsrc; p->L = L; p->k = k; p->s = s; p->pad = pad;
  p->Lout = Lout; p->Lout_p = Lout_p; p->Cout = Cout; p->slope = slope;
  return XVA_OK;
}

}  // namespace

int conv_c1_fwd(const float* x, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, const float* w, const float* bias,
                int k, int s, int pad, int Z, int Lout, int Lout_p, int Cout, float slope, float* out, cudaStream_t stream) {
  C1 p;
  int rc = fill(&p, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Lout, Lout_p, Cout, slope);
  if (rc != XVA_OK) return rc;
  XVA_CHECK_ARG(Cout <= 128 && s <= 4, "conv_c1 fwd: Cout=%d (max 128), stride=%d (max 4)", Cout, s);
  const int tiles = ceil_div(Lout_p, kRowsFwd);
  static const bool tiled = [] {  // XVA_C1_TILED=0: the original kernels (A/B and debugging)
    const char* e = getenv("XVA_C1_TILED");
    return !(e && e[0] == '0');
  }();
  const bool fits = tiled && Cout <= 256 && 256 % Cout == 0;
  if (fits && k == 15 && s == 1) conv_c1_fwd_tiled_kernel<15, 1><<<Z * tiles, 256, 0, stream>>>(x, w, bias, p, tiles, out);
  else if (fits && k == 5 && s == 3) conv_c1_fwd_tiled_kernel<5, 3><<<Z * tiles, 256, 0, stream>>>(x, w, bias, p, tiles, out);
  else conv_c1_fwd_kernel<<<Z * tiles, 256, 0, stream>>>(x, w, bias, p, tiles, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int conv_c1_bwd_w(const float* dpre, const float* x, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k, int s,
                  int pad, int Z, int Lout, int Lout_p, int Cout, float* dw, float* db, cudaStream_t stream) {
  XVA_CHECK_ARG(Cout <= 128 && s <= 4, "conv_c1 bwd: Cout=%d (max 128), stride=%d (max 4)", Cout, s);
  C1 p;
  int rc = fill(&p, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Lout, Lout_p, Cout, 0.0f);
  if (rc != XVA_OK) return rc;
  const int tiles = ceil_div(Lout, kRowsBwdW);
  conv_c1_bwd_w_kernel<<<Z * tiles, 256, 0, stream>>>(dpre, x, p, tiles, dw, db);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int conv_c1_bwd_x(const float* dpre, const float* w, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k, int s,
                  int pad, int Z, int Lout, int Lout_p, int Cout, float scale, float* dx, cudaStream_t stream) {
  C1 p;
  int rc = fill(&p, xs_b, xs_q, xs_c, P, Lsrc, L, k, s, pad, Lout, Lout_p, Cout, 0.0f);
  if (rc != XVA_OK) return rc;
  const long items = static_cast<long>(Z) * L;
  static const bool tiled = [] {
    const char* e = getenv("XVA_C1_TILED");
    return !(e && e[0] == '0');
  }();
  // dpre rows one tile of kTileQ samples can touch: ceil((kTileQ - 1 + k - 1) / s) + 1
  const int t_rows = (kTileQ + k - 2) / s + 2;
  const size_t smem = (static_cast<size_t>(k) + t_rows) * Cout * sizeof(float);
  if (tiled && smem <= 64 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      XVA_CHECK_CUDA(cudaFuncSetAttribute(conv_c1_bwd_x_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_done = true;
    }
    const int tiles = ceil_div(L, kTileQ);
    conv_c1_bwd_x_tiled_kernel<<<Z * tiles, 256, smem, stream>>>(dpre, w, p, tiles, t_rows, scale, dx);
  } else {
    conv_c1_bwd_x_kernel<<<static_cast<int>(ceil_div_l(items, 8)), 256, 0, stream>>>(dpre, w, p, items, scale, dx);
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int avgpool4_fwd(const float* x, int B, int L, float* out, cudaStream_t stream) {
  const int Lout = L / 2 + 1;
  const long total = static_cast<long>(B) * Lout;
  avgpool4_fwd_kernel<<<grid_for(total), 256, 0, stream>>>(x, L, Lout, total, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int avgpool4_bwd(const float* dout, int B, int L, float* dx, cudaStream_t stream) {
  const int Lout = L / 2 + 1;
  const long total = static_cast<long>(B) * L;
  avgpool4_bwd_kernel<<<grid_for(total), 256, 0, stream>>>(dout, L, Lout, total, dx);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int zero_tail_rows(float* x, int Z, int Lp, int Lvalid, int C, cudaStream_t stream) {
  XVA_CHECK_ARG(Lvalid <= Lp, "zero_tail_rows: %d > %d", Lvalid, Lp);
  if (Lvalid == Lp) return XVA_OK;
  const long total = static_cast<long>(Z) * (Lp - Lvalid) * C;
  zero_tail_rows_kernel<<<grid_for(total), 256, 0, stream>>>(x, Lp, Lvalid, C, total);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(disc)

}  // namespace xva
