// FastPitch training stage 1, the aligner: everything between the two projection stacks of
// ConvAttention and the monotonic alignment search, plus the two attention losses and their gradients. Replaces, paths
// relative to python/fastpitch1_1/fastpitch/ :
//   attention.py:203-219            the isotropic-Gaussian score  -0.0005 * sum_c (q[t,c] - k[j,c])^2, log_softmax over text,
//                                   + log(prior + 1e-8)  (= attn_logprob), masked softmax over text (= attn_soft)
//   attn_loss_function.py:20-44     AttentionCTCLoss: per utterance, log_softmax over [blank = -1, keys < key_len], CTC with
//                                   the target 1..key_len, reduction 'mean' (divide by key_len), zero_infinity; mean over B
//   attn_loss_function.py:47-54     AttentionBinarizationLoss: -sum_{hard == 1} log(clamp(soft, 1e-12)) / sum(hard)
// and the autograd of all three. The reference builds a [B, C, Tm, Tt] broadcast difference (1.44 GB at 32 x 80 x 880 x
// 160), loops over the batch in Python for the CTC calls and syncs 32 times; here the score is computed from the key
// matrix held in shared memory (one pass over q, one write of each output), the CTC forward and backward recursions of an
// utterance are two blocks walking the mel axis in opposite directions at the same time with one thread per
// extended-label state, and the score gradients come out of two kernels.
//
// All of it is fp32 CUDA-core work on a 80-channel contraction (0.7 GFLOP per pass at 32 x 880 x 160) and a few 18 MB
// tensors: HBM / latency bound, not a tensor-core shape. The differences q - k are formed explicitly in fp32, exactly as
// the reference does (the expanded form |q|^2 - 2 q.k + |k|^2 on tf32 tensor cores would cancel catastrophically once the
// aligner has converged and q ~ k). The CTC recursion keeps alpha / beta in fp64 (the magnitudes reach thousands after
// 880 frames; torch's fp32 recursion carries ~1e-3 of absolute noise there) and evaluates the log-sum-exp of the small
// differences in fp32.
#include <cmath>

#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

constexpr int kWarps = 8;           // warps per block of the row kernels
constexpr int kRowsPerBlock = 16;   // mel rows per block: the key matrix is staged once for all of them
constexpr int kMaxChunks = 16;      // text positions per lane in registers -> Tt <= 512
constexpr int kMaxC = 96;           // attention channels (n_att_channels = 80)
constexpr float kScoreScale = 0.0005f;  // attention.py:209
constexpr float kPriorEps = 1e-8f;      // attention.py:211

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

// keys of one utterance [Tt, C] (row pitch ldk) -> shared memory with row pitch C + 1: lanes that walk the text axis
// (forward) and lanes that walk the channel axis (backward) both hit 32 different banks
__device__ __forceinline__ void stage_keys(const float* __restrict__ kb, int Tt, int C, long ldk, float* ks) {
  const int pitch = C + 1;
  for (int i = threadIdx.x; i < Tt * C; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    ks[j * pitch + c] = kb[static_cast<long>(j) * ldk + c];
  }
}

// ------------------------------------------------------------------------------------------------ score, forward
// logprob[b,t,j] = log_softmax_j(-0.0005 |q[b,t] - k[b,j]|^2) + log(prior[b,t,j] + 1e-8)            all j < Tt
// soft[b,t,j]    = softmax_j(logprob[b,t,j] for j < in_lens[b]), 0 for the padded keys
__global__ void __launch_bounds__(kWarps * 32)
attn_score_fwd_kernel(const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk,
                      const float* __restrict__ prior, const int* __restrict__ in_lens, int Tm, int Tt, int C,
                      float* __restrict__ logprob, float* __restrict__ soft) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.y;
  const int pitch = C + 1;
  float* ks = smem_f;                        // [Tt][C + 1]
  float* qs = ks + Tt * pitch;               // [kWarps][C]
  stage_keys(k + static_cast<long>(b) * Tt * ldk, Tt, C, ldk, ks);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_key = min(max(in_lens[b], 0), Tt);
  float* qw = qs + warp * C;
  for (int r = warp; r < kRowsPerBlock; r += kWarps) {
    const int t = blockIdx.x * kRowsPerBlock + r;
    if (t >= Tm) break;                      // warp-uniform
    const long row = static_cast<long>(b) * Tm + t;
    __syncwarp();
    for (int c = lane; c < C; c += 32) qw[c] = q[row * ldq + c];
    __syncwarp();
    float d[kMaxChunks];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      d[i] = -INFINITY;
      const int j = i * 32 + lane;
      if (i * 32 < Tt && j < Tt) {
        const float* kr = ks + j * pitch;
        float acc = 0.0f;
        for (int c = 0; c < C; ++c) {
          const float df = qw[c] - kr[c];
          acc = fmaf(df, df, acc);
        }
        d[i] = -kScoreScale * acc;
        mx = fmaxf(mx, d[i]);
      }
    }
    mx = warp_max(mx);
    float se = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i)
      if (i * 32 < Tt && i * 32 + lane < Tt) se += expf(d[i] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    float mx2 = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int j = i * 32 + lane;
      if (i * 32 < Tt && j < Tt) {
        const float lp = (d[i] - lse) + logf(prior[row * Tt + j] + kPriorEps);
        logprob[row * Tt + j] = lp;
        d[i] = lp;
        if (j < n_key) mx2 = fmaxf(mx2, lp);
      }
    }
    mx2 = warp_max(mx2);
    float s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int j = i * 32 + lane;
      if (i * 32 < Tt && j < n_key) {
        d[i] = expf(d[i] - mx2);
        s2 += d[i];
      }
    }
    s2 = warp_sum(s2);
    const float inv = 1.0f / s2;
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) {
      const int j = i * 32 + lane;
      if (i * 32 < Tt && j < Tt) soft[row * Tt + j] = (j < n_key) ? d[i] * inv : 0.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ score, backward
// g = d(loss)/d(logprob) (every path into it already summed). Per mel row:
//   dD[j]   = g[j] - softmax_j(score) * sum_j g[j],   softmax_j(score) = exp(logprob[j] - log(prior[j] + 1e-8))
//   dq[t,c] = 0.001 * sum_j dD[j] * (k[j,c] - q[t,c])                       (d/dq of -0.0005 (q - k)^2)
// dD is written out for the key gradient (it may alias g).
__global__ void __launch_bounds__(kWarps * 32)
attn_score_bwd_kernel(const float* g, const float* __restrict__ logprob, const float* __restrict__ prior,
                      const float* __restrict__ q, long ldq, const float* __restrict__ k, long ldk, int Tm, int Tt,
                      int C, float* dD, float* __restrict__ dq, long lddq) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.y;
  const int pitch = C + 1;
  float* ks = smem_f;                        // [Tt][C + 1]
  float* qs = ks + Tt * pitch;               // [kWarps][C]
  float* ds = qs + kWarps * C;               // [kWarps][Tt]
  stage_keys(k + static_cast<long>(b) * Tt * ldk, Tt, C, ldk, ks);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qw = qs + warp * C;
  float* dw = ds + warp * Tt;
  for (int r = warp; r < kRowsPerBlock; r += kWarps) {
    const int t = blockIdx.x * kRowsPerBlock + r;
    if (t >= Tm) break;
    const long row = static_cast<long>(b) * Tm + t;
    __syncwarp();
    for (int c = lane; c < C; c += 32) qw[c] = q[row * ldq + c];
    float gs = 0.0f;
    for (int j = lane; j < Tt; j += 32) {
      const float gj = g[row * Tt + j];
      dw[j] = gj;
      gs += gj;
    }
    gs = warp_sum(gs);
    for (int j = lane; j < Tt; j += 32) {
      const float sm = expf(logprob[row * Tt + j] - logf(prior[row * Tt + j] + kPriorEps));
      const float v = dw[j] - sm * gs;
      dw[j] = v;
      dD[row * Tt + j] = v;
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      const float qc = qw[c];
      float acc = 0.0f;
      for (int j = 0; j < Tt; ++j) acc = fmaf(dw[j], ks[j * pitch + c] - qc, acc);
      dq[row * lddq + c] = tf32_rn(2.0f * kScoreScale * acc);   // A operand of the query stack's gradient GEMMs
    }
  }
}

// dk[j,c] = 0.001 * sum_t dD[t,j] * (q[t,c] - k[j,c]). One block per (utterance, 16 text positions); the mel axis is
// walked in chunks of 32 rows staged in shared memory; each thread owns up to 6 of the 16 x C outputs.
constexpr int kKgJ = 16, kKgT = 32, kKgThreads = 256;
constexpr int kKgOut = (kKgJ * kMaxC + kKgThreads - 1) / kKgThreads;

__global__ void __launch_bounds__(kKgThreads)
attn_key_grad_kernel(const float* __restrict__ dD, const float* __restrict__ q, long ldq, const float* __restrict__ k,
                     long ldk, int Tm, int Tt, int C, float* __restrict__ dk, long lddk) {
  __shared__ float dds[kKgT][kKgJ];
  __shared__ float qs[kKgT][kMaxC];
  const int b = blockIdx.y, j0 = blockIdx.x * kKgJ;
  const int n_out = kKgJ * C;
  float acc[kKgOut], kreg[kKgOut];
  int oj[kKgOut], oc[kKgOut];
#pragma unroll
  for (int n = 0; n < kKgOut; ++n) {
    const int o = threadIdx.x + n * kKgThreads;
    oj[n] = o / C;
    oc[n] = o - oj[n] * C;
    acc[n] = 0.0f;
    const bool live = o < n_out && j0 + oj[n] < Tt;
    kreg[n] = live ? k[(static_cast<long>(b) * Tt + j0 + oj[n]) * ldk + oc[n]] : 0.0f;
    if (!live) {          // park dead slots on a valid shared-memory address; their result is never stored
      oj[n] = 0;
      oc[n] = 0;
    }
  }
  for (int t0 = 0; t0 < Tm; t0 += kKgT) {
    __syncthreads();
    for (int i = threadIdx.x; i < kKgT * kKgJ; i += kKgThreads) {
      const int tt = i / kKgJ, jj = i - tt * kKgJ;
      const bool in = t0 + tt < Tm && j0 + jj < Tt;
      dds[tt][jj] = in ? dD[(static_cast<long>(b) * Tm + t0 + tt) * Tt + j0 + jj] : 0.0f;
    }
    for (int i = threadIdx.x; i < kKgT * C; i += kKgThreads) {
      const int tt = i / C, c = i - tt * C;
      qs[tt][c] = (t0 + tt < Tm) ? q[(static_cast<long>(b) * Tm + t0 + tt) * ldq + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int tt = 0; tt < kKgT; ++tt) {
#pragma unroll
      for (int n = 0; n < kKgOut; ++n) acc[n] = fmaf(dds[tt][oj[n]], qs[tt][oc[n]] - kreg[n], acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < kKgOut; ++n) {
    const int o = threadIdx.x + n * kKgThreads;
    const int jj = o / C, c = o - jj * C;
    if (o < n_out && j0 + jj < Tt)
      dk[(static_cast<long>(b) * Tt + j0 + jj) * lddk + c] = tf32_rn(2.0f * kScoreScale * acc[n]);
  }
}

// ------------------------------------------------------------------------------------------------ CTC
// Extended label sequence of the target 1..L: state s (0 <= s < S = 2L+1) is the blank for even s and key (s-1)/2 for
// odd s; all keys are distinct, so the skip s-2 -> s is open for every odd s >= 3.
//   z[t, .]    = [blank_logprob, logprob[t, 0..L)] ;  n[t, .] = log_softmax(z[t, .])       (attn_loss_function.py:28-36)
//   alpha_t(s) = n[t, lab(s)] + lse(alpha_{t-1}(s), alpha_{t-1}(s-1), alpha_{t-1}(s-2)*),  alpha_0 = n[0, .] on s <= 1
//   beta_t(s)  = n[t, lab(s)] + lse(beta_{t+1}(s),  beta_{t+1}(s+1),  beta_{t+1}(s+2)*),   beta_{T-1} = n[T-1, .] on s >= S-2
//   nll        = -lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2)),   cost = nll / max(L, 1)  (0 if nll is infinite: zero_infinity)
//   d cost / d logprob[t, j] = (exp(n[t,j+1]) - exp(alpha_t(s) + beta_t(s) - n[t,j+1] + nll)) / max(L, 1),  s = 2j+1
//   (the gradient of -log p through the row's log_softmax; the blank column is a constant).
// Three launches. (1) the row normalisers lse[b,t], one warp per row. (2) the two recursions: they are independent, so
// block (b, 0) walks alpha forward and block (b, 1) walks beta backward at the same time -- the mel axis is the only
// sequential dimension of the whole stage, and this halves it; one thread per state, the emission of step t+1 is loaded
// while step t is computed (the global-load latency was 1/3 of the first version's step time, the barrier another 1/3),
// both tables go to the workspace in fp64. (3) the gradient, fully parallel over (b, t, j).
// Measured at 32 x 880 x 160: 0.45 ms for the three launches. An eight-step register look-ahead instead of one was not
// faster (0.55 ms, scripts/gpu_call_v.sh): after the first prefetch the step is bound by the ~150 instructions per state
// (three expf, one logf, fp64 adds) that one SM issues for the 11 warps of an utterance, not by load latency.
// First version (one block per utterance doing normaliser, alpha, then beta + gradient): 2.35 ms at 32 x 880 x 160.

// log(exp(a) + exp(b) + exp(c)): the maximum is carried in fp64, the (small) differences go through fp32 exp / log
__device__ __forceinline__ double lse3(double a, double b, double c) {
  const double m = fmax(a, fmax(b, c));
  if (m == -INFINITY) return m;
  const float s = expf(static_cast<float>(a - m)) + expf(static_cast<float>(b - m)) + expf(static_cast<float>(c - m));
  return m + static_cast<double>(logf(s));
}

// lse[b,t] = log(exp(blank) + sum_{j < L} exp(logprob[b,t,j])) for t < T (0 elsewhere)
__global__ void __launch_bounds__(kWarps * 32)
ctc_row_lse_kernel(const float* __restrict__ logprob, const int* __restrict__ in_lens, const int* __restrict__ out_lens,
                   int Tm, int Tt, float blank, long rows, float* __restrict__ lse) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarps + warp;
  if (row >= rows) return;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row - static_cast<long>(b) * Tm);
  const int L = min(max(in_lens[b], 0), Tt), T = min(max(out_lens[b], 0), Tm);
  if (t >= T) {
    if (lane == 0) lse[row] = 0.0f;
    return;
  }
  const float* lp = logprob + row * Tt;
  float mx = blank;
  for (int j = lane; j < L; j += 32) mx = fmaxf(mx, lp[j]);
  mx = warp_max(mx);
  float se = (lane == 0) ? expf(blank - mx) : 0.0f;
  for (int j = lane; j < L; j += 32) se += expf(lp[j] - mx);
  se = warp_sum(se);
  if (lane == 0) lse[row] = mx + logf(se);
}

constexpr int kCtcMaxThreads = 1024;
constexpr int kCtcNS = 2;   // states per thread: S_max = 2 Tt + 1 <= 2 * 1024

// blockIdx.y = 0: alpha, forward in time, -> table[0]; blockIdx.y = 1: beta, backward in time, -> table[1].
// table[d][b][t][s] fp64, row pitch S_max. The alpha block also writes nll[b] and cost[b].
__global__ void __launch_bounds__(kCtcMaxThreads)
ctc_recursion_kernel(const float* __restrict__ logprob, const int* __restrict__ in_lens, const int* __restrict__ out_lens,
                     const float* __restrict__ lse, int B, int Tm, int Tt, float blank, double* __restrict__ table,
                     double* __restrict__ nll_out, double* __restrict__ cost) {
  extern __shared__ double smem_d[];
  const int b = blockIdx.x;
  const bool fwd = blockIdx.y == 0;
  const int L = min(max(in_lens[b], 0), Tt), T = min(max(out_lens[b], 0), Tm);
  const int S = 2 * L + 1, S_max = 2 * Tt + 1;
  double* buf0 = smem_d;                       // [S_max + 4], state s at index s + 2, two -inf guards on either side
  double* buf1 = buf0 + (S_max + 4);
  const float* lp = logprob + static_cast<long>(b) * Tm * Tt;
  const float* ls = lse + static_cast<long>(b) * Tm;
  double* tab = table + (static_cast<long>(fwd ? 0 : 1) * B + b) * Tm * S_max;
  const double ninf = -INFINITY;
  if (T == 0) {
    if (fwd && threadIdx.x == 0) {
      nll_out[b] = 0.0;
      cost[b] = 0.0;
    }
    return;
  }
  for (int i = threadIdx.x; i < 2 * (S_max + 4); i += blockDim.x) buf0[i] = ninf;
  __syncthreads();
  const int dt = fwd ? 1 : -1, t_first = fwd ? 0 : T - 1, off1 = fwd ? -1 : 1;
  // raw emission of (t, s): logprob of key (s-1)/2 for odd s, the blank constant for even s; minus lse[t] when used
  float raw[kCtcNS], raw_next[kCtcNS];
  float lse_cur = ls[t_first], lse_next = 0.0f;
#pragma unroll
  for (int k = 0; k < kCtcNS; ++k) {
    const int s = threadIdx.x + k * blockDim.x;
    raw[k] = (s < S && (s & 1)) ? lp[static_cast<long>(t_first) * Tt + (s >> 1)] : blank;
  }
  double* prev = buf0 + 2;
  double* cur = buf1 + 2;
  for (int step = 0; step < T; ++step) {
    const int t = t_first + step * dt;
    const bool more = step + 1 < T;
    if (more) {                                   // loads of the next step, in flight while this one is computed
      lse_next = ls[t + dt];
#pragma unroll
      for (int k = 0; k < kCtcNS; ++k) {
        const int s = threadIdx.x + k * blockDim.x;
        raw_next[k] = (s < S && (s & 1)) ? lp[static_cast<long>(t + dt) * Tt + (s >> 1)] : blank;
      }
    }
#pragma unroll
    for (int k = 0; k < kCtcNS; ++k) {
      const int s = threadIdx.x + k * blockDim.x;
      if (s < S) {
        const double e = static_cast<double>(raw[k] - lse_cur);
        double v;
        if (step == 0) {
          v = (fwd ? (s <= 1) : (s >= S - 2)) ? e : ninf;
        } else {
          const double c2 = (s & 1) ? prev[s + 2 * off1] : ninf;
          v = e + lse3(prev[s], prev[s + off1], c2);
        }
        cur[s] = v;
        tab[static_cast<long>(t) * S_max + s] = v;
      }
    }
    __syncthreads();
    double* sw = prev;
    prev = cur;
    cur = sw;
    lse_cur = lse_next;
#pragma unroll
    for (int k = 0; k < kCtcNS; ++k) raw[k] = raw_next[k];
  }
  if (fwd && threadIdx.x == 0) {
    const double nll = -lse3(prev[S - 1], prev[S - 2], ninf);   // S = 1: prev[-1] is a guard
    const bool ok = nll < INFINITY && nll == nll;               // zero_infinity: no cost (and no gradient) otherwise
    nll_out[b] = nll;
    cost[b] = ok ? nll / static_cast<double>(max(L, 1)) : 0.0;
  }
}

// grad[b,t,j] = (1 / (B max(L,1))) * (exp(n) - exp(alpha_t(2j+1) + beta_t(2j+1) - n + nll)),  n = logprob[b,t,j] - lse[b,t];
// zero outside [T, L] and for utterances whose alignment is impossible
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ logprob, const int* __restrict__ in_lens, const int* __restrict__ out_lens,
                const float* __restrict__ lse, const double* __restrict__ table, const double* __restrict__ nll_in, int B,
                int Tm, int Tt, float* __restrict__ grad) {
  const long total = static_cast<long>(B) * Tm * Tt;
  const int S_max = 2 * Tt + 1;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / Tt;
    const int j = static_cast<int>(i - row * Tt);
    const int b = static_cast<int>(row / Tm), t = static_cast<int>(row - static_cast<long>(b) * Tm);
    const int L = min(max(in_lens[b], 0), Tt), T = min(max(out_lens[b], 0), Tm);
    const double nll = nll_in[b];
    float g = 0.0f;
    if (t < T && j < L && nll < INFINITY && nll == nll) {
      const long at = (static_cast<long>(b) * Tm + t) * S_max + 2 * j + 1;
      const double al = table[at], be = table[static_cast<long>(B) * Tm * S_max + at];
      const float n = logprob[i] - lse[row];
      const float post = expf(static_cast<float>(al + be + nll) - n);
      g = (expf(n) - post) / (static_cast<float>(B) * static_cast<float>(max(L, 1)));
    }
    grad[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ binarization loss
// acc[0] += sum_{hard == 1} log(max(soft, eps)),  acc[1] += sum hard        (attn_loss_function.py:51-54)
__global__ void __launch_bounds__(kWarps * 32)
attn_bin_loss_kernel(const float* __restrict__ hard, const float* __restrict__ soft, long rows, int Tt, float eps,
                     double* __restrict__ acc) {
  __shared__ float part[2][kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ls = 0.0f, cnt = 0.0f;
  for (long row = static_cast<long>(blockIdx.x) * kWarps + warp; row < rows; row += static_cast<long>(gridDim.x) * kWarps) {
    for (int j = lane; j < Tt; j += 32) {
      const float h = hard[row * Tt + j];
      cnt += h;
      if (h == 1.0f) ls += logf(fmaxf(soft[row * Tt + j], eps));
    }
  }
  ls = warp_sum(ls);
  cnt = warp_sum(cnt);
  if (lane == 0) {
    part[0][warp] = ls;
    part[1][warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < kWarps; ++i) {
      a += part[0][i];
      c += part[1][i];
    }
    atomicAdd(acc, a);
    atomicAdd(acc + 1, c);
  }
}

// g[t,j] = a * gctc[t,j] + (bw / N) * (soft[t,j] * sum_j h'[t,j] - h'[t,j]),  h' = hard * [soft >= eps],  N = acc[1]
// (the second term is d/d logprob of bw * (-sum_{hard} log clamp(soft, eps) / N) through the masked softmax; the padded
// keys have soft = hard = 0). hard == nullptr: g = a * gctc.
__global__ void __launch_bounds__(kWarps * 32)
attn_grad_combine_kernel(const float* __restrict__ gctc, const float* __restrict__ hard, const float* __restrict__ soft,
                         const double* __restrict__ acc, float a, float bw, float eps, long rows, int Tt,
                         float* __restrict__ g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * kWarps + warp;
  if (row >= rows) return;
  float kw = 0.0f;
  if (hard != nullptr && bw != 0.0f) {
    const double n = acc[1];
    kw = n > 0.0 ? static_cast<float>(static_cast<double>(bw) / n) : 0.0f;
  }
  float hs = 0.0f;
  if (kw != 0.0f) {
    for (int j = lane; j < Tt; j += 32)
      hs += (hard[row * Tt + j] == 1.0f && soft[row * Tt + j] >= eps) ? 1.0f : 0.0f;
    hs = warp_sum(hs);
  }
  for (int j = lane; j < Tt; j += 32) {
    float v = a * gctc[row * Tt + j];
    if (kw != 0.0f) {
      const float s = soft[row * Tt + j];
      const float hp = (hard[row * Tt + j] == 1.0f && s >= eps) ? 1.0f : 0.0f;
      v += kw * (s * hs - hp);
    }
    g[row * Tt + j] = v;
  }
}

size_t score_smem(int Tt, int C, bool bwd) {
  return (static_cast<size_t>(Tt) * (C + 1) + static_cast<size_t>(kWarps) * C + (bwd ? static_cast<size_t>(kWarps) * Tt : 0)) * 4;
}

int check_score_shape(const char* what, int B, int Tm, int Tt, int C, long ldq, long ldk) {
  XVA_CHECK_ARG(B >= 1 && Tm >= 1 && Tt >= 1 && C >= 1, "%s: B=%d Tm=%d Tt=%d C=%d", what, B, Tm, Tt, C);
  XVA_CHECK_ARG(Tt <= 32 * kMaxChunks, "%s: Tt=%d exceeds %d text positions", what, Tt, 32 * kMaxChunks);
  XVA_CHECK_ARG(C <= kMaxC && ldq >= C && ldk >= C, "%s: C=%d (max %d), ldq=%ld, ldk=%ld", what, C, kMaxC, ldq, ldk);
  XVA_CHECK_ARG(score_smem(Tt, C, true) <= 200 * 1024, "%s: Tt=%d x C=%d needs %zu bytes of shared memory (max 200 KiB)",
                what, Tt, C, score_smem(Tt, C, true));
  return XVA_OK;
}

}  // namespace

int attn_score_fwd(const float* q, long ldq, const float* k, long ldk, const float* prior, const int* in_lens, int B,
                   int Tm, int Tt, int C, float* logprob, float* soft, cudaStream_t stream) {
  XVA_CHECK_ARG(q && k && prior && in_lens && logprob && soft, "attn_score_fwd: null pointer");
  if (int rc = check_score_shape("attn_score_fwd", B, Tm, Tt, C, ldq, ldk)) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(attn_score_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const dim3 grid(ceil_div(Tm, kRowsPerBlock), B);
  attn_score_fwd_kernel<<<grid, kWarps * 32, score_smem(Tt, C, false), stream>>>(q, ldq, k, ldk, prior, in_lens, Tm, Tt,
                                                                                  C, logprob, soft);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int attn_score_bwd(const float* g, const float* logprob, const float* prior, const float* q, long ldq, const float* k,
                   long ldk, int B, int Tm, int Tt, int C, float* dD, float* dq, long lddq, float* dk, long lddk,
                   cudaStream_t stream) {
  XVA_CHECK_ARG(g && logprob && prior && q && k && dD && dq && dk, "attn_score_bwd: null pointer");
  if (int rc = check_score_shape("attn_score_bwd", B, Tm, Tt, C, ldq, ldk)) return rc;
  XVA_CHECK_ARG(lddq >= C && lddk >= C, "attn_score_bwd: lddq=%ld lddk=%ld < C=%d", lddq, lddk, C);
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(attn_score_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const dim3 grid(ceil_div(Tm, kRowsPerBlock), B);
  attn_score_bwd_kernel<<<grid, kWarps * 32, score_smem(Tt, C, true), stream>>>(g, logprob, prior, q, ldq, k, ldk, Tm, Tt,
                                                                                 C, dD, dq, lddq);
  XVA_CHECK_LAUNCH();
  const dim3 grid_k(ceil_div(Tt, kKgJ), B);
  attn_key_grad_kernel<<<grid_k, kKgThreads, 0, stream>>>(dD, q, ldq, k, ldk, Tm, Tt, C, dk, lddk);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

// workspace: alpha and beta tables (fp64, B x Tm x (2 Tt + 1) each), nll (fp64, B), the row normalisers (fp32, B x Tm)
long long attn_ctc_workspace_bytes(int B, int Tm, int Tt) {
  const long long tables = 2LL * B * Tm * (2LL * Tt + 1) * static_cast<long long>(sizeof(double));
  return tables + static_cast<long long>(B) * 8 + (static_cast<long long>(B) * Tm * 4 + 7) / 8 * 8;
}

int attn_ctc(const float* logprob, const int* in_lens, const int* out_lens, int B, int Tm, int Tt, float blank_logprob,
             void* workspace, long long workspace_bytes, double* cost, float* grad, cudaStream_t stream) {
  XVA_CHECK_ARG(logprob && in_lens && out_lens && workspace && cost && grad, "attn_ctc: null pointer");
  XVA_CHECK_ARG(B >= 1 && Tm >= 1 && Tt >= 1, "attn_ctc: B=%d Tm=%d Tt=%d", B, Tm, Tt);
  XVA_CHECK_ARG(2 * Tt + 1 <= kCtcNS * kCtcMaxThreads, "attn_ctc: Tt=%d exceeds %d text positions", Tt,
                (kCtcNS * kCtcMaxThreads - 1) / 2);
  XVA_CHECK_ARG(workspace_bytes >= attn_ctc_workspace_bytes(B, Tm, Tt), "attn_ctc: workspace of %lld bytes, need %lld",
                workspace_bytes, attn_ctc_workspace_bytes(B, Tm, Tt));
  XVA_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "attn_ctc: workspace not 8-byte aligned");
  const int S_max = 2 * Tt + 1;
  double* table = static_cast<double*>(workspace);
  double* nll = table + 2LL * B * Tm * S_max;
  float* lse = reinterpret_cast<float*>(nll + B);
  const long rows = static_cast<long>(B) * Tm;
  ctc_row_lse_kernel<<<static_cast<unsigned>(ceil_div_l(rows, kWarps)), kWarps * 32, 0, stream>>>(
      logprob, in_lens, out_lens, Tm, Tt, blank_logprob, rows, lse);
  XVA_CHECK_LAUNCH();
  const size_t smem = 2 * static_cast<size_t>(S_max + 4) * sizeof(double);
  const int threads = S_max >= kCtcMaxThreads ? kCtcMaxThreads : round_up(S_max, 32);
  ctc_recursion_kernel<<<dim3(B, 2), threads, smem, stream>>>(logprob, in_lens, out_lens, lse, B, Tm, Tt, blank_logprob,
                                                              table, nll, cost);
  XVA_CHECK_LAUNCH();
  const long total = rows * Tt;
  const long blocks = ceil_div_l(total, 256);
  const int grid = static_cast<int>(blocks < 8L * num_sms() ? blocks : 8L * num_sms());
  ctc_grad_kernel<<<grid, 256, 0, stream>>>(logprob, in_lens, out_lens, lse, table, nll, B, Tm, Tt, grad);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int attn_bin_loss(const float* hard, const float* soft, long rows, int Tt, float eps, double* acc, cudaStream_t stream) {
  XVA_CHECK_ARG(hard && soft && acc && rows >= 1 && Tt >= 1, "attn_bin_loss: bad arguments (rows=%ld Tt=%d)", rows, Tt);
  const long blocks = ceil_div_l(rows, kWarps);
  const int grid = static_cast<int>(blocks < 4L * num_sms() ? blocks : 4L * num_sms());
  attn_bin_loss_kernel<<<grid, kWarps * 32, 0, stream>>>(hard, soft, rows, Tt, eps, acc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int attn_grad_combine(const float* gctc, const float* hard, const float* soft, const double* acc, float a, float bw,
                      float eps, long rows, int Tt, float* g, cudaStream_t stream) {
  XVA_CHECK_ARG(gctc && g && rows >= 1 && Tt >= 1, "attn_grad_combine: bad arguments (rows=%ld Tt=%d)", rows, Tt);
  XVA_CHECK_ARG(hard == nullptr || (soft && acc), "attn_grad_combine: hard without soft / acc");
  attn_grad_combine_kernel<<<static_cast<unsigned>(ceil_div_l(rows, kWarps)), kWarps * 32, 0, stream>>>(
      gctc, hard, soft, acc, a, bw, eps, rows, Tt, g);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(align)

}  // namespace xva
