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
// matrix held in shared memory (one pass over q, one write of each output), the CTC forward-backward of an utterance is
// one block walking the mel axis with one thread per extended-label state, and the gradients come out of two kernels.
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
// log(exp(a) + exp(b) + exp(c)): the maximum is carried in fp64, the (small) differences go through fp32 exp / log
__device__ __forceinline__ double lse3(double a, double b, double c) {
  const double m = fmax(a, fmax(b, c));
  if (m == -INFINITY) return m;
  const float s = expf(static_cast<float>(a - m)) + expf(static_cast<float>(b - m)) + expf(static_cast<float>(c - m));
  return m + static_cast<double>(logf(s));
}

constexpr int kCtcThreads = 256;

// One block per utterance. Extended label sequence of the target 1..L: state s (0 <= s < S = 2L+1) is the blank for
// even s and key (s-1)/2 for odd s; all keys are distinct, so s-2 -> s is allowed for every odd s >= 3.
//   z[t, .]   = [blank_logprob, logprob[t, 0..L)] ;  n[t, .] = log_softmax(z[t, .])       (attn_loss_function.py:28-36)
//   alpha_t(s) = n[t, lab(s)] + lse(alpha_{t-1}(s), alpha_{t-1}(s-1), alpha_{t-1}(s-2)*)
//   nll        = -lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2)),   cost = nll / max(L, 1)  (0 if nll is infinite)
//   beta likewise from the end;  d cost / d logprob[t, j] = (exp(n[t,j+1]) - exp(alpha_t(s) + beta_t(s) - n[t,j+1] + nll))
//   / max(L, 1) with s = 2j+1  (the gradient of -log p through the row's log_softmax; the blank column is a constant).
// grad receives d(mean_b cost_b)/d logprob, i.e. the above times 1/B, zero outside [T, L].
__global__ void __launch_bounds__(kCtcThreads)
attn_ctc_kernel(const float* __restrict__ logprob, const int* __restrict__ in_lens, const int* __restrict__ out_lens,
                int B, int Tm, int Tt, float blank, double* __restrict__ alpha_ws, double* __restrict__ cost,
                float* __restrict__ grad) {
  extern __shared__ double smem_d[];
  __shared__ double nll_s;
  const int b = blockIdx.x;
  const int L = min(max(in_lens[b], 0), Tt), T = min(max(out_lens[b], 0), Tm);
  const int S = 2 * L + 1, S_max = 2 * Tt + 1;
  double* buf0 = smem_d;                       // [S_max + 4], state s at index s + 2, two -inf guards on either side
  double* buf1 = buf0 + (S_max + 4);
  float* lse = reinterpret_cast<float*>(buf1 + (S_max + 4));   // [Tm]
  const float* lp = logprob + static_cast<long>(b) * Tm * Tt;
  float* gr = grad + static_cast<long>(b) * Tm * Tt;
  double* aw = alpha_ws + static_cast<long>(b) * Tm * S_max;
  const double ninf = -INFINITY;
  for (long i = threadIdx.x; i < static_cast<long>(Tm) * Tt; i += blockDim.x) gr[i] = 0.0f;
  if (T == 0) {
    if (threadIdx.x == 0) cost[b] = 0.0;
    return;
  }
  // ---- per-row normaliser of [blank, keys < L]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int t = warp; t < T; t += n_warps) {
    float mx = blank;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, lp[static_cast<long>(t) * Tt + j]);
    mx = warp_max(mx);
    float se = (lane == 0) ? expf(blank - mx) : 0.0f;
    for (int j = lane; j < L; j += 32) se += expf(lp[static_cast<long>(t) * Tt + j] - mx);
    se = warp_sum(se);
    if (lane == 0) lse[t] = mx + logf(se);
  }
  for (int i = threadIdx.x; i < 2 * (S_max + 4); i += blockDim.x) buf0[i] = ninf;
  __syncthreads();
  auto emit = [&](int t, int s) -> double {
    const float v = (s & 1) ? lp[static_cast<long>(t) * Tt + (s >> 1)] : blank;
    return static_cast<double>(v - lse[t]);
  };
  // ---- alpha
  double* prev = buf0 + 2;
  double* cur = buf1 + 2;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const double v = (s <= 1) ? emit(0, s) : ninf;
    prev[s] = v;
    aw[s] = v;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      const double c2 = (s & 1) ? prev[s - 2] : ninf;
      const double v = emit(t, s) + lse3(prev[s], prev[s - 1], c2);
      cur[s] = v;
      aw[static_cast<long>(t) * S_max + s] = v;
    }
    __syncthreads();
    double* sw = prev;
    prev = cur;
    cur = sw;
  }
  if (threadIdx.x == 0) nll_s = -lse3(prev[S - 1], prev[S - 2], ninf);   // S = 1: prev[-1] is a guard
  __syncthreads();
  const double nll = nll_s;
  const double inv_len = 1.0 / static_cast<double>(max(L, 1));
  if (!(nll < INFINITY) || nll != nll) {        // zero_infinity (nn.CTCLoss(zero_infinity=True)): no cost, no gradient
    if (threadIdx.x == 0) cost[b] = 0.0;
    return;
  }
  if (threadIdx.x == 0) cost[b] = nll * inv_len;
  const float w = static_cast<float>(inv_len / static_cast<double>(B));
  // ---- beta and the gradient, last frame first. The buffers are reused: reset the guards' neighbours first.
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (S_max + 4); i += blockDim.x) buf0[i] = ninf;
  __syncthreads();
  double* nxt = buf0 + 2;
  cur = buf1 + 2;
  for (int t = T - 1; t >= 0; --t) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      double v;
      if (t == T - 1) {
        v = (s >= S - 2) ? emit(t, s) : ninf;
      } else {
        const double c2 = (s & 1) ? nxt[s + 2] : ninf;
        v = emit(t, s) + lse3(nxt[s], nxt[s + 1], c2);
      }
      cur[s] = v;
      if (s & 1) {
        const int j = s >> 1;
        const float n_tj = lp[static_cast<long>(t) * Tt + j] - lse[t];
        const float post = expf(static_cast<float>(aw[static_cast<long>(t) * S_max + s] + v + nll) - n_tj);
        gr[static_cast<long>(t) * Tt + j] = w * (expf(n_tj) - post);
      }
    }
    __syncthreads();
    double* sw = nxt;
    nxt = cur;
    cur = sw;
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

long long attn_ctc_workspace_bytes(int B, int Tm, int Tt) {
  return static_cast<long long>(B) * Tm * (2LL * Tt + 1) * static_cast<long long>(sizeof(double));
}

int attn_ctc(const float* logprob, const int* in_lens, const int* out_lens, int B, int Tm, int Tt, float blank_logprob,
             void* workspace, long long workspace_bytes, double* cost, float* grad, cudaStream_t stream) {
  XVA_CHECK_ARG(logprob && in_lens && out_lens && workspace && cost && grad, "attn_ctc: null pointer");
  XVA_CHECK_ARG(B >= 1 && Tm >= 1 && Tt >= 1, "attn_ctc: B=%d Tm=%d Tt=%d", B, Tm, Tt);
  XVA_CHECK_ARG(workspace_bytes >= attn_ctc_workspace_bytes(B, Tm, Tt), "attn_ctc: workspace of %lld bytes, need %lld",
                workspace_bytes, attn_ctc_workspace_bytes(B, Tm, Tt));
  XVA_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "attn_ctc: workspace not 8-byte aligned");
  const size_t smem = 2 * (2 * static_cast<size_t>(Tt) + 5) * sizeof(double) + static_cast<size_t>(Tm) * sizeof(float);
  XVA_CHECK_ARG(smem <= 200 * 1024, "attn_ctc: Tm=%d Tt=%d needs %zu bytes of shared memory (max 200 KiB)", Tm, Tt, smem);
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(attn_ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  attn_ctc_kernel<<<B, kCtcThreads, smem, stream>>>(logprob, in_lens, out_lens, B, Tm, Tt, blank_logprob,
                                                    static_cast<double*>(workspace), cost, grad);
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
