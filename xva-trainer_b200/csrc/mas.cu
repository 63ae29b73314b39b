// Monotonic alignment search (Viterbi over a mel x text soft alignment) -- the integer index path of FastPitch training
// stage 1: replaces b_mas / mas_width1, python/fastpitch1_1/fastpitch/alignment.py:79-118, which the reference runs with
// numba on the CPU after a device->host copy of the [B, 1, Tm, Tt] attention (model.py:283-294) and copies back.
//
// One block per utterance, one thread per text position j; the mel axis is sequential:
//   log_p[0, j]  = log a[0, j] for j = 0, -inf otherwise                                  (alignment.py:86-88)
//   log_p[i, j]  = log a[i, j] + (log_p[i-1, j-1] >= log_p[i-1, j] ? log_p[i-1, j-1] : log_p[i-1, j])   (:91-100)
//   back-track from (Tm-1, Tt-1) along the stored choices                                  (:103-108)
// The recurrence is two fp32 operations per cell (one compare, one add, no fused multiply), so given the same fp32
// log-probabilities the path is bit-identical to the reference's; the stay / advance choice is kept as one bit per cell
// in shared memory (Tm x Tt / 8 bytes: 17.6 KB at 880 x 160). Outputs: the hard alignment (0 / 1, same layout as the
// input, zero outside [out_len, in_len]) and the per-token durations (its column sums, model.py:318) as int32.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

constexpr int kMasThreads = 256;

__global__ void __launch_bounds__(kMasThreads)
mas_kernel(const float* __restrict__ attn, const int* __restrict__ in_lens, const int* __restrict__ out_lens, int Tm,
           int Tt, int is_log, float* __restrict__ hard, int* __restrict__ durs) {
  extern __shared__ unsigned int smem_u[];
  const int b = blockIdx.x;
  const int n_txt = min(in_lens[b], Tt), n_mel = min(out_lens[b], Tm);
  const int words = (Tt + 31) / 32;                       // choice bits per mel row
  unsigned int* bits = smem_u;                            // [Tm][words]
  float* row = reinterpret_cast<float*>(smem_u + static_cast<size_t>(Tm) * words);  // [2][Tt + 1], row[.][0] = -inf guard
  const float* a = attn + static_cast<long>(b) * Tm * Tt;
  float* h = hard + static_cast<long>(b) * Tm * Tt;
  for (long i = threadIdx.x; i < static_cast<long>(Tm) * Tt; i += blockDim.x) h[i] = 0.0f;
  for (int j = threadIdx.x; j < Tt; j += blockDim.x) durs[static_cast<long>(b) * Tt + j] = 0;
  if (n_txt <= 0 || n_mel <= 0) return;
  const float ninf = -INFINITY;
  // log of a probability: the reference takes np.log of the fp32 array (alignment.py:85); here the double-precision
  // logarithm rounded to fp32 (correctly rounded in all but ~1e-8 of cases), so near-ties resolve as they would with an
  // accurate libm. Pass is_log = 1 to hand in log-probabilities and make the path independent of any logarithm.
  // is_log carries two flags: bit 0 = the input already holds log-probabilities; bit 1 = xVAPitch's maximum_path
  // (python/xvapitch/util.py:14-53): on an exact tie between "stay" and "advance" it stays (v1 >= v0, :35) where
  // FastPitch advances (alignment.py:96), and it has no extra mark in row 0 when the back-track stops short of token 0.
  const bool stay_on_tie = (is_log & 2) != 0;
  is_log &= 1;
  auto lg = [&](float v) { return is_log ? v : static_cast<float>(log(static_cast<double>(v))); };
  float* cur = row;
  float* nxt = row + (Tt + 1);
  for (int j = threadIdx.x; j <= Tt; j += blockDim.x) {
    cur[j] = (j == 1) ? lg(a[0]) : ninf;                   // text position j lives at index j + 1; index 0 = -inf guard
    nxt[j] = ninf;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // the mel axis is sequential and each step is a handful of instructions, so a global load inside the step would be
  // most of its latency: the value of row i + 1 (first 256 text positions) is fetched while row i is processed
  const bool pre_lane = static_cast<int>(threadIdx.x) < n_txt;
  float a_pre = (pre_lane && n_mel > 1) ? a[Tt + threadIdx.x] : 0.0f;
  for (int i = 1; i < n_mel; ++i) {
    const float a_cur = a_pre;
    if (pre_lane && i + 1 < n_mel) a_pre = a[static_cast<long>(i + 1) * Tt + threadIdx.x];
    for (int jb = 0; jb < n_txt; jb += blockDim.x) {       // uniform trip count: every thread reaches the ballot
      const int j = jb + threadIdx.x;
      bool take_adv = false;
      if (j < n_txt) {
        const float la = lg(jb == 0 ? a_cur : a[static_cast<long>(i) * Tt + j]);
        const float stay = cur[j + 1], adv = cur[j];       // adv = log_p[i-1, j-1]
        take_adv = (j >= 1) && (stay_on_tie ? (adv > stay) : (adv >= stay));   // alignment.py:96 | util.py:35
        nxt[j + 1] = __fadd_rn(la, take_adv ? adv : stay);
      }
      const unsigned int m = __ballot_sync(0xffffffffu, take_adv);
      if (lane == 0 && j < n_txt) bits[static_cast<size_t>(i) * words + (j >> 5)] = m;
    }
    __syncthreads();
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int j = n_txt - 1;
    int* d = durs + static_cast<long>(b) * Tt;
    int run = 0;                                            // marks in the current column (the path never returns to one)
    for (int i = n_mel - 1; i >= 0; --i) {
      h[static_cast<long>(i) * Tt + j] = 1.0f;
      ++run;
      if (i > 0 && ((bits[static_cast<size_t>(i) * words + (j >> 5)] >> (j & 31)) & 1u)) {
        d[j] = run;                                         // stores only: a read-modify-write per frame was half the kernel
        run = 0;
        --j;
      }
    }
    d[j] = run;
    // alignment.py:106-108: prev_ind[0, :] is 0, so after the loop curr_text_idx = 0 and opt[0, 0] is set as well -- a
    // second mark in row 0 when the back-track did not reach text position 0 (fewer mel frames than tokens)
    if (!stay_on_tie && h[0] == 0.0f) {
      h[0] = 1.0f;
      d[0] += 1;
    }
  }
}

// out = (float) log((double) a): the logarithm mas_kernel takes with is_log = 0, for every element at once
__global__ void __launch_bounds__(256) mas_log_kernel(const float* __restrict__ a, long n, float* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(log(static_cast<double>(a[i])));
}

}  // namespace

int mas_log(const float* attn, long n, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(attn && out && n >= 1, "mas_log: bad arguments (n=%ld)", n);
  const long blocks = ceil_div_l(n, 256);
  const int grid = static_cast<int>(blocks < 16L * num_sms() ? blocks : 16L * num_sms());
  mas_log_kernel<<<grid, 256, 0, stream>>>(attn, n, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int mas_width1(const float* attn, const int* in_lens, const int* out_lens, int B, int Tm, int Tt, int is_log, float* hard,
               int* durs, cudaStream_t stream) {
  XVA_CHECK_ARG(B >= 1 && Tm >= 1 && Tt >= 1, "mas: B=%d Tm=%d Tt=%d", B, Tm, Tt);
  const size_t smem = (static_cast<size_t>(Tm) * ((Tt + 31) / 32) + 2 * (Tt + 1)) * 4;
  XVA_CHECK_ARG(smem <= 200 * 1024, "mas: Tm=%d x Tt=%d needs %zu bytes of shared memory (max 200 KiB)", Tm, Tt, smem);
  static bool attr_done = false;
  if (!attr_done) {
    XVA_CHECK_CUDA(cudaFuncSetAttribute(mas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  mas_kernel<<<B, kMasThreads, smem, stream>>>(attn, in_lens, out_lens, Tm, Tt, is_log, hard, durs);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

}  // namespace xva
