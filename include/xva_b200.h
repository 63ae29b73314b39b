/* libxva_b200 -- C ABI of the B200-native FastPitch 1.1 / HiFi-GAN training hot path.
 *
 * The reference (DanRuta/xva-trainer) has no FFI of its own: its operator API is Python (nn.Module methods and
 * free functions). Each entry point below names the reference interface it stands in for (file:line relative to
 * the reference tree); the Python host side in xva-trainer_b200/ keeps the reference's names and signatures and
 * calls these through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator in practice), unless the
 *     parameter name ends in _host; the library allocates no device memory and keeps no pointer after return;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the device;
 *   - return 0 on success, negative on failure; the message is in xva_last_error() (thread-local);
 *   - activations are channels-last: [batch, time, channels], fp32; "lens" are int32;
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every op fails.
 */
#ifndef XVA_B200_H_
#define XVA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XVA_ABI_VERSION 1
#define XVA_MAX_TAPS 48

int xva_abi_version(void);
const char* xva_last_error(void);
/* 0 if `device` is a compute-capability 10.x GPU, negative otherwise. */
int xva_device_check(int device);

/* ------------------------------------------------------------------------------------------------------------
 * Tap-GEMM: the one dense-contraction kernel family (tcgen05 + TMA, tf32 operands, fp32 accumulate in TMEM).
 *
 *  mode 0  out[z,r,n] = alpha * sum_j sum_k A[z, r+shift[j], k] * B[zb, n, k]      zb = j*b_tap_z + z*b_batch_z
 *          replaces nn.Linear / nn.Conv1d forward: fastpitch/transformer.py:46-52,109,138 ; common/layers.py:94 ;
 *          fastpitch/model.py:118-122,386 ; hifigan/models.py:21-37,87,111 ; and torch.bmm(q, k^T) transformer.py:118
 *  mode 1  out[z,r,n] = alpha * sum_j sum_k A[z, r+shift[j], k] * B[zb, k, n]
 *          replaces the autograd input-gradient of the same layers and torch.bmm(attn_prob, v) transformer.py:130
 *  mode 2  out[zo,j,m,n] (+)= sum_{zr<ZR} sum_t A[z,t,m] * B[z, t+shift[j], n]        z = zo*ZR + zr
 *          replaces the autograd weight-gradient of the same layers, and dK / dV of the attention
 *
 *  Rows addressed outside an operand's [0, rows) read as zero -- this is the Conv1d zero padding.
 *  MN-major operands (B in mode 1; A and B in mode 2) are fetched in 32-column chunks: when their column count
 *  (N, or M) is not a multiple of 32 the row stride must be >= the count rounded up to 32 (pad columns are
 *  multiplied into output rows/columns that are never stored).
 *  Epilogue (mode 0/1), in order: alpha, +bias[n], (leaky) ReLU, gate ((leaky-)ReLU backward), dropout(pre),
 *  +residual, [LayerNorm(gamma,beta) over n, dropout(post)], [tanh], zero rows >= lens[z], [round to tf32];
 *  out_act (optional) additionally receives leaky_relu(out).
 *  Operand precision: the MMA reads fp32 operands as tf32 by TRUNCATION. Every producer of a GEMM operand in this
 *  library (XVA_GEMM_ROUND_OUT, the softmax / LayerNorm-backward / embedding / loss-gradient kernels, and the tf32
 *  weight copy kept by xva_lamb_step / xva_round_tf32) therefore stores it already rounded to nearest.
 * ---------------------------------------------------------------------------------------------------------- */
enum {
  XVA_GEMM_RELU = 1 << 0,
  XVA_GEMM_LN = 1 << 1,
  XVA_GEMM_DROP_PRE = 1 << 2,
  XVA_GEMM_DROP_POST = 1 << 3,
  XVA_GEMM_ATOMIC = 1 << 4,
  XVA_GEMM_LRELU_GATE = 1 << 5,
  XVA_GEMM_TANH = 1 << 7,     /* out = tanh(value) as the last step (Generator.forward, hifigan/models.py:126) */
  XVA_GEMM_HALO = 1 << 9,     /* k-tap convolution (mode 0/1, un-segmented tiles, equal a_col): fetch the activation tile once
                                 per k-block with halo rows and read the taps through row-shifted descriptors. Parity-green
                                 but measured 5 % SLOWER than one fetch per tap on B200 (13.76 vs 14.33 ms/step), so opt-in */
  XVA_GEMM_SOFTMAX_BWD = 1 << 8, /* softmax + attention-dropout backward fused into the dP = dO.V^T product
                                    (autograd of transformer.py:120-128): out = alpha * P * (acc * dropmask - rowvec[row])
                                    with P read through the `gate` slot, rowvec[z*R + r] = sum_j P_d[r,j] dP_d[r,j] = dO[r].O[r],
                                    the dropout mask of xva_softmax_fwd (element index (z*R + r) * drop_ld + n) */
  XVA_GEMM_ROUND_OUT = 1 << 6 /* store `out` rounded to tf32 (nearest): set when the result is a later GEMM operand */
};

typedef struct xva_gemm_args {
  int32_t mode;
  int32_t Z;      /* batch items */
  int32_t R;      /* rows per item: output rows (mode 0/1) or contraction rows (mode 2) */
  int32_t M;      /* mode 2: output rows */
  int32_t N;      /* output columns */
  int32_t K;      /* mode 0/1: contraction length per tap */
  int32_t taps;
  int32_t shift[XVA_MAX_TAPS];
  int32_t ZR;     /* mode 2: consecutive z reduced into one output batch */
  int32_t split;  /* mode 2: CTAs sharing one output tile (needs XVA_GEMM_ATOMIC when > 1) */

  const float* a;
  int64_t a_rs, a_zs; /* element strides: row, batch item */
  int32_t a_rows;     /* valid rows of A per item (0 = R) */
  int32_t _pad0;
  const float* b;
  int64_t b_rs, b_zs;
  int32_t b_rows;     /* valid rows of B per item: mode 0 rows n (0 = N), mode 1 rows k (0 = K), mode 2 rows t (0 = R);
                         rows past it read as zero */
  int32_t b_nz;       /* z slices in B */
  int32_t b_tap_z, b_batch_z;

  float* out;
  int64_t o_rs, o_zs, o_js;

  float alpha;
  int32_t flags;
  const float* bias;
  const float* residual;
  int64_t r_rs, r_zs;
  const float* gate;
  int64_t g_rs, g_zs;
  float gate_slope;
  float act_slope;    /* XVA_GEMM_RELU computes v > 0 ? v : act_slope * v (0 = ReLU, 0.1 = the HiFi-GAN leaky ReLU) */
  const int32_t* lens;
  const float* gamma;
  const float* beta;
  float ln_eps;
  int32_t _pad2;
  float* out_pre;   /* pre-LayerNorm value, same strides as out (optional) */
  float* ln_mean;   /* [Z*R] (optional) */
  float* ln_rstd;
  float drop_p;
  float out_act_slope;
  uint64_t seed;
  float* out_act;   /* optional second output, same strides as out: leaky_relu(out, out_act_slope) rounded to tf32 --
                       the operand the next convolution reads (ResBlock1.forward, hifigan/models.py:41-48) */
  int32_t a_col[XVA_MAX_TAPS]; /* per-tap column offset of the activation operand (A in mode 0/1, B in mode 2, there a
                                  multiple of 32): strided convolutions read a [T/stride, stride*C] view of their
                                  input, tap = (row shift, phase*C + group offset) */
  const uint64_t* seed_dev; /* optional device counter added to `seed` (x odd constant) at run time, so a captured
                               CUDA graph draws a fresh dropout mask on every replay */
  int32_t groups;   /* grouped convolution (DiscriminatorS, hifigan/models.py:207-213) in ONE launch; 0 / 1 = dense.
                       With G = groups the output-tile index along n (mode 0/1) or m (mode 2) is the group g:
                         mode 0: N = G*Og, K = Cg.  out[.., g*Og+n] contracts A columns a_col[j] + g*grp_step + [0,K)
                                 with B rows g*Og+n (B = packed weights [taps, N, K]); Og % 16 == 0, Og <= 256
                         mode 1: N = G*Cg, K = Og.  out[.., g*Cg+n] contracts A columns a_col[j] + g*K + [0,K) with B rows
                                 g*K + [0,K), B columns [0,Cg) (B = the same packed weights); K % 32 == 0, Cg % 32 == 0
                         mode 2: M = G*Og, N = Cg.  out[j, g*Og+m, n] contracts A columns g*Og+m with B columns
                                 a_col[j] + g*grp_step + n; Og % 32 == 0, Og <= 128, grp_step % 32 == 0 */
  int32_t grp_step; /* column step per group in the activation operand (mode 0: Cg of the input; mode 2: Cg) */
  const float* rowvec; /* XVA_GEMM_SOFTMAX_BWD: per-row scalar [Z*R] */
  int32_t drop_ld;     /* XVA_GEMM_SOFTMAX_BWD: row pitch of the dropout element index (0 = N) */
  int32_t _pad3;
} xva_gemm_args;

/* sizeof(xva_gemm_args) as compiled into the library, so a binding can verify its struct layout. */
int xva_sizeof_gemm_args(void);
int xva_gemm(const xva_gemm_args* args, void* stream);
/* Bring-up aid: eight cycle counters of CTA 0 from the last xva_gemm launched with XVA_GEMM_DBG & 32 in the
 * environment (role wait / work times; layout in csrc/gemm_tc.cu). Synchronises the device. */
int xva_gemm_debug_counters(long long* out8_host);
/* Same contract on CUDA cores in exact fp32: a checker for the tests, never used by the product path. */
int xva_gemm_ref(const xva_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Length regulator -- replaces regulate_len(), fastpitch/model.py:59-79 (integer index path, bit-exact).
 *   scan : reps = trunc(durs*pace + 0.5); cum[b,0..Tt] = exclusive int32 prefix sum; dec_lens[b] = min(sum, mel_max_len)
 *          (mel_max_len < 0: no clamp)
 *   fwd  : out[b,t,:] = enc[b,j,:] with cum[b,j] <= t < cum[b,j+1], zero for t >= cum[b,Tt]; idx[b,t] = j or -1 (optional)
 *   bwd  : denc[b,j,:] (+)= sum_{t in token j, t < T_out} dout[b,t,:]
 * ---------------------------------------------------------------------------------------------------------- */
int xva_regulate_len_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int32_t* cum,
                          int32_t* dec_lens, void* stream);
int xva_regulate_len_fwd(const float* enc, const int32_t* cum, int B, int Tt, int C, int T_out, float* out,
                         int32_t* idx, void* stream);
int xva_regulate_len_bwd(const float* dout, const int32_t* cum, int B, int Tt, int C, int T_out, float* denc,
                         int accumulate, void* stream);

/* average_pitch(), fastpitch/model.py:82-100: per-token mean of the non-zero frame values. pitch [B,F,Tm],
 * durs [B,Tt] -> out [B,F,Tt]. log1p_out = 1 writes log(1 + mean) (the energy target, model.py:415). */
int xva_average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, int log1p_out,
                      void* stream);

/* out[row] = dot(a[row, 0..C), b[row, 0..C)) with row pitches a_ld / b_ld: the per-row term of the fused softmax backward
 * (XVA_GEMM_SOFTMAX_BWD): sum_j P_d[r,j] * dP_d[r,j] = dO[r] . O[r] for O = P_d V. */
int xva_rowdot2(const float* a, const float* b, int64_t rows, int C, int64_t a_ld, int64_t b_ld, float* out, void* stream);

/* Fused single-head attention of an FFT block -- replaces the body of MultiHeadAttn._forward between the qkv projection
 * and the output projection, fastpitch/transformer.py:113-133 (n_head = 1, d_head = 64): scores q.k^T * scale, key mask
 * (keys >= lens[b]), softmax, attention dropout and the product with v in ONE kernel; the [B, T, T] score / probability
 * tensors live only in tensor memory (tcgen05.mma accumulates S there, the softmax threads overwrite it with P, a second
 * tcgen05.mma reads P from there). qkv [B, T, >= 192] fp32 holding tf32-rounded values, columns q 0..63 | k 64..127 |
 * v 128..191, row stride rs, item stride zs (elements). out [B, T, 64] (strides o_rs, o_zs; stored tf32-rounded: it is the
 * operand of the output projection), lse [B * T] = log sum_j exp(scale * q.k_j) over the unmasked keys (+inf for a row
 * without any), what the backward recomputes P from. Dropout: element (b, row, key) is dropped by the shared counter
 * hash of ((b * T + row) * drop_ld + key) with (seed, *seed_dev) -- the index the unfused xva_softmax kernels use. */
int xva_attn_fwd(const float* qkv, int64_t rs, int64_t zs, int B, int T, const int32_t* lens, float scale, float drop_p,
                 uint64_t seed, const uint64_t* seed_dev, int drop_ld, float* out, int64_t o_rs, int64_t o_zs, float* lse,
                 void* stream);

/* Backward of xva_attn_fwd (the autograd of transformer.py:113-133), two launches: dq per 128-row query tile, dk / dv per
 * 128-key tile, both recomputing P = exp(scale * q.k - lse[row]) and the dropout mask, with S, dP, dS and P in tensor
 * memory. dout [B, T, 64] = gradient of the attention output (strides d_rs, d_zs; tf32-rounded GEMM operand), lse from the
 * forward, dsum [B * T] = dot(dout[row], out[row]) (xva_rowdot2). dqkv [B, T, >= 192] (strides g_rs, g_zs) receives
 * dq | dk | dv in the column layout of qkv, tf32-rounded. Same dropout arguments as the forward call. */
int xva_attn_bwd(const float* qkv, int64_t rs, int64_t zs, const float* dout, int64_t d_rs, int64_t d_zs, const float* lse,
                 const float* dsum, int B, int T, const int32_t* lens, float scale, float drop_p, uint64_t seed,
                 const uint64_t* seed_dev, int drop_ld, float* dqkv, int64_t g_rs, int64_t g_zs, void* stream);

/* Monotonic alignment search -- replaces b_mas / mas_width1, fastpitch/alignment.py:79-118 (called through
 * FastPitch.binarize_attention_parallel, model.py:283-294, after a device->host copy; training stage 1): Viterbi path
 * through the soft alignment attn [B, Tm, Tt] (mel x text, the reference's [B, 1, Tm, Tt]) restricted to
 * [out_lens[b], in_lens[b]]. hard [B, Tm, Tt] gets the 0/1 alignment (zero elsewhere), durs [B, Tt] its column sums
 * (attn_hard.sum(2), model.py:318) as int32. is_log = 0: attn holds probabilities (the reference's input; log taken in
 * double and rounded to fp32); is_log = 1: attn holds fp32 log-probabilities and the path is bit-identical to the
 * reference recurrence on the same values. is_log | 2: the search of xVAPitch (maximum_path, xvapitch/util.py:14-53; SURVEY
 * 8f rank 1) on the same layout -- log-likelihoods in, "stay" preferred on an exact tie (util.py:35) where FastPitch
 * advances, no second mark in row 0. Tm x ceil(Tt/32) x 4 bytes of shared memory (<= 200 KiB). */
int xva_mas_width1(const float* attn, const int32_t* in_lens, const int32_t* out_lens, int B, int Tm, int Tt, int is_log,
                   float* hard, int32_t* durs, void* stream);
/* out[i] = (float) log((double) attn[i]), i < n: the logarithm xva_mas_width1 takes element by element with is_log = 0,
 * as one parallel pass, so that the (sequential) search can run on log-probabilities (is_log = 1) with the identical
 * result -- the double-precision log sat on the critical path of every one of its Tm steps. */
int xva_mas_log(const float* attn, int64_t n, float* out, void* stream);

/* Stage-1 aligner score -- replaces the body of ConvAttention.forward after the two projection stacks,
 * fastpitch/attention.py:203-219. q [B, Tm, C] (row pitch ldq) = query_proj(mel), k [B, Tt, C] (row pitch ldk) =
 * key_proj(text embedding), prior [B, Tm, Tt], in_lens [B] (the key mask of model.py:303):
 *   logprob[b,t,j] = log_softmax_j(-0.0005 * sum_c (q[b,t,c] - k[b,j,c])^2) + log(prior[b,t,j] + 1e-8)    (attn_logprob)
 *   soft[b,t,j]    = softmax over j < in_lens[b] of logprob[b,t,:], 0 for the padded keys                 (attn_soft)
 * both [B, Tm, Tt] (the reference's [B, 1, Tm, Tt]). C <= 96, Tt <= 512; the keys of one utterance sit in shared memory.
 * bwd: g = d(loss)/d(logprob) with every contribution summed (xva_attn_grad_combine) -> dD (the gradient of the raw
 * score, [B, Tm, Tt], may alias g), dq [B, Tm, C] (row pitch lddq), dk [B, Tt, C] (row pitch lddk); two launches. */
int xva_attn_score_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* prior, const int32_t* in_lens,
                       int B, int Tm, int Tt, int C, float* logprob, float* soft, void* stream);
int xva_attn_score_bwd(const float* g, const float* logprob, const float* prior, const float* q, int64_t ldq,
                       const float* k, int64_t ldk, int B, int Tm, int Tt, int C, float* dD, float* dq, int64_t lddq,
                       float* dk, int64_t lddk, void* stream);

/* AttentionCTCLoss.forward, fastpitch/attn_loss_function.py:20-44, and its autograd, for the whole batch in one launch
 * (the reference loops over the batch in Python: 32 F.ctc_loss calls and as many host syncs). Per utterance b, with
 * L = in_lens[b], T = out_lens[b]: rows t < T of [blank_logprob, logprob[b,t,0..L)] are log_softmax-ed and scored by CTC
 * against the target 1..L; cost[b] = -log p / max(L, 1) (nn.CTCLoss reduction 'mean' on a batch of one), 0 when the
 * alignment is impossible (zero_infinity=True). The reference's loss is mean_b cost[b]; grad [B, Tm, Tt] receives
 * d(mean_b cost[b]) / d logprob (zero outside [T, L]). workspace: xva_attn_ctc_workspace_bytes(B, Tm, Tt) bytes, 8-byte
 * aligned (the fp64 forward and backward variables, 2 x B x Tm x (2 Tt + 1), and the row normalisers); caller-owned
 * like every other buffer. Three launches: row normalisers, the alpha and beta recursions side by side (2 B blocks),
 * the gradient. Tt <= 1023. */
int64_t xva_attn_ctc_workspace_bytes(int B, int Tm, int Tt);
int xva_attn_ctc(const float* logprob, const int32_t* in_lens, const int32_t* out_lens, int B, int Tm, int Tt,
                 float blank_logprob, void* workspace, int64_t workspace_bytes, double* cost, float* grad, void* stream);

/* AttentionBinarizationLoss.forward, fastpitch/attn_loss_function.py:47-54: acc (double[2], ACCUMULATED) +=
 * {sum over hard == 1 of log(max(soft, eps)), sum of hard}; the loss is -acc[0] / acc[1]. hard, soft [rows, Tt]. */
int xva_attn_bin_loss(const float* hard, const float* soft, int64_t rows, int Tt, float eps, double* acc, void* stream);

/* d(loss)/d(logprob) of FastPitchTrainer.iteration's stage-1 loss, xva_train.py:790-798:
 *   g = a * gctc + (bw / acc[1]) * (soft * rowsum(h') - h'),  h' = hard * [soft >= eps]
 * a = loss scale * attn_loss_scale, bw = loss scale * kl_weight (the binarization term flows back through the masked
 * softmax that produced soft). hard = NULL (or bw = 0): g = a * gctc. All [rows, Tt]. */
int xva_attn_grad_combine(const float* gctc, const float* hard, const float* soft, const double* acc, float a, float bw,
                          float eps, int64_t rows, int Tt, float* g, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention softmax -- replaces masked_fill + F.softmax + dropatt, fastpitch/transformer.py:120-127, and its autograd.
 *   fwd : s [Z,R,N] = alpha*q.k^T (from xva_gemm) -> p (softmax over n with keys n >= lens[z] masked), and
 *         pd = p * dropout (optional; pass NULL when drop_p == 0 and use p for the P*V product)
 *   bwd : dpd [Z,R,N] (gradient wrt pd) is overwritten with alpha * ds (gradient wrt q.k^T)
 *   ld  : row stride of s / p / pd / dpd in elements (0 = N). Pad columns [N, ld) are written as zero, so a stride
 *         rounded up to 32 makes the tensors legal MN-major operands of xva_gemm.
 * ---------------------------------------------------------------------------------------------------------- */
int xva_softmax_fwd(const float* s, const int32_t* lens, int Z, int R, int N, int ld, float* p, float* pd,
                    float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream);
int xva_softmax_bwd(const float* p, float* dpd, int Z, int R, int N, int ld, float alpha, float drop_p,
                    uint64_t seed, const uint64_t* seed_dev, void* stream);

/* LayerNorm backward for the LayerNorm epilogue of xva_gemm (nn.LayerNorm autograd, transformer.py:75,148 and
 * common/layers.py:96). x/mean/rstd are the out_pre/ln_mean/ln_rstd the forward saved; rows >= lens[z] get zero.
 * dx_drop (optional) = dx * dropout(pre) mask of the forward; dgamma/dbeta/dbias (optional) are ACCUMULATED.
 * relu_gate = 1: x is a ReLU output (ConvReLUNorm), dx is zeroed where x <= 0. seed_dev: see xva_gemm_args. */
int xva_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const int32_t* lens, int Z, int R, int C, float* dx, float* dx_drop, float* dgamma,
                      float* dbeta, float* dbias, float drop_post_p, uint64_t seed_post, float drop_pre_p,
                      uint64_t seed_pre, const uint64_t* seed_dev, int relu_gate, void* stream);

/* nn.LayerNorm forward (transformer.py:75,148) on a stored pre-LN tensor x [Z,R,C] (C % 4 == 0, <= 1024):
 * y = ((x - mean) * rstd * gamma + beta) for rows r < lens[z] (lens optional), zero otherwise, stored tf32-rounded (it is
 * the next GEMM's operand); mean / rstd [Z*R] are saved for xva_layernorm_bwd. The FFT blocks use this after a GEMM
 * whose epilogue did bias + dropout + residual: un-fused, the GEMM keeps double-buffered accumulators. */
int xva_layernorm_fwd(const float* x, const float* gamma, const float* beta, const int32_t* lens, int Z, int R, int C,
                      float eps, float* y, float* mean, float* rstd, void* stream);

/* Test switch, default on: 0 makes every kernel store GEMM operands unrounded (and xva_round_tf32 a plain copy), so
 * that the exact-fp32 checker xva_gemm_ref reproduces an fp32 reference to rounding. Not for production use:
 * the tensor-core path then truncates its operands. Synchronous (cudaMemcpyToSymbol). */
int xva_set_operand_rounding(int on);

/* dst[i] = tf32-rounded src[i] (round to nearest): refreshes the GEMM-operand copy of the parameters after a load. */
int xva_round_tf32(const float* src, float* dst, int64_t n, void* stream);

/* *counter += inc on the stream: the per-step dropout counter every `seed_dev` argument points at. */
int xva_counter_add(uint64_t* counter, uint64_t inc, void* stream);

/* out[n] += sum over rows of x[row*ld + n]   (bias gradients). */
int xva_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* stream);

/* FFTransformer input stage, transformer.py:212-227: out = (tokens ? emb[tokens] : in) + pos_emb(t)*mask.
 * Encoder: tokens int64 [B,T] + emb [n,C] (mask = token != 0). Decoder: in [B,T,C] + lens (mask = t < lens[b]).
 * inv_freq = NULL: no positional term -- the plain encoder.word_emb(inputs) the stage-1 aligner reads (model.py:299). */
int xva_embed_pos(const int64_t* tokens, const float* emb, const float* in, const int32_t* lens,
                  const float* inv_freq, int B, int T, int C, float* out, void* stream);
int xva_embed_bwd(const int64_t* tokens, const float* dout, int B, int T, int C, float* demb, void* stream);

/* pitch_emb / energy_emb = nn.Conv1d(1, C, 3, padding=1), model.py:403-404,417-418:
 *   io[b,t,:] += bias + sum_j w[:,j] * x[b,t+j-1] for t < lens[b] (lens optional; padded token rows are left as they
 *   are -- nothing downstream reads them: the predictors mask their input, model.py:119, and padded tokens have zero
 *   duration) ; bwd accumulates dw [C,3] and dbias [C]. */
int xva_scalar_conv_add(float* io, const float* x, const float* w, const float* bias, const int32_t* lens, int B,
                        int T, int C, void* stream);
int xva_scalar_conv_bwd(const float* dout, const float* x, int B, int T, int C, float* dw, float* dbias, void* stream);

/* TemporalPredictor.fc (C -> 1) * mask, model.py:121. bwd writes dx and ACCUMULATES dw [C], db [1]. */
int xva_rowdot_fwd(const float* x, const float* w, const float* bias, const int32_t* lens, int Z, int R, int C,
                   float* out, void* stream);
int xva_rowdot_bwd(const float* dout, const float* x, const float* w, const int32_t* lens, int Z, int R, int C,
                   float* dx, float* dw, float* db, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * FastPitchLoss masked MSE terms, fastpitch/loss_function.py:81-136. acc is double[2] = {sum sq err, mask count},
 * ACCUMULATED (zero it first); loss = acc[0]/acc[1]. The grad kernels write scale * d(loss)/d(pred).
 *   mel : pred [B,T_out,C] (rows >= T_out count as zero), tgt [B,C,Tm], mask = tgt != 0; dpred row stride ldd >= C
 *   lens: pred,tgt [B,T], mask = t < lens[b]; log1p_tgt = 1 compares against log(tgt + 1) (duration loss)
 * ---------------------------------------------------------------------------------------------------------- */
int xva_mel_mse(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, double* acc, void* stream);
int xva_mel_mse_grad(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, int ldd, const double* acc,
                     float scale, float* dpred, void* stream);
int xva_lens_mse(const float* pred, const float* tgt, const int32_t* lens, int B, int T, int log1p_tgt, double* acc,
                 void* stream);
int xva_lens_mse_grad(const float* pred, const float* tgt, const int32_t* lens, int B, int T, int log1p_tgt,
                      const double* acc, float scale, float* dpred, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-tensor LAMB -- replaces Lamb.step, fastpitch1_1/lamb.py:40-106, plus clip_grad_norm_ (xva_train.py:857).
 * p/g/m/v are flat fp32 arenas; `chunks` is a device array of {int64 start; int32 len; int32 tensor} records (one
 * CUDA block each; every chunk lies inside one tensor). norms is double[2*n_tensors], zeroed by the caller.
 * gnorm_sq (optional) = sum of squared gradients from xva_grad_sqnorm: gradients are scaled by
 * min(1, max_norm/(sqrt(gnorm_sq)+1e-6)) on the fly. lr is read from device memory (CUDA-graph friendly).
 * A non-finite gnorm_sq (NaN / Inf anywhere in the gradients) makes the whole call a no-op on the device: the
 * skip-the-step rule of xva_train.py:825-832 without a host round trip (p, m, v and p_tf32 are left untouched).
 * p_tf32 (optional, same layout as p) receives the updated parameters rounded to tf32: the copy the GEMMs read.
 * ---------------------------------------------------------------------------------------------------------- */
int xva_grad_sqnorm(const float* g, const void* chunks, int n_chunks, double* out, void* stream);
int xva_lamb_step(float* p, const float* g, float* m, float* v, const void* chunks, int n_chunks, double* norms,
                  const double* gnorm_sq, float max_norm, const float* lr_dev, float beta1, float beta2, float eps,
                  float weight_decay, float* p_tf32, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * HiFi-GAN element-wise steps (hifigan/models.py:110-128) and optimizer (hifigan/xva_train.py:298-300).
 *   mean3_lrelu : out = leaky_relu((y0 + y1 + y2) / 3, slope)   -- `x = xs / num_kernels` followed by the leaky ReLU of
 *                 the next stage (slope 0.1, models.py:115) or of conv_post (slope 0.01, models.py:124); tf32-rounded
 *   sum3        : out = a + b + c, tf32-rounded (gradient of a tensor read by the three ResBlocks of a stage)
 *   tanh_bwd    : out[r, 0] = dy[r] * (1 - y[r]^2), out[r, 1..ld) = 0   (models.py:126; ld pads the single channel so
 *                 the buffer is a legal MN-major wgrad operand)
 *   adamw_step  : torch.optim.AdamW over a flat arena; step is 1-based (taken from *step_dev when that is given, so a
 *                 captured CUDA graph advances it with xva_counter_add), lr read from device memory
 * ---------------------------------------------------------------------------------------------------------- */
int xva_mean3_lrelu(const float* y0, const float* y1, const float* y2, int64_t n, float slope, float* out, void* stream);
int xva_sum3(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream);
int xva_tanh_bwd(const float* dy, const float* y, int64_t rows, int ld, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * xVAPitch posterior encoder, element-wise parts of its WaveNet stack.
 *   xva_gated_act_fwd: acts[r, c] = tanh(x_in[r, c]) * sigmoid(x_in[r, H + c]), tf32-rounded -- replaces
 *     fused_add_tanh_sigmoid_multiply (python/xvapitch/wavenet.py:6-13; the add of the conditioning slice is done by
 *     the epilogue of the GEMM that produced x_in). x_in rows have pitch ld_in >= 2H, acts [rows, H]; H % 4 == 0.
 *   xva_gated_act_bwd: dx_in [rows, 2H] = [d * s * (1 - t^2) | d * t * s * (1 - s)], t and s recomputed from x_in.
 *   xva_vits_sample_fwd: z = (mean + eps * exp(log_scale)) on rows t < lens[b], 0 elsewhere, with stats [B, T, 2C] =
 *     [mean | log_scale] (already masked by the projection's epilogue) and eps [B, T, C] the caller's N(0, 1) draw --
 *     replaces python/xvapitch/model.py:1473-1474.  xva_vits_sample_bwd: dstats = [dz | dz * eps * exp(log_scale)].
 *   xva_vits_logp_operands: the two operands of the alignment log-likelihood of xVAPitch.train_step
 *     (python/xvapitch/model.py:766-771: four terms, two einsums) as ONE batched product of K = 2C + 32 columns:
 *     tok [B, Tt, K] = [exp(-2 logs_p) | m_p exp(-2 logs_p) | r | 0..], r = sum_c(-log(2 pi)/2 - logs_p - m_p^2 exp(-2 logs_p)/2);
 *     frm [B, Ts, K] = [-z_p^2 / 2 | z_p | 1 | 0..]; logp[b, j, i] = <frm[b, j], tok[b, i]> (xva_gemm_ref: exact fp32, the
 *     path search that follows -- xva_mas_width1 -- compares these values). Inputs channels-last [B, T, C].
 *   xva_vits_kl: VitsGeneratorLoss.kl_loss (python/xvapitch/losses.py:86-103) and its four gradients in one pass:
 *     *acc += sum over valid frames and channels of logs_p - logs_q - 1/2 + (z_p - m_p)^2 exp(-2 logs_p) / 2 (divide by
 *     sum(lens) for the loss); dz_p, dlogs_q, dm_p, dlogs_p = scale / sum(lens) times the partial derivatives, zero on
 *     frames >= lens[b]. All tensors [B, T, C].
 *   xva_colsum_items: out[z * out_ld + c] += sum_t x[z * z_stride + t * ld + c] for z < Z, t < rows, c < C -- the
 *     gradient of a per-utterance vector the forward broadcast over the frames (wavenet.py:99 g_l, hifigan.py:250).
 * ---------------------------------------------------------------------------------------------------------- */
int xva_gated_act_fwd(const float* x_in, int64_t rows, int H, int64_t ld_in, float* acts, void* stream);
int xva_gated_act_bwd(const float* dacts, const float* x_in, int64_t rows, int H, int64_t ld_in, float* dx_in, void* stream);
int xva_colsum_items(const float* x, int Z, int rows, int C, int64_t ld, int64_t z_stride, float* out, int64_t out_ld,
                     void* stream);
int xva_vits_logp_operands(const float* m_p, const float* logs_p, const float* z_p, int B, int Tt, int Ts, int C, float* tok,
                            float* frm, void* stream);
int xva_vits_kl(const float* z_p, const float* logs_q, const float* m_p, const float* logs_p, const int32_t* lens, int B, int T,
                int C, float scale, double* acc, float* dz_p, float* dlogs_q, float* dm_p, float* dlogs_p, void* stream);
int xva_vits_sample_fwd(const float* stats, const float* eps, const int32_t* lens, int B, int T, int C, float* z, void* stream);
int xva_vits_sample_bwd(const float* dz, const float* eps, const float* stats, const int32_t* lens, int B, int T, int C,
                        float* dstats, void* stream);
int xva_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                   float eps, float weight_decay, int step, const uint64_t* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * xVAPitch text encoder: TextEncoder, python/xvapitch/model.py:1089-1170, over
 * RelativePositionTransformer, python/xvapitch/glow_tts.py:373-485. Its convolutions, score / value products, softmax
 * and LayerNorm are the xva_gemm / xva_softmax / xva_layernorm entries above; these are the remaining steps.
 *   xva_text_embed_fwd : out[b, t, :] = [emb[tokens[b, t]] * scale | lang[b]] for t < lens[b], zero rows otherwise and zero
 *                        in the pad columns [C + L, ld) -- model.py:1152-1165 (scale = sqrt(C), lang = the language
 *                        embedding of the utterance, [B, L]); stored tf32-rounded (it is a GEMM operand). x_emb [B, T, C]
 *                        (optional) = emb[tokens] * scale at EVERY position, the second value the reference returns.
 *   xva_text_embed_bwd : demb[tokens[b, t], c] += scale * dout[b, t, c] for t < lens[b], c < C; dout has row pitch ld.
 *                        (d lang = per-utterance column sums of dout[:, :, C : C + L): xva_colsum_items.)
 *   xva_rel_band_add   : s[z, t, t + r - W] += rel[z, t, r] for r in [0, 2 W] with 0 <= t + r - W < T -- the relative-position
 *                        logits q . E_k^T of glow_tts.py:178-186 added onto the scores (the reference pads and reshapes,
 *                        :260-277); s [Z, T, ld], rel [Z, T, ldr]. The backward uses it on dP with rel = dO . E_v^T.
 *   xva_rel_band_gather: out[z, t, r] = p[z, t, t + r - W] inside the band and the sequence, 0 elsewhere and for
 *                        r in [2 W + 1, ldo) -- the attention weights in relative indexing, glow_tts.py:192-195 (:279-292);
 *                        tf32-rounded: out is the operand of the product with E_v (and, on dS, with E_k).
 *   xva_pad_cols       : dst [rows, ld] = src [rows, C] followed by zero columns (row pitch for MN-major GEMM operands).
 * ---------------------------------------------------------------------------------------------------------- */
int xva_text_embed_fwd(const int64_t* tokens, const float* emb, const float* lang, const int32_t* lens, int B, int T, int C,
                       int L, int ld, float scale, float* out, float* x_emb, void* stream);
int xva_text_embed_bwd(const int64_t* tokens, const float* dout, const int32_t* lens, int B, int T, int C, int ld, float scale,
                       float* demb, void* stream);
int xva_rel_band_add(float* s, const float* rel, int Z, int T, int W, int ld, int ldr, void* stream);
int xva_rel_band_gather(const float* p, int Z, int T, int W, int ld, int ldo, float* out, void* stream);
int xva_pad_cols(const float* src, int64_t rows, int C, int ld, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Mel-spectrogram extractor -- mel_spectrogram(), hifigan/meldataset.py:217-240 (forward and the backward the 45 * L1
 * mel loss needs, hifigan/xva_train.py:480,504). The STFT and the mel projection are xva_gemm launches (a 4-tap GEMM
 * over the [len/hop, hop] view of the padded signal, and a plain GEMM); these are the steps around them.
 *   reflect_pad : out[b, i] = y[b, reflect(i - pad)], length n + 2 pad (F.pad mode='reflect', meldataset.py:229), tf32
 *   spec_mag    : mag[r, c] = sqrt(re^2 + im^2 + eps) for c < nb, 0 for nb <= c < ld_m; spec rows hold re in columns
 *                 [0, nb) and im in [nb, 2 nb) (meldataset.py:235). eps < 0 selects sqrt(max(re^2 + im^2, -eps)), the
 *                 form of xVAPitch's TorchSTFT (python/xvapitch/audio.py:172; no gradient below the floor)
 *   log_clamp   : out = log(max(x, lo))  (spectral_normalize_torch, meldataset.py:238); bwd: dy / x where x >= lo
 *   reduce_loss : kind 0: acc += sum |a - b| (F.l1_loss, feature_loss models.py:263-269); kind 1: acc += sum (c - a)^2
 *                 (discriminator_loss / generator_loss, models.py:272-294). acc is a device double.
 *   loss_grad   : kind 0: out (+)= scale * sign(b - a) (gradient wrt b), times gate_slope where b <= 0 (b a leaky-ReLU
 *                 output: gradient wrt its pre-activation; pass 1 for none); kind 1: out (+)= 2 scale (a - c) (wrt a)
 * ---------------------------------------------------------------------------------------------------------- */
int xva_reflect_pad_fwd(const float* y, int B, int64_t n, int pad, float* out, void* stream);
int xva_reflect_pad_bwd(const float* dyp, int B, int64_t n, int pad, float* dy, void* stream);
int xva_spec_mag_fwd(const float* spec, int64_t rows, int nb, int ld_s, int ld_m, float eps, float* mag, void* stream);
int xva_spec_mag_bwd(const float* dmag, const float* spec, int64_t rows, int nb, int ld_s, int ld_m, float eps,
                     float* dspec, void* stream);
int xva_log_clamp_fwd(const float* x, int64_t n, float lo, float* out, void* stream);
int xva_log_clamp_bwd(const float* dy, const float* x, int64_t n, float lo, float* dx, void* stream);
int xva_reduce_loss(const float* a, const float* b, int64_t n, int kind, float c, double* acc, void* stream);
int xva_loss_grad(const float* a, const float* b, int64_t n, int kind, float c, float scale, float gate_slope,
                  int accumulate, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * HiFi-GAN discriminators, the parts that are not GEMMs (hifigan/models.py:140-260).
 *   conv_c1_* : first convolution of a discriminator (one input channel) applied to Z = B*P sequences cut out of the
 *               raw waveform: element q of sequence (b, c) is sample q*xs_q + c*xs_c of item b (reflected about the last
 *               sample when >= Lsrc: the period padding of DiscriminatorP.forward, models.py:159-163). MPD period p:
 *               P = p, xs_q = p, xs_c = 1, L = ceil(Lsrc / p); MSD: P = 1, xs_q = 1, xs_c = 0, L = Lsrc.
 *               out [Z, Lout_p, Cout] = leaky_relu(conv + bias) for rows < Lout, zero rows up to Lout_p (tf32-rounded).
 *               bwd_w accumulates dw [Cout, k], db [Cout] from dpre (gradient wrt the pre-activation); bwd_x ACCUMULATES
 *               scale * dL/d(waveform) with atomics (every discriminator adds into the same buffer).
 *   avgpool4  : AvgPool1d(4, 2, padding=2) (models.py:241) on [B, L] -> [B, L/2 + 1]; bwd overwrites dx.
 *   zero_tail_rows : x[z, Lvalid.., :] = 0 for x [Z, Lp, C].
 * ---------------------------------------------------------------------------------------------------------- */
/* feature_loss term and its gradient in one pass (models.py:263-269): *acc += sum |a - b|, out = scale * sign(b - a),
 * times gate_slope where b <= 0 (b a leaky-ReLU output: gradient wrt its pre-activation; pass 1 for none). */
int xva_l1_loss_grad(const float* a, const float* b, int64_t n, float scale, float gate_slope, double* acc, float* out,
                     void* stream);
int xva_conv_c1_fwd(const float* x, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, const float* w,
                    const float* bias, int k, int s, int pad, int Z, int Lout, int Lout_p, int Cout, float slope, float* out,
                    void* stream);
int xva_conv_c1_bwd_w(const float* dpre, const float* x, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k,
                      int s, int pad, int Z, int Lout, int Lout_p, int Cout, float* dw, float* db, void* stream);
int xva_conv_c1_bwd_x(const float* dpre, const float* w, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k,
                      int s, int pad, int Z, int Lout, int Lout_p, int Cout, float scale, float* dx, void* stream);
int xva_avgpool4_fwd(const float* x, int B, int L, float* out, void* stream);
int xva_avgpool4_bwd(const float* dout, int B, int L, float* dx, void* stream);
int xva_zero_tail_rows(float* x, int Z, int Lp, int Lvalid, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Weight-norm reparametrisation + packing of every convolution of a model in one launch -- replaces the
 * torch.nn.utils.weight_norm forward hook of each conv (hifigan/models.py:21-38, 87-105, 144-150, 207-215:
 * w = g * v / ||v||, norm over all dims but 0) plus the re-layout of w into the tap-GEMM's operand format, and their
 * autograd. One descriptor per convolution, the table lives in device memory.
 *   element (r, c, j) of v [rows, inner / k, k] goes, scaled by g[r] / ||v[r]|| and tf32-rounded, to
 *     dst[tap_off[j] + r * ld + ((r / og) % f) * cg + c]      (Conv1d / Conv2d; f > 1: block-diagonal super-groups)
 *     dst[tap_off[j] + c * ld + r]                            (XVA_WN_TRANSPOSED: ConvTranspose1d, v is [Cin, Cout, k])
 *   bwd: dv (+)= g/||v|| * (dW - v (v.dW)/||v||^2), dg (+)= (v.dW)/||v||, with dW gathered from ddst (same layout as dst).
 * max_inner = the largest `inner` of the table (shared-memory staging of one row; <= 12288).
 * ---------------------------------------------------------------------------------------------------------- */
#define XVA_WN_TRANSPOSED 1
#define XVA_WN_NO_ROUND 2
/* a plain (not weight-normed) convolution in the same table: w = v (g, dg unused, may be null), bwd: dv (+)= dW. The
 * xVAPitch waveform decoder removes weight norm from conv_pre / conv_post and adds a plain cond_layer
 * (python/xvapitch/hifigan.py:224-232, xvapitch/model.py:134-149). */
#define XVA_WN_PLAIN 4
typedef struct xva_wn_desc {
  const float* v;
  const float* g;
  float* dv;          /* accumulated (bwd) */
  float* dg;
  float* dst;         /* packed-weight arena base (fwd) */
  const float* ddst;  /* gradient arena base (bwd) */
  int32_t rows, inner, k, flags;
  int32_t ld, og, f, cg;
  int32_t row_start;  /* sum of `rows` of the preceding descriptors */
  int32_t _pad;
  int64_t tap_off[XVA_MAX_TAPS];
} xva_wn_desc;
int xva_sizeof_wn_desc(void);

/* Spectral norm + packing of every convolution of a spectral-normed model -- replaces torch.nn.utils.spectral_norm's
 * forward pre-hook (one power iteration per training forward, eps 1e-12; hifigan/models.py:207-215 with
 * use_spectral_norm=True) plus the re-layout, and their autograd. With W = weight_orig [rows, inner]:
 *   fwd, training: v <- normalize(W^T u); u <- normalize(W v); sigma = u . (W v); dst = W / sigma (packed, tf32-rounded).
 *                  u / v (the module's buffers) are replaced by the new iterates, which are also stored in u_sav / v_sav.
 *   fwd, eval:     sigma = u . (W v) with the stored vectors (copied to u_sav / v_sav).
 *   bwd:           dw (+)= dW_eff / sigma - (<dW_eff, W> / sigma^2) u_sav v_sav^T, dW_eff gathered from ddst.
 * Layout fields (k, flags, ld, og, f, cg, tap_off) as in xva_wn_desc. `work` is per-descriptor scratch of
 * ceil(rows / 64) * inner + rows + 2 floats that must survive from a forward to its backward (it holds sigma).
 * blk_start = sum over the preceding descriptors of ceil(inner / 256) * ceil(rows / 64); total_blocks = that sum over all. */
typedef struct xva_sn_desc {
  const float* w;
  float* u;
  float* v;
  float* u_sav;
  float* v_sav;
  float* dw;          /* accumulated (bwd); may be null when no backward runs */
  float* dst;
  const float* ddst;
  float* work;
  int32_t rows, inner, k, flags;
  int32_t ld, og, f, cg;
  int32_t row_start, blk_start;
  int64_t tap_off[XVA_MAX_TAPS];
} xva_sn_desc;
int xva_sizeof_sn_desc(void);
int xva_sn_pack_fwd(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, int training,
                    void* stream);
int xva_sn_pack_bwd(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, void* stream);
int xva_wn_pack_fwd(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, void* stream);
int xva_wn_pack_bwd(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XVA_B200_H_ */
