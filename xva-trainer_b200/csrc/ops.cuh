// Launchers of the non-GEMM kernels (HBM-bound row / index / reduction work). Definitions in the .cu files named
// next to each group; the C ABI in capi.cu forwards to these.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/xva_b200.h"

namespace xva {

// ---- test switch for the tf32 operand rounding, one setter per translation unit (common.cuh)
int set_operand_rounding_gemm_tc(int on);
int set_operand_rounding_attn_fused(int on);
int set_operand_rounding_gemm_ref(int on);
int set_operand_rounding_rowops(int on);
int set_operand_rounding_loss_optim(int on);
int set_operand_rounding_elemwise(int on);
int set_operand_rounding_vits(int on);
int set_operand_rounding_melspec(int on);
int set_operand_rounding_disc(int on);
int set_operand_rounding_wnpack(int on);
int set_operand_rounding_align(int on);
int set_operand_rounding_relattn(int on);

// ---- regulate.cu
int duration_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int* cum, int* dec_lens,
                  cudaStream_t stream);
int regulate_gather(const float* enc, const int* cum, int B, int Tt, int C, int T_out, float* out, int* idx_out,
                    cudaStream_t stream);
int regulate_scatter(const float* dout, const int* cum, int B, int Tt, int C, int T_out, float* denc, int accumulate,
                     cudaStream_t stream);
int average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, int log1p_out,
                  cudaStream_t stream);

// ---- rowops.cu
int softmax_fwd(const float* s, const int* lens, int Z, int R, int N, int ld, float* p_out, float* pd_out, float drop_p,
                uint64_t seed, const uint64_t* seed_dev, cudaStream_t stream);
int softmax_bwd(const float* p, float* dpd, int Z, int R, int N, int ld, float alpha, float drop_p, uint64_t seed,
                const uint64_t* seed_dev, cudaStream_t stream);
int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                  const int* lens, int Z, int R, int C, float* dx, float* dx_drop, float* dgamma, float* dbeta,
                  float* dbias, float drop_post_p, uint64_t seed_post, float drop_pre_p, uint64_t seed_pre,
                  const uint64_t* seed_dev, int relu_gate, cudaStream_t stream);
int mas_width1(const float* attn, const int* in_lens, const int* out_lens, int B, int Tm, int Tt, int is_log, float* hard,
               int* durs, cudaStream_t stream);
int mas_log(const float* attn, long n, float* out, cudaStream_t stream);
int layernorm_fwd(const float* x, const float* gamma, const float* beta, const int* lens, int Z, int R, int C, float eps,
                  float* y, float* mean, float* rstd, cudaStream_t stream);
int counter_add(unsigned long long* counter, unsigned long long inc, cudaStream_t stream);
int round_tf32(const float* src, float* dst, long n, cudaStream_t stream);
int rowdot2(const float* a, const float* b, long rows, int C, long a_ld, long b_ld, float* out, cudaStream_t stream);
int colsum(const float* x, long rows, int C, long ld, float* out, cudaStream_t stream);
int embed_pos(const long long* tokens, const float* emb, const float* in, const int* lens, const float* inv_freq,
              int B, int T, int C, float* out, cudaStream_t stream);
int embed_bwd(const long long* tokens, const float* dout, int B, int T, int C, float* demb, cudaStream_t stream);
int scalar_conv_add(float* io, const float* x, const float* w, const float* bias, const int* lens, int B, int T,
                    int C, cudaStream_t stream);
int scalar_conv_bwd(const float* dout, const float* x, int B, int T, int C, float* dw, float* dbias,
                    cudaStream_t stream);
int rowdot_fwd(const float* x, const float* w, const float* bias, const int* lens, int Z, int R, int C, float* out,
               cudaStream_t stream);
int rowdot_bwd(const float* dout, const float* x, const float* w, const int* lens, int Z, int R, int C, float* dx,
               float* dw, float* db, cudaStream_t stream);

// ---- align.cu (FastPitch stage-1 aligner: score, CTC, binarization loss)
int attn_score_fwd(const float* q, long ldq, const float* k, long ldk, const float* prior, const int* in_lens, int B,
                   int Tm, int Tt, int C, float* logprob, float* soft, cudaStream_t stream);
int attn_score_bwd(const float* g, const float* logprob, const float* prior, const float* q, long ldq, const float* k,
                   long ldk, int B, int Tm, int Tt, int C, float* dD, float* dq, long lddq, float* dk, long lddk,
                   cudaStream_t stream);
long long attn_ctc_workspace_bytes(int B, int Tm, int Tt);
int attn_ctc(const float* logprob, const int* in_lens, const int* out_lens, int B, int Tm, int Tt, float blank_logprob,
             void* workspace, long long workspace_bytes, double* cost, float* grad, cudaStream_t stream);
int attn_bin_loss(const float* hard, const float* soft, long rows, int Tt, float eps, double* acc, cudaStream_t stream);
int attn_grad_combine(const float* gctc, const float* hard, const float* soft, const double* acc, float a, float bw,
                      float eps, long rows, int Tt, float* g, cudaStream_t stream);

// ---- vits.cu
int gated_act_fwd(const float* x_in, long rows, int H, long ld_in, float* acts, cudaStream_t stream);
int gated_act_bwd(const float* dacts, const float* x_in, long rows, int H, long ld_in, float* dx_in, cudaStream_t stream);
int colsum_items(const float* x, int Z, int rows, int C, long ld, long zs, float* out, long out_ld, cudaStream_t stream);
int vits_logp_operands(const float* m, const float* logs, const float* z, int B, int Tt, int Ts, int C, float* tok, float* frm,
                       cudaStream_t stream);
int vits_kl(const float* z, const float* lq, const float* m, const float* lp, const int* lens, int B, int T, int C, float scale,
            double* acc, float* dz, float* dlq, float* dm, float* dlp, cudaStream_t stream);
int vits_sample_fwd(const float* stats, const float* eps, const int* lens, int B, int T, int C, float* z, cudaStream_t stream);
int vits_sample_bwd(const float* dz, const float* eps, const float* stats, const int* lens, int B, int T, int C, float* dstats,
                    cudaStream_t stream);

// ---- relattn.cu (xVAPitch text encoder: embedding, relative-position band, padded copies)
int text_embed_fwd(const long long* tokens, const float* emb, const float* lang, const int* lens, int B, int T, int C, int L,
                   int ld, float scale, float* out, float* x_emb, cudaStream_t stream);
int text_embed_bwd(const long long* tokens, const float* dout, const int* lens, int B, int T, int C, int ld, float scale,
                   float* demb, cudaStream_t stream);
int rel_band_add(float* s, const float* rel, int Z, int T, int W, int ld, int ldr, cudaStream_t stream);
int rel_band_gather(const float* p, int Z, int T, int W, int ld, int ldo, float* out, cudaStream_t stream);
int pad_cols(const float* src, long rows, int C, int ld, float* dst, cudaStream_t stream);

// ---- elemwise.cu
int mean3_lrelu(const float* y0, const float* y1, const float* y2, long n, float slope, float* out, cudaStream_t stream);
int sum3(const float* a, const float* b, const float* c, long n, float* out, cudaStream_t stream);
int tanh_bwd(const float* dy, const float* y, long rows, int ld, float* out, cudaStream_t stream);
int wn_pack(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, int backward, cudaStream_t stream);
int sn_pack(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, int training,
            int backward, cudaStream_t stream);
int l1_loss_grad(const float* a, const float* b, long n, float scale, float gate_slope, double* acc, float* out,
                 cudaStream_t stream);
int adamw_step(float* p, const float* g, float* m, float* v, long n, const float* lr_dev, float b1, float b2, float eps,
               float wd, int step, const unsigned long long* step_dev, cudaStream_t stream);

// ---- melspec.cu
int reflect_pad_fwd(const float* y, int B, long n, int pad, float* out, cudaStream_t stream);
int reflect_pad_bwd(const float* dyp, int B, long n, int pad, float* dy, cudaStream_t stream);
int spec_mag_fwd(const float* spec, long rows, int nb, int ld_s, int ld_m, float eps, float* mag, cudaStream_t stream);
int spec_mag_bwd(const float* dmag, const float* spec, long rows, int nb, int ld_s, int ld_m, float eps, float* dspec,
                 cudaStream_t stream);
int log_clamp_fwd(const float* x, long n, float lo, float* out, cudaStream_t stream);
int log_clamp_bwd(const float* dy, const float* x, long n, float lo, float* dx, cudaStream_t stream);
int reduce_loss(const float* a, const float* b, long n, int kind, float c, double* acc, cudaStream_t stream);
int loss_grad(const float* a, const float* b, long n, int kind, float c, float scale, float gate_slope, int accumulate,
              float* out, cudaStream_t stream);

// ---- disc.cu
int conv_c1_fwd(const float* x, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, const float* w, const float* bias,
                int k, int s, int pad, int Z, int Lout, int Lout_p, int Cout, float slope, float* out, cudaStream_t stream);
int conv_c1_bwd_w(const float* dpre, const float* x, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k, int s,
                  int pad, int Z, int Lout, int Lout_p, int Cout, float* dw, float* db, cudaStream_t stream);
int conv_c1_bwd_x(const float* dpre, const float* w, long xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k, int s,
                  int pad, int Z, int Lout, int Lout_p, int Cout, float scale, float* dx, cudaStream_t stream);
int avgpool4_fwd(const float* x, int B, int L, float* out, cudaStream_t stream);
int avgpool4_bwd(const float* dout, int B, int L, float* dx, cudaStream_t stream);
int zero_tail_rows(float* x, int Z, int Lp, int Lvalid, int C, cudaStream_t stream);

// ---- loss_optim.cu
int mel_mse(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, double* acc, cudaStream_t stream);
int mel_mse_grad(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, int ldd, const double* acc,
                 float scale, float* dpred, cudaStream_t stream);
int lens_mse(const float* pred, const float* tgt, const int* lens, int B, int T, int log1p_tgt, double* acc,
             cudaStream_t stream);
int lens_mse_grad(const float* pred, const float* tgt, const int* lens, int B, int T, int log1p_tgt, const double* acc,
                  float scale, float* dpred, cudaStream_t stream);
int grad_sqnorm(const float* g, const void* chunks, int n_chunks, double* out, cudaStream_t stream);
int lamb_step(float* p, const float* g, float* m, float* v, const void* chunks, int n_chunks, double* norms,
              const double* gnorm_sq, float max_norm, const float* lr_dev, float b1, float b2, float eps, float wd,
              float* p_tf32, cudaStream_t stream);

// attn_fused.cu: fused single-head attention of the FFT blocks (scores, mask, softmax, dropout, P.V in tensor memory)
int attn_fused_fwd(const float* qkv, long rs, long zs, int B, int T, const int* lens, float scale, float drop_p,
                   uint64_t seed, const uint64_t* seed_dev, int drop_ld, float* out, long o_rs, long o_zs, float* lse,
                   cudaStream_t stream);
int attn_fused_bwd(const float* qkv, long rs, long zs, const float* dout, long d_rs, long d_zs, const float* lse,
                   const float* dsum, int B, int T, const int* lens, float scale, float drop_p, uint64_t seed,
                   const uint64_t* seed_dev, int drop_ld, float* dqkv, long g_rs, long g_zs, cudaStream_t stream);

}  // namespace xva
