// extern "C" surface of libxva_b200.so (declared in include/xva_b200.h). Thin: argument checks live in the launchers.
#include "../../include/xva_b200.h"

#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"

using namespace xva;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int xva_abi_version(void) { return XVA_ABI_VERSION; }

const char* xva_last_error(void) { return last_error(); }

int xva_device_check(int device) {
  cudaDeviceProp prop;
  XVA_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d (%s) is compute capability %d.%d; libxva_b200 needs 10.x (sm_100a)", device, prop.name,
              prop.major, prop.minor);
    return XVA_ERR_DEVICE;
  }
  return XVA_OK;
}

int xva_sizeof_gemm_args(void) { return static_cast<int>(sizeof(xva_gemm_args)); }

int xva_gemm(const xva_gemm_args* args, void* stream) {
  XVA_CHECK_ARG(args != nullptr, "xva_gemm: null args");
  return gemm_tc_launch(*args, S(stream));
}

int xva_gemm_debug_counters(long long* out8) { return gemm_debug_counters(out8); }

int xva_gemm_ref(const xva_gemm_args* args, void* stream) {
  XVA_CHECK_ARG(args != nullptr, "xva_gemm_ref: null args");
  return gemm_ref_launch(*args, S(stream));
}

int xva_regulate_len_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int32_t* cum,
                          int32_t* dec_lens, void* stream) {
  return duration_scan(durs, B, Tt, pace, mel_max_len, cum, dec_lens, S(stream));
}

int xva_regulate_len_fwd(const float* enc, const int32_t* cum, int B, int Tt, int C, int T_out, float* out,
                         int32_t* idx, void* stream) {
  return regulate_gather(enc, cum, B, Tt, C, T_out, out, idx, S(stream));
}

int xva_regulate_len_bwd(const float* dout, const int32_t* cum, int B, int Tt, int C, int T_out, float* denc,
                         int accumulate, void* stream) {
  return regulate_scatter(dout, cum, B, Tt, C, T_out, denc, accumulate, S(stream));
}

int xva_average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, int log1p_out,
                      void* stream) {
  return average_pitch(pitch, durs, B, F, Tm, Tt, out, log1p_out, S(stream));
}

int xva_rowdot2(const float* a, const float* b, int64_t rows, int C, int64_t a_ld, int64_t b_ld, float* out, void* stream) {
  return rowdot2(a, b, static_cast<long>(rows), C, static_cast<long>(a_ld), static_cast<long>(b_ld), out, S(stream));
}

int xva_attn_fwd(const float* qkv, int64_t rs, int64_t zs, int B, int T, const int32_t* lens, float scale, float drop_p,
                 uint64_t seed, const uint64_t* seed_dev, int drop_ld, float* out, int64_t o_rs, int64_t o_zs, float* lse,
                 void* stream) {
  return attn_fused_fwd(qkv, static_cast<long>(rs), static_cast<long>(zs), B, T, lens, scale, drop_p, seed, seed_dev, drop_ld,
                        out, static_cast<long>(o_rs), static_cast<long>(o_zs), lse, S(stream));
}

int xva_attn_bwd(const float* qkv, int64_t rs, int64_t zs, const float* dout, int64_t d_rs, int64_t d_zs, const float* lse,
                 const float* dsum, int B, int T, const int32_t* lens, float scale, float drop_p, uint64_t seed,
                 const uint64_t* seed_dev, int drop_ld, float* dqkv, int64_t g_rs, int64_t g_zs, void* stream) {
  return attn_fused_bwd(qkv, static_cast<long>(rs), static_cast<long>(zs), dout, static_cast<long>(d_rs),
                        static_cast<long>(d_zs), lse, dsum, B, T, lens, scale, drop_p, seed, seed_dev, drop_ld, dqkv,
                        static_cast<long>(g_rs), static_cast<long>(g_zs), S(stream));
}

int xva_mas_width1(const float* attn, const int32_t* in_lens, const int32_t* out_lens, int B, int Tm, int Tt, int is_log,
                   float* hard, int32_t* durs, void* stream) {
  return mas_width1(attn, in_lens, out_lens, B, Tm, Tt, is_log, hard, durs, S(stream));
}

int xva_attn_score_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* prior, const int32_t* in_lens,
                       int B, int Tm, int Tt, int C, float* logprob, float* soft, void* stream) {
  return attn_score_fwd(q, static_cast<long>(ldq), k, static_cast<long>(ldk), prior, in_lens, B, Tm, Tt, C, logprob, soft,
                        S(stream));
}

int xva_attn_score_bwd(const float* g, const float* logprob, const float* prior, const float* q, int64_t ldq,
                       const float* k, int64_t ldk, int B, int Tm, int Tt, int C, float* dD, float* dq, int64_t lddq,
                       float* dk, int64_t lddk, void* stream) {
  return attn_score_bwd(g, logprob, prior, q, static_cast<long>(ldq), k, static_cast<long>(ldk), B, Tm, Tt, C, dD, dq,
                        static_cast<long>(lddq), dk, static_cast<long>(lddk), S(stream));
}

int64_t xva_attn_ctc_workspace_bytes(int B, int Tm, int Tt) { return attn_ctc_workspace_bytes(B, Tm, Tt); }

int xva_attn_ctc(const float* logprob, const int32_t* in_lens, const int32_t* out_lens, int B, int Tm, int Tt,
                 float blank_logprob, void* workspace, int64_t workspace_bytes, double* cost, float* grad, void* stream) {
  return attn_ctc(logprob, in_lens, out_lens, B, Tm, Tt, blank_logprob, workspace, workspace_bytes, cost, grad, S(stream));
}

int xva_attn_bin_loss(const float* hard, const float* soft, int64_t rows, int Tt, float eps, double* acc, void* stream) {
  return attn_bin_loss(hard, soft, static_cast<long>(rows), Tt, eps, acc, S(stream));
}

int xva_attn_grad_combine(const float* gctc, const float* hard, const float* soft, const double* acc, float a, float bw,
                          float eps, int64_t rows, int Tt, float* g, void* stream) {
  return attn_grad_combine(gctc, hard, soft, acc, a, bw, eps, static_cast<long>(rows), Tt, g, S(stream));
}

int xva_mas_log(const float* attn, int64_t n, float* out, void* stream) {
  return mas_log(attn, static_cast<long>(n), out, S(stream));
}

int xva_softmax_fwd(const float* s, const int32_t* lens, int Z, int R, int N, int ld, float* p, float* pd,
                    float drop_p, uint64_t seed, const uint64_t* seed_dev, void* stream) {
  return softmax_fwd(s, lens, Z, R, N, ld, p, pd, drop_p, seed, seed_dev, S(stream));
}

int xva_softmax_bwd(const float* p, float* dpd, int Z, int R, int N, int ld, float alpha, float drop_p,
                    uint64_t seed, const uint64_t* seed_dev, void* stream) {
  return softmax_bwd(p, dpd, Z, R, N, ld, alpha, drop_p, seed, seed_dev, S(stream));
}

int xva_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const int32_t* lens, int Z, int R, int C, float* dx, float* dx_drop, float* dgamma,
                      float* dbeta, float* dbias, float drop_post_p, uint64_t seed_post, float drop_pre_p,
                      uint64_t seed_pre, const uint64_t* seed_dev, int relu_gate, void* stream) {
  return layernorm_bwd(dy, x, mean, rstd, gamma, lens, Z, R, C, dx, dx_drop, dgamma, dbeta, dbias, drop_post_p,
                       seed_post, drop_pre_p, seed_pre, seed_dev, relu_gate, S(stream));
}

int xva_layernorm_fwd(const float* x, const float* gamma, const float* beta, const int32_t* lens, int Z, int R, int C,
                      float eps, float* y, float* mean, float* rstd, void* stream) {
  return layernorm_fwd(x, gamma, beta, lens, Z, R, C, eps, y, mean, rstd, S(stream));
}

int xva_set_operand_rounding(int on) {
  int rc;
  if ((rc = set_operand_rounding_gemm_tc(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_gemm_ref(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_attn_fused(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_rowops(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_elemwise(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_vits(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_melspec(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_disc(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_wnpack(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_align(on)) != XVA_OK) return rc;
  if ((rc = set_operand_rounding_relattn(on)) != XVA_OK) return rc;
  return set_operand_rounding_loss_optim(on);
}

int xva_round_tf32(const float* src, float* dst, int64_t n, void* stream) {
  return round_tf32(src, dst, static_cast<long>(n), S(stream));
}

int xva_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
  return counter_add(reinterpret_cast<unsigned long long*>(counter), inc, S(stream));
}

int xva_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* stream) {
  return colsum(x, rows, C, ld, out, S(stream));
}

int xva_embed_pos(const int64_t* tokens, const float* emb, const float* in, const int32_t* lens,
                  const float* inv_freq, int B, int T, int C, float* out, void* stream) {
  return embed_pos(reinterpret_cast<const long long*>(tokens), emb, in, lens, inv_freq, B, T, C, out, S(stream));
}

int xva_embed_bwd(const int64_t* tokens, const float* dout, int B, int T, int C, float* demb, void* stream) {
  return embed_bwd(reinterpret_cast<const long long*>(tokens), dout, B, T, C, demb, S(stream));
}

int xva_scalar_conv_add(float* io, const float* x, const float* w, const float* bias, const int32_t* lens, int B,
                        int T, int C, void* stream) {
  return scalar_conv_add(io, x, w, bias, lens, B, T, C, S(stream));
}

int xva_scalar_conv_bwd(const float* dout, const float* x, int B, int T, int C, float* dw, float* dbias,
                        void* stream) {
  return scalar_conv_bwd(dout, x, B, T, C, dw, dbias, S(stream));
}

int xva_rowdot_fwd(const float* x, const float* w, const float* bias, const int32_t* lens, int Z, int R, int C,
                   float* out, void* stream) {
  return rowdot_fwd(x, w, bias, lens, Z, R, C, out, S(stream));
}

int xva_rowdot_bwd(const float* dout, const float* x, const float* w, const int32_t* lens, int Z, int R, int C,
                   float* dx, float* dw, float* db, void* stream) {
  return rowdot_bwd(dout, x, w, lens, Z, R, C, dx, dw, db, S(stream));
}

int xva_mel_mse(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, double* acc, void* stream) {
  return mel_mse(pred, tgt, B, T_out, Tm, C, acc, S(stream));
}

int xva_mel_mse_grad(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, int ldd, const double* acc,
                     float scale, float* dpred, void* stream) {
  return mel_mse_grad(pred, tgt, B, T_out, Tm, C, ldd, acc, scale, dpred, S(stream));
}

int xva_lens_mse(const float* pred, const float* tgt, const int32_t* lens, int B, int T, int log1p_tgt, double* acc,
                 void* stream) {
  return lens_mse(pred, tgt, lens, B, T, log1p_tgt, acc, S(stream));
}

int xva_lens_mse_grad(const float* pred, const float* tgt, const int32_t* lens, int B, int T, int log1p_tgt,
                      const double* acc, float scale, float* dpred, void* stream) {
  return lens_mse_grad(pred, tgt, lens, B, T, log1p_tgt, acc, scale, dpred, S(stream));
}

int xva_grad_sqnorm(const float* g, const void* chunks, int n_chunks, double* out, void* stream) {
  return grad_sqnorm(g, chunks, n_chunks, out, S(stream));
}

int xva_lamb_step(float* p, const float* g, float* m, float* v, const void* chunks, int n_chunks, double* norms,
                  const double* gnorm_sq, float max_norm, const float* lr_dev, float beta1, float beta2, float eps,
                  float weight_decay, float* p_tf32, void* stream) {
  return lamb_step(p, g, m, v, chunks, n_chunks, norms, gnorm_sq, max_norm, lr_dev, beta1, beta2, eps, weight_decay,
                   p_tf32, S(stream));
}

int xva_mean3_lrelu(const float* y0, const float* y1, const float* y2, int64_t n, float slope, float* out, void* stream) {
  return mean3_lrelu(y0, y1, y2, static_cast<long>(n), slope, out, S(stream));
}

int xva_sum3(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream) {
  return sum3(a, b, c, static_cast<long>(n), out, S(stream));
}

int xva_gated_act_fwd(const float* x_in, int64_t rows, int H, int64_t ld_in, float* acts, void* stream) {
  return gated_act_fwd(x_in, static_cast<long>(rows), H, static_cast<long>(ld_in), acts, S(stream));
}

int xva_gated_act_bwd(const float* dacts, const float* x_in, int64_t rows, int H, int64_t ld_in, float* dx_in, void* stream) {
  return gated_act_bwd(dacts, x_in, static_cast<long>(rows), H, static_cast<long>(ld_in), dx_in, S(stream));
}

int xva_colsum_items(const float* x, int Z, int rows, int C, int64_t ld, int64_t z_stride, float* out, int64_t out_ld,
                     void* stream) {
  return colsum_items(x, Z, rows, C, static_cast<long>(ld), static_cast<long>(z_stride), out, static_cast<long>(out_ld), S(stream));
}

int xva_vits_logp_operands(const float* m_p, const float* logs_p, const float* z_p, int B, int Tt, int Ts, int C, float* tok,
                            float* frm, void* stream) {
  return vits_logp_operands(m_p, logs_p, z_p, B, Tt, Ts, C, tok, frm, S(stream));
}

int xva_vits_kl(const float* z_p, const float* logs_q, const float* m_p, const float* logs_p, const int32_t* lens, int B, int T,
                int C, float scale, double* acc, float* dz_p, float* dlogs_q, float* dm_p, float* dlogs_p, void* stream) {
  return vits_kl(z_p, logs_q, m_p, logs_p, lens, B, T, C, scale, acc, dz_p, dlogs_q, dm_p, dlogs_p, S(stream));
}

int xva_vits_sample_fwd(const float* stats, const float* eps, const int32_t* lens, int B, int T, int C, float* z, void* stream) {
  return vits_sample_fwd(stats, eps, lens, B, T, C, z, S(stream));
}

int xva_vits_sample_bwd(const float* dz, const float* eps, const float* stats, const int32_t* lens, int B, int T, int C,
                        float* dstats, void* stream) {
  return vits_sample_bwd(dz, eps, stats, lens, B, T, C, dstats, S(stream));
}

int xva_text_embed_fwd(const int64_t* tokens, const float* emb, const float* lang, const int32_t* lens, int B, int T, int C,
                       int L, int ld, float scale, float* out, float* x_emb, void* stream) {
  return text_embed_fwd(reinterpret_cast<const long long*>(tokens), emb, lang, lens, B, T, C, L, ld, scale, out, x_emb, S(stream));
}

int xva_text_embed_bwd(const int64_t* tokens, const float* dout, const int32_t* lens, int B, int T, int C, int ld, float scale,
                       float* demb, void* stream) {
  return text_embed_bwd(reinterpret_cast<const long long*>(tokens), dout, lens, B, T, C, ld, scale, demb, S(stream));
}

int xva_rel_band_add(float* s, const float* rel, int Z, int T, int W, int ld, int ldr, void* stream) {
  return rel_band_add(s, rel, Z, T, W, ld, ldr, S(stream));
}

int xva_rel_band_gather(const float* p, int Z, int T, int W, int ld, int ldo, float* out, void* stream) {
  return rel_band_gather(p, Z, T, W, ld, ldo, out, S(stream));
}

int xva_pad_cols(const float* src, int64_t rows, int C, int ld, float* dst, void* stream) {
  return pad_cols(src, static_cast<long>(rows), C, ld, dst, S(stream));
}

int xva_tanh_bwd(const float* dy, const float* y, int64_t rows, int ld, float* out, void* stream) {
  return tanh_bwd(dy, y, static_cast<long>(rows), ld, out, S(stream));
}

int xva_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                   float eps, float weight_decay, int step, const uint64_t* step_dev, void* stream) {
  return adamw_step(p, g, m, v, static_cast<long>(n), lr_dev, beta1, beta2, eps, weight_decay, step,
                    reinterpret_cast<const unsigned long long*>(step_dev), S(stream));
}

int xva_sizeof_wn_desc(void) { return static_cast<int>(sizeof(xva_wn_desc)); }
int xva_sizeof_sn_desc(void) { return static_cast<int>(sizeof(xva_sn_desc)); }
int xva_sn_pack_fwd(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, int training,
                    void* stream) {
  return sn_pack(table_dev, n_desc, total_rows, total_blocks, max_inner, training, 0, S(stream));
}
int xva_sn_pack_bwd(const xva_sn_desc* table_dev, int n_desc, int total_rows, int total_blocks, int max_inner, void* stream) {
  return sn_pack(table_dev, n_desc, total_rows, total_blocks, max_inner, 0, 1, S(stream));
}
int xva_l1_loss_grad(const float* a, const float* b, int64_t n, float scale, float gate_slope, double* acc, float* out,
                     void* stream) {
  return l1_loss_grad(a, b, static_cast<long>(n), scale, gate_slope, acc, out, S(stream));
}
int xva_wn_pack_fwd(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, void* stream) {
  return wn_pack(table_dev, n_desc, total_rows, max_inner, 0, S(stream));
}
int xva_wn_pack_bwd(const xva_wn_desc* table_dev, int n_desc, int total_rows, int max_inner, void* stream) {
  return wn_pack(table_dev, n_desc, total_rows, max_inner, 1, S(stream));
}

int xva_reflect_pad_fwd(const float* y, int B, int64_t n, int pad, float* out, void* stream) {
  return reflect_pad_fwd(y, B, static_cast<long>(n), pad, out, S(stream));
}
int xva_reflect_pad_bwd(const float* dyp, int B, int64_t n, int pad, float* dy, void* stream) {
  return reflect_pad_bwd(dyp, B, static_cast<long>(n), pad, dy, S(stream));
}
int xva_spec_mag_fwd(const float* spec, int64_t rows, int nb, int ld_s, int ld_m, float eps, float* mag, void* stream) {
  return spec_mag_fwd(spec, static_cast<long>(rows), nb, ld_s, ld_m, eps, mag, S(stream));
}
int xva_spec_mag_bwd(const float* dmag, const float* spec, int64_t rows, int nb, int ld_s, int ld_m, float eps,
                     float* dspec, void* stream) {
  return spec_mag_bwd(dmag, spec, static_cast<long>(rows), nb, ld_s, ld_m, eps, dspec, S(stream));
}
int xva_log_clamp_fwd(const float* x, int64_t n, float lo, float* out, void* stream) {
  return log_clamp_fwd(x, static_cast<long>(n), lo, out, S(stream));
}
int xva_log_clamp_bwd(const float* dy, const float* x, int64_t n, float lo, float* dx, void* stream) {
  return log_clamp_bwd(dy, x, static_cast<long>(n), lo, dx, S(stream));
}
int xva_reduce_loss(const float* a, const float* b, int64_t n, int kind, float c, double* acc, void* stream) {
  return reduce_loss(a, b, static_cast<long>(n), kind, c, acc, S(stream));
}
int xva_loss_grad(const float* a, const float* b, int64_t n, int kind, float c, float scale, float gate_slope,
                  int accumulate, float* out, void* stream) {
  return loss_grad(a, b, static_cast<long>(n), kind, c, scale, gate_slope, accumulate, out, S(stream));
}

int xva_conv_c1_fwd(const float* x, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, const float* w,
                    const float* bias, int k, int s, int pad, int Z, int Lout, int Lout_p, int Cout, float slope, float* out,
                    void* stream) {
  return conv_c1_fwd(x, static_cast<long>(xs_b), xs_q, xs_c, P, Lsrc, L, w, bias, k, s, pad, Z, Lout, Lout_p, Cout, slope,
                     out, S(stream));
}
int xva_conv_c1_bwd_w(const float* dpre, const float* x, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k,
                      int s, int pad, int Z, int Lout, int Lout_p, int Cout, float* dw, float* db, void* stream) {
  return conv_c1_bwd_w(dpre, x, static_cast<long>(xs_b), xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout, dw, db,
                       S(stream));
}
int xva_conv_c1_bwd_x(const float* dpre, const float* w, int64_t xs_b, int xs_q, int xs_c, int P, int Lsrc, int L, int k,
                      int s, int pad, int Z, int Lout, int Lout_p, int Cout, float scale, float* dx, void* stream) {
  return conv_c1_bwd_x(dpre, w, static_cast<long>(xs_b), xs_q, xs_c, P, Lsrc, L, k, s, pad, Z, Lout, Lout_p, Cout, scale, dx,
                       S(stream));
}
int xva_avgpool4_fwd(const float* x, int B, int L, float* out, void* stream) { return avgpool4_fwd(x, B, L, out, S(stream)); }
int xva_avgpool4_bwd(const float* dout, int B, int L, float* dx, void* stream) { return avgpool4_bwd(dout, B, L, dx, S(stream)); }
int xva_zero_tail_rows(float* x, int Z, int Lp, int Lvalid, int C, void* stream) {
  return zero_tail_rows(x, Z, Lp, Lvalid, C, S(stream));
}

}  // extern "C"
