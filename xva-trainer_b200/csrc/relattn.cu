// xVAPitch text encoder, the steps that are not tap-GEMMs, softmax or LayerNorm:
//   text_embed       cat(emb[tokens] * sqrt(C), language embedding) * mask       python/xvapitch/model.py:1152-1165
//   rel_band_add     relative-position logits added to the score band             python/xvapitch/glow_tts.py:178-186
//   rel_band_gather  the band of the attention weights in relative indexing       python/xvapitch/glow_tts.py:192-195
//   pad_cols         row pitch rounded up to 32 for the weight-gradient operands
// All five are element-wise, HBM-bound and tiny (a text encoder step touches ~B x T x 204 floats per tensor): one
// thread per output element, coalesced along the channel / key dimension, grid sized to the data and capped at a few
// waves. The per-element functions live in relattn_body.h, which the CPU tests also compile for the host.
#include "common.cuh"
#include "ops.cuh"

#define XVA_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define XVA_RN(x) ::xva::tf32_rn(x)
#define XVA_ADD(p, v) atomicAdd((p), (v))
#else  // host pass of nvcc: the functions are never called there
#define XVA_RN(x) (x)
#define XVA_ADD(p, v) (*(p) += (v))
#endif
#include "relattn_body.h"

namespace xva {

namespace {

inline int elem_grid(long total) {
  long b = ceil_div_l(total, 256);
  const long cap = 16L * num_sms();
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

#define XVA_ELEM_LOOP(total) \
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < (total); i += static_cast<long>(gridDim.x) * blockDim.x)

__global__ void __launch_bounds__(256)
text_embed_fwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb, const float* __restrict__ lang,
                      const int* __restrict__ lens, long total, int T, int C, int L, int ld, float scale,
                      float* __restrict__ out, float* __restrict__ x_emb) {
  XVA_ELEM_LOOP(total) relattn::text_embed_fwd_elem(i, tokens, emb, lang, lens, T, C, L, ld, scale, out, x_emb);
}

__global__ void __launch_bounds__(256)
text_embed_bwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ dout, const int* __restrict__ lens,
                      long total, int T, int C, int ld, float scale, float* __restrict__ demb) {
  XVA_ELEM_LOOP(total) relattn::text_embed_bwd_elem(i, tokens, dout, lens, T, C, ld, scale, demb);
}

__global__ void __launch_bounds__(256)
rel_band_add_kernel(float* __restrict__ s, const float* __restrict__ rel, long total, int T, int W, int ld, int ldr) {
  XVA_ELEM_LOOP(total) relattn::rel_band_add_elem(i, s, rel, T, W, ld, ldr);
}

__global__ void __launch_bounds__(256)
rel_band_gather_kernel(const float* __restrict__ p, long total, int T, int W, int ld, int ldo, float* __restrict__ out) {
  XVA_ELEM_LOOP(total) relattn::rel_band_gather_elem(i, p, T, W, ld, ldo, out);
}

__global__ void __launch_bounds__(256)
pad_cols_kernel(const float* __restrict__ src, long total, int C, int ld, float* __restrict__ dst) {
  XVA_ELEM_LOOP(total) relattn::pad_cols_elem(i, src, C, ld, dst);
}

}  // namespace

int text_embed_fwd(const long long* tokens, const float* emb, const float* lang, const int* lens, int B, int T, int C, int L,
                   int ld, float scale, float* out, float* x_emb, cudaStream_t stream) {
  XVA_CHECK_ARG(tokens && emb && out, "text_embed fwd: null tokens / emb / out");
  XVA_CHECK_ARG(B >= 1 && T >= 1 && C >= 1 && L >= 0 && ld >= C + L, "text_embed fwd: B=%d T=%d C=%d L=%d ld=%d", B, T, C, L, ld);
  XVA_CHECK_ARG(L == 0 || lang, "text_embed fwd: L=%d without a language embedding", L);
  const long total = static_cast<long>(B) * T * ld;
  text_embed_fwd_kernel<<<elem_grid(total), 256, 0, stream>>>(tokens, emb, lang, lens, total, T, C, L, ld, scale, out, x_emb);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int text_embed_bwd(const long long* tokens, const float* dout, const int* lens, int B, int T, int C, int ld, float scale,
                   float* demb, cudaStream_t stream) {
  XVA_CHECK_ARG(tokens && dout && demb, "text_embed bwd: null tokens / dout / demb");
  XVA_CHECK_ARG(B >= 1 && T >= 1 && C >= 1 && ld >= C, "text_embed bwd: B=%d T=%d C=%d ld=%d", B, T, C, ld);
  const long total = static_cast<long>(B) * T * C;
  text_embed_bwd_kernel<<<elem_grid(total), 256, 0, stream>>>(tokens, dout, lens, total, T, C, ld, scale, demb);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int rel_band_add(float* s, const float* rel, int Z, int T, int W, int ld, int ldr, cudaStream_t stream) {
  XVA_CHECK_ARG(s && rel, "rel_band_add: null operand");
  XVA_CHECK_ARG(Z >= 1 && T >= 1 && W >= 0 && ld >= T && ldr >= 2 * W + 1, "rel_band_add: Z=%d T=%d W=%d ld=%d ldr=%d", Z, T, W, ld, ldr);
  const long total = static_cast<long>(Z) * T * (2 * W + 1);
  rel_band_add_kernel<<<elem_grid(total), 256, 0, stream>>>(s, rel, total, T, W, ld, ldr);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int rel_band_gather(const float* p, int Z, int T, int W, int ld, int ldo, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(p && out, "rel_band_gather: null operand");
  XVA_CHECK_ARG(Z >= 1 && T >= 1 && W >= 0 && ld >= T && ldo >= 2 * W + 1, "rel_band_gather: Z=%d T=%d W=%d ld=%d ldo=%d", Z, T, W, ld, ldo);
  const long total = static_cast<long>(Z) * T * ldo;
  rel_band_gather_kernel<<<elem_grid(total), 256, 0, stream>>>(p, total, T, W, ld, ldo, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int pad_cols(const float* src, long rows, int C, int ld, float* dst, cudaStream_t stream) {
  XVA_CHECK_ARG(src && dst && rows >= 0 && C >= 1 && ld >= C, "pad_cols: rows=%ld C=%d ld=%d", rows, C, ld);
  if (rows == 0) return XVA_OK;
  const long total = rows * ld;
  pad_cols_kernel<<<elem_grid(total), 256, 0, stream>>>(src, total, C, ld, dst);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(relattn)

}  // namespace xva
