// Element functions of the xVAPitch text-encoder kernels (csrc/relattn.cu): one call computes ONE output element, no
// shared memory, no warp collectives. The __global__ kernels in relattn.cu are grid-stride loops over these functions.
// The same functions compile as plain C++: the CPU tests build them with g++ (tests/relattn_host.cpp) and run the loops on
// the host, so the index arithmetic below is checked against the CPU restatement of the reference without a GPU.
//   XVA_HD          function qualifiers (__host__ __device__ __forceinline__ under nvcc, nothing under g++)
//   XVA_RN(x)       store rounding of a GEMM operand (tf32_rn on the device; the host build rounds or not by a switch)
//   XVA_ADD(p, v)   accumulation into memory several elements may hit (atomicAdd on the device, += on the host)
#pragma once
#include <stdint.h>

namespace xva {
namespace relattn {

// TextEncoder.forward, stats=False, python/xvapitch/model.py:1152-1165: x = cat(emb[tokens] * sqrt(C), lang) * mask.
// Element i of out [B, T, ld] (ld >= C + L: columns [C + L, ld) are zero so the tensor is a legal MN-major operand).
// x_emb [B, T, C] (optional) receives the scaled embedding of EVERY position, masked or not, as the reference returns it.
XVA_HD void text_embed_fwd_elem(long i, const long long* tokens, const float* emb, const float* lang, const int* lens,
                                int T, int C, int L, int ld, float scale, float* out, float* x_emb) {
  const long row = i / ld;
  const int c = static_cast<int>(i - row * ld);
  const int b = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long>(b) * T);
  const bool live = (lens == nullptr) || (t < lens[b]);
  float v = 0.0f;
  if (c < C) {
    const float e = emb[static_cast<long>(tokens[row]) * C + c] * scale;
    if (x_emb) x_emb[row * C + c] = e;
    if (live) v = e;
  } else if (c < C + L) {
    if (live) v = lang[static_cast<long>(b) * L + (c - C)];
  }
  out[i] = XVA_RN(v);
}

// Backward of the embedding half: element i of [B, T, C]; demb[tokens[row], c] += scale * dout[row, c] on live rows.
// (The language half is a per-item column sum: xva_colsum_items on the column slice [C, C + L).)
XVA_HD void text_embed_bwd_elem(long i, const long long* tokens, const float* dout, const int* lens, int T, int C, int ld,
                                float scale, float* demb) {
  const long row = i / C;
  const int c = static_cast<int>(i - row * C);
  const int b = static_cast<int>(row / T);
  const int t = static_cast<int>(row - static_cast<long>(b) * T);
  if (lens != nullptr && t >= lens[b]) return;
  XVA_ADD(demb + static_cast<long>(tokens[row]) * C + c, scale * dout[row * ld + c]);
}

// Relative-position term of the attention scores, python/xvapitch/glow_tts.py:178-186 (the pad / reshape skewing of
// :260-277 written as an index shift): s[z, t, t + r - W] += rel[z, t, r] for r in [0, 2W], key index inside [0, T).
// Element i of [Z, T, 2W + 1]. Used twice: on the scores (rel = q . E_k^T) and, in the backward, on dP (rel = dO . E_v^T).
XVA_HD void rel_band_add_elem(long i, float* s, const float* rel, int T, int W, int ld, int ldr) {
  const int nb = 2 * W + 1;
  const long row = i / nb;
  const int r = static_cast<int>(i - row * nb);
  const int t = static_cast<int>(row % T);
  const int j = t + r - W;
  if (j < 0 || j >= T) return;
  s[row * ld + j] += rel[row * ldr + r];
}

// The inverse view, glow_tts.py:192-193 (:279-292): out[z, t, r] = p[z, t, t + r - W] for r in [0, 2W] with the key
// inside [0, T), zero otherwise and in the pad columns [2W + 1, ldo). Element i of out [Z, T, ldo]; the result is the
// K-major operand of the product with E_v (forward) / E_k (backward) and the MN-major operand of dE_v / dE_k.
XVA_HD void rel_band_gather_elem(long i, const float* p, int T, int W, int ld, int ldo, float* out) {
  const long row = i / ldo;
  const int r = static_cast<int>(i - row * ldo);
  const int t = static_cast<int>(row % T);
  const int j = t + r - W;
  float v = 0.0f;
  if (r <= 2 * W && j >= 0 && j < T) v = p[row * ld + j];
  out[i] = XVA_RN(v);
}

// dst [rows, ld] = src [rows, C] with zero pad columns: the weight-gradient GEMM reads its operands MN-major in 32-column
// chunks, so a tensor whose channel count is not a multiple of 32 (204 = 192 + 12 here) needs a row pitch rounded up.
XVA_HD void pad_cols_elem(long i, const float* src, int C, int ld, float* dst) {
  const long row = i / ld;
  const int c = static_cast<int>(i - row * ld);
  dst[i] = (c < C) ? src[row * C + c] : 0.0f;
}

}  // namespace relattn
}  // namespace xva
