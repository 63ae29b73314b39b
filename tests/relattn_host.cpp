// Host build of csrc/relattn_body.h for the CPU tests (tests/cabi_emu.py): the per-element functions the CUDA kernels of
// csrc/relattn.cu loop over, compiled with g++ and driven by plain loops, behind the extern "C" signatures of
// include/xva_b200.h. Test infrastructure only -- never loaded by the product package.
#include <stdint.h>
#include <string.h>

#define XVA_HD static inline
// Store rounding of GEMM operands: off = the exact-arithmetic test mode (xva_set_operand_rounding(0)); on = cvt.rna.tf32.f32
// (round to nearest, ties away, 13 low mantissa bits cleared), the product path's precision.
static int g_round = 0;
static inline float emu_rn(float x) {
  if (!g_round) return x;
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
#define XVA_RN(x) emu_rn(x)
#define XVA_ADD(p, v) (*(p) += (v))
#include "../xva-trainer_b200/csrc/relattn_body.h"

using namespace xva::relattn;

extern "C" {

void xva_emu_set_rounding(int on) { g_round = on; }

int xva_text_embed_fwd(const int64_t* tokens, const float* emb, const float* lang, const int32_t* lens, int B, int T, int C,
                       int L, int ld, float scale, float* out, float* x_emb, void*) {
  if (!tokens || !emb || !out || ld < C + L || (L > 0 && !lang)) return -1;
  const long total = static_cast<long>(B) * T * ld;
  for (long i = 0; i < total; ++i)
    text_embed_fwd_elem(i, reinterpret_cast<const long long*>(tokens), emb, lang, lens, T, C, L, ld, scale, out, x_emb);
  return 0;
}

int xva_text_embed_bwd(const int64_t* tokens, const float* dout, const int32_t* lens, int B, int T, int C, int ld, float scale,
                       float* demb, void*) {
  if (!tokens || !dout || !demb || ld < C) return -1;
  const long total = static_cast<long>(B) * T * C;
  for (long i = 0; i < total; ++i) text_embed_bwd_elem(i, reinterpret_cast<const long long*>(tokens), dout, lens, T, C, ld, scale, demb);
  return 0;
}

int xva_rel_band_add(float* s, const float* rel, int Z, int T, int W, int ld, int ldr, void*) {
  if (!s || !rel || ld < T || ldr < 2 * W + 1) return -1;
  const long total = static_cast<long>(Z) * T * (2 * W + 1);
  for (long i = 0; i < total; ++i) rel_band_add_elem(i, s, rel, T, W, ld, ldr);
  return 0;
}

int xva_rel_band_gather(const float* p, int Z, int T, int W, int ld, int ldo, float* out, void*) {
  if (!p || !out || ld < T || ldo < 2 * W + 1) return -1;
  const long total = static_cast<long>(Z) * T * ldo;
  for (long i = 0; i < total; ++i) rel_band_gather_elem(i, p, T, W, ld, ldo, out);
  return 0;
}

int xva_pad_cols(const float* src, int64_t rows, int C, int ld, float* dst, void*) {
  if (!src || !dst || ld < C) return -1;
  const long total = static_cast<long>(rows) * ld;
  for (long i = 0; i < total; ++i) pad_cols_elem(i, src, C, ld, dst);
  return 0;
}

}  // extern "C"
