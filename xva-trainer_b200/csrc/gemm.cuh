// Internal description of one tap-GEMM launch (tcgen05 kernel in gemm_tc.cu, SIMT checker in gemm_ref.cu).
//
// One kernel family covers every dense contraction of the FastPitch / HiFi-GAN hot path:
//
//   mode 0 (A K-major, B K-major)    out[z,r,n] = alpha * sum_j sum_k A[z, r+shift_j, k] * B[zb(j,z), n, k]
//        Linear / Conv1d forward on channels-last activations: A = activations [Z,R,K], B = packed weights
//        [taps,N,K]; also Q*K^T (B batched over z).
//   mode 1 (A K-major, B MN-major)   out[z,r,n] = alpha * sum_j sum_k A[z, r+shift_j, k] * B[zb(j,z), k, n]
//        Conv1d dgrad against the same packed weights (contraction over Cout), P*V, dS*K.
//   mode 2 (A MN-major, B MN-major)  out[zo,j,m,n] (+)= sum_{zr<ZR} sum_t A[z,t,m] * B[z, t+shift_j, n],  z = zo*ZR+zr
//        Conv1d / Linear wgrad (ZO=1, ZR=batch, split over z with fp32 atomics), dK = dS^T Q, dV = P^T dO (ZR=1).
//
// Rows addressed outside [0,rows) of an operand read as zero (TMA out-of-bounds fill) -- that IS the conv zero
// padding and the per-utterance halo of the reference (transformer.py:46-52 Conv1d padding=k//2).
#pragma once
#include <cstdint>
#include "common.cuh"
#include "../../include/xva_b200.h"

namespace xva {

constexpr int kMaxTaps = XVA_MAX_TAPS;

enum GemmFlags : int {
  GEMM_RELU = XVA_GEMM_RELU,              // v = max(v, 0) after bias
  GEMM_LN = XVA_GEMM_LN,                  // LayerNorm over the N columns of each row (one n-tile: N <= 512)
  GEMM_DROP_PRE = XVA_GEMM_DROP_PRE,      // dropout before the residual add / LN (transformer.py:51,139)
  GEMM_DROP_POST = XVA_GEMM_DROP_POST,    // dropout after LN (ConvReLUNorm: common/layers.py:94-97)
  GEMM_ATOMIC = XVA_GEMM_ATOMIC,          // mode 2: accumulate into `out` with fp32 atomics (split-z)
  GEMM_LRELU_GATE = XVA_GEMM_LRELU_GATE,  // reserved
  GEMM_TANH = XVA_GEMM_TANH,
  GEMM_SOFTMAX_BWD = XVA_GEMM_SOFTMAX_BWD,
  GEMM_HALO = XVA_GEMM_HALO,  // see include/xva_b200.h
  GEMM_ROUND_OUT = XVA_GEMM_ROUND_OUT,    // out is a later GEMM operand: store it rounded to tf32
};

// The argument block IS the C-ABI struct (include/xva_b200.h); field meaning is documented there.
using GemmArgs = xva_gemm_args;

inline GemmArgs gemm_args() {
  GemmArgs g{};
  g.Z = 1;
  g.taps = 1;
  g.ZR = 1;
  g.split = 1;
  g.b_nz = 1;
  g.alpha = 1.0f;
  g.ln_eps = 1e-5f;
  return g;
}

// tcgen05 + TMA implementation (the product path).
int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream);
// fp32 tensor map with zero out-of-bounds fill (cuTensorMapEncodeTiled): dims / strides in elements, innermost first,
// strides[0] implicit; `map` points at a CUtensorMap; swizzle is a CUtensorMapSwizzle value. Shared with attn_fused.cu.
int tma_encode_f32(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, int swizzle);

// Bring-up: cycle counters of CTA 0 from the last launch made with XVA_GEMM_DBG & 32 (see gemm_tc.cu).
int gemm_debug_counters(long long* out8);
// Plain fp32 SIMT implementation of the same contract; used by the tests to separate "descriptor/layout bug"
// from "host wiring bug". Never called by the product path.
int gemm_ref_launch(const GemmArgs& g, cudaStream_t stream);

}  // namespace xva
