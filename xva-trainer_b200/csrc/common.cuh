// Shared host-side helpers for libxva_b200: error reporting across the C ABI, launch checks.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace xva {

// Thread-local last-error text returned by xva_last_error() (see include/xva_b200.h).
void set_error(const char* fmt, ...);
const char* last_error();

// Error codes of the C ABI (negative = failure).
enum : int {
  XVA_OK = 0,
  XVA_ERR_ARG = -1,      // invalid argument / unsupported shape
  XVA_ERR_CUDA = -2,     // CUDA runtime or driver failure (text in xva_last_error)
  XVA_ERR_DEVICE = -3,   // not a compute-capability 10.x device
};

#define XVA_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::xva::set_error(__VA_ARGS__);        \
      return ::xva::XVA_ERR_ARG;            \
    }                                       \
  } while (0)

#define XVA_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::xva::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::xva::XVA_ERR_CUDA;                                                              \
    }                                                                                          \
  } while (0)

#define XVA_CHECK_LAUNCH() XVA_CHECK_CUDA(cudaGetLastError())

int num_sms();  // SM count of the current device (cached)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Round-to-nearest (ties away) fp32 -> tf32, returned as fp32 with the low 13 mantissa bits cleared. The tensor core
// TRUNCATES fp32 operands to tf32; storing every GEMM operand pre-rounded makes the conversion unbiased, so operand
// rounding errors average out over a dot product instead of accumulating as a systematic -2^-12 relative bias.
// Test switch (xva_set_operand_rounding): with rounding off the fp32 checker GEMM reproduces a strict-fp32 reference to fp32
// rounding, which is how the tests prove the wiring independently of tensor-core precision. One copy per
// translation unit (the library is built without relocatable device code); XVA_DEFINE_ROUNDING_SWITCH(tag) emits
// the setter each .cu that rounds registers with capi.cu.
static __device__ int g_xva_round_operands = 1;
#define XVA_DEFINE_ROUNDING_SWITCH(tag)                                                        \
  int set_operand_rounding_##tag(int on) {                                                     \
    XVA_CHECK_CUDA(cudaMemcpyToSymbol(g_xva_round_operands, &on, sizeof(int)));                \
    return XVA_OK;                                                                             \
  }
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return g_xva_round_operands ? __uint_as_float(r) : x;
}
__device__ __forceinline__ float4 tf32_rn4(float4 v) {
  return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}

// Counter-based RNG shared by every kernel that applies dropout: the keep/drop decision for element
// `idx` of a tensor is a pure function of (seed, idx), so backward re-derives the mask instead of storing it.
// One 64-bit hash serves the four consecutive elements idx & ~3 .. (idx & ~3) + 3 (16 bits each): the row kernels and
// the GEMM epilogue handle float4 groups, and a per-element 64-bit hash made them instruction-bound (measured:
// softmax / LayerNorm backward at 2.4 TB/s instead of HBM speed).
__host__ __device__ __forceinline__ uint64_t hash_u64(uint64_t seed, uint64_t idx4) {
  uint64_t x = idx4 * 0x9E3779B97F4A7C15ull + seed;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  return x;
}
// Returns the multiplier of inverted dropout: 0 if dropped, 1/(1-p) if kept.  thresh = p * 2^32 (its top 16 bits are
// compared with the element's 16-bit field); thresh == 0 (p = 0) skips the hash.
__host__ __device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh,
                                                        float inv_keep) {
  if (thresh == 0u) return inv_keep;
  const uint32_t f = static_cast<uint32_t>(hash_u64(seed, idx >> 2) >> (16 * (idx & 3))) & 0xFFFFu;
  return f >= (thresh >> 16) ? inv_keep : 0.0f;
}
#ifdef __CUDACC__
// The multipliers of elements idx .. idx+3. One hash when idx is a multiple of 4 (the float4 case).
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint64_t idx, uint32_t thresh, float inv_keep) {
  if (thresh == 0u) return make_float4(inv_keep, inv_keep, inv_keep, inv_keep);
  if ((idx & 3) == 0) {
    const uint64_t h = hash_u64(seed, idx >> 2);
    const uint32_t t16 = thresh >> 16, lo = static_cast<uint32_t>(h), hi = static_cast<uint32_t>(h >> 32);
    return make_float4((lo & 0xFFFFu) >= t16 ? inv_keep : 0.0f, (lo >> 16) >= t16 ? inv_keep : 0.0f,
                       (hi & 0xFFFFu) >= t16 ? inv_keep : 0.0f, (hi >> 16) >= t16 ? inv_keep : 0.0f);
  }
  return make_float4(dropout_scale(seed, idx, thresh, inv_keep), dropout_scale(seed, idx + 1, thresh, inv_keep),
                     dropout_scale(seed, idx + 2, thresh, inv_keep), dropout_scale(seed, idx + 3, thresh, inv_keep));
}
#endif

}  // namespace xva
