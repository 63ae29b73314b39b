// Fused single-head attention for the FFT blocks (MultiHeadAttn._forward, fastpitch/transformer.py:100-152, n_head = 1,
// d_head = 64): scores, key mask, softmax, attention dropout and P.V in ONE kernel per direction. The score tile lives
// in tensor memory: S = Q.K^T is accumulated there by tcgen05.mma, read back by the softmax threads (one query row per
// thread = one TMEM lane), overwritten in place by the tf32-rounded, dropout-scaled probabilities and consumed from there
// as the A operand of the second tcgen05.mma (O += P.V) -- the [B, T, T] score / probability tensors (2 x 101 MB per
// layer at 32 x 880, written and re-read six times by the unfused path) never exist. Online softmax over 128-key chunks
// (running row maximum and sum; the O accumulator is rescaled in tensor memory when the maximum moves); the backward
// kernels recompute P from the saved row log-sum-exp.
//
// Forward CTA = (utterance, 128 query rows), 128 threads, two CTAs per SM (96 KiB of shared memory, 256 TMEM columns each):
// the sequential steps of one CTA (TMA -> MMA -> softmax -> MMA) overlap with the other's.
//   smem: Q tile 128 x 64 (K-major, SWIZZLE_128B, two 32-column k-blocks), K chunk 128 x 64 (same), V chunk 128 keys x 64
//         (MN-major operand of P.V: four 32-key k-blocks of [2 column chunks][32 rows][32 floats], 128B_ATOM_32B swizzle)
//   tmem: columns [0, 128) S / P, [128, 192) O
// Dropout: keep / drop of element (b, row, key) is the shared counter hash of ((b * T + row) * drop_ld + key), the same
// index the unfused softmax kernels use, so both paths draw identical masks from one seed.
#include <cuda.h>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include "gemm.cuh"
#include "ptx.cuh"

namespace xva {

namespace {

constexpr int kAttnRows = 128;   // query rows per CTA (= TMEM lanes)
constexpr int kAttnKeys = 128;   // keys per chunk
constexpr int kAttnD = 64;
constexpr int kAttnThreads = 128;
constexpr int kTileBytes = kAttnRows * kAttnD * 4;          // 32 KiB: one 128 x 64 fp32 operand tile
constexpr int kFwdSmem = 3 * kTileBytes + 1024 + 128;       // Q, K, V + alignment slack + barriers
constexpr int kFwdTmemCols = 256;

struct AttnDev {
  int B, T, drop_ld;
  const int* lens;
  float* out;
  long o_rs, o_zs;
  float* lse;
  float scale_log2;  // softmax scale * log2(e): probabilities are exp2(s * scale_log2 - m)
  uint32_t drop_thresh;
  float inv_keep;
  uint64_t seed;
  const uint64_t* seed_dev;
  int round_on;
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// tf32 x tf32 -> f32, M = 128
__device__ __forceinline__ uint32_t instr_desc(int n, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= static_cast<uint32_t>(b_mn_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(kAttnRows >> 4) << 24;
  return d;
}
__device__ __forceinline__ uint32_t rna_tf32(float x, int on) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return on ? r : __float_as_uint(x);
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v,
                const __grid_constant__ AttnDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, q0 = blockIdx.x * kAttnRows;
  const int T = p.T;
  const int lim = p.lens ? (p.lens[b] < T ? p.lens[b] : T) : T;   // keys [0, lim) are attended to
  const int n_chunks = (lim + kAttnKeys - 1) / kAttnKeys;

  if (warp == 0 && ptx::elect_one()) {
    ptx::tma_prefetch_desc(&map_qk);
    ptx::tma_prefetch_desc(&map_v);
    for (int i = 0; i < 5; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, kFwdTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t tS = tmem_base + lane_base;          // this thread's row of S / P
  const uint32_t tO = tmem_base + lane_base + 128;    // ... of O

  auto issue_k = [&](int c) {
    ptx::mbar_arrive_expect_tx(bar_k, kTileBytes);
    ptx::tma_load_3d(sK, &map_qk, bar_k, kAttnD, c * kAttnKeys, b);
    ptx::tma_load_3d(sK + kTileBytes / 2, &map_qk, bar_k, kAttnD + 32, c * kAttnKeys, b);
  };
  auto issue_v = [&](int c) {
    ptx::mbar_arrive_expect_tx(bar_v, kTileBytes);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
      ptx::tma_load_4d(sV + kb * (kTileBytes / 4), &map_v, bar_v, 0, c * kAttnKeys + kb * 32, (2 * kAttnD) / 32, b);
  };

  if (n_chunks > 0 && warp == 0 && ptx::elect_one()) {
    ptx::mbar_arrive_expect_tx(bar_q, kTileBytes);
    ptx::tma_load_3d(sQ, &map_qk, bar_q, 0, q0, b);
    ptx::tma_load_3d(sQ + kTileBytes / 2, &map_qk, bar_q, 32, q0, b);
    issue_k(0);
    issue_v(0);
  }
  __syncwarp();

  const uint32_t idesc_s = instr_desc(kAttnKeys, 0);
  const uint32_t idesc_o = instr_desc(kAttnD, 1);
  const uint64_t dq0 = smem_desc(ptx::smem_u32(sQ), 16, 1024, 2);
  const uint64_t dk0 = smem_desc(ptx::smem_u32(sK), 16, 1024, 2);
  const uint64_t dv0 = smem_desc(ptx::smem_u32(sV), 32 * 128, 512, 1);
  const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) * 0xA24BAED4963EE407ull : 0ull);
  const int row = q0 + tid;
  const float c2 = p.scale_log2;
  float m_run = -INFINITY, l_run = 0.0f;

  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t ph = static_cast<uint32_t>(c & 1);
    // ---- S = Q.K^T into tensor memory
    if (warp == 0) {
      if (c == 0) ptx::mbar_wait(bar_q, 0);
      ptx::mbar_wait(bar_k, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            ptx::mma_tf32(tmem_base, dq0 + kb * (kTileBytes / 2 / 16) + k4 * 2, dk0 + kb * (kTileBytes / 2 / 16) + k4 * 2,
                          idesc_s, (kb | k4) ? 1u : 0u);
        ptx::mma_commit(bar_s);
      }
      __syncwarp();
    }
    ptx::mbar_wait(bar_s, ph);
    ptx::tc_fence_after();
    if (warp == 0 && c + 1 < n_chunks && ptx::elect_one()) issue_k(c + 1);   // the K tile is free again
    __syncwarp();

    // ---- online softmax of this thread's row over the chunk's keys
    const int key0 = c * kAttnKeys;
    const int nvalid = (lim - key0) < kAttnKeys ? (lim - key0) : kAttnKeys;   // >= 1
    float mx = -INFINITY;
    // (not unrolled: the fully unrolled chunk loop was 124 KB of SASS and stalled on instruction fetch, profiles/r02_ncu_attn.txt)
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
      if (blk * 32 < nvalid) {
        uint32_t v[32];
        ptx::tmem_ld32(tS + blk * 32, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (blk * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
    }
    const float m_new = fmaxf(m_run, mx * c2);
    const float alpha = ptx::ex2_approx(m_run - m_new);   // 0 on the first chunk (m_run = -inf)
    float lsum = 0.0f;
    const uint64_t idx_row = (static_cast<uint64_t>(b) * T + row) * static_cast<uint64_t>(p.drop_ld) + key0;
    // (not unrolled: the fully unrolled chunk loop was 124 KB of SASS and stalled on instruction fetch, profiles/r02_ncu_attn.txt)
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
      uint32_t v[32];
      if (blk * 32 < nvalid) {
        ptx::tmem_ld32(tS + blk * 32, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 ds = dropout_scale4(seed, idx_row + blk * 32 + 4 * i4, p.drop_thresh, p.inv_keep);
          const float dsv[4] = {ds.x, ds.y, ds.z, ds.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            const float pv = (blk * 32 + i < nvalid) ? ptx::ex2_approx(__uint_as_float(v[i]) * c2 - m_new) : 0.0f;
            lsum += pv;
            v[i] = rna_tf32(pv * dsv[e], p.round_on);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      ptx::tmem_st32(tS + blk * 32, v);
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    if (c > 0) {   // the running maximum moved: rescale the O accumulator in place
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld32(tO + h * 32, v);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        ptx::tmem_st32(tO + h * 32, v);
      }
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();

    // ---- O += P.V, P read from tensor memory
    if (warp == 0) {
      ptx::mbar_wait(bar_v, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            ptx::mma_tf32_ts(tmem_base + 128, tmem_base + kb * 32 + k4 * 8, dv0 + kb * (kTileBytes / 4 / 16) + k4 * (1024 / 16),
                             idesc_o, (c > 0 || (kb | k4)) ? 1u : 0u);
        ptx::mma_commit(bar_o);
      }
      __syncwarp();
    }
    ptx::mbar_wait(bar_o, ph);
    ptx::tc_fence_after();
    if (warp == 0 && c + 1 < n_chunks && ptx::elect_one()) issue_v(c + 1);
    __syncwarp();
  }

  // ---- O / l -> vec (tf32-rounded: it is the operand of the output projection), log-sum-exp for the backward
  const bool valid_row = row < T;
  float* dst = p.out + static_cast<long>(b) * p.o_zs + static_cast<long>(row) * p.o_rs;
  if (n_chunks > 0) {
    const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      ptx::tmem_ld32(tO + h * 32, v);
      ptx::tmem_wait_ld();
      if (valid_row) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          uint4 o;
          o.x = rna_tf32(__uint_as_float(v[i]) * inv, p.round_on);
          o.y = rna_tf32(__uint_as_float(v[i + 1]) * inv, p.round_on);
          o.z = rna_tf32(__uint_as_float(v[i + 2]) * inv, p.round_on);
          o.w = rna_tf32(__uint_as_float(v[i + 3]) * inv, p.round_on);
          *reinterpret_cast<uint4*>(dst + h * 32 + i) = o;
        }
      }
    }
  } else if (valid_row) {
#pragma unroll
    for (int i = 0; i < kAttnD; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (valid_row && p.lse)
    p.lse[static_cast<long>(b) * T + row] = l_run > 0.0f ? (m_run + log2f(l_run)) * 0.6931471805599453f : INFINITY;

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kFwdTmemCols);
}

int g_attn_round_host = 1;

// ================================================================================================ backward
// Two kernels, both recomputing P from the saved row log-sum-exp (nothing score-sized is ever stored):
//   attn_bwd_dq_kernel   CTA = (utterance, 128 query rows), loop over 128-key chunks:
//        S = Q.K^T, dP = dO.V^T  (tensor memory)  ->  dS = scale * P o (dP o M/keep - D[row])  (in place of S)  ->  dQ += dS.K
//   attn_bwd_dkv_kernel  CTA = (utterance, 128 keys), loop over 128-query chunks, everything transposed:
//        S^T = K.Q^T, dP^T = V.dO^T  ->  P_d^T (in place of S^T), dS^T (in place of dP^T)  ->  dV += P_d^T.dO, dK += dS^T.Q
// with P = exp(scale * S - lse[row]), M the dropout keep mask (same counter hash as the forward), D[row] = dO[row].O[row]
// (xva_rowdot2). The A operand of every second-stage product is read from tensor memory; its B operand (K, Q or dO as
// [d, rows]) is an MN-major tile, so those three tensors are staged twice per chunk: K-major for the first-stage
// products and MN-major for the second. 256 threads: two warps per TMEM lane quarter split a row's 128 columns.
constexpr int kBwdThreads = 512;   // 16 warps: four per TMEM lane quarter, each owns 32 of a chunk's 128 columns
constexpr int kBwdTmemCols = 512;
constexpr int kDqSmem = 6 * kTileBytes + 1024 + 128;     // Q, dO, K, V, K (MN-major) x 2
constexpr int kDkvSmem = 6 * kTileBytes + 1024 + 128 + 2 * kAttnRows * 4;   // K, V, Q, dO, Q (MN), dO (MN) + lse / D of the chunk

struct AttnBwdDev {
  int B, T, drop_ld;
  const int* lens;
  const float* lse;
  const float* dsum;   // D[b * T + row] = dO[row] . O[row]
  float* dqkv;
  long g_rs, g_zs;
  float scale, scale_log2;
  uint32_t drop_thresh;
  float inv_keep;
  uint64_t seed;
  const uint64_t* seed_dev;
  int round_on;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_m,
                   const __grid_constant__ CUtensorMap map_do, const __grid_constant__ AttnBwdDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + kTileBytes;
  uint8_t* sK = smem + 2 * kTileBytes;
  uint8_t* sV = smem + 3 * kTileBytes;
  uint8_t* sKm = smem + 4 * kTileBytes;   // two buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);
  uint64_t* bar_qd = bars;
  uint64_t* bar_kv = bars + 1;
  uint64_t* bar_m = bars + 2;   // [2]
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int b = blockIdx.y, q0 = blockIdx.x * kAttnRows;
  const int T = p.T;
  const int lim = p.lens ? (p.lens[b] < T ? p.lens[b] : T) : T;
  const int n_chunks = (lim + kAttnKeys - 1) / kAttnKeys;

  if (warp == 0 && ptx::elect_one()) {
    ptx::tma_prefetch_desc(&map_qk);
    ptx::tma_prefetch_desc(&map_m);
    ptx::tma_prefetch_desc(&map_do);
    for (int i = 0; i < 6; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, kBwdTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
  const uint32_t tS = tmem_base + lane_base;            // [0, 128)   S, then dS
  const uint32_t tdP = tmem_base + lane_base + 128;     // [128, 256) dP
  const uint32_t tdQ = tmem_base + lane_base + 256;     // [256, 320) dQ

  auto issue_kv = [&](int c) {
    ptx::mbar_arrive_expect_tx(bar_kv, 2 * kTileBytes);
    ptx::tma_load_3d(sK, &map_qk, bar_kv, kAttnD, c * kAttnKeys, b);
    ptx::tma_load_3d(sK + kTileBytes / 2, &map_qk, bar_kv, kAttnD + 32, c * kAttnKeys, b);
    ptx::tma_load_3d(sV, &map_qk, bar_kv, 2 * kAttnD, c * kAttnKeys, b);
    ptx::tma_load_3d(sV + kTileBytes / 2, &map_qk, bar_kv, 2 * kAttnD + 32, c * kAttnKeys, b);
  };
  auto issue_m = [&](int c) {
    uint64_t* bar = &bar_m[c & 1];
    uint8_t* dst = sKm + (c & 1) * kTileBytes;
    ptx::mbar_arrive_expect_tx(bar, kTileBytes);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
      ptx::tma_load_4d(dst + kb * (kTileBytes / 4), &map_m, bar, 0, c * kAttnKeys + kb * 32, kAttnD / 32, b);
  };

  if (n_chunks > 0 && warp == 0 && ptx::elect_one()) {
    ptx::mbar_arrive_expect_tx(bar_qd, 2 * kTileBytes);
    ptx::tma_load_3d(sQ, &map_qk, bar_qd, 0, q0, b);
    ptx::tma_load_3d(sQ + kTileBytes / 2, &map_qk, bar_qd, 32, q0, b);
    ptx::tma_load_3d(sdO, &map_do, bar_qd, 0, q0, b);
    ptx::tma_load_3d(sdO + kTileBytes / 2, &map_do, bar_qd, 32, q0, b);
    issue_kv(0);
    issue_m(0);
  }
  __syncwarp();

  const uint32_t idesc_s = instr_desc(kAttnKeys, 0);
  const uint32_t idesc_o = instr_desc(kAttnD, 1);
  const uint64_t dq0 = smem_desc(ptx::smem_u32(sQ), 16, 1024, 2);
  const uint64_t ddo0 = smem_desc(ptx::smem_u32(sdO), 16, 1024, 2);
  const uint64_t dk0 = smem_desc(ptx::smem_u32(sK), 16, 1024, 2);
  const uint64_t dv0 = smem_desc(ptx::smem_u32(sV), 16, 1024, 2);
  const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) * 0xA24BAED4963EE407ull : 0ull);
  const int row = q0 + quarter * 32 + lane;
  const bool valid_row = row < T;
  const float lse2 = valid_row ? __ldg(p.lse + static_cast<long>(b) * T + row) * 1.4426950408889634f : INFINITY;
  const float Dr = valid_row ? __ldg(p.dsum + static_cast<long>(b) * T + row) : 0.0f;
  const float c2 = p.scale_log2, sc = p.scale;
  constexpr uint32_t kb_step = kTileBytes / 2 / 16;   // K-major k-block (32 columns) in descriptor units

  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t ph = static_cast<uint32_t>(c & 1);
    if (warp == 0) {
      if (c == 0) ptx::mbar_wait(bar_qd, 0);
      ptx::mbar_wait(bar_kv, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            ptx::mma_tf32(tmem_base, dq0 + kb * kb_step + k4 * 2, dk0 + kb * kb_step + k4 * 2, idesc_s, (kb | k4) ? 1u : 0u);
            ptx::mma_tf32(tmem_base + 128, ddo0 + kb * kb_step + k4 * 2, dv0 + kb * kb_step + k4 * 2, idesc_s, (kb | k4) ? 1u : 0u);
          }
        ptx::mma_commit(bar_s);
      }
      __syncwarp();
    }
    ptx::mbar_wait(bar_s, ph);
    ptx::tc_fence_after();
    if (warp == 0 && c + 1 < n_chunks && ptx::elect_one()) issue_kv(c + 1);
    __syncwarp();

    const int key0 = c * kAttnKeys;
    const int nvalid = (lim - key0) < kAttnKeys ? (lim - key0) : kAttnKeys;
    const uint64_t idx_row = (static_cast<uint64_t>(b) * T + row) * static_cast<uint64_t>(p.drop_ld) + key0;
    {
      const int col0 = half * 32;            // `half` is the warp's column part 0..3 here (16 warps)
      uint32_t s[32];
      if (col0 < nvalid) {
        uint32_t dp[32];
        ptx::tmem_ld32(tS + col0, s);
        ptx::tmem_ld32(tdP + col0, dp);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 ds = dropout_scale4(seed, idx_row + col0 + 4 * i4, p.drop_thresh, p.inv_keep);
          const float dsv[4] = {ds.x, ds.y, ds.z, ds.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            const float pv = (col0 + i < nvalid) ? ptx::ex2_approx(__uint_as_float(s[i]) * c2 - lse2) : 0.0f;
            s[i] = rna_tf32(sc * pv * (__uint_as_float(dp[i]) * dsv[e] - Dr), p.round_on);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = 0u;
      }
      ptx::tmem_st32(tS + col0, s);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();

    if (warp == 0) {
      ptx::mbar_wait(&bar_m[c & 1], static_cast<uint32_t>((c >> 1) & 1));
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dm0 = smem_desc(ptx::smem_u32(sKm + (c & 1) * kTileBytes), 32 * 128, 512, 1);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            ptx::mma_tf32_ts(tmem_base + 256, tmem_base + kb * 32 + k4 * 8, dm0 + kb * (kTileBytes / 4 / 16) + k4 * (1024 / 16),
                             idesc_o, (c > 0 || (kb | k4)) ? 1u : 0u);
        ptx::mma_commit(bar_o);
        if (c + 1 < n_chunks) issue_m(c + 1);   // the other buffer: its last reader (chunk c - 1) finished before bar_o(c - 1)
      }
      __syncwarp();
    }
    ptx::mbar_wait(bar_o, ph);
    ptx::tc_fence_after();
  }

  // ---- dQ -> dqkv[:, :, 0:64] (tf32-rounded: operand of the qkv weight gradient and input gradient)
  float* dst = p.dqkv + static_cast<long>(b) * p.g_zs + static_cast<long>(row) * p.g_rs + (half & 1) * 32;
  if (half >= 2) {
    // column parts 2, 3 have nothing to store (dQ is 64 wide)
  } else if (n_chunks > 0) {
    uint32_t v[32];
    ptx::tmem_ld32(tdQ + half * 32, v);
    ptx::tmem_wait_ld();
    if (valid_row) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        uint4 o;
        o.x = rna_tf32(__uint_as_float(v[i]), p.round_on);
        o.y = rna_tf32(__uint_as_float(v[i + 1]), p.round_on);
        o.z = rna_tf32(__uint_as_float(v[i + 2]), p.round_on);
        o.w = rna_tf32(__uint_as_float(v[i + 3]), p.round_on);
        *reinterpret_cast<uint4*>(dst + i) = o;
      }
    }
  } else if (valid_row) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kBwdTmemCols);
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_m,
                    const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_dom,
                    const __grid_constant__ AttnBwdDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytes;
  uint8_t* sQ = smem + 2 * kTileBytes;
  uint8_t* sdO = smem + 3 * kTileBytes;
  uint8_t* sQm = smem + 4 * kTileBytes;
  uint8_t* sdOm = smem + 5 * kTileBytes;
  float* s_lse = reinterpret_cast<float*>(smem + 6 * kTileBytes);   // [128] lse * log2(e) of the chunk's query rows
  float* s_D = s_lse + kAttnRows;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_D + kAttnRows);
  uint64_t* bar_kvt = bars;       // K, V tiles (once)
  uint64_t* bar_qd = bars + 1;    // Q, dO chunk, K-major
  uint64_t* bar_m = bars + 2;     // Q, dO chunk, MN-major
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int b = blockIdx.y, k0 = blockIdx.x * kAttnKeys;
  const int T = p.T;
  const int lim = p.lens ? (p.lens[b] < T ? p.lens[b] : T) : T;
  const bool live = k0 < lim;                       // a key tile past the utterance gets no gradient at all
  const int n_chunks = live ? (T + kAttnRows - 1) / kAttnRows : 0;

  if (warp == 0 && ptx::elect_one()) {
    ptx::tma_prefetch_desc(&map_qk);
    ptx::tma_prefetch_desc(&map_m);
    ptx::tma_prefetch_desc(&map_do);
    ptx::tma_prefetch_desc(&map_dom);
    for (int i = 0; i < 5; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, kBwdTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
  const uint32_t tS = tmem_base + lane_base;            // [0, 128)   S^T, then P_d^T
  const uint32_t tdP = tmem_base + lane_base + 128;     // [128, 256) dP^T, then dS^T
  const uint32_t tdV = tmem_base + lane_base + 256;     // [256, 320)
  const uint32_t tdK = tmem_base + lane_base + 320;     // [320, 384)

  auto issue_qd = [&](int c) {
    ptx::mbar_arrive_expect_tx(bar_qd, 2 * kTileBytes);
    ptx::tma_load_3d(sQ, &map_qk, bar_qd, 0, c * kAttnRows, b);
    ptx::tma_load_3d(sQ + kTileBytes / 2, &map_qk, bar_qd, 32, c * kAttnRows, b);
    ptx::tma_load_3d(sdO, &map_do, bar_qd, 0, c * kAttnRows, b);
    ptx::tma_load_3d(sdO + kTileBytes / 2, &map_do, bar_qd, 32, c * kAttnRows, b);
  };
  auto issue_m = [&](int c) {
    ptx::mbar_arrive_expect_tx(bar_m, 2 * kTileBytes);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      ptx::tma_load_4d(sQm + kb * (kTileBytes / 4), &map_m, bar_m, 0, c * kAttnRows + kb * 32, 0, b);
      ptx::tma_load_4d(sdOm + kb * (kTileBytes / 4), &map_dom, bar_m, 0, c * kAttnRows + kb * 32, 0, b);
    }
  };

  if (n_chunks > 0 && warp == 0 && ptx::elect_one()) {
    ptx::mbar_arrive_expect_tx(bar_kvt, 2 * kTileBytes);
    ptx::tma_load_3d(sK, &map_qk, bar_kvt, kAttnD, k0, b);
    ptx::tma_load_3d(sK + kTileBytes / 2, &map_qk, bar_kvt, kAttnD + 32, k0, b);
    ptx::tma_load_3d(sV, &map_qk, bar_kvt, 2 * kAttnD, k0, b);
    ptx::tma_load_3d(sV + kTileBytes / 2, &map_qk, bar_kvt, 2 * kAttnD + 32, k0, b);
    issue_qd(0);
    issue_m(0);
  }
  __syncwarp();

  const uint32_t idesc_s = instr_desc(kAttnRows, 0);
  const uint32_t idesc_o = instr_desc(kAttnD, 1);
  const uint64_t dk0 = smem_desc(ptx::smem_u32(sK), 16, 1024, 2);
  const uint64_t dv0 = smem_desc(ptx::smem_u32(sV), 16, 1024, 2);
  const uint64_t dq0 = smem_desc(ptx::smem_u32(sQ), 16, 1024, 2);
  const uint64_t ddo0 = smem_desc(ptx::smem_u32(sdO), 16, 1024, 2);
  const uint64_t dqm0 = smem_desc(ptx::smem_u32(sQm), 32 * 128, 512, 1);
  const uint64_t ddom0 = smem_desc(ptx::smem_u32(sdOm), 32 * 128, 512, 1);
  const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) * 0xA24BAED4963EE407ull : 0ull);
  const int key = k0 + quarter * 32 + lane;        // this thread's key (one TMEM lane)
  const bool key_on = key < lim;
  const float c2 = p.scale_log2, sc = p.scale;
  const uint32_t t16 = p.drop_thresh >> 16;
  const int field = 16 * (key & 3);
  constexpr uint32_t kb_step = kTileBytes / 2 / 16;

  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t ph = static_cast<uint32_t>(c & 1);
    const int r0 = c * kAttnRows;
    // row statistics of the chunk's queries (rows past T: lse = +inf -> P = 0, D = 0)
    if (tid < kAttnRows) {
      const int r = r0 + tid;
      s_lse[tid] = r < T ? __ldg(p.lse + static_cast<long>(b) * T + r) * 1.4426950408889634f : INFINITY;
      s_D[tid] = r < T ? __ldg(p.dsum + static_cast<long>(b) * T + r) : 0.0f;
    }
    if (warp == 0) {
      if (c == 0) ptx::mbar_wait(bar_kvt, 0);
      ptx::mbar_wait(bar_qd, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            ptx::mma_tf32(tmem_base, dk0 + kb * kb_step + k4 * 2, dq0 + kb * kb_step + k4 * 2, idesc_s, (kb | k4) ? 1u : 0u);
            ptx::mma_tf32(tmem_base + 128, dv0 + kb * kb_step + k4 * 2, ddo0 + kb * kb_step + k4 * 2, idesc_s, (kb | k4) ? 1u : 0u);
          }
        ptx::mma_commit(bar_s);
      }
      __syncwarp();
    }
    __syncthreads();          // s_lse / s_D visible
    ptx::mbar_wait(bar_s, ph);
    ptx::tc_fence_after();
    if (warp == 0 && c + 1 < n_chunks && ptx::elect_one()) issue_qd(c + 1);
    __syncwarp();

    {
      const int col0 = half * 32;                       // query columns of this warp (`half` = column part 0..3, 16 warps)
      uint32_t s[32], dp[32];
      ptx::tmem_ld32(tS + col0, s);
      ptx::tmem_ld32(tdP + col0, dp);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        // dropout: one 64-bit hash serves keys 4g .. 4g+3 of a query row; the four lanes of a key group each hash one
        // of the four query rows of this step and exchange the results
        uint64_t h_mine = 0;
        if (p.drop_thresh) {
          const int rq = r0 + col0 + 4 * i4 + (lane & 3);
          const uint64_t idx = (static_cast<uint64_t>(b) * T + rq) * static_cast<uint64_t>(p.drop_ld) + key;
          h_mine = hash_u64(seed, idx >> 2);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * i4 + e;
          float dsc = p.inv_keep;
          if (p.drop_thresh) {
            const uint64_t h = __shfl_sync(0xffffffffu, h_mine, (lane & ~3) + e);
            dsc = (static_cast<uint32_t>(h >> field) & 0xFFFFu) >= t16 ? p.inv_keep : 0.0f;
          }
          const float pv = key_on ? ptx::ex2_approx(__uint_as_float(s[i]) * c2 - s_lse[col0 + i]) : 0.0f;
          const float g = sc * pv * (__uint_as_float(dp[i]) * dsc - s_D[col0 + i]);
          s[i] = rna_tf32(pv * dsc, p.round_on);
          dp[i] = rna_tf32(g, p.round_on);
        }
      }
      ptx::tmem_st32(tS + col0, s);
      ptx::tmem_st32(tdP + col0, dp);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();

    if (warp == 0) {
      ptx::mbar_wait(bar_m, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t acc = (c > 0 || (kb | k4)) ? 1u : 0u;
            ptx::mma_tf32_ts(tmem_base + 256, tmem_base + kb * 32 + k4 * 8, ddom0 + kb * (kTileBytes / 4 / 16) + k4 * (1024 / 16),
                             idesc_o, acc);
            ptx::mma_tf32_ts(tmem_base + 320, tmem_base + 128 + kb * 32 + k4 * 8, dqm0 + kb * (kTileBytes / 4 / 16) + k4 * (1024 / 16),
                             idesc_o, acc);
          }
        ptx::mma_commit(bar_o);
      }
      __syncwarp();
    }
    ptx::mbar_wait(bar_o, ph);
    ptx::tc_fence_after();
    if (warp == 0 && c + 1 < n_chunks && ptx::elect_one()) issue_m(c + 1);
    __syncwarp();
  }

  // ---- dK -> dqkv[:, :, 64:128], dV -> dqkv[:, :, 128:192]. The tensor-memory loads are .sync.aligned: every lane of the
  // warp executes them (n_chunks is uniform over the CTA); only the stores depend on the lane's key being a real row.
  {
    const bool store = key < T;
    float* dst = p.dqkv + static_cast<long>(b) * p.g_zs + static_cast<long>(store ? key : 0) * p.g_rs;
    if (n_chunks > 0) {
      // 16 warps, 2 x 64 output columns per row: column parts 0, 1 store dK, parts 2, 3 store dV, 32 columns each
      {
        const int which = half >> 1, hc = half & 1;
        uint32_t v[32];
        ptx::tmem_ld32((which == 0 ? tdK : tdV) + hc * 32, v);
        ptx::tmem_wait_ld();
        if (store) {
          float* d = dst + (which + 1) * kAttnD + hc * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            uint4 o;
            o.x = rna_tf32(__uint_as_float(v[i]), p.round_on);
            o.y = rna_tf32(__uint_as_float(v[i + 1]), p.round_on);
            o.z = rna_tf32(__uint_as_float(v[i + 2]), p.round_on);
            o.w = rna_tf32(__uint_as_float(v[i + 3]), p.round_on);
            *reinterpret_cast<uint4*>(d + i) = o;
          }
        }
      }
    } else if (store) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + kAttnD + half * 32 + i) = make_float4(0.f, 0.f, 0.f, 0.f);   // 4 parts x 32 = dK | dV
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kBwdTmemCols);
}

}  // namespace

// qkv [B, T, 192] (columns: q 0..63 | k 64..127 | v 128..191, row stride rs, item stride zs, fp32 holding tf32-rounded
// values) -> out [B, T, 64] = dropout(softmax(q.k^T * scale + key mask)) . v, lse [B * T] (natural log of the row sum of
// exp(scaled scores)); lens [B] int32 or null (keys >= lens[b] are masked).
int attn_fused_fwd(const float* qkv, long rs, long zs, int B, int T, const int* lens, float scale, float drop_p,
                   uint64_t seed, const uint64_t* seed_dev, int drop_ld, float* out, long o_rs, long o_zs, float* lse,
                   cudaStream_t stream) {
  XVA_CHECK_ARG(qkv && out && B >= 1 && T >= 1, "attn_fwd: null / empty argument (B=%d T=%d)", B, T);
  XVA_CHECK_ARG(rs >= 3 * kAttnD && (rs % 4) == 0 && (zs % 4) == 0, "attn_fwd: qkv row stride %ld / item stride %ld", rs, zs);
  XVA_CHECK_ARG((o_rs % 4) == 0 && (o_zs % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "attn_fwd: output not 16-byte aligned / strides not a multiple of 4");
  XVA_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "attn_fwd: dropout p=%f", drop_p);
  CUtensorMap map_qk, map_v;
  int rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3 * kAttnD), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t str[3] = {1, static_cast<uint64_t>(rs), static_cast<uint64_t>(B > 1 ? zs : rs * T)};
    uint32_t box[3] = {32, kAttnRows, 1};
    if ((rc = tma_encode_f32(&map_qk, qkv, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  }
  {
    uint64_t dims[4] = {32, static_cast<uint64_t>(T), static_cast<uint64_t>(3 * kAttnD / 32), static_cast<uint64_t>(B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(rs), 32, static_cast<uint64_t>(B > 1 ? zs : rs * T)};
    uint32_t box[4] = {32, 32, kAttnD / 32, 1};
    if ((rc = tma_encode_f32(&map_v, qkv, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != XVA_OK) return rc;
  }
  AttnDev p{};
  p.B = B;
  p.T = T;
  p.drop_ld = drop_ld > 0 ? drop_ld : T;
  p.lens = lens;
  p.out = out;
  p.o_rs = o_rs;
  p.o_zs = o_zs;
  p.lse = lse;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (drop_p > 0.0f) {
    p.drop_thresh = static_cast<uint32_t>(static_cast<double>(drop_p) * 4294967296.0);
    p.inv_keep = 1.0f / (1.0f - drop_p);
  } else {
    p.drop_thresh = 0;
    p.inv_keep = 1.0f;
  }
  p.seed = seed;
  p.seed_dev = seed_dev;
  p.round_on = g_attn_round_host;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
  });
  XVA_CHECK_CUDA(attr_err);
  dim3 grid(ceil_div(T, kAttnRows), B);
  attn_fwd_kernel<<<grid, kAttnThreads, kFwdSmem, stream>>>(map_qk, map_v, p);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

// Backward of attn_fused_fwd. dout [B, T, 64] = gradient of the attention output (tf32-rounded: it is a GEMM operand),
// lse from the forward, dsum [B * T] = rowwise dot(dout, out) (xva_rowdot2). dqkv [B, T, >= 192] receives dq | dk | dv in
// the column layout of qkv, tf32-rounded (they are the operands of the qkv projection's weight and input gradients).
int attn_fused_bwd(const float* qkv, long rs, long zs, const float* dout, long d_rs, long d_zs, const float* lse,
                   const float* dsum, int B, int T, const int* lens, float scale, float drop_p, uint64_t seed,
                   const uint64_t* seed_dev, int drop_ld, float* dqkv, long g_rs, long g_zs, cudaStream_t stream) {
  XVA_CHECK_ARG(qkv && dout && lse && dsum && dqkv && B >= 1 && T >= 1, "attn_bwd: null / empty argument (B=%d T=%d)", B, T);
  XVA_CHECK_ARG(rs >= 3 * kAttnD && (rs % 4) == 0 && (zs % 4) == 0 && d_rs >= kAttnD && (d_rs % 4) == 0 && (d_zs % 4) == 0,
                "attn_bwd: operand strides");
  XVA_CHECK_ARG((g_rs % 4) == 0 && (g_zs % 4) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                "attn_bwd: dqkv not 16-byte aligned / strides not a multiple of 4");
  XVA_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "attn_bwd: dropout p=%f", drop_p);
  CUtensorMap map_qk, map_m, map_do, map_dom;
  int rc;
  const uint64_t zq = static_cast<uint64_t>(B > 1 ? zs : rs * T), zd = static_cast<uint64_t>(B > 1 ? d_zs : d_rs * T);
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3 * kAttnD), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t str[3] = {1, static_cast<uint64_t>(rs), zq};
    uint32_t box[3] = {32, kAttnRows, 1};
    if ((rc = tma_encode_f32(&map_qk, qkv, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  }
  {
    uint64_t dims[4] = {32, static_cast<uint64_t>(T), static_cast<uint64_t>(3 * kAttnD / 32), static_cast<uint64_t>(B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(rs), 32, zq};
    uint32_t box[4] = {32, 32, kAttnD / 32, 1};
    if ((rc = tma_encode_f32(&map_m, qkv, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != XVA_OK) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(kAttnD), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t str[3] = {1, static_cast<uint64_t>(d_rs), zd};
    uint32_t box[3] = {32, kAttnRows, 1};
    if ((rc = tma_encode_f32(&map_do, dout, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  }
  {
    uint64_t dims[4] = {32, static_cast<uint64_t>(T), static_cast<uint64_t>(kAttnD / 32), static_cast<uint64_t>(B)};
    uint64_t str[4] = {1, static_cast<uint64_t>(d_rs), 32, zd};
    uint32_t box[4] = {32, 32, kAttnD / 32, 1};
    if ((rc = tma_encode_f32(&map_dom, dout, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != XVA_OK) return rc;
  }
  AttnBwdDev p{};
  p.B = B;
  p.T = T;
  p.drop_ld = drop_ld > 0 ? drop_ld : T;
  p.lens = lens;
  p.lse = lse;
  p.dsum = dsum;
  p.dqkv = dqkv;
  p.g_rs = g_rs;
  p.g_zs = g_zs;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (drop_p > 0.0f) {
    p.drop_thresh = static_cast<uint32_t>(static_cast<double>(drop_p) * 4294967296.0);
    p.inv_keep = 1.0f / (1.0f - drop_p);
  } else {
    p.drop_thresh = 0;
    p.inv_keep = 1.0f;
  }
  p.seed = seed;
  p.seed_dev = seed_dev;
  p.round_on = g_attn_round_host;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem);
  });
  XVA_CHECK_CUDA(attr_err);
  dim3 grid(ceil_div(T, kAttnRows), B);
  attn_bwd_dq_kernel<<<grid, kBwdThreads, kDqSmem, stream>>>(map_qk, map_m, map_do, p);
  XVA_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<<<grid, kBwdThreads, kDkvSmem, stream>>>(map_qk, map_m, map_do, map_dom, p);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int set_operand_rounding_attn_fused(int on) {
  g_attn_round_host = on;
  return XVA_OK;
}

}  // namespace xva
