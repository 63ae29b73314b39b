// tcgen05 / TMA tap-GEMM for sm_100a (see gemm.cuh for the contract).
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer,
// warp 2 = TMEM allocator, warps 4..7 = epilogue (one thread per accumulator row / TMEM lane).
// Operands are fp32 in HBM, staged by TMA into 128B-swizzled shared-memory tiles and multiplied as
// kind::tf32 (fp32 accumulate in TMEM). Accumulators are double-buffered in TMEM (2 x 256 columns) so the
// epilogue of tile i overlaps the main loop of tile i+1; the LayerNorm epilogue owns up to 512 columns.
//
// The k-tap convolution is a sum of `taps` GEMMs whose A tiles are the same activation rows shifted by
// shift[j]; the shift is a TMA coordinate, the zero padding is TMA out-of-bounds fill: no im2col, no halo copy.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include "gemm.cuh"
#include "ptx.cuh"

namespace xva {

namespace {

constexpr int kBlockM = 128;        // accumulator rows per tile (= TMEM lanes)
constexpr int kBlockK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;           // tf32 MMA K
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kThreads = 256;
constexpr int kATileBytes = kBlockM * kBlockK * 4;  // 16 KiB
constexpr int kSmemBudget = 220 * 1024;

struct GemmDev {
  int mode, Z, R, M, N, K, taps, ZR, split, zper;
  int n_tile, n_sub, n_mma, tiles_n, tiles_m, k_chunks, stages, acc_stages, num_tiles;
  int b_tap_z, b_batch_z;
  uint32_t mn_layout, mn_lbo, mn_sbo;  // MN-major descriptor constants (debug-overridable, see gemm_tc_launch)
  int shift[kMaxTaps];
  float* out;
  long o_rs, o_zs, o_js;
  float alpha;
  int flags;
  const float* bias;
  const float* residual;
  long r_rs, r_zs;
  const float* gate;
  long g_rs, g_zs;
  float gate_slope;
  const int* lens;
  const float* gamma;
  const float* beta;
  float ln_eps;
  float* out_pre;
  float* ln_mean;
  float* ln_rstd;
  uint32_t drop_thresh;
  float inv_keep;
  uint64_t seed;
  const uint64_t* seed_dev;
};

struct TileCoord {
  int z;      // mode 0/1: batch item; mode 2: first z of the reduced range
  int z_end;  // mode 2: one past the last z of the reduced range
  int zo;     // output batch index
  int j;      // mode 2: tap of this output tile
  int m0;     // first output row of the tile
  int n0;     // first output column of the tile
  int iters;  // k-iterations of the main loop
};

__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int t) {
  TileCoord c;
  int n_t = t % p.tiles_n;
  t /= p.tiles_n;
  int m_t = t % p.tiles_m;
  t /= p.tiles_m;
  c.n0 = n_t * p.n_tile;
  c.m0 = m_t * kBlockM;
  if (p.mode != 2) {
    c.z = t;
    c.z_end = t + 1;
    c.zo = t;
    c.j = 0;
    c.iters = p.taps * p.k_chunks;
  } else {
    c.j = t % p.taps;
    t /= p.taps;
    int s = t % p.split;
    c.zo = t / p.split;
    c.z = c.zo * p.ZR + s * p.zper;
    int zend = c.z + p.zper;
    int zlim = (c.zo + 1) * p.ZR;
    c.z_end = zend < zlim ? zend : zlim;
    c.iters = (c.z_end - c.z) * p.k_chunks;
  }
  return c;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout), version 1.
//   K-major : SWIZZLE_128B (layout type 2); 8-row groups of 128B rows, SBO = 1024 B between groups, LBO unused (1).
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (layout type 1; TMA 128B_ATOM_32B): atoms of
//             4 k-rows x 128 B whose 32-byte chunks are XOR-ed with (row & 3); LBO = byte distance between
//             32-element MN chunks, SBO = 512 B between 4-row k groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): tf32 x tf32 -> f32, M = 128, runtime N.
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // c_format = F32
  d |= 2u << 7;                                   // a_format = TF32
  d |= 2u << 10;                                  // b_format = TF32
  d |= static_cast<uint32_t>(a_mn_major) << 15;   // a_major
  d |= static_cast<uint32_t>(b_mn_major) << 16;   // b_major
  d |= static_cast<uint32_t>(n >> 3) << 17;       // n_dim
  d |= static_cast<uint32_t>(kBlockM >> 4) << 24; // m_dim
  return d;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  // 1024-byte alignment of the tile ring (SWIZZLE_128B atoms are 1024 B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_tile_bytes = p.n_tile * kBlockK * 4;
  const int stage_bytes = kATileBytes + b_tile_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap_a);
    ptx::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bar_full[s], 1);
      ptx::mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bar_tmem_full[a], 1);
      ptx::mbar_init(&bar_tmem_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&tmem_base_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const bool a_mn = (p.mode == 2);
  const bool b_mn = (p.mode != 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord c = decode_tile(p, tile);
        for (int it = 0; it < c.iters; ++it) {
          ptx::mbar_wait(&bar_empty[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + kATileBytes;
          ptx::mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
          if (p.mode != 2) {
            const int j = it / p.k_chunks;
            const int kc = it - j * p.k_chunks;
            const int zb = j * p.b_tap_z + c.z * p.b_batch_z;
            ptx::tma_load_3d(sa, &tmap_a, &bar_full[s], kc * kBlockK, c.m0 + p.shift[j], c.z);
            if (p.mode == 0) {
              for (int sub = 0; sub < p.n_mma; ++sub)
                ptx::tma_load_3d(sb + sub * p.n_sub * kBlockK * 4, &tmap_b, &bar_full[s], kc * kBlockK,
                                 c.n0 + sub * p.n_sub, zb);
            } else {
              ptx::tma_load_4d(sb, &tmap_b, &bar_full[s], 0, kc * kBlockK, c.n0 / 32, zb);
            }
          } else {
            const int zi = it / p.k_chunks;
            const int tc = it - zi * p.k_chunks;
            const int z = c.z + zi;
            ptx::tma_load_4d(sa, &tmap_a, &bar_full[s], 0, tc * kBlockK, c.m0 / 32, z);
            ptx::tma_load_4d(sb, &tmap_b, &bar_full[s], 0, tc * kBlockK + p.shift[c.j], c.n0 / 32, z);
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.n_sub, a_mn ? 1 : 0, b_mn ? 1 : 0);
      const uint32_t a_lbo = a_mn ? p.mn_lbo : 16;
      const uint32_t b_lbo = b_mn ? p.mn_lbo : 16;
      const uint32_t a_kstep = a_mn ? 1024 : kUmmaK * 4;  // bytes to advance per UMMA_K
      const uint32_t b_kstep = b_mn ? 1024 : kUmmaK * 4;
      const uint32_t b_sub_bytes = b_mn ? (p.n_sub / 32) * (kBlockK * 128) : p.n_sub * kBlockK * 4;
      int s = 0;
      uint32_t ph = 0;
      int tile_iter = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const TileCoord c = decode_tile(p, tile);
        const int acc = tile_iter % p.acc_stages;
        const uint32_t acc_ph = (tile_iter / p.acc_stages) & 1;
        ptx::mbar_wait(&bar_tmem_empty[acc], acc_ph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * 256;
        for (int it = 0; it < c.iters; ++it) {
          ptx::mbar_wait(&bar_full[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
#pragma unroll
          for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
            const uint64_t da = make_smem_desc(sa + k4 * a_kstep, a_lbo, a_mn ? p.mn_sbo : 1024, a_mn ? p.mn_layout : 2);
            for (int sub = 0; sub < p.n_mma; ++sub) {
              const uint64_t db = make_smem_desc(sb + sub * b_sub_bytes + k4 * b_kstep, b_lbo, b_mn ? p.mn_sbo : 1024, b_mn ? p.mn_layout : 2);
              ptx::mma_tf32(tmem_acc + sub * p.n_sub, da, db, idesc, (it > 0 || k4 > 0) ? 1u : 0u);
            }
          }
          ptx::mma_commit(&bar_empty[s]);  // frees the smem stage once these MMAs have read it
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        ptx::mma_commit(&bar_tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (128 threads <-> 128 TMEM lanes)
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    int tile_iter = 0;
    const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) * 0xA24BAED4963EE407ull : 0ull);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
      const TileCoord c = decode_tile(p, tile);
      const int acc = tile_iter % p.acc_stages;
      const uint32_t acc_ph = (tile_iter / p.acc_stages) & 1;
      ptx::mbar_wait(&bar_tmem_full[acc], acc_ph);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + acc * 256 + lane_base;

      const int row = c.m0 + row_in_tile;
      const int row_limit = (p.mode == 2) ? p.M : p.R;
      const bool row_ok = row < row_limit;
      const int n_cols = (p.N - c.n0) < p.n_tile ? (p.N - c.n0) : p.n_tile;  // valid columns of this tile
      const int n_chunks = (n_cols + 15) / 16;

      if (p.mode == 2) {
        float* orow = p.out + c.zo * p.o_zs + c.j * p.o_js + static_cast<long>(row) * p.o_rs + c.n0;
        for (int ch = 0; ch < n_chunks; ++ch) {
          uint32_t v[16];
          ptx::tmem_ld16(tacc + ch * 16, v);
          ptx::tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = ch * 16 + i;
              if (n < n_cols) {
                const float val = p.alpha * __uint_as_float(v[i]);
                if (p.flags & GEMM_ATOMIC) atomicAdd(orow + n, val);
                else orow[n] = val;
              }
            }
          }
        }
      } else {
        const long rowoff_o = c.z * p.o_zs + static_cast<long>(row) * p.o_rs + c.n0;
        const long rowoff_r = c.z * p.r_zs + static_cast<long>(row) * p.r_rs + c.n0;
        const long rowoff_g = c.z * p.g_zs + static_cast<long>(row) * p.g_rs + c.n0;
        const float keep_row = (p.lens == nullptr || row < p.lens[c.z]) ? 1.0f : 0.0f;
        const uint64_t drop_row = (static_cast<uint64_t>(c.z) * p.R + row) * static_cast<uint64_t>(p.N) + c.n0;
        const bool vec_ok = ((p.o_rs & 3) == 0) && ((c.n0 & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

        // v = alpha*acc + bias -> relu -> gate -> dropout(pre) -> + residual      (column n of this row)
        auto pre_value = [&](float accv, int n) -> float {
          float v = p.alpha * accv;
          if (p.bias) v += __ldg(p.bias + c.n0 + n);
          if (p.flags & GEMM_RELU) v = fmaxf(v, 0.0f);
          if (p.gate) {
            const float gv = __ldg(p.gate + rowoff_g + n);
            v *= (gv > 0.0f) ? 1.0f : p.gate_slope;
          }
          if (p.flags & GEMM_DROP_PRE) v *= dropout_scale(seed, drop_row + n, p.drop_thresh, p.inv_keep);
          if (p.residual) v += __ldg(p.residual + rowoff_r + n);
          return v;
        };

        if (!(p.flags & GEMM_LN)) {
          for (int ch = 0; ch < n_chunks; ++ch) {
            uint32_t v[16];
            ptx::tmem_ld16(tacc + ch * 16, v);
            ptx::tmem_wait_ld();
            if (row_ok) {
              float o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = ch * 16 + i;
                o[i] = (n < n_cols) ? pre_value(__uint_as_float(v[i]), n) * keep_row : 0.0f;
              }
              float* dst = p.out + rowoff_o + ch * 16;
              if (vec_ok && (ch * 16 + 16 <= n_cols)) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (ch * 16 + i < n_cols) dst[i] = o[i];
              }
            }
          }
        } else {
          // LayerNorm over the n_cols (= N) columns held by this thread's TMEM lane. Three TMEM passes:
          // (1) finish the pre-LN value in place and sum it, (2) centred second moment, (3) normalise + store.
          float sum = 0.0f;
          for (int ch = 0; ch < n_chunks; ++ch) {
            uint32_t v[16];
            ptx::tmem_ld16(tacc + ch * 16, v);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float x = row_ok ? pre_value(__uint_as_float(v[i]), ch * 16 + i) : 0.0f;
              sum += x;
              v[i] = __float_as_uint(x);
            }
            ptx::tmem_st16(tacc + ch * 16, v);
          }
          ptx::tmem_wait_st();
          const float mean = sum / static_cast<float>(n_cols);
          float sq = 0.0f;
          for (int ch = 0; ch < n_chunks; ++ch) {
            uint32_t v[16];
            ptx::tmem_ld16(tacc + ch * 16, v);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float d = __uint_as_float(v[i]) - mean;
              sq += d * d;
            }
          }
          const float rstd = rsqrtf(sq / static_cast<float>(n_cols) + p.ln_eps);
          if (row_ok) {
            const long srow = static_cast<long>(c.z) * p.R + row;
            if (p.ln_mean) p.ln_mean[srow] = mean;
            if (p.ln_rstd) p.ln_rstd[srow] = rstd;
          }
          for (int ch = 0; ch < n_chunks; ++ch) {
            uint32_t v[16];
            ptx::tmem_ld16(tacc + ch * 16, v);
            ptx::tmem_wait_ld();
            if (row_ok) {
              float o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = ch * 16 + i;
                const float x = __uint_as_float(v[i]);
                float y = (x - mean) * rstd * __ldg(p.gamma + n) + __ldg(p.beta + n);
                if (p.flags & GEMM_DROP_POST) y *= dropout_scale(seed, drop_row + n, p.drop_thresh, p.inv_keep);
                o[i] = y * keep_row;
              }
              float* dst = p.out + rowoff_o + ch * 16;
              float* dpre = p.out_pre ? p.out_pre + rowoff_o + ch * 16 : nullptr;
              if (vec_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                if (dpre) {
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(dpre + 4 * q) =
                        make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                    __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  dst[i] = o[i];
                  if (dpre) dpre[i] = __uint_as_float(v[i]);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// fp32 tensor map, zero out-of-bounds fill; 128B swizzle (K-major tiles) or 128B swizzle with 32B atoms (MN-major tiles). dims/strides in elements, innermost first;
// strides[0] is implicit (1).
int encode_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
               const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return XVA_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t gbox[5], estride[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estride[i] = 1;
    if (i > 0) {
      gstride[i - 1] = strides_elems[i] * sizeof(float);
      if (gstride[i - 1] % 16 != 0) {
        set_error("tensor-map stride %llu B of dim %d is not a multiple of 16", (unsigned long long)gstride[i - 1], i);
        return XVA_ERR_ARG;
      }
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) {
    set_error("tensor-map base address not 16-byte aligned");
    return XVA_ERR_ARG;
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), gdim, gstride, gbox, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  static_cast<CUtensorMapSwizzle>(swizzle),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)(rank > 2 ? gdim[2] : 0),
              gbox[0], gbox[1], rank > 2 ? gbox[2] : 0);
    return XVA_ERR_CUDA;
  }
  return XVA_OK;
}

}  // namespace

int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream) {
  XVA_CHECK_ARG(g.mode >= 0 && g.mode <= 2, "gemm: bad mode %d", g.mode);
  XVA_CHECK_ARG(g.taps >= 1 && g.taps <= kMaxTaps, "gemm: taps %d out of range", g.taps);
  XVA_CHECK_ARG(g.Z >= 1 && g.R >= 1 && g.N >= 1, "gemm: empty problem Z=%d R=%d N=%d", g.Z, g.R, g.N);
  XVA_CHECK_ARG(g.a && g.b && g.out, "gemm: null operand");

  GemmDev p{};
  p.mode = g.mode;
  p.Z = g.Z;
  p.R = g.R;
  p.M = g.M;
  p.N = g.N;
  p.K = g.K;
  p.taps = g.taps;
  for (int j = 0; j < g.taps; ++j) p.shift[j] = g.shift[j];
  p.b_tap_z = g.b_tap_z;
  p.b_batch_z = g.b_batch_z;

  // ---- tile shape along N
  const bool ln = (g.flags & GEMM_LN) != 0;
  const int n_gran = (g.mode == 0) ? 16 : 32;  // MN-major B tiles are made of 32-element chunks
  if (ln) {
    XVA_CHECK_ARG(g.mode != 2, "gemm: LayerNorm epilogue not available in wgrad mode");
    XVA_CHECK_ARG(g.N % 16 == 0 && g.N <= 512, "gemm: LayerNorm epilogue needs N %% 16 == 0 and N <= 512 (N=%d)", g.N);
    XVA_CHECK_ARG(g.gamma && g.beta, "gemm: LayerNorm epilogue needs gamma/beta");
    p.n_tile = round_up(g.N, n_gran);
    p.tiles_n = 1;
  } else {
    p.tiles_n = ceil_div(g.N, 256);
    p.n_tile = round_up(ceil_div(g.N, p.tiles_n), n_gran);
  }
  p.n_mma = ceil_div(p.n_tile, 256);
  p.n_sub = p.n_tile / p.n_mma;
  if (p.n_sub % n_gran != 0) {  // e.g. 400 -> 2 x 200: round the tile up so both halves are legal
    p.n_sub = round_up(p.n_sub, n_gran);
    p.n_tile = p.n_sub * p.n_mma;
  }
  XVA_CHECK_ARG(p.n_tile <= 512 && p.n_sub <= 256 && p.n_sub % 16 == 0, "gemm: bad n tiling %d/%d", p.n_tile, p.n_sub);
  p.acc_stages = (p.n_tile <= 256) ? 2 : 1;

  const int stage_bytes = kATileBytes + p.n_tile * kBlockK * 4;
  p.stages = kSmemBudget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  XVA_CHECK_ARG(p.stages >= 2, "gemm: tile too large for shared memory");
  const int smem_bytes = p.stages * stage_bytes + 1024;

  // ---- tiling along M and the k loop
  if (g.mode != 2) {
    XVA_CHECK_ARG(g.K >= 1, "gemm: K=%d", g.K);
    // K itself may be ragged (the last 32-wide k-block is zero-filled by TMA); only row strides need 16-byte
    // alignment, which encode_map checks.
    p.tiles_m = ceil_div(g.R, kBlockM);
    p.k_chunks = ceil_div(g.K, kBlockK);
    p.ZR = 1;
    p.split = 1;
    p.zper = 1;
    p.num_tiles = g.Z * p.tiles_m * p.tiles_n;
  } else {
    XVA_CHECK_ARG(g.M >= 1, "gemm: wgrad M=%d", g.M);
    XVA_CHECK_ARG(g.M % 32 == 0 || g.a_rs >= round_up(g.M, 32),
                  "gemm: MN-major A with M=%d needs M %% 32 == 0 or a row stride >= %d (got %lld)", g.M,
                  round_up(g.M, 32), (long long)g.a_rs);
    XVA_CHECK_ARG(g.N % 32 == 0 || g.b_rs >= round_up(g.N, 32),
                  "gemm: MN-major B with N=%d needs N %% 32 == 0 or a row stride >= %d (got %lld)", g.N,
                  round_up(g.N, 32), (long long)g.b_rs);
    XVA_CHECK_ARG(g.ZR >= 1 && g.Z % g.ZR == 0, "gemm: Z=%d not divisible by ZR=%d", g.Z, g.ZR);
    p.tiles_m = ceil_div(g.M, kBlockM);
    p.k_chunks = ceil_div(g.R, kBlockK);
    p.ZR = g.ZR;
    int split = g.split < 1 ? 1 : g.split;
    if (split > g.ZR) split = g.ZR;
    p.zper = ceil_div(g.ZR, split);
    p.split = ceil_div(g.ZR, p.zper);
    XVA_CHECK_ARG(p.split == 1 || (g.flags & GEMM_ATOMIC), "gemm: split > 1 needs GEMM_ATOMIC");
    p.num_tiles = (g.Z / g.ZR) * p.split * g.taps * p.tiles_m * p.tiles_n;
  }

  // MN-major tiles: SWIZZLE_128B_BASE32B descriptors + TMA 128B_ATOM_32B. XVA_MN_DEBUG=layout:lbo:sbo:tma_swizzle
  // overrides the four constants (bring-up aid for tests/gpu_mn_probe.py; unset in production).
  p.mn_layout = 1;
  p.mn_lbo = kBlockK * 128;
  p.mn_sbo = 512;
  int mn_swizzle = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (const char* dbg = getenv("XVA_MN_DEBUG")) {
    unsigned a, b, c;
    int d;
    if (sscanf(dbg, "%u:%u:%u:%d", &a, &b, &c, &d) == 4) {
      p.mn_layout = a;
      p.mn_lbo = b;
      p.mn_sbo = c;
      mn_swizzle = d;
    }
  }

  // ---- tensor maps
  CUtensorMap map_a, map_b;
  int rc;
  const int a_rows = g.a_rows ? g.a_rows : g.R;
  if (g.mode != 2) {
    uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)a_rows, (uint64_t)g.Z};
    uint64_t str[3] = {1, (uint64_t)g.a_rs, (uint64_t)g.a_zs};
    uint32_t box[3] = {kBlockK, kBlockM, 1};
    if (g.Z == 1 || str[2] == 0) str[2] = (uint64_t)g.a_rs * a_rows;
    if ((rc = encode_map(&map_a, g.a, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  } else {
    uint64_t dims[4] = {32, (uint64_t)a_rows, (uint64_t)ceil_div(g.M, 32), (uint64_t)g.Z};
    uint64_t str[4] = {1, (uint64_t)g.a_rs, 32, (uint64_t)g.a_zs};
    uint32_t box[4] = {32, kBlockK, kBlockM / 32, 1};
    if (g.Z == 1 || str[3] == 0) str[3] = (uint64_t)g.a_rs * a_rows;
    if ((rc = encode_map(&map_a, g.a, 4, dims, str, box, mn_swizzle)) != XVA_OK) return rc;
  }
  if (g.mode == 0) {
    const int bn_rows = g.b_rows ? g.b_rows : g.N;
    uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)bn_rows, (uint64_t)g.b_nz};
    uint64_t str[3] = {1, (uint64_t)g.b_rs, (uint64_t)g.b_zs};
    uint32_t box[3] = {kBlockK, (uint32_t)p.n_sub, 1};
    if (g.b_nz == 1 || str[2] == 0) str[2] = (uint64_t)g.b_rs * bn_rows;
    if ((rc = encode_map(&map_b, g.b, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  } else if (g.mode == 1) {
    XVA_CHECK_ARG(g.N % 32 == 0 || g.b_rs >= round_up(g.N, 32),
                  "gemm: MN-major B with N=%d needs N %% 32 == 0 or a row stride >= %d (got %lld)", g.N,
                  round_up(g.N, 32), (long long)g.b_rs);
    const int bk_rows = g.b_rows ? g.b_rows : g.K;
    uint64_t dims[4] = {32, (uint64_t)bk_rows, (uint64_t)ceil_div(g.N, 32), (uint64_t)g.b_nz};
    uint64_t str[4] = {1, (uint64_t)g.b_rs, 32, (uint64_t)g.b_zs};
    uint32_t box[4] = {32, kBlockK, (uint32_t)(p.n_tile / 32), 1};
    if (g.b_nz == 1 || str[3] == 0) str[3] = (uint64_t)g.b_rs * bk_rows;
    if ((rc = encode_map(&map_b, g.b, 4, dims, str, box, mn_swizzle)) != XVA_OK) return rc;
  } else {
    const int b_rows = g.b_rows ? g.b_rows : g.R;
    uint64_t dims[4] = {32, (uint64_t)b_rows, (uint64_t)ceil_div(g.N, 32), (uint64_t)g.Z};
    uint64_t str[4] = {1, (uint64_t)g.b_rs, 32, (uint64_t)g.b_zs};
    uint32_t box[4] = {32, kBlockK, (uint32_t)(p.n_tile / 32), 1};
    if (g.Z == 1 || str[3] == 0) str[3] = (uint64_t)g.b_rs * b_rows;
    if ((rc = encode_map(&map_b, g.b, 4, dims, str, box, mn_swizzle)) != XVA_OK) return rc;
  }

  // ---- epilogue
  p.out = g.out;
  p.o_rs = g.o_rs;
  p.o_zs = g.o_zs;
  p.o_js = g.o_js;
  p.alpha = g.alpha;
  p.flags = g.flags;
  p.bias = g.bias;
  p.residual = g.residual;
  p.r_rs = g.r_rs;
  p.r_zs = g.r_zs;
  p.gate = g.gate;
  p.g_rs = g.g_rs;
  p.g_zs = g.g_zs;
  p.gate_slope = g.gate_slope;
  p.lens = g.lens;
  p.gamma = g.gamma;
  p.beta = g.beta;
  p.ln_eps = g.ln_eps;
  p.out_pre = g.out_pre;
  p.ln_mean = g.ln_mean;
  p.ln_rstd = g.ln_rstd;
  p.seed = g.seed;
  p.seed_dev = g.seed_dev;
  if ((g.flags & (GEMM_DROP_PRE | GEMM_DROP_POST)) && g.drop_p > 0.0f) {
    XVA_CHECK_ARG(g.drop_p < 1.0f, "gemm: dropout p=%f", g.drop_p);
    p.drop_thresh = static_cast<uint32_t>(static_cast<double>(g.drop_p) * 4294967296.0);
    p.inv_keep = 1.0f / (1.0f - g.drop_p);
  } else {
    p.flags &= ~(GEMM_DROP_PRE | GEMM_DROP_POST);
    p.drop_thresh = 0;
    p.inv_keep = 1.0f;
  }

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
  });
  XVA_CHECK_CUDA(attr_err);

  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  gemm_tc_kernel<<<grid, kThreads, smem_bytes, stream>>>(map_a, map_b, p);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

}  // namespace xva
