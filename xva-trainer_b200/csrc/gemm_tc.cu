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
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "gemm.cuh"
#include "ptx.cuh"

namespace xva {

static int g_round_host = 1;  // host mirror of g_xva_round_operands: travels to the kernel as GemmDev::round_on

namespace {

constexpr int kBlockM = 128;        // accumulator rows per tile (= TMEM lanes)
constexpr int kBlockK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;           // tf32 MMA K
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kThreads = 384;        // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
constexpr int kATileBytes = kBlockM * kBlockK * 4;  // 16 KiB
constexpr int kEpiSmemBytes = 8 * 4096;  // one 32x32 fp32 transpose tile per epilogue warp
constexpr int kSmemBudget = 192 * 1024;   // operand ring; + kEpiSmemBytes + 1 KiB alignment slack <= 227 KiB
constexpr int kLnSmemBytes = 2 * 4 * 32 * 2 * 4;
constexpr int kMaxStagesA = 4;       // activation ring of the halo mode
constexpr int kBarSmemBytes = (2 * kMaxStages + 4) * 8 + 8 + 2 * kMaxStagesA * 8;
constexpr int kAlignSlack = 768;
constexpr int kTailSmemBytes = kEpiSmemBytes + kLnSmemBytes + kBarSmemBytes + kAlignSlack;
enum { EPI_WGRAD = 0, EPI_PLAIN = 1, EPI_FULL = 2, EPI_LN = 3 };

struct GemmDev {
  int mode, Z, R, M, N, K, taps, ZR, split, zper;
  // Halo mode (mode 0/1, k-tap convolutions on un-segmented tiles): the activation tile of a k-block is loaded ONCE with
  // hp halo rows on either side and every tap reads it through a descriptor whose start address is advanced by
  // (hp + shift) rows (SWIZZLE_128B is a function of the absolute shared-memory address, so any row offset is legal with
  // base offset 0: scripts/rowshift_probe.cu). The weights keep their own ring, one stage per (tap, k-block).
  int halo, hp, a_rows, stages_a, ring_bytes;
  int tap_inner;  // mode 0/1 loop order of the operand stream: k-block outer, tap inner (see the producer)
  // mode 2, multi-tap tiles: one tile accumulates tp taps side by side in TMEM (tap jj at columns jj * n_tile), so dy is
  // fetched once per k-block for all of them and the tap-shifted windows of x hit L2 -- a small-channel weight gradient
  // (HiFi-GAN ResBlocks: 32 / 64 channels, 3..11 taps) was HBM-bound on re-reading both operands once per tap.
  int tp, tgroups;
  int rsplit, chunk_rows;  // mode 2: every item's R contraction rows are cut into rsplit chunks of chunk_rows (a multiple
                           // of 32); the reduction units (item, chunk) -- ZR of them per output -- are what `split` divides
  int n_tile, n_sub, n_mma, tiles_n, tiles_m, k_chunks, stages, acc_stages, num_tiles;
  int b_tap_z, b_batch_z;
  int row_tiles;  // Z * tiles_m: 128-row tiles of the output (mode 0/1)
  int groups;     // > 1: grouped convolution, the n tile (mode 0/1) or m tile (mode 2) index is the group
  int grp_a;      // mode 0/1: A column step per group; mode 2: B column step per group
  int grp_bk;     // mode 1: B row (contraction) step per group
  int gpt;        // grouped mode 2: groups per 128-row m tile (= 128 / Og); the tile computes the dense
                  // [gpt*Og, gpt*Cg] product of its groups and stores only the gpt diagonal [Og, Cg] blocks
  int og, cg;     // grouped mode 2: output rows / columns per group
  // Segmented row tiles (mode 0/1): a 128-row tile is made of 128/seg segments of `seg` consecutive rows, taken in
  // (item, segment) order, so short sequences (R = 160 tokens, or the 10..83-row period columns of DiscriminatorP)
  // share tiles instead of padding each item to a multiple of 128 rows. seg = 128: one segment = the classic tile.
  int seg, segs, tot_segs;  // rows per segment (32 / 64 / 128), segments per item, Z * segs
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
  float act_slope;
  float out_act_slope;
  float* out_act;
  int a_col[kMaxTaps];
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
  const float* rowvec;  // GEMM_SOFTMAX_BWD: per-row scalar
  int drop_ld;          // GEMM_SOFTMAX_BWD: row pitch of the dropout index
  int dbg;     // bring-up knobs (XVA_GEMM_DBG): 1 = no epilogue stores, 2 = no TMA after the first ring fill, 4 = no MMA
  int vec_ok;  // every epilogue pointer is 16-byte aligned and every stride a multiple of 4: float4 accesses
  int round_on;  // host mirror of the operand-rounding test switch (xva_set_operand_rounding), read once per tile
};

struct TileCoord {
  int z;      // mode 0/1: batch item; mode 2: first z of the reduced range
  int z_end;  // mode 2: one past the last z of the reduced range
  int zo;     // output batch index
  int j;      // mode 2: tap of this output tile
  int m0;     // first output row of the tile
  int n0;     // first output column of the tile
  int iters;  // k-iterations of the main loop
  int g;      // group (0 when the launch is not grouped)
  int s0;     // segmented tiles: first segment (in item-major order) of this tile
  bool dup;   // CTA pair, odd row-tile count: this CTA repeats its partner's rows and stores nothing
};

// kCG = 2: a tile belongs to a CTA pair; the two CTAs take consecutive 128-row tiles (possibly of different batch
// items -- they only have to share the B operand) and the same n tile.
template <int kCG>
__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int t, int cta_rank) {
  TileCoord c;
  c.dup = false;
  int n_t = t % p.tiles_n;
  t /= p.tiles_n;
  if (kCG == 2) {
    t = 2 * t + cta_rank;
    if (t >= p.row_tiles) {
      t = p.row_tiles - 1;
      c.dup = true;
    }
  }
  c.s0 = 0;
  int m_t;
  if (p.seg < kBlockM && p.mode != 2) {  // tile t covers segments [t * 128/seg, (t+1) * 128/seg)
    c.s0 = t * (kBlockM / p.seg);
    m_t = 0;
    t = 0;  // c.z / c.m0 are per segment (see seg_coord)
  } else {
    m_t = t % p.tiles_m;
    t /= p.tiles_m;
  }
  c.n0 = n_t * p.n_tile;
  c.m0 = m_t * kBlockM;
  c.g = (p.groups > 1) ? ((p.mode != 2) ? n_t : m_t * p.gpt) : 0;
  if (p.mode != 2) {
    c.z = t;
    c.z_end = t + 1;
    c.zo = t;
    c.j = 0;
    c.iters = p.taps * p.k_chunks;
  } else {
    c.j = (t % p.tgroups) * p.tp;  // first tap of the tile's tap group (tp = 1: the tap)
    t /= p.tgroups;
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

// Segment S (item-major) -> (item, first row). Segments past the last item map to item Z: TMA zero-fills them and
// the epilogue skips them.
__device__ __forceinline__ void seg_coord(const GemmDev& p, int S, int& z, int& r0) {
  z = S / p.segs;
  r0 = (S - z * p.segs) * p.seg;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout), version 1.
//   K-major : SWIZZLE_128B (layout type 2); 8-row groups of 128B rows, SBO = 1024 B between groups, LBO unused (1).
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (layout type 1; TMA 128B_ATOM_32B): atoms of
//             4 k-rows x 128 B whose 32-byte chunks are XOR-ed with (row & 3); LBO = byte distance between
//             32-element MN chunks, SBO = 512 B between 4-row k groups.
// Bring-up counters (XVA_GEMM_DBG & 32): cycles CTA 0 spends in each role state, summed over its tiles.
//  [0] MMA warp waiting for a free accumulator   [1] MMA warp waiting for operands   [2] MMA warp issuing
//  [3] epilogue warp 4 waiting for an accumulator [4] epilogue warp 4 working        [5] tiles of CTA 0   [6] total
__device__ long long g_gemm_dbg[8];

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
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // c_format = F32
  d |= 2u << 7;                                   // a_format = TF32
  d |= 2u << 10;                                  // b_format = TF32
  d |= static_cast<uint32_t>(a_mn_major) << 15;   // a_major
  d |= static_cast<uint32_t>(b_mn_major) << 16;   // b_major
  d |= static_cast<uint32_t>(n >> 3) << 17;       // n_dim
  d |= static_cast<uint32_t>(m >> 4) << 24;       // m_dim (256 = both CTAs of a pair)
  return d;
}


template <int kEpi, int kCG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ GemmDev p) {
  // One dynamic allocation, carved by hand (the ring + staging leave < 1 KiB of the 227 KiB an SM offers):
  //   [operand ring: stages * stage_bytes, 1024-aligned][epilogue transpose tiles 32 KiB][LN exchange 2 KiB][barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];

  // 1024-byte alignment of the tile ring (SWIZZLE_128B atoms are 1024 B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) > kAlignSlack) {
    printf("gemm_tc_kernel: dynamic shared memory base needs %d B of alignment slack (have %d)\n",
           static_cast<int>(smem - smem_raw), kAlignSlack);
    __trap();
  }
  const int b_tile_bytes = p.n_tile * kBlockK * 4 / kCG;  // a CTA pair splits every B tile along n
  const int stage_bytes = kATileBytes + b_tile_bytes * (p.mode == 2 ? p.tp : 1);  // multi-tap wgrad: tp B tiles per stage
  const uint32_t cta_rank = (kCG == 2) ? ptx::cluster_ctarank() : 0u;
  const int cta_id = (kCG == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_ctas = (kCG == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  uint8_t* tail = smem + p.ring_bytes + kEpiSmemBytes;
  float (*ln_part)[4][32][2] = reinterpret_cast<float (*)[4][32][2]>(tail);  // LayerNorm (mean, M2) exchange
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(tail + kLnSmemBytes);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_tmem_full = bar_empty + kMaxStages;
  uint64_t* bar_tmem_empty = bar_tmem_full + 2;
  uint32_t& tmem_base_slot = *reinterpret_cast<uint32_t*>(bar_tmem_empty + 2);
  uint64_t* bar_afull = bar_tmem_empty + 3;
  uint64_t* bar_aempty = bar_afull + kMaxStagesA;
  const int a_stage_bytes = p.a_rows * kBlockK * 4;               // halo mode
  uint8_t* ring_b = smem + p.stages_a * a_stage_bytes;            // halo mode: the weight ring follows the activation ring

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
      ptx::mbar_init(&bar_tmem_empty[a], 8 * kCG);
    }
    for (int a = 0; a < p.stages_a; ++a) {
      ptx::mbar_init(&bar_afull[a], 1);
      ptx::mbar_init(&bar_aempty[a], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (kCG == 2) {
      ptx::tmem_alloc_cg2(&tmem_base_slot, kTmemCols);
      ptx::tmem_relinquish_cg2();
    } else {
      ptx::tmem_alloc(&tmem_base_slot, kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (kCG == 2) ptx::cluster_sync_all();  // the peer's barriers must exist before anything is signalled across
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touches no
  // global memory and may overlap the tail of the previous kernel in the stream; from here on this grid reads and
  // writes tensors, so it waits for the previous grid to complete and flush. The trigger lets the next kernel's CTAs
  // start their own prologue on SMs this grid no longer occupies (a no-op pair when launched without the attribute).
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const bool a_mn = (p.mode == 2);
  const bool b_mn = (p.mode != 0);

  if (warp == 0 && p.halo) {
    // ------------------------------------------------------------------ TMA producer, halo mode
    int s = 0, sa_i = 0;
    uint32_t ph = 0, pha = 0;
    for (int tile = cta_id; tile < p.num_tiles; tile += n_ctas) {
      const TileCoord c = decode_tile<kCG>(p, tile, cta_rank);
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        ptx::mbar_wait(&bar_aempty[sa_i], pha ^ 1);
        if (ptx::elect_one()) {
          uint8_t* sa = smem + sa_i * a_stage_bytes;
          if constexpr (kCG == 2) {
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&bar_afull[sa_i], 2 * a_stage_bytes);
            const uint32_t full = ptx::mapa(ptx::smem_u32(&bar_afull[sa_i]), 0);
            ptx::tma_load_3d_cg2(sa, &tmap_a, full, p.a_col[0] + kc * kBlockK, c.m0 - p.hp, c.z);
          } else {
            ptx::mbar_arrive_expect_tx(&bar_afull[sa_i], a_stage_bytes);
            ptx::tma_load_3d(sa, &tmap_a, &bar_afull[sa_i], p.a_col[0] + kc * kBlockK, c.m0 - p.hp, c.z);
          }
        }
        __syncwarp();
        if (++sa_i == p.stages_a) {
          sa_i = 0;
          pha ^= 1;
        }
        for (int j = 0; j < p.taps; ++j) {
          ptx::mbar_wait(&bar_empty[s], ph ^ 1);
          if (ptx::elect_one()) {
            uint8_t* sb = ring_b + s * b_tile_bytes;
            const int zb = j * p.b_tap_z;
            if constexpr (kCG == 2) {
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&bar_full[s], 2 * b_tile_bytes);
              const uint32_t full = ptx::mapa(ptx::smem_u32(&bar_full[s]), 0);
              const int half_n = p.n_sub >> 1;
              for (int sub = 0; sub < p.n_mma; ++sub) {
                const int nb = c.n0 + sub * p.n_sub + static_cast<int>(cta_rank) * half_n;
                if (p.mode == 0)
                  ptx::tma_load_3d_cg2(sb + sub * half_n * kBlockK * 4, &tmap_b, full, kc * kBlockK, nb, zb);
                else
                  ptx::tma_load_4d_cg2(sb + sub * half_n * kBlockK * 4, &tmap_b, full, 0, kc * kBlockK, nb / 32, zb);
              }
            } else {
              ptx::mbar_arrive_expect_tx(&bar_full[s], b_tile_bytes);
              if (p.mode == 0) {
                for (int sub = 0; sub < p.n_mma; ++sub)
                  ptx::tma_load_3d(sb + sub * p.n_sub * kBlockK * 4, &tmap_b, &bar_full[s], kc * kBlockK,
                                   c.n0 + sub * p.n_sub, zb);
              } else {
                ptx::tma_load_4d(sb, &tmap_b, &bar_full[s], 0, kc * kBlockK, c.n0 / 32, zb);
              }
            }
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = cta_id; tile < p.num_tiles; tile += n_ctas) {
        const TileCoord c = decode_tile<kCG>(p, tile, cta_rank);
        if (p.dbg & 16) continue;  // probe: raw MMA issue rate, no operand pipeline at all
        // outer index: tap (mode 0/1) or batch item of the reduced range (mode 2); inner: 32-wide k block
        const int n_outer = c.iters / p.k_chunks;
        // grouped mode 1: the group's contraction rows start at g*K of B and its B columns at 0 (not at n0)
        const int bk0 = c.g * p.grp_bk;
        const int bn0 = (p.groups > 1 && p.mode == 1) ? 0 : c.n0;
        const int n_seg = (p.mode != 2) ? kBlockM / p.seg : 1;
        int seg_z[4], seg_r[4];
        seg_z[0] = c.z;
        seg_r[0] = c.m0;
        if (n_seg > 1) {
#pragma unroll
          for (int sg = 0; sg < 4; ++sg)
            if (sg < n_seg) seg_coord(p, c.s0 + sg, seg_z[sg], seg_r[sg]);
        }
        const int seg_bytes = p.seg * kBlockK * 4;
        int it = 0;
        // Loop order. Mode 2: reduction unit outer, k-block inner. Mode 0/1 with p.tap_inner: k-block outer, TAP INNER --
        // the taps of one k-block re-read the same activation lines (shifted by a row) back to back, so they hit L2; with
        // the tap outer, a CTA streams its whole 128 x K row block once per tap and at K = 1536 the 148 CTAs push 116 MB
        // through the 126 MB L2 between two reads of a line (ncu: 459 MB read from DRAM for 223 MB of operands).
        const bool tap_inner = (p.mode != 2) && p.tap_inner;
        const int n1 = tap_inner ? p.k_chunks : n_outer, n2 = tap_inner ? n_outer : p.k_chunks;
        for (int o1 = 0; o1 < n1; ++o1) {
          for (int o2 = 0; o2 < n2; ++o2, ++it) {
            const int jo = tap_inner ? o2 : o1, kc = tap_inner ? o1 : o2;
            const int shift_j = (p.mode != 2) ? p.shift[jo] : p.shift[c.j];
            const int acol_j = (p.mode != 2) ? p.a_col[jo] : 0;
            ptx::mbar_wait(&bar_empty[s], ph ^ 1);
            uint8_t* sa = smem + s * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            if (!ptx::elect_one()) {
            } else if ((p.dbg & 2) && (tile != cta_id || it >= p.stages)) {
              if (kCG == 1 || cta_rank == 0) ptx::mbar_arrive(&bar_full[s]);
            } else if constexpr (kCG == 2) {
              // Both CTAs load into their own shared memory; all completion bytes land on the LEADER's full
              // barrier, which alone expects them (its MMA thread is the only consumer).
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&bar_full[s], 2 * stage_bytes);
              const uint32_t full = ptx::mapa(ptx::smem_u32(&bar_full[s]), 0);
              const int zb = jo * p.b_tap_z;
              const int half_n = p.n_sub >> 1;
#pragma unroll
              for (int sg = 0; sg < 4; ++sg)
                if (sg < n_seg)
                  ptx::tma_load_3d_cg2(sa + sg * seg_bytes, &tmap_a, full, acol_j + c.g * p.grp_a + kc * kBlockK,
                                       seg_r[sg] + shift_j, seg_z[sg]);
              for (int sub = 0; sub < p.n_mma; ++sub) {
                const int nb = bn0 + sub * p.n_sub + static_cast<int>(cta_rank) * half_n;
                if (p.mode == 0)
                  ptx::tma_load_3d_cg2(sb + sub * half_n * kBlockK * 4, &tmap_b, full, kc * kBlockK, nb, zb);
                else
                  ptx::tma_load_4d_cg2(sb + sub * half_n * kBlockK * 4, &tmap_b, full, 0, bk0 + kc * kBlockK, nb / 32, zb);
              }
            } else if (p.mode != 2) {
              ptx::mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
              const int zb = jo * p.b_tap_z + c.z * p.b_batch_z;
#pragma unroll
              for (int sg = 0; sg < 4; ++sg)
                if (sg < n_seg)
                  ptx::tma_load_3d(sa + sg * seg_bytes, &tmap_a, &bar_full[s], acol_j + c.g * p.grp_a + kc * kBlockK,
                                   seg_r[sg] + shift_j, seg_z[sg]);
              if (p.mode == 0) {
                for (int sub = 0; sub < p.n_mma; ++sub)
                  ptx::tma_load_3d(sb + sub * p.n_sub * kBlockK * 4, &tmap_b, &bar_full[s], kc * kBlockK,
                                   c.n0 + sub * p.n_sub, zb);
              } else {
                ptx::tma_load_4d(sb, &tmap_b, &bar_full[s], 0, bk0 + kc * kBlockK, bn0 / 32, zb);
              }
            } else {
              const int nt_here = (p.taps - c.j) < p.tp ? (p.taps - c.j) : p.tp;
              ptx::mbar_arrive_expect_tx(&bar_full[s], kATileBytes + nt_here * b_tile_bytes);
              const int u = c.z + jo;  // reduction unit = (item, row chunk)
              const int z = u / p.rsplit;
              const int r0 = (u - z * p.rsplit) * p.chunk_rows + kc * kBlockK;
              ptx::tma_load_4d(sa, &tmap_a, &bar_full[s], 0, r0, c.m0 / 32, z);
              for (int jj = 0; jj < nt_here; ++jj)
                ptx::tma_load_4d(sb + jj * b_tile_bytes, &tmap_b, &bar_full[s], 0, r0 + p.shift[c.j + jj],
                                 (c.n0 + p.a_col[c.j + jj] + c.g * p.grp_a) / 32, z);
            }
            __syncwarp();
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (cta_rank == 0) {  // whole warp, uniform control flow; one elected lane issues
      const uint32_t idesc = make_idesc(p.n_sub, a_mn ? 1 : 0, b_mn ? 1 : 0, kBlockM * kCG);
      // descriptor = constant fields | (shared address >> 4); advancing along k only touches the address field
      const uint64_t da_hi = make_smem_desc(0, a_mn ? p.mn_lbo : 16, a_mn ? p.mn_sbo : 1024, a_mn ? p.mn_layout : 2);
      const uint64_t db_hi = make_smem_desc(0, b_mn ? p.mn_lbo : 16, b_mn ? p.mn_sbo : 1024, b_mn ? p.mn_layout : 2);
      const uint32_t a_kstep = (a_mn ? 1024 : kUmmaK * 4) >> 4;  // address-field units (16 B) per UMMA_K
      const uint32_t b_kstep = (b_mn ? 1024 : kUmmaK * 4) >> 4;
      const uint32_t b_sub = ((b_mn ? (p.n_sub / 32) * (kBlockK * 128) : p.n_sub * kBlockK * 4) / kCG) >> 4;
      const uint32_t ring = ptx::smem_u32(smem) >> 4;
      const uint32_t stage_u = static_cast<uint32_t>(stage_bytes) >> 4;
      const int n_mma = p.n_mma, n_sub = p.n_sub, dbg = p.dbg;
      int s = 0;
      uint32_t ph = 0;
      uint32_t a_lo = ring;  // address field (16-byte units) of the current stage's A tile
      int tile_iter = 0;
      long long cyc_wait_acc = 0, cyc_wait_ops = 0;
      const long long t_start = clock64();
      if (p.halo) {
        // halo mode: k-block outer (one activation stage), taps inner (one weight stage each); tap j reads the
        // activation stage starting (hp + shift_j) rows in
        const uint32_t ring_b_u = ptx::smem_u32(ring_b) >> 4;
        const uint32_t bstage_u = static_cast<uint32_t>(b_tile_bytes) >> 4;
        const uint32_t astage_u = static_cast<uint32_t>(a_stage_bytes) >> 4;
        int sa_i = 0;
        uint32_t pha = 0;
        for (int tile = cta_id; tile < p.num_tiles; tile += n_ctas, ++tile_iter) {
          const int acc = tile_iter % p.acc_stages;
          const uint32_t acc_ph = (tile_iter / p.acc_stages) & 1;
          ptx::mbar_wait(&bar_tmem_empty[acc], acc_ph ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_acc = tmem_base + acc * 256;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            ptx::mbar_wait(&bar_afull[sa_i], pha);
            ptx::tc_fence_after();
            for (int j = 0; j < p.taps; ++j) {
              ptx::mbar_wait(&bar_full[s], ph);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t a_u = ring + sa_i * astage_u + static_cast<uint32_t>((p.hp + p.shift[j]) * (kBlockK * 4 / 16));
                const uint64_t da0 = da_hi | static_cast<uint64_t>(a_u);
                const uint64_t db0 = db_hi | static_cast<uint64_t>(ring_b_u + s * bstage_u);
                const uint32_t first = (kc > 0 || j > 0) ? 1u : 0u;
#pragma unroll
                for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                  for (int sub = 0; sub < n_mma; ++sub) {
                    if (kCG == 2)
                      ptx::mma_tf32_cg2(tmem_acc + sub * n_sub, da0 + k4 * a_kstep, db0 + sub * b_sub + k4 * b_kstep, idesc,
                                        k4 ? 1u : first);
                    else
                      ptx::mma_tf32(tmem_acc + sub * n_sub, da0 + k4 * a_kstep, db0 + sub * b_sub + k4 * b_kstep, idesc,
                                    k4 ? 1u : first);
                  }
                }
                if (kCG == 2) ptx::mma_commit_cg2(&bar_empty[s], 3);
                else ptx::mma_commit(&bar_empty[s]);
              }
              __syncwarp();
              if (++s == p.stages) {
                s = 0;
                ph ^= 1;
              }
            }
            if (ptx::elect_one()) {  // every tap of this k-block has been issued: the activation stage frees with them
              if (kCG == 2) ptx::mma_commit_cg2(&bar_aempty[sa_i], 3);
              else ptx::mma_commit(&bar_aempty[sa_i]);
            }
            __syncwarp();
            if (++sa_i == p.stages_a) {
              sa_i = 0;
              pha ^= 1;
            }
          }
          if (ptx::elect_one()) {
            if (kCG == 2) ptx::mma_commit_cg2(&bar_tmem_full[acc], 3);
            else ptx::mma_commit(&bar_tmem_full[acc]);
          }
          __syncwarp();
        }
      } else
      for (int tile = cta_id; tile < p.num_tiles; tile += n_ctas, ++tile_iter) {
        const TileCoord c = decode_tile<kCG>(p, tile, 0);
        const int acc = tile_iter % p.acc_stages;
        const uint32_t acc_ph = (tile_iter / p.acc_stages) & 1;
        long long t_a = clock64();
        ptx::mbar_wait(&bar_tmem_empty[acc], acc_ph ^ 1);
        ptx::tc_fence_after();
        long long t_b = clock64();
        cyc_wait_acc += t_b - t_a;
        const uint32_t tmem_acc = tmem_base + acc * 256;
        for (int it = 0; it < c.iters; ++it) {
          if (!(dbg & 16)) {
            long long t_c = clock64();
            ptx::mbar_wait(&bar_full[s], ph);
            ptx::tc_fence_after();
            cyc_wait_ops += clock64() - t_c;
          }
          if (ptx::elect_one()) {
            // straight-line issue: 4 k-slices x n_mma column halves, descriptors differ only in the address field
            const uint64_t da0 = da_hi | static_cast<uint64_t>(a_lo);
            const uint64_t db0 = db_hi | static_cast<uint64_t>(a_lo + (kATileBytes >> 4));
            const uint32_t first = it > 0 ? 1u : 0u;
            if (dbg & 4) {
            } else if (n_mma == 1) {
              // (multi-tap weight gradient: the stage holds one B tile per tap of the group, each with its own accumulator)
              const int nt_here = (p.mode == 2 && p.tp > 1) ? ((p.taps - c.j) < p.tp ? (p.taps - c.j) : p.tp) : 1;
              const uint32_t btile_u = static_cast<uint32_t>(b_tile_bytes) >> 4;
              for (int jj = 0; jj < nt_here; ++jj) {
#pragma unroll
                for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                  if (kCG == 2)
                    ptx::mma_tf32_cg2(tmem_acc + jj * p.n_tile, da0 + k4 * a_kstep, db0 + jj * btile_u + k4 * b_kstep, idesc,
                                      k4 ? 1u : first);
                  else
                    ptx::mma_tf32(tmem_acc + jj * p.n_tile, da0 + k4 * a_kstep, db0 + jj * btile_u + k4 * b_kstep, idesc,
                                  k4 ? 1u : first);
                }
              }
            } else {
#pragma unroll
              for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                  if (kCG == 2)
                    ptx::mma_tf32_cg2(tmem_acc + sub * n_sub, da0 + k4 * a_kstep, db0 + sub * b_sub + k4 * b_kstep, idesc,
                                      k4 ? 1u : first);
                  else
                    ptx::mma_tf32(tmem_acc + sub * n_sub, da0 + k4 * a_kstep, db0 + sub * b_sub + k4 * b_kstep, idesc,
                                  k4 ? 1u : first);
                }
              }
            }
            // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
            if (dbg & 16) {
            } else if (kCG == 2) ptx::mma_commit_cg2(&bar_empty[s], 3);
            else ptx::mma_commit(&bar_empty[s]);
          }
          __syncwarp();
          a_lo += stage_u;
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
            a_lo = ring;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (ptx::elect_one()) {
          if (kCG == 2) ptx::mma_commit_cg2(&bar_tmem_full[acc], 3);
          else ptx::mma_commit(&bar_tmem_full[acc]);
        }
        __syncwarp();
      }
      if ((dbg & 32) && blockIdx.x == 0 && lane == 0) {
        const long long total = clock64() - t_start;
        g_gemm_dbg[0] = cyc_wait_acc;
        g_gemm_dbg[1] = cyc_wait_ops;
        g_gemm_dbg[2] = total - cyc_wait_acc - cyc_wait_ops;
        g_gemm_dbg[5] = tile_iter;
        g_gemm_dbg[6] = total;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    // TMEM lanes 32q..32q+31 are only visible to warps with (warp & 3) == q, so two warps share each 32-row quarter
    // and split the tile's 32-column chunks between them (even / odd). A 32x32 block is read with one tcgen05.ld
    // (thread = row), transposed through a 4 KiB XOR-swizzled shared-memory tile, and from there on every thread
    // holds float4 pieces in a COALESCED layout: lane l <-> columns 4*(l&7)..+3 of rows 4k + (l>>3), k = 0..7, so
    // eight consecutive lanes cover one full 128-byte line of a row for every global load (bias, gate, residual),
    // store and reduction. All global loads of a chunk are issued before the first use (8 x 16 B in flight per
    // lane): the epilogue of the memory-bound launches is a bandwidth problem, not a latency chain.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int c4 = lane & 7, rsub = lane >> 3;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    float4* tbuf = reinterpret_cast<float4*>(smem + p.ring_bytes) + (warp - 4) * 256;
    int tile_iter = 0;
    long long epi_wait = 0, epi_work = 0;
    const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) * 0xA24BAED4963EE407ull : 0ull);
    const bool vec = p.vec_ok != 0;

    auto load_chunk = [&](uint32_t taddr, float4 (&t)[8]) {
      uint32_t v[32];
      ptx::tmem_ld32(taddr, v);
      ptx::tmem_wait_ld();
      __syncwarp();  // everyone has finished reading the previous block out of tbuf
#pragma unroll
      for (int c = 0; c < 8; ++c)
        tbuf[lane * 8 + (c ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                        __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = 4 * k + rsub;
        t[k] = tbuf[r * 8 + (c4 ^ (r & 7))];
      }
    };
    // Batched access to one float4 column slot of the 8 rows this lane owns. `full`: all four columns valid and
    // float4-aligned (the common case); otherwise `nv` (< 4 or unaligned) columns are touched one by one.
    auto load8 = [&](const float* base, long rs, int row0, int row_limit, bool full, int nv, float4 (&o)[8]) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = row0 + 4 * k;
        o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_limit) {
          const float* ptr = base + static_cast<long>(row) * rs;
          if (full) {
            o[k] = __ldg(reinterpret_cast<const float4*>(ptr));
          } else {
            if (nv > 0) o[k].x = __ldg(ptr);
            if (nv > 1) o[k].y = __ldg(ptr + 1);
            if (nv > 2) o[k].z = __ldg(ptr + 2);
            if (nv > 3) o[k].w = __ldg(ptr + 3);
          }
        }
      }
    };
    auto store1 = [&](float* ptr, bool full, int nv, const float4 v) {
      if (full) {
        *reinterpret_cast<float4*>(ptr) = v;
      } else {
        if (nv > 0) ptr[0] = v.x;
        if (nv > 1) ptr[1] = v.y;
        if (nv > 2) ptr[2] = v.z;
        if (nv > 3) ptr[3] = v.w;
      }
    };

    for (int tile = cta_id; tile < p.num_tiles; tile += n_ctas, ++tile_iter) {
      const TileCoord c = decode_tile<kCG>(p, tile, cta_rank);
      const int acc = tile_iter % p.acc_stages;
      const uint32_t acc_ph = (tile_iter / p.acc_stages) & 1;
      const long long t_e0 = clock64();
      ptx::mbar_wait(&bar_tmem_full[acc], acc_ph);
      ptx::tc_fence_after();
      const long long t_e1 = clock64();
      epi_wait += t_e1 - t_e0;
      const uint32_t tacc = tmem_base + acc * 256 + lane_base;

      int row_limit = (c.dup || (p.dbg & 1)) ? 0 : ((kEpi == EPI_WGRAD) ? p.M : p.R);
      const int n_cols = (p.N - c.n0) < p.n_tile ? (p.N - c.n0) : p.n_tile;  // valid columns of this tile
      const int n_chunks = (n_cols + 31) / 32;
      // item and first row of this warp's 32-row quarter (a quarter never straddles a segment: seg % 32 == 0)
      int zq = c.z, rq0 = c.m0 + q * 32;
      if (kEpi != EPI_WGRAD && p.seg < kBlockM) {
        seg_coord(p, c.s0 + (q * 32) / p.seg, zq, rq0);
        rq0 += (q * 32) % p.seg;
        if (zq >= p.Z) {
          zq = 0;
          row_limit = 0;
        }
      }
      const int row0 = rq0 + rsub;  // this lane's rows are row0 + 4k
      const int last_ch = ((n_chunks - 1 - half) & ~1) + half;  // last chunk of this warp (< half: none)
      bool released = false;
      auto release_tmem = [&]() {  // accumulator fully read: the MMA warp may start the next tile into it
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCG == 2) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bar_tmem_empty[acc]), 0));
          else ptx::mbar_arrive(&bar_tmem_empty[acc]);
        }
        released = true;
      };

      if constexpr (kEpi == EPI_WGRAD) {
        float* obase = p.out + c.zo * p.o_zs + c.j * p.o_js + c.n0;  // tap c.j; re-pointed per tap of a multi-tap tile
        // grouped: this warp's 32 rows belong to group gi of the tile; only its diagonal block is stored
        int ch_lo = 0, ch_hi = n_chunks, col_shift = 0;
        if (p.groups > 1) {
          const int gi = (q * 32) / p.og;
          ch_lo = gi * p.cg / 32;
          ch_hi = ch_lo + p.cg / 32;
          col_shift = gi * p.cg;
        }
        const int nt_here = (p.tp > 1) ? ((p.taps - c.j) < p.tp ? (p.taps - c.j) : p.tp) : 1;
        for (int qi = half; qi < nt_here * n_chunks; qi += 2) {
          const int jj = qi / n_chunks, ch = qi - jj * n_chunks;  // tap of the group, 32-column chunk of its accumulator
          if (ch < ch_lo || ch >= ch_hi) continue;  // warp-uniform
          float4 t[8];
          load_chunk(tacc + jj * p.n_tile + ch * 32, t);
          if (p.tp == 1 && ch == last_ch) release_tmem();
          obase = p.out + c.zo * p.o_zs + (c.j + jj) * p.o_js + c.n0;
          const int n = ch * 32 + 4 * c4;
          const int nv = n < n_cols ? ((n_cols - n) < 4 ? (n_cols - n) : 4) : 0;
          const bool full = vec && nv == 4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = row0 + 4 * k;
            if (row >= row_limit || nv == 0) continue;
            float* dst = obase + static_cast<long>(row) * p.o_rs + (n - col_shift);
            const float4 o = make_float4(p.alpha * t[k].x, p.alpha * t[k].y, p.alpha * t[k].z, p.alpha * t[k].w);
            if (p.flags & GEMM_ATOMIC) {
              if (full) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z),
                             "f"(o.w)
                             : "memory");
              } else {
                if (nv > 0) atomicAdd(dst, o.x);
                if (nv > 1) atomicAdd(dst + 1, o.y);
                if (nv > 2) atomicAdd(dst + 2, o.z);
                if (nv > 3) atomicAdd(dst + 3, o.w);
              }
            } else {
              store1(dst, full, nv, (p.flags & GEMM_ROUND_OUT) ? tf32_rn4(o) : o);
            }
          }
        }
      } else {
        const int len_z = p.lens ? p.lens[zq] : 0x7fffffff;
        const long zoff_o = zq * p.o_zs + c.n0, zoff_r = zq * p.r_zs + c.n0, zoff_g = zq * p.g_zs + c.n0;

        // Every feature test below is uniform over the launch: it is evaluated once per tile and each feature is its own
        // loop over the lane's 8 rows. (With the tests inside one fused per-row loop the epilogue executed ~19 thread
        // instructions per output element and the epilogue-bound launches -- attention scores, o_net, gated input
        // gradients -- ran instruction-bound at ~1.8 TB/s of stores: profiles/r01_epilogue_stalls.txt.)
        const int flags = p.flags;
        const bool f_bias = p.bias != nullptr, f_relu = (flags & GEMM_RELU) != 0, f_tanh = (flags & GEMM_TANH) != 0;
        const bool f_round = (flags & GEMM_ROUND_OUT) != 0 && p.round_on != 0, f_lens = p.lens != nullptr;
        const bool f_act = p.out_act != nullptr, f_scale = p.alpha != 1.0f;
        const bool f_gate = (kEpi != EPI_PLAIN) && p.gate != nullptr, f_res = (kEpi != EPI_PLAIN) && p.residual != nullptr;
        const bool f_drop = (kEpi != EPI_PLAIN) && (flags & GEMM_DROP_PRE) != 0;
        const bool f_sbwd = (kEpi != EPI_PLAIN) && (flags & GEMM_SOFTMAX_BWD) != 0;
        const float alpha = p.alpha, act_slope = p.act_slope, gate_slope = p.gate_slope;
        auto cvt4 = [](float4 v) {  // fp32 -> tf32, round to nearest (the switch was read once: f_round)
          uint32_t a, b, c2, d;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v.x));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.y));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c2) : "f"(v.z));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(v.w));
          return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c2), __uint_as_float(d));
        };

        // x = alpha*acc + bias -> relu -> gate -> dropout(pre) -> + residual   for the 8 float4 of one chunk
        auto finish_chunk = [&](float4 (&t)[8], int n, bool full, int nv) {
          // One auxiliary row set: the gate (or, without a gate, the residual) is requested before the arithmetic on
          // the accumulator so its latency overlaps; with both, the residual is fetched after the gate is consumed
          // (two live sets next to the accumulator spill at 168 registers).
          float4 aux[8];
          if constexpr (kEpi != EPI_PLAIN) {
            if (f_gate) load8(p.gate + zoff_g + n, p.g_rs, row0, row_limit, full, nv, aux);
            else if (f_res) load8(p.residual + zoff_r + n, p.r_rs, row0, row_limit, full, nv, aux);
          }
          if constexpr (kEpi != EPI_PLAIN) {
            if (f_sbwd) {  // dS = alpha * P * (dP_dropped * mask - D[row]); aux holds P (requested above via the gate slot)
              const uint64_t d0 = (static_cast<uint64_t>(zq) * p.R + row0) * static_cast<uint64_t>(p.drop_ld) + c.n0 + n;
              const uint64_t dstep = 4ull * static_cast<uint64_t>(p.drop_ld);
              const float* dv = p.rowvec + static_cast<long>(zq) * p.R;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int row = row0 + 4 * k;
                const float dr = row < row_limit ? __ldg(dv + row) : 0.0f;
                const float4 ds = dropout_scale4(seed, d0 + k * dstep, p.drop_thresh, p.inv_keep);
                t[k] = make_float4(alpha * aux[k].x * (t[k].x * ds.x - dr), alpha * aux[k].y * (t[k].y * ds.y - dr),
                                   alpha * aux[k].z * (t[k].z * ds.z - dr), alpha * aux[k].w * (t[k].w * ds.w - dr));
              }
              return;
            }
          }
          if (f_bias) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nv) {
              const float* bp = p.bias + c.n0 + n;
              if (full) b = __ldg(reinterpret_cast<const float4*>(bp));
              else {
                b.x = __ldg(bp);
                if (nv > 1) b.y = __ldg(bp + 1);
                if (nv > 2) b.z = __ldg(bp + 2);
                if (nv > 3) b.w = __ldg(bp + 3);
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              t[k] = make_float4(alpha * t[k].x + b.x, alpha * t[k].y + b.y, alpha * t[k].z + b.z, alpha * t[k].w + b.w);
          } else if (f_scale) {
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = make_float4(alpha * t[k].x, alpha * t[k].y, alpha * t[k].z, alpha * t[k].w);
          }
          if (f_relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              t[k].x = t[k].x > 0.f ? t[k].x : act_slope * t[k].x; t[k].y = t[k].y > 0.f ? t[k].y : act_slope * t[k].y;
              t[k].z = t[k].z > 0.f ? t[k].z : act_slope * t[k].z; t[k].w = t[k].w > 0.f ? t[k].w : act_slope * t[k].w;
            }
          }
          if constexpr (kEpi != EPI_PLAIN) {
            if (f_gate) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                t[k].x *= aux[k].x > 0.f ? 1.f : gate_slope; t[k].y *= aux[k].y > 0.f ? 1.f : gate_slope;
                t[k].z *= aux[k].z > 0.f ? 1.f : gate_slope; t[k].w *= aux[k].w > 0.f ? 1.f : gate_slope;
              }
              if (f_res) load8(p.residual + zoff_r + n, p.r_rs, row0, row_limit, full, nv, aux);
            }
            if (f_drop) {
              const uint64_t d0 = (static_cast<uint64_t>(zq) * p.R + row0) * static_cast<uint64_t>(p.N) + c.n0 + n;
              const uint64_t dstep = 4ull * static_cast<uint64_t>(p.N);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 ds = dropout_scale4(seed, d0 + k * dstep, p.drop_thresh, p.inv_keep);
                t[k].x *= ds.x; t[k].y *= ds.y; t[k].z *= ds.z; t[k].w *= ds.w;
              }
            }
            if (f_res) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                t[k].x += aux[k].x; t[k].y += aux[k].y; t[k].z += aux[k].z; t[k].w += aux[k].w;
              }
            }
          }
        };

        if constexpr (kEpi != EPI_LN) {
          const long ostep = 4L * p.o_rs;
          for (int ch = half; ch < n_chunks; ch += 2) {
            float4 t[8];
            load_chunk(tacc + ch * 32, t);
            if (ch == last_ch) release_tmem();
            const int n = ch * 32 + 4 * c4;
            const int nv = n < n_cols ? ((n_cols - n) < 4 ? (n_cols - n) : 4) : 0;
            const bool full = vec && nv == 4;
            finish_chunk(t, n, full, nv);
            if (f_tanh) {
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = make_float4(tanhf(t[k].x), tanhf(t[k].y), tanhf(t[k].z), tanhf(t[k].w));
            }
            if (f_lens) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (row0 + 4 * k >= len_z) t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // rows this lane may store: k < kmax (none for an empty column slot)
            const int kmax = nv == 0 ? 0 : (row_limit - row0 + 3) / 4;
            const long off0 = zoff_o + static_cast<long>(row0) * p.o_rs + n;
            if (f_act) {
              const float sl = p.out_act_slope;
              float* ap = p.out_act + off0;
#pragma unroll
              for (int k = 0; k < 8; ++k, ap += ostep) {
                if (k >= kmax) continue;
                float4 a = make_float4(t[k].x > 0.f ? t[k].x : sl * t[k].x, t[k].y > 0.f ? t[k].y : sl * t[k].y,
                                       t[k].z > 0.f ? t[k].z : sl * t[k].z, t[k].w > 0.f ? t[k].w : sl * t[k].w);
                if (p.round_on) a = cvt4(a);
                store1(ap, full, nv, a);
              }
            }
            if (f_round) {
#pragma unroll
              for (int k = 0; k < 8; ++k) t[k] = cvt4(t[k]);
            }
            float* op = p.out + off0;
#pragma unroll
            for (int k = 0; k < 8; ++k, op += ostep) {
              if (k < kmax) store1(op, full, nv, t[k]);
            }
          }
        } else {
          // LayerNorm over the n_cols (= N) columns of each row. Pass 1 finishes the pre-LN value, stores it (to
          // out_pre, or to out as scratch) and accumulates moments about the first element this warp sees of each
          // row (shifted sums: no cancellation when |mean| >> std). The two warps of a quarter then merge their
          // (count, mean, M2) through shared memory, and pass 2 re-reads the values this same thread wrote
          // (L2-resident), normalises and stores. The accumulator is released between the passes.
          float* pre_dst = p.out_pre ? p.out_pre : p.out;
          float s1[8], s2[8], x0[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) s1[k] = s2[k] = x0[k] = 0.0f;
          int cnt = 0;  // columns this warp has accumulated per row
          for (int ch = half; ch < n_chunks; ch += 2) {
            float4 t[8];
            load_chunk(tacc + ch * 32, t);
            if (ch == last_ch) release_tmem();
            const int n = ch * 32 + 4 * c4;
            const int nv = n < n_cols ? ((n_cols - n) < 4 ? (n_cols - n) : 4) : 0;
            const bool full = vec && nv == 4;
            finish_chunk(t, n, full, nv);
            cnt += (n_cols - ch * 32) < 32 ? (n_cols - ch * 32) : 32;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int row = row0 + 4 * k;
              if (row < row_limit && nv) store1(pre_dst + zoff_o + static_cast<long>(row) * p.o_rs + n, full, nv, t[k]);
              if (ch == half) x0[k] = __shfl_sync(0xffffffffu, t[k].x, lane & ~7);
              const float d0 = nv > 0 ? t[k].x - x0[k] : 0.f, d1 = nv > 1 ? t[k].y - x0[k] : 0.f;
              const float d2 = nv > 2 ? t[k].z - x0[k] : 0.f, d3 = nv > 3 ? t[k].w - x0[k] : 0.f;
              s1[k] += (d0 + d1) + (d2 + d3);
              s2[k] += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
          }
          if (!released) release_tmem();
          // per-warp (count, mean, M2) per row -> shared memory -> merge with the partner warp (Chan et al.)
          float mean[8], rstd[8];
          float* my = ln_part[half][q][0];
          const float* other = ln_part[half ^ 1][q][0];
          const float fc = static_cast<float>(cnt);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
              s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
            const float md = cnt ? s1[k] / fc : 0.f;
            mean[k] = x0[k] + md;
            rstd[k] = fmaxf(s2[k] - s1[k] * md, 0.f);  // M2 for now
            if (c4 == 0) {
              my[(4 * k + rsub) * 2] = mean[k];
              my[(4 * k + rsub) * 2 + 1] = rstd[k];
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          const int cnt_o = n_cols - cnt;
          const float fo = static_cast<float>(cnt_o), fn = static_cast<float>(n_cols);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float mo = other[(4 * k + rsub) * 2], m2o = other[(4 * k + rsub) * 2 + 1];
            float m = mean[k], m2 = rstd[k];
            if (cnt_o > 0) {
              const float delta = mo - m;
              m2 = m2 + m2o + delta * delta * (fc * fo / fn);
              m = (fc * m + fo * mo) / fn;
            }
            mean[k] = m;
            rstd[k] = rsqrtf(m2 / fn + p.ln_eps);
            const int row = row0 + 4 * k;
            if (half == 0 && c4 == 0 && row < row_limit) {
              const long srow = static_cast<long>(zq) * p.R + row;
              if (p.ln_mean) p.ln_mean[srow] = mean[k];
              if (p.ln_rstd) p.ln_rstd[srow] = rstd[k];
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // partner has read my stats: slot reusable
          for (int ch = half; ch < n_chunks; ch += 2) {
            const int n = ch * 32 + 4 * c4;
            const int nv = n < n_cols ? ((n_cols - n) < 4 ? (n_cols - n) : 4) : 0;
            if (nv == 0) continue;
            const bool full = vec && nv == 4;
            float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), bt = gm;
            if (full) {
              gm = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
              bt = __ldg(reinterpret_cast<const float4*>(p.beta + n));
            } else {
              gm.x = p.gamma[n]; bt.x = p.beta[n];
              if (nv > 1) { gm.y = p.gamma[n + 1]; bt.y = p.beta[n + 1]; }
              if (nv > 2) { gm.z = p.gamma[n + 2]; bt.z = p.beta[n + 2]; }
              if (nv > 3) { gm.w = p.gamma[n + 3]; bt.w = p.beta[n + 3]; }
            }
            float4 x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // plain (coherent) loads: these addresses were written by this thread
              const int row = row0 + 4 * k;
              x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < row_limit) {
                const float* ptr = pre_dst + zoff_o + static_cast<long>(row) * p.o_rs + n;
                if (full) x[k] = *reinterpret_cast<const float4*>(ptr);
                else {
                  x[k].x = ptr[0];
                  if (nv > 1) x[k].y = ptr[1];
                  if (nv > 2) x[k].z = ptr[2];
                  if (nv > 3) x[k].w = ptr[3];
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int row = row0 + 4 * k;
              if (row >= row_limit) continue;
              float4 y = make_float4((x[k].x - mean[k]) * rstd[k] * gm.x + bt.x, (x[k].y - mean[k]) * rstd[k] * gm.y + bt.y,
                                     (x[k].z - mean[k]) * rstd[k] * gm.z + bt.z, (x[k].w - mean[k]) * rstd[k] * gm.w + bt.w);
              if (p.flags & GEMM_DROP_POST) {
                const uint64_t di = (static_cast<uint64_t>(zq) * p.R + row) * static_cast<uint64_t>(p.N) + c.n0 + n;
                const float4 ds = dropout_scale4(seed, di, p.drop_thresh, p.inv_keep);
                y.x *= ds.x; y.y *= ds.y; y.z *= ds.z; y.w *= ds.w;
              }
              if (row >= len_z) y = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.flags & GEMM_ROUND_OUT) y = tf32_rn4(y);
              store1(p.out + zoff_o + static_cast<long>(row) * p.o_rs + n, full, nv, y);
            }
          }
        }
      }
      if (!released) release_tmem();
      epi_work += clock64() - t_e1;
    }
    if ((p.dbg & 32) && blockIdx.x == 0 && warp == 4 && lane == 0) {
      g_gemm_dbg[3] = epi_wait;
      g_gemm_dbg[4] = epi_work;
    }
  }

  ptx::tc_fence_before();
  if (kCG == 2) {
    ptx::cluster_sync_all();  // neither CTA may retire (or free TMEM) while the pair's MMAs can still touch it
    if (warp == 2) ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
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

int tma_encode_f32(void* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box, int swizzle) {
  return encode_map(static_cast<CUtensorMap*>(map), base, rank, dims, strides_elems, box, swizzle);
}

int gemm_debug_counters(long long* out8) {
  XVA_CHECK_CUDA(cudaMemcpyFromSymbol(out8, g_gemm_dbg, sizeof(long long) * 8));
  return XVA_OK;
}

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
  const int G = g.groups > 1 ? g.groups : 1;
  p.groups = G;
  if (G > 1) {
    XVA_CHECK_ARG(!(g.flags & GEMM_LN) && g.b_batch_z == 0, "gemm: groups with LayerNorm / batched B");
    if (g.mode == 0) {
      XVA_CHECK_ARG(g.N % G == 0 && (g.N / G) % 16 == 0 && g.N / G <= 256, "gemm: grouped fwd needs N/G %% 16 == 0, <= 256 (N=%d G=%d)", g.N, G);
      p.grp_a = g.grp_step;
    } else if (g.mode == 1) {
      XVA_CHECK_ARG(g.N % G == 0 && (g.N / G) % 32 == 0 && g.N / G <= 256 && g.K % 32 == 0,
                    "gemm: grouped dgrad needs N/G %% 32 == 0, <= 256 and K %% 32 == 0 (N=%d K=%d G=%d)", g.N, g.K, G);
      p.grp_a = g.K;
      p.grp_bk = g.K;
    } else {
      const int og = g.M / G;
      XVA_CHECK_ARG(g.M % G == 0 && (og == 32 || og == 64 || og == 128) && g.grp_step == g.N && g.N % 32 == 0 &&
                        (kBlockM / og) * g.N <= 256,
                    "gemm: grouped wgrad needs M/G in {32, 64, 128}, N %% 32 == 0, grp_step == N, (128/Og)*N <= 256 "
                    "(M=%d N=%d G=%d grp_step=%d)", g.M, g.N, G, g.grp_step);
      p.grp_a = g.grp_step;
      p.og = og;
      p.cg = g.N;
      p.gpt = kBlockM / og;
      p.N = p.gpt * g.N;  // columns of one tile: the groups of a 128-row m tile side by side
    }
  }

  // ---- segmented row tiles (mode 0/1, B shared by all items): pick the segment length that needs the fewest tiles
  p.seg = kBlockM;
  p.segs = ceil_div(g.R, kBlockM);
  int row_tiles = (g.mode != 2) ? g.Z * ceil_div(g.R, kBlockM) : 0;
  static const bool seg_enabled = [] {
    const char* e = getenv("XVA_GEMM_SEG");
    return !(e && e[0] == '0');
  }();
  if (seg_enabled && g.mode != 2 && g.b_batch_z == 0 && g.Z > 1) {
    for (int sg = 64; sg >= 32; sg >>= 1) {
      const int t = ceil_div(g.Z * ceil_div(g.R, sg), kBlockM / sg);
      if (t < row_tiles) {
        row_tiles = t;
        p.seg = sg;
        p.segs = ceil_div(g.R, sg);
      }
    }
  }
  p.tot_segs = g.Z * p.segs;

  // ---- CTA pair (cta_group::2): two 128-row tiles share one B tile, each CTA stages half of it. Halves the B bytes
  // every SM pulls through L2 (the fp32 operand stream is L2-bandwidth-bound at 128x256 tiles). Needs a B operand
  // that does not depend on the batch item (weights). Measured per shape on B200: it pays for long main loops over at
  // least one wave of row tiles (decoder ConvFF: 4-7 % faster) and costs 7-15 % on short-K or sub-wave launches.
  static const bool pair_enabled = [] {
    const char* e = getenv("XVA_GEMM_PAIR");
    return !(e && e[0] == '0');
  }();
  static const bool pair_forced = [] {
    const char* e = getenv("XVA_GEMM_PAIR");
    return e && e[0] == '2';
  }();
  auto want_pair = [&](long iters) {  // at least one full wave of row tiles and a main loop worth sharing B for
    if (!pair_enabled || g.mode == 2 || g.b_batch_z != 0 || row_tiles < 2) return false;
    return pair_forced || (row_tiles >= num_sms() && iters >= 16);
  };

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
    int nt_max = 256;
    if (const char* e = getenv("XVA_GEMM_NTILE")) nt_max = atoi(e) >= 16 ? atoi(e) : 256;
    // Small problems leave SMs idle at 256-wide tiles; narrower tiles add parallelism but re-read A once per n tile;
    // a 257..512-column tile reads A once but its accumulator cannot be double-buffered (the epilogue is exposed).
    // Pick the width that minimises  waves x (k-steps x max(operand fetch, MMA) + exposed epilogue)  in SM clocks:
    // fetch = (16 KiB of A + the B tile, halved when a CTA pair shares it) at ~64 B/clk from L2; MMA = 2 clk per column;
    // exposed epilogue ~ 40 clk per column of a 128-row tile. (Measured, B = 32 x 880: N = 384, K = 3 x 1536 runs 12 %
    // faster as one 384-column tile than as two 192-column tiles; N = 1536 is fastest at 256.)
    static const bool fill_enabled = [] {
      const char* e = getenv("XVA_GEMM_FILL");
      return !(e && e[0] == '0');
    }();
    if (g.mode != 2 && fill_enabled && G == 1 && !getenv("XVA_GEMM_NTILE")) {
      const long iters = static_cast<long>(g.taps) * ceil_div(g.K, kBlockK);
      long best = -1;
      for (int cand = 512; cand >= 64; cand >>= 1) {
        const int tn = ceil_div(g.N, cand);
        const int nt = round_up(ceil_div(g.N, tn), n_gran);
        if (cand == 512 && (nt <= 256 || round_up(nt / 2, g.mode == 0 ? 16 : 64) * 2 > 512)) continue;
        const bool pr = want_pair(iters);
        const int slots = pr ? num_sms() / 2 : num_sms();
        const int units = pr ? ceil_div(row_tiles, 2) : row_tiles;
        const long fetch = 256 + (pr ? nt : 2 * nt), mma = 2L * nt;
        const long cost = static_cast<long>(ceil_div(units * tn, slots)) *
                          (iters * (fetch > mma ? fetch : mma) + (nt > 256 ? 40L * nt : 0));
        if (best < 0 || cost < best) {
          best = cost;
          nt_max = cand;
        }
      }
    } else if (g.mode == 2 && G == 1 && g.N > 256 && g.N <= 512 && !getenv("XVA_GEMM_NTILE")) {
      nt_max = 512;  // weight gradient with 257..512 columns: one tile (the dy operand is read once)
    }
    p.tiles_n = ceil_div(p.N, nt_max);
    p.n_tile = round_up(ceil_div(p.N, p.tiles_n), n_gran);
    if (G > 1 && g.mode != 2) {  // one n tile per group
      p.tiles_n = G;
      p.n_tile = g.N / G;
    }
  }
  p.n_mma = ceil_div(p.n_tile, 256);
  p.n_sub = p.n_tile / p.n_mma;
  if (p.n_sub % n_gran != 0) {  // e.g. 400 -> 2 x 200: round the tile up so both halves are legal
    p.n_sub = round_up(p.n_sub, n_gran);
    p.n_tile = p.n_sub * p.n_mma;
  }
  XVA_CHECK_ARG(p.n_tile <= 512 && p.n_sub <= 256 && p.n_sub % 16 == 0, "gemm: bad n tiling %d/%d", p.n_tile, p.n_sub);
  p.acc_stages = (p.n_tile <= 256) ? 2 : 1;

  // (half-tiles of a pair must be whole swizzle atoms / 32-column chunks)
  int dbg_stages = 0;
  const bool pair = g.mode != 2 &&
                    want_pair(static_cast<long>(g.taps) * ceil_div(g.K, kBlockK)) &&
                    (g.mode == 0 ? (p.n_sub % 16 == 0) : (p.n_sub % 64 == 0));
  const int cg = pair ? 2 : 1;
  p.row_tiles = row_tiles;
  if (const char* e = getenv("XVA_GEMM_DBG")) p.dbg = atoi(e);
  if (const char* e = getenv("XVA_GEMM_STAGES")) {
    if (atoi(e) >= 2) dbg_stages = atoi(e);
  }

  static const bool tap_inner_enabled = [] {
    const char* e = getenv("XVA_GEMM_TAP_INNER");
    return !(e && e[0] == '0');
  }();
  p.tap_inner = tap_inner_enabled ? 1 : 0;

  // ---- multi-tap weight-gradient tiles (see GemmDev::tp)
  p.tp = 1;
  p.tgroups = g.taps;
  static const bool mtap_enabled = [] {
    const char* e = getenv("XVA_GEMM_MTAP");
    return !(e && e[0] == '0');
  }();
  if (mtap_enabled && g.mode == 2 && G == 1 && g.taps > 1 && p.tiles_n == 1 && p.n_tile <= 128) {
    int tp = kTmemCols / p.n_tile;
    const int smem_tp = (kSmemBudget / 2 - kATileBytes) / (p.n_tile * kBlockK * 4);  // at least two stages
    if (tp > smem_tp) tp = smem_tp;
    if (tp > g.taps) tp = g.taps;
    // spread the taps evenly over the groups (11 taps, 8 per tile -> 6 + 5)
    const int groups = ceil_div(g.taps, tp);
    tp = ceil_div(g.taps, groups);
    if (tp > 1) {
      p.tp = tp;
      p.tgroups = groups;
      if (tp * p.n_tile > 256) p.acc_stages = 1;
    }
  }
  const int stage_bytes = kATileBytes + p.tp * p.n_tile * kBlockK * 4 / cg;
  p.stages = kSmemBudget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (dbg_stages && dbg_stages < p.stages) p.stages = dbg_stages;
  XVA_CHECK_ARG(p.stages >= 2, "gemm: tile too large for shared memory");
  p.ring_bytes = p.stages * stage_bytes;
  p.a_rows = kBlockM;
  p.stages_a = 0;

  // ---- halo mode: a k-tap convolution on un-segmented tiles whose taps all read the same columns fetches its activation
  // tile once per k-block (with halo rows) instead of once per tap -- the fp32 operand stream of these launches is
  // L2 -> SM bandwidth bound, and the activation tile is 1/2 (3 taps, 256 columns) to 9/10 (11 taps, 32 columns) of it.
  // Measured on B200 (same box, back to back): FastPitch step 13.76 ms without, 14.33-14.47 ms with (either halo
  // alignment); HiFi-GAN generator at 880 frames 95 -> 123 ms. The launches are bound by the latency x bytes-in-flight
  // of the operand stream rather than by the bytes themselves, and one large activation box per k-block pipelines worse
  // than three small ones. Kept behind XVA_GEMM_HALO (flag) / XVA_GEMM_HALO=1 (environment) with its tests.
  static const bool halo_enabled = [] {
    const char* e = getenv("XVA_GEMM_HALO");
    return e && e[0] == '1';
  }();
  if ((halo_enabled || (g.flags & GEMM_HALO)) && g.mode != 2 && g.taps > 1 && G == 1 && p.seg == kBlockM && g.b_batch_z == 0 && p.dbg == 0) {
    int h = 0;
    bool same_cols = true;
    for (int j = 0; j < g.taps; ++j) {
      const int a = g.shift[j] < 0 ? -g.shift[j] : g.shift[j];
      h = a > h ? a : h;
      same_cols = same_cols && g.a_col[j] == g.a_col[0];
    }
    int hp_align = 4;
    if (const char* e = getenv("XVA_GEMM_HALO_ALIGN")) hp_align = atoi(e) >= 4 ? atoi(e) : 4;
    const int hp = round_up(h, hp_align);
    const int a_rows = kBlockM + 2 * hp;
    const int b_tile = p.n_tile * kBlockK * 4 / cg;
    const int stages_a = 3;
    const int a_ring = stages_a * a_rows * kBlockK * 4;
    int stages_b = (kSmemBudget - a_ring) / b_tile;
    if (stages_b > kMaxStages) stages_b = kMaxStages;
    if (same_cols && a_rows <= 256 && stages_b >= 3) {
      p.halo = 1;
      p.hp = hp;
      p.a_rows = a_rows;
      p.stages_a = stages_a;
      p.stages = stages_b;
      p.ring_bytes = a_ring + stages_b * b_tile;
    }
  }
  const int smem_bytes = p.ring_bytes + kTailSmemBytes;

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
    p.rsplit = 1;
    p.chunk_rows = 0;
    p.num_tiles = (pair ? ceil_div(row_tiles, 2) : row_tiles) * p.tiles_n;
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
    // A caller asking for more CTAs per output tile than there are items gets the contraction rows of every item cut
    // into chunks too (few long sequences: the HiFi-GAN generator has 16 items x 8192..225 280 rows)
    int split = g.split < 1 ? 1 : g.split;
    if (p.tp > 1 && (g.flags & GEMM_ATOMIC)) {  // fewer, fatter tiles: keep every SM busy by splitting the rows further
      const int want = ceil_div(num_sms(), p.tgroups * p.tiles_m);
      if (split < want) split = want;
    }
    p.rsplit = 1;
    p.chunk_rows = round_up(g.R, kBlockK);
    if (split > g.ZR) {
      const int want = ceil_div(split, g.ZR);
      p.chunk_rows = round_up(ceil_div(g.R, want), kBlockK);
      if (p.chunk_rows < 8 * kBlockK) p.chunk_rows = 8 * kBlockK;  // at least 8 k-steps per unit
      p.rsplit = ceil_div(g.R, p.chunk_rows);
    }
    p.k_chunks = p.chunk_rows / kBlockK;
    p.ZR = g.ZR * p.rsplit;
    if (split > p.ZR) split = p.ZR;
    p.zper = ceil_div(p.ZR, split);
    p.split = ceil_div(p.ZR, p.zper);
    XVA_CHECK_ARG(p.split == 1 || (g.flags & GEMM_ATOMIC), "gemm: split > 1 needs GEMM_ATOMIC");
    p.num_tiles = (g.Z / g.ZR) * p.split * p.tgroups * p.tiles_m * p.tiles_n;
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
    int a_cols = g.K;  // with per-tap column offsets each tap reads columns [a_col, a_col + K) of a wider row
    for (int j = 0; j < g.taps; ++j) a_cols = g.a_col[j] + g.K > a_cols ? g.a_col[j] + g.K : a_cols;
    a_cols += (G - 1) * p.grp_a;
    // (a ragged last k-block then reads real neighbouring columns of A; they meet zero-filled rows of B)
    uint64_t dims[3] = {(uint64_t)a_cols, (uint64_t)a_rows, (uint64_t)g.Z};
    uint64_t str[3] = {1, (uint64_t)g.a_rs, (uint64_t)g.a_zs};
    uint32_t box[3] = {kBlockK, (uint32_t)(p.halo ? p.a_rows : p.seg), 1};
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
    uint32_t box[3] = {kBlockK, (uint32_t)(p.n_sub / cg), 1};
    if (g.b_nz == 1 || str[2] == 0) str[2] = (uint64_t)g.b_rs * bn_rows;
    if ((rc = encode_map(&map_b, g.b, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != XVA_OK) return rc;
  } else if (g.mode == 1) {
    XVA_CHECK_ARG(g.N % 32 == 0 || g.b_rs >= round_up(g.N, 32),
                  "gemm: MN-major B with N=%d needs N %% 32 == 0 or a row stride >= %d (got %lld)", g.N,
                  round_up(g.N, 32), (long long)g.b_rs);
    const int bk_rows = g.b_rows ? g.b_rows : g.K * G;
    uint64_t dims[4] = {32, (uint64_t)bk_rows, (uint64_t)ceil_div(G > 1 ? g.N / G : g.N, 32), (uint64_t)g.b_nz};
    uint64_t str[4] = {1, (uint64_t)g.b_rs, 32, (uint64_t)g.b_zs};
    uint32_t box[4] = {32, kBlockK, (uint32_t)(pair ? p.n_sub / 64 : p.n_tile / 32), 1};
    if (g.b_nz == 1 || str[3] == 0) str[3] = (uint64_t)g.b_rs * bk_rows;
    if ((rc = encode_map(&map_b, g.b, 4, dims, str, box, mn_swizzle)) != XVA_OK) return rc;
  } else {
    const int b_rows = g.b_rows ? g.b_rows : g.R;
    int b_cols = g.N;  // mode 2: a_col[j] offsets the columns of B per tap (strided convolution views); 32-aligned
    for (int j = 0; j < g.taps; ++j) {
      XVA_CHECK_ARG(g.a_col[j] % 32 == 0, "gemm: wgrad column offset %d of tap %d is not a multiple of 32", g.a_col[j], j);
      b_cols = g.a_col[j] + g.N > b_cols ? g.a_col[j] + g.N : b_cols;
    }
    b_cols += (G - 1) * p.grp_a;
    uint64_t dims[4] = {32, (uint64_t)b_rows, (uint64_t)ceil_div(b_cols, 32), (uint64_t)g.Z};
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
  p.act_slope = g.act_slope;
  p.out_act_slope = g.out_act_slope;
  p.out_act = g.out_act;
  for (int j = 0; j < g.taps; ++j) p.a_col[j] = g.a_col[j];
  p.lens = g.lens;
  p.gamma = g.gamma;
  p.beta = g.beta;
  p.ln_eps = g.ln_eps;
  p.out_pre = g.out_pre;
  p.ln_mean = g.ln_mean;
  p.ln_rstd = g.ln_rstd;
  p.seed = g.seed;
  p.seed_dev = g.seed_dev;
  p.rowvec = g.rowvec;
  p.drop_ld = g.drop_ld > 0 ? g.drop_ld : g.N;
  if (g.flags & GEMM_SOFTMAX_BWD)
    XVA_CHECK_ARG(g.mode != 2 && g.gate && g.rowvec && !(g.flags & GEMM_LN), "gemm: SOFTMAX_BWD needs gate (P), rowvec, mode 0/1");
  if ((g.flags & (GEMM_DROP_PRE | GEMM_DROP_POST | GEMM_SOFTMAX_BWD)) && g.drop_p > 0.0f) {
    XVA_CHECK_ARG(g.drop_p < 1.0f, "gemm: dropout p=%f", g.drop_p);
    p.drop_thresh = static_cast<uint32_t>(static_cast<double>(g.drop_p) * 4294967296.0);
    p.inv_keep = 1.0f / (1.0f - g.drop_p);
  } else {
    p.flags &= ~(GEMM_DROP_PRE | GEMM_DROP_POST);
    p.drop_thresh = 0;
    p.inv_keep = 1.0f;
  }

  {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    auto m4 = [](long v) { return (v & 3) == 0; };
    bool ok = al(g.out) && m4(g.o_rs) && m4(g.o_zs) && m4(g.o_js);
    if (g.bias) ok = ok && al(g.bias);
    if (g.residual) ok = ok && al(g.residual) && m4(g.r_rs) && m4(g.r_zs);
    if (g.gate) ok = ok && al(g.gate) && m4(g.g_rs) && m4(g.g_zs);
    if (g.out_pre) ok = ok && al(g.out_pre);
    if (g.out_act) ok = ok && al(g.out_act);
    if (g.flags & GEMM_LN) ok = ok && al(g.gamma) && al(g.beta);
    p.vec_ok = ok ? 1 : 0;
    p.round_on = g_round_host;
  }

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    const int mx = kSmemBudget + kTailSmemBytes;
    const void* fns[] = {(const void*)gemm_tc_kernel<EPI_WGRAD, 1>, (const void*)gemm_tc_kernel<EPI_PLAIN, 1>,
                         (const void*)gemm_tc_kernel<EPI_FULL, 1>,  (const void*)gemm_tc_kernel<EPI_LN, 1>,
                         (const void*)gemm_tc_kernel<EPI_PLAIN, 2>, (const void*)gemm_tc_kernel<EPI_FULL, 2>,
                         (const void*)gemm_tc_kernel<EPI_LN, 2>};
    for (const void* f : fns)
      if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  });
  XVA_CHECK_CUDA(attr_err);

  const int epi = (g.mode == 2) ? EPI_WGRAD
                  : (p.flags & GEMM_LN) ? EPI_LN
                  : (p.gate || p.residual || (p.flags & GEMM_DROP_PRE)) ? EPI_FULL : EPI_PLAIN;
  // measured: no gain under CUDA-graph replay (13.87 vs 13.91 ms/step), so opt-in
  static const bool pdl_enabled = [] {
    const char* e = getenv("XVA_GEMM_PDL");
    return e && e[0] == '1';
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (pair) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (pdl_enabled) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  if (!pair) {
    cfg.gridDim = dim3(p.num_tiles < num_sms() ? p.num_tiles : num_sms());
    switch (epi) {
      case EPI_WGRAD: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_WGRAD, 1>, map_a, map_b, p)); break;
      case EPI_LN: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_LN, 1>, map_a, map_b, p)); break;
      case EPI_FULL: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_FULL, 1>, map_a, map_b, p)); break;
      default: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_PLAIN, 1>, map_a, map_b, p)); break;
    }
  } else {
    const int pairs = num_sms() / 2;
    cfg.gridDim = dim3(2 * (p.num_tiles < pairs ? p.num_tiles : pairs));
    switch (epi) {
      case EPI_LN: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_LN, 2>, map_a, map_b, p)); break;
      case EPI_FULL: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_FULL, 2>, map_a, map_b, p)); break;
      default: XVA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI_PLAIN, 2>, map_a, map_b, p)); break;
    }
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int set_operand_rounding_gemm_tc(int on) {
  XVA_CHECK_CUDA(cudaMemcpyToSymbol(g_xva_round_operands, &on, sizeof(int)));
  g_round_host = on;
  return XVA_OK;
}

}  // namespace xva
