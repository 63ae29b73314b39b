// fp32 SIMT implementation of the tap-GEMM contract of gemm.cuh: one thread per output element (mode 0/1:
// one 128-thread block per output row, so the LayerNorm epilogue can reduce in shared memory).
// Test infrastructure for the tcgen05 kernel: same arguments, same epilogue semantics, exact fp32 products.
// Not on the product path.
#include "gemm.cuh"

namespace xva {

namespace {

struct RefDev {
  GemmArgs g;
  uint32_t drop_thresh;
  float inv_keep;
};

__device__ __forceinline__ uint64_t eff_seed(const GemmArgs& g) {
  return g.seed + (g.seed_dev ? *g.seed_dev * 0xA24BAED4963EE407ull : 0ull);
}

__device__ float ref_dot(const GemmArgs& g, int z, int r, int n) {
  float acc = 0.0f;
  const int a_rows = g.a_rows ? g.a_rows : g.R;
  // grouped launch (xva_gemm_args.groups): column n belongs to group grp
  const int G = g.groups > 1 ? g.groups : 1;
  const int n_per = g.N / G;
  const int grp = G > 1 ? n / n_per : 0;
  const int a_goff = G > 1 ? grp * (g.mode == 0 ? g.grp_step : g.K) : 0;
  for (int j = 0; j < g.taps; ++j) {
    const int ar = r + g.shift[j];
    if (ar < 0 || ar >= a_rows) continue;
    const float* arow = g.a + z * g.a_zs + static_cast<long>(ar) * g.a_rs + g.a_col[j] + a_goff;
    const int zb = j * g.b_tap_z + z * g.b_batch_z;
    if (g.mode == 0) {
      if (g.b_rows && n >= g.b_rows) continue;
      const float* brow = g.b + zb * g.b_zs + static_cast<long>(n) * g.b_rs;
      for (int k = 0; k < g.K; ++k) acc = fmaf(arow[k], brow[k], acc);
    } else if (G > 1) {
      const float* bcol = g.b + zb * g.b_zs + static_cast<long>(grp) * g.K * g.b_rs + (n - grp * n_per);
      for (int k = 0; k < g.K; ++k) acc = fmaf(arow[k], bcol[static_cast<long>(k) * g.b_rs], acc);
    } else {
      const float* bcol = g.b + zb * g.b_zs + n;
      const int kmax = (g.b_rows && g.b_rows < g.K) ? g.b_rows : g.K;
      for (int k = 0; k < kmax; ++k) acc = fmaf(arow[k], bcol[static_cast<long>(k) * g.b_rs], acc);
    }
  }
  return acc;
}

__global__ void gemm_ref_fwd_kernel(const RefDev d) {
  const GemmArgs& g = d.g;
  const int r = blockIdx.x, z = blockIdx.y;
  __shared__ float red[128];
  __shared__ float s_mean, s_rstd;
  const long rowoff_o = z * g.o_zs + static_cast<long>(r) * g.o_rs;
  const long rowoff_r = z * g.r_zs + static_cast<long>(r) * g.r_rs;
  const long rowoff_g = z * g.g_zs + static_cast<long>(r) * g.g_rs;
  const float keep_row = (g.lens == nullptr || r < g.lens[z]) ? 1.0f : 0.0f;
  const uint64_t drop_row = (static_cast<uint64_t>(z) * g.R + r) * static_cast<uint64_t>(g.N);
  const bool ln = g.flags & GEMM_LN;

  auto pre_value = [&](int n) -> float {
    if (g.flags & GEMM_SOFTMAX_BWD) {
      const int ld = g.drop_ld > 0 ? g.drop_ld : g.N;
      const uint64_t di = (static_cast<uint64_t>(z) * g.R + r) * static_cast<uint64_t>(ld) + n;
      const float dp = ref_dot(g, z, r, n) * dropout_scale(eff_seed(g), di, d.drop_thresh, d.inv_keep);
      return g.alpha * g.gate[rowoff_g + n] * (dp - g.rowvec[static_cast<long>(z) * g.R + r]);
    }
    float v = g.alpha * ref_dot(g, z, r, n);
    if (g.bias) v += g.bias[n];
    if (g.flags & GEMM_RELU) v = v > 0.0f ? v : g.act_slope * v;
    if (g.gate) v *= (g.gate[rowoff_g + n] > 0.0f) ? 1.0f : g.gate_slope;
    if (g.flags & GEMM_DROP_PRE) v *= dropout_scale(eff_seed(g), drop_row + n, d.drop_thresh, d.inv_keep);
    if (g.residual) v += g.residual[rowoff_r + n];
    return v;
  };

  if (!ln) {
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
      float v = pre_value(n);
      if (g.flags & GEMM_TANH) v = tanhf(v);
      v *= keep_row;
      if (g.out_act) g.out_act[rowoff_o + n] = tf32_rn(v > 0.0f ? v : g.out_act_slope * v);
      g.out[rowoff_o + n] = (g.flags & GEMM_ROUND_OUT) ? tf32_rn(v) : v;
    }
    return;
  }
  // LayerNorm: N <= 512, each thread owns up to 4 columns.
  float vals[4];
  float sum = 0.0f;
  int cnt = 0;
  for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
    vals[cnt] = pre_value(n);
    sum += vals[cnt++];
  }
  red[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < blockDim.x; ++i) s += red[i];
    s_mean = s / g.N;
  }
  __syncthreads();
  const float mean = s_mean;
  float sq = 0.0f;
  for (int i = 0; i < cnt; ++i) sq += (vals[i] - mean) * (vals[i] - mean);
  red[threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < blockDim.x; ++i) s += red[i];
    s_rstd = rsqrtf(s / g.N + g.ln_eps);
    if (g.ln_mean) g.ln_mean[static_cast<long>(z) * g.R + r] = mean;
    if (g.ln_rstd) g.ln_rstd[static_cast<long>(z) * g.R + r] = s_rstd;
  }
  __syncthreads();
  const float rstd = s_rstd;
  cnt = 0;
  for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
    const float x = vals[cnt++];
    float y = (x - mean) * rstd * g.gamma[n] + g.beta[n];
    if (g.flags & GEMM_DROP_POST) y *= dropout_scale(eff_seed(g), drop_row + n, d.drop_thresh, d.inv_keep);
    g.out[rowoff_o + n] = (g.flags & GEMM_ROUND_OUT) ? tf32_rn(y * keep_row) : y * keep_row;
    if (g.out_pre) g.out_pre[rowoff_o + n] = x;
  }
}

// mode 2: grid (N/… , M, ZO*taps); one thread per (m, n).
__global__ void gemm_ref_wgrad_kernel(const RefDev d) {
  const GemmArgs& g = d.g;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  const int j = blockIdx.z % g.taps;
  const int zo = blockIdx.z / g.taps;
  if (n >= g.N) return;
  const int b_rows = g.b_rows ? g.b_rows : g.R;
  const int a_rows = g.a_rows ? g.a_rows : g.R;
  float acc = 0.0f;
  const int goff = g.groups > 1 ? (m / (g.M / g.groups)) * g.grp_step : 0;
  for (int zr = 0; zr < g.ZR; ++zr) {
    const int z = zo * g.ZR + zr;
    for (int t = 0; t < g.R && t < a_rows; ++t) {
      const int bt = t + g.shift[j];
      if (bt < 0 || bt >= b_rows) continue;
      acc = fmaf(g.a[z * g.a_zs + static_cast<long>(t) * g.a_rs + m],
                 g.b[z * g.b_zs + static_cast<long>(bt) * g.b_rs + g.a_col[j] + goff + n], acc);
    }
  }
  float* o = g.out + zo * g.o_zs + j * g.o_js + static_cast<long>(m) * g.o_rs + n;
  if (g.flags & GEMM_ATOMIC) *o += g.alpha * acc;  // each element has exactly one writer here
  else *o = (g.flags & GEMM_ROUND_OUT) ? tf32_rn(g.alpha * acc) : g.alpha * acc;
}

}  // namespace

int gemm_ref_launch(const GemmArgs& g, cudaStream_t stream) {
  XVA_CHECK_ARG(g.mode >= 0 && g.mode <= 2, "gemm_ref: bad mode %d", g.mode);
  XVA_CHECK_ARG(g.taps >= 1 && g.taps <= kMaxTaps, "gemm_ref: taps %d", g.taps);
  RefDev d;
  d.g = g;
  if ((g.flags & (GEMM_DROP_PRE | GEMM_DROP_POST | GEMM_SOFTMAX_BWD)) && g.drop_p > 0.0f) {
    d.drop_thresh = static_cast<uint32_t>(static_cast<double>(g.drop_p) * 4294967296.0);
    d.inv_keep = 1.0f / (1.0f - g.drop_p);
  } else {
    d.g.flags &= ~(GEMM_DROP_PRE | GEMM_DROP_POST);
    d.drop_thresh = 0;
    d.inv_keep = 1.0f;
  }
  if (g.mode != 2) {
    XVA_CHECK_ARG(!(g.flags & GEMM_LN) || g.N <= 512, "gemm_ref: LN needs N <= 512");
    dim3 grid(g.R, g.Z);
    gemm_ref_fwd_kernel<<<grid, 128, 0, stream>>>(d);
  } else {
    dim3 grid(ceil_div(g.N, 128), g.M, (g.Z / g.ZR) * g.taps);
    gemm_ref_wgrad_kernel<<<grid, 128, 0, stream>>>(d);
  }
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(gemm_ref)

}  // namespace xva
