// Masked-MSE losses of FastPitchLoss (fastpitch/loss_function.py:81-136) and the LAMB optimizer of the reference
// (lamb.py:40-106) as multi-tensor kernels over one flat parameter arena.
//
// Reductions accumulate in fp64 (block partial -> one atomicAdd(double) per block) so the scalar losses and the
// LAMB trust ratios do not depend on the reduction order beyond ~1e-16 relative.
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
  return t;  // valid in thread 0
}

// mel loss: pred [B,T_out,C] (rows t >= T_out count as 0: F.pad at loss_function.py:103), tgt [B,C,Tm],
// mask = tgt != 0.  acc[0] += sum mask*(pred-tgt)^2, acc[1] += sum mask.
__global__ void __launch_bounds__(256)
mel_mse_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int T_out, int Tm, int C,
                      long total, double* __restrict__ acc) {
  __shared__ double sh[8];
  double se = 0.0, cnt = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long bt = i / C;
    const int t = static_cast<int>(bt % Tm), b = static_cast<int>(bt / Tm);
    const float y = tgt[(static_cast<long>(b) * C + c) * Tm + t];
    if (y != 0.0f) {
      const float p = t < T_out ? pred[(static_cast<long>(b) * T_out + t) * C + c] : 0.0f;
      const float d = p - y;
      se += static_cast<double>(d * d);
      cnt += 1.0;
    }
  }
  se = block_sum_d(se, sh);
  cnt = block_sum_d(cnt, sh);
  if (threadIdx.x == 0) {
    atomicAdd(acc, se);
    atomicAdd(acc + 1, cnt);
  }
}

// dpred[b,t,c] = scale * 2*(pred-tgt)*mask / count, written with row stride ldd (columns C..ldd are zeroed so the
// buffer can feed the MN-major wgrad, whose channel count must be a multiple of 32).
__global__ void __launch_bounds__(256)
mel_mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int T_out, int Tm, int C, int ldd,
                    long total, const double* __restrict__ acc, float scale, float* __restrict__ dpred) {
  const float k = static_cast<float>(2.0 * scale / acc[1]);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ldd);
    const long bt = i / ldd;
    const int t = static_cast<int>(bt % T_out), b = static_cast<int>(bt / T_out);
    float g = 0.0f;
    if (c < C) {
      const float y = tgt[(static_cast<long>(b) * C + c) * Tm + t];
      if (y != 0.0f) g = tf32_rn(k * (pred[(static_cast<long>(b) * T_out + t) * C + c] - y));  // operand of proj dgrad/wgrad
    }
    dpred[i] = g;
  }
}

// token-level losses (pitch, energy, log-duration): pred/tgt [B,T], mask = t < lens[b]
__global__ void __launch_bounds__(256)
lens_mse_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int* __restrict__ lens,
                       int T, long total, int log1p_tgt, double* __restrict__ acc) {
  __shared__ double sh[8];
  double se = 0.0, cnt = 0.0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % T), b = static_cast<int>(i / T);
    if (t < lens[b]) {
      const float y = log1p_tgt ? logf(tgt[i] + 1.0f) : tgt[i];
      const float d = pred[i] - y;
      se += static_cast<double>(d * d);
      cnt += 1.0;
    }
  }
  se = block_sum_d(se, sh);
  cnt = block_sum_d(cnt, sh);
  if (threadIdx.x == 0) {
    atomicAdd(acc, se);
    atomicAdd(acc + 1, cnt);
  }
}

__global__ void __launch_bounds__(256)
lens_mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, const int* __restrict__ lens, int T,
                     long total, int log1p_tgt, const double* __restrict__ acc, float scale, float* __restrict__ dpred) {
  const float k = static_cast<float>(2.0 * scale / acc[1]);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % T), b = static_cast<int>(i / T);
    float g = 0.0f;
    if (t < lens[b]) g = k * (pred[i] - (log1p_tgt ? logf(tgt[i] + 1.0f) : tgt[i]));
    dpred[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ LAMB
// Chunk table: one block per chunk {start, len, tensor}. All chunks of a tensor share norms[tensor*2 .. +1].
struct LambChunk {
  long long start;
  int len;
  int tensor;
};

// sum of squares of the gradients of the listed chunks (clip_grad_norm_, xva_train.py:857)
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const float* __restrict__ g, const LambChunk* __restrict__ chunks, double* __restrict__ out) {
  __shared__ double sh[8];
  const LambChunk ck = chunks[blockIdx.x];
  const float* gp = g + ck.start;
  double s = 0.0;
  for (int i = threadIdx.x; i < ck.len; i += blockDim.x) {
    const float v = gp[i];
    s += static_cast<double>(v) * v;
  }
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__device__ __forceinline__ float clip_coef(const double* gnorm_sq, float max_norm) {
  if (gnorm_sq == nullptr || max_norm <= 0.0f) return 1.0f;
  const float c = max_norm / (static_cast<float>(sqrt(*gnorm_sq)) + 1e-6f);
  return c < 1.0f ? c : 1.0f;
}

// stage 1: m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; r = m/(sqrt(v)+eps) + wd*p ;
//          norms[t][0] += sum p^2 ; norms[t][1] += sum r^2          (lamb.py:77-92)
__global__ void __launch_bounds__(256)
lamb_stage1_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                   float* __restrict__ v, const LambChunk* __restrict__ chunks, double* __restrict__ norms,
                   const double* __restrict__ gnorm_sq, float max_norm, float b1, float b2, float eps, float wd) {
  __shared__ double sh[8];
  // A non-finite gradient norm (a NaN / Inf anywhere in the stage's gradients) turns the whole update into a no-op:
  // the skip of xva_train.py:825-832 decided on the device, so weights, moments and the tf32 weight copy never see it
  // and a captured step needs no host round trip before the optimizer.
  if (gnorm_sq != nullptr && !isfinite(*gnorm_sq)) return;
  const LambChunk ck = chunks[blockIdx.x];
  const float coef = clip_coef(gnorm_sq, max_norm);
  double sp = 0.0, sr = 0.0;
  for (int i = threadIdx.x; i < ck.len; i += blockDim.x) {
    const long long k = ck.start + i;
    const float gi = g[k] * coef;
    const float pi = p[k];
    const float mi = b1 * m[k] + (1.0f - b1) * gi;
    const float vi = b2 * v[k] + (1.0f - b2) * gi * gi;
    m[k] = mi;
    v[k] = vi;
    const float r = mi / (sqrtf(vi) + eps) + wd * pi;
    sp += static_cast<double>(pi) * pi;
    sr += static_cast<double>(r) * r;
  }
  sp = block_sum_d(sp, sh);
  sr = block_sum_d(sr, sh);
  if (threadIdx.x == 0) {
    atomicAdd(norms + 2 * ck.tensor, sp);
    atomicAdd(norms + 2 * ck.tensor + 1, sr);
  }
}

// stage 2: trust = clamp(||p||,0,10)/||r|| (1 if either is 0) ; p -= lr*trust*r      (lamb.py:85-104)
__global__ void __launch_bounds__(256)
lamb_stage2_kernel(float* __restrict__ p, const float* __restrict__ m, const float* __restrict__ v,
                   const LambChunk* __restrict__ chunks, const double* __restrict__ norms,
                   const double* __restrict__ gnorm_sq, const float* __restrict__ lr_dev, float eps, float wd,
                   float* __restrict__ p_tf32) {
  if (gnorm_sq != nullptr && !isfinite(*gnorm_sq)) return;  // see stage 1
  const LambChunk ck = chunks[blockIdx.x];
  const float wn = fminf(static_cast<float>(sqrt(norms[2 * ck.tensor])), 10.0f);
  const float rn = static_cast<float>(sqrt(norms[2 * ck.tensor + 1]));
  const float trust = (wn == 0.0f || rn == 0.0f) ? 1.0f : wn / rn;
  const float step = lr_dev[0] * trust;
  for (int i = threadIdx.x; i < ck.len; i += blockDim.x) {
    const long long k = ck.start + i;
    const float pi = p[k];
    const float r = m[k] / (sqrtf(v[k]) + eps) + wd * pi;
    const float pn = pi - step * r;
    p[k] = pn;
    if (p_tf32) p_tf32[k] = tf32_rn(pn);
  }
}

inline int grid_for(long total) {
  long b = ceil_div_l(total, 256 * 4);
  const long cap = 8L * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int mel_mse(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, double* acc, cudaStream_t stream) {
  XVA_CHECK_ARG(T_out <= Tm, "mel loss: prediction longer than target (%d > %d)", T_out, Tm);
  const long total = static_cast<long>(B) * Tm * C;
  mel_mse_reduce_kernel<<<grid_for(total), 256, 0, stream>>>(pred, tgt, T_out, Tm, C, total, acc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int mel_mse_grad(const float* pred, const float* tgt, int B, int T_out, int Tm, int C, int ldd, const double* acc,
                 float scale, float* dpred, cudaStream_t stream) {
  XVA_CHECK_ARG(ldd >= C, "mel loss grad: ldd=%d < C=%d", ldd, C);
  const long total = static_cast<long>(B) * T_out * ldd;
  mel_mse_grad_kernel<<<grid_for(total), 256, 0, stream>>>(pred, tgt, T_out, Tm, C, ldd, total, acc, scale, dpred);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int lens_mse(const float* pred, const float* tgt, const int* lens, int B, int T, int log1p_tgt, double* acc,
             cudaStream_t stream) {
  const long total = static_cast<long>(B) * T;
  lens_mse_reduce_kernel<<<grid_for(total), 256, 0, stream>>>(pred, tgt, lens, T, total, log1p_tgt, acc);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int lens_mse_grad(const float* pred, const float* tgt, const int* lens, int B, int T, int log1p_tgt, const double* acc,
                  float scale, float* dpred, cudaStream_t stream) {
  const long total = static_cast<long>(B) * T;
  lens_mse_grad_kernel<<<grid_for(total), 256, 0, stream>>>(pred, tgt, lens, T, total, log1p_tgt, acc, scale, dpred);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int grad_sqnorm(const float* g, const void* chunks, int n_chunks, double* out, cudaStream_t stream) {
  if (n_chunks == 0) return XVA_OK;
  grad_sqnorm_kernel<<<n_chunks, 256, 0, stream>>>(g, static_cast<const LambChunk*>(chunks), out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int lamb_step(float* p, const float* g, float* m, float* v, const void* chunks, int n_chunks, double* norms,
              const double* gnorm_sq, float max_norm, const float* lr_dev, float b1, float b2, float eps, float wd,
              float* p_tf32, cudaStream_t stream) {
  if (n_chunks == 0) return XVA_OK;
  const LambChunk* ck = static_cast<const LambChunk*>(chunks);
  lamb_stage1_kernel<<<n_chunks, 256, 0, stream>>>(p, g, m, v, ck, norms, gnorm_sq, max_norm, b1, b2, eps, wd);
  XVA_CHECK_LAUNCH();
  lamb_stage2_kernel<<<n_chunks, 256, 0, stream>>>(p, m, v, ck, norms, gnorm_sq, lr_dev, eps, wd, p_tf32);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(loss_optim)

}  // namespace xva
