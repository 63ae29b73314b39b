// Element-wise HBM-bound kernels of the HiFi-GAN path (float4 grid-stride; every one is a single read + write pass).
//   mean3_lrelu   x = (y0+y1+y2)/3 ; out = leaky_relu(x)      MRF average + the activation that follows it
//                                                            (Generator.forward, hifigan/models.py:117-124)
//   sum3          out = a + b + c                             gradient of a tensor consumed by the three ResBlocks
//   tanh_bwd      d(pre) = dy * (1 - y^2), written into column 0 of a zero-padded [rows, ld] buffer
//   adamw         torch.optim.AdamW semantics over a flat arena (hifigan/xva_train.py:298-300)
#include "common.cuh"
#include "ops.cuh"

namespace xva {

namespace {

__device__ __forceinline__ float lrelu(float v, float s) { return v > 0.0f ? v : s * v; }

__global__ void __launch_bounds__(256)
mean3_lrelu_kernel(const float4* __restrict__ y0, const float4* __restrict__ y1, const float4* __restrict__ y2, long n4,
                   float slope, float4* __restrict__ out) {
  const float k = 1.0f / 3.0f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 a = y0[i], b = y1[i], c = y2[i];
    // (a + b) + c, then / 3: the reference's accumulation order (xs = r0; xs += r1; xs += r2; x = xs / 3)
    float4 m = make_float4(((a.x + b.x) + c.x) * k, ((a.y + b.y) + c.y) * k, ((a.z + b.z) + c.z) * k, ((a.w + b.w) + c.w) * k);
    out[i] = tf32_rn4(make_float4(lrelu(m.x, slope), lrelu(m.y, slope), lrelu(m.z, slope), lrelu(m.w, slope)));
  }
}

__global__ void __launch_bounds__(256)
sum3_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c, long n4,
            float4* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 x = a[i], y = b[i], z = c[i];
    out[i] = tf32_rn4(make_float4(x.x + y.x + z.x, x.y + y.y + z.y, x.z + y.z + z.z, x.w + y.w + z.w));
  }
}

__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long rows, int ld, float* __restrict__ out) {
  const long total = rows * ld;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / ld;
    const int c = static_cast<int>(i - r * ld);
    float v = 0.0f;
    if (c == 0) {
      const float t = y[r];
      v = tf32_rn(dy[r] * (1.0f - t * t));
    }
    out[i] = v;
  }
}

// p = p*(1 - lr*wd) - lr * mhat / (sqrt(vhat) + eps), mhat = m/(1-b1^t), vhat = v/(1-b2^t)    (torch AdamW)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
             const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd, float bc1, float bc2,
             const unsigned long long* __restrict__ step_dev) {
  const float lr = lr_dev[0];
  if (step_dev) {  // step count kept on the device (CUDA-graph replay): bias corrections computed here
    const float t = static_cast<float>(*step_dev);
    bc1 = 1.0f - powf(b1, t);
    bc2 = 1.0f - powf(b2, t);
  }
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = p[i] * (1.0f - lr * wd) - (lr / bc1) * (mi / denom);
  }
}

inline int grid_for(long n) {
  long b = ceil_div_l(n, 256 * 4);
  const long cap = 16L * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int mean3_lrelu(const float* y0, const float* y1, const float* y2, long n, float slope, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(n % 4 == 0, "mean3_lrelu: n=%ld must be a multiple of 4", n);
  if (n == 0) return XVA_OK;
  mean3_lrelu_kernel<<<grid_for(n / 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(y0), reinterpret_cast<const float4*>(y1),
                                                         reinterpret_cast<const float4*>(y2), n / 4, slope,
                                                         reinterpret_cast<float4*>(out));
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int sum3(const float* a, const float* b, const float* c, long n, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(n % 4 == 0, "sum3: n=%ld must be a multiple of 4", n);
  if (n == 0) return XVA_OK;
  sum3_kernel<<<grid_for(n / 4), 256, 0, stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                  reinterpret_cast<const float4*>(c), n / 4, reinterpret_cast<float4*>(out));
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int tanh_bwd(const float* dy, const float* y, long rows, int ld, float* out, cudaStream_t stream) {
  XVA_CHECK_ARG(ld >= 1, "tanh_bwd: ld=%d", ld);
  if (rows == 0) return XVA_OK;
  tanh_bwd_kernel<<<grid_for(rows * ld), 256, 0, stream>>>(dy, y, rows, ld, out);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

int adamw_step(float* p, const float* g, float* m, float* v, long n, const float* lr_dev, float b1, float b2, float eps,
               float wd, int step, const unsigned long long* step_dev, cudaStream_t stream) {
  XVA_CHECK_ARG(step >= 1 || step_dev, "adamw: step=%d (1-based)", step);
  if (step < 1) step = 1;
  if (n == 0) return XVA_OK;
  const float bc1 = 1.0f - powf(b1, static_cast<float>(step)), bc2 = 1.0f - powf(b2, static_cast<float>(step));
  adamw_kernel<<<grid_for(n), 256, 0, stream>>>(p, g, m, v, n, lr_dev, b1, b2, eps, wd, bc1, bc2, step_dev);
  XVA_CHECK_LAUNCH();
  return XVA_OK;
}

XVA_DEFINE_ROUNDING_SWITCH(elemwise)

}  // namespace xva
