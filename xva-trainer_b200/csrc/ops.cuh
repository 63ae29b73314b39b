// Launchers of the non-GEMM kernels (HBM-bound row / index / reduction work). Definitions in the .cu files named
// next to each group; the C ABI in capi.cu forwards to these.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xva {

// ---- regulate.cu
int duration_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int* cum, int* dec_lens,
                  cudaStream_t stream);
int regulate_gather(const float* enc, const int* cum, int B, int Tt, int C, int T_out, float* out, int* idx_out,
                    cudaStream_t stream);
int regulate_scatter(const float* dout, const int* cum, int B, int Tt, int C, int T_out, float* denc, int accumulate,
                     cudaStream_t stream);
int average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, cudaStream_t stream);

}  // namespace xva
