// extern "C" surface of libxva_b200.so (declared in include/xva_b200.h). Thin: argument checks live in the launchers.
#include "../../include/xva_b200.h"

#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"

using namespace xva;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int xva_abi_version(void) { return XVA_ABI_VERSION; }

const char* xva_last_error(void) { return last_error(); }

int xva_device_check(int device) {
  cudaDeviceProp prop;
  XVA_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d (%s) is compute capability %d.%d; libxva_b200 needs 10.x (sm_100a)", device, prop.name,
              prop.major, prop.minor);
    return XVA_ERR_DEVICE;
  }
  return XVA_OK;
}

int xva_sizeof_gemm_args(void) { return static_cast<int>(sizeof(xva_gemm_args)); }

int xva_gemm(const xva_gemm_args* args, void* stream) {
  XVA_CHECK_ARG(args != nullptr, "xva_gemm: null args");
  return gemm_tc_launch(*args, S(stream));
}

int xva_gemm_ref(const xva_gemm_args* args, void* stream) {
  XVA_CHECK_ARG(args != nullptr, "xva_gemm_ref: null args");
  return gemm_ref_launch(*args, S(stream));
}

int xva_regulate_len_scan(const float* durs, int B, int Tt, float pace, int mel_max_len, int32_t* cum,
                          int32_t* dec_lens, void* stream) {
  return duration_scan(durs, B, Tt, pace, mel_max_len, cum, dec_lens, S(stream));
}

int xva_regulate_len_fwd(const float* enc, const int32_t* cum, int B, int Tt, int C, int T_out, float* out,
                         int32_t* idx, void* stream) {
  return regulate_gather(enc, cum, B, Tt, C, T_out, out, idx, S(stream));
}

int xva_regulate_len_bwd(const float* dout, const int32_t* cum, int B, int Tt, int C, int T_out, float* denc,
                         int accumulate, void* stream) {
  return regulate_scatter(dout, cum, B, Tt, C, T_out, denc, accumulate, S(stream));
}

int xva_average_pitch(const float* pitch, const float* durs, int B, int F, int Tm, int Tt, float* out, void* stream) {
  return average_pitch(pitch, durs, B, F, Tm, Tt, out, S(stream));
}

}  // extern "C"
