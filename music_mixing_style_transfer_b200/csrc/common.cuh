// common.cuh -- shared host/device helpers for libmst_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/mst_b200.h"

namespace mst {

// thread-local error slot behind mst_last_error()
char* error_buffer();
int fail(const char* fmt, ...);

#define MST_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::mst::fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define MST_CHECK(cond, ...)                    \
  do {                                          \
    if (!(cond)) return ::mst::fail(__VA_ARGS__); \
  } while (0)

inline int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int sm_count();  // cached multiProcessorCount of the current device

// ---- driver entry point for TMA descriptors (no -lcuda link: resolved through the runtime) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled tensor_map_encoder();  // nullptr (with last_error set) if the driver lacks it

// ---- encoder tensor-core path (enc_umma.cu), used by encoder.cu ----
bool enc_umma_eligible(int c_in, int c_out, int k, int stride);
size_t enc_umma_weight_bytes(int c_in, int c_out, int k);
int enc_umma_pack(const float* w, const float* b, const float* bn_w, const float* bn_b, const float* bn_mean,
                  const float* bn_var, int c_out, int c_in, int k, void* w_out, float* b_out, cudaStream_t st);
int enc_umma_conv(const void* x, const void* w_packed, const float* bias, void* y, int B, int c_in, int t_in, int c_out,
                  int k, int stride, bool residual, int out_l, int out_r, cudaStream_t st);
int enc_split_from_f32(const float* x, void* y, int B, int C, int T, int halo_l, int halo_r, cudaStream_t st);
int enc_pool_split(const void* x, float* emb, int B, int C, int T, int halo_l, int t_pad, cudaStream_t st);

// rows between two segments of a TCN activation buffer: the segment length rounded up to the 256-row work item, so that the
// interleaved (5-D) tensor maps of the small dilations exist for every length; rows [T, tcn_seg_rows(T)) are kept at zero
static inline int tcn_seg_rows(int T) { return (T + 255) / 256 * 256; }

// ---- TCN "f16 + 2 x e4m3" operand format (tcn_f8.cu): packer and converters used by tcn.cu ----
size_t tcn_f8_weight_bytes();
int tcn_f8_pack_layer(const float* conv_w, const float* bn_w, const float* bn_var, void* w_out, float* inv_scale,
                      unsigned int* scratch, cudaStream_t st);
int tcn_f8_act_pack(const float* x, void* act, int B, int T, cudaStream_t st);
int tcn_f8_act_unpack(const void* act, float* y, int B, int T, cudaStream_t st);

}  // namespace mst
