// tcn_f8.cu -- the "f16f8" operand format of the MixFXcloner TCN (MST_TCN_F16F8, the default precision): weight packer,
// block 0 (CUDA cores, writes the format) and the fp32 <-> format converters.  The tcgen05 kernel that consumes the
// format is tcn_block_umma_kernel<.., FMT = 1> in tcn.cu.
//
// Same computation as the bf16 x 3 split (TCNBlock.forward, architectures.py:222-234), different operand split: bf16 x 3
// spends 3 bf16 MMAs per algorithmic MMA (Xhi*Whi + Xlo*Whi + Xhi*Wlo).  Here
//     X*W*(S*2^11)  =  fp16(X) * fp16(W*S*2^11)                       kind::f16     (1 unit)
//                    + e4m3((X - fp16 X)*2^11) * e4m3(W*S)            kind::f8f6f4  (1/2 unit: FP8 runs at twice the rate)
//                    + e4m3(X) * e4m3(W*S*2^11 - fp16(W*S*2^11))      kind::f8f6f4  (1/2 unit)
// The main weights are pre-multiplied by the exact power of two 2^11 (and a per-layer power of two S that puts
// max|W*S| into [4,8)), so the two correction products come out at the SAME scale as the main product and all three
// accumulate into ONE fp32 TMEM accumulator; the epilogue multiplies by 1/(S*2^11).  The corrections are 2^-11 of the
// main term, so E4M3's 2^-4 relative rounding costs 2^-15 overall -- CPU emulation of the 14-block TCN: 6.2e-6 RMS against
// fp32 (bf16 x 3: 2.6e-6; single-pass fp16: 1.8e-4; budget 1e-4).
//
// Activation row (512 B per time step, same bytes as fp32):
//   [ fp16 hi, ch 0-63 | fp16 hi, ch 64-127 | e4m3 (x - hi)*2^11, ch 0-127 | e4m3 x, ch 0-127 ]      4 planes x 128 B
// Weights per tap (64 KB):  [ fp16 W*S*2^11 [co][ci 0-63] | same, ci 64-127 | e4m3 W*S [co][ci 0-127] | e4m3 lo [co][ci 0-127] ]
// Every plane / weight tile is 128 rows x 128 B -> one TMA box, one SWIZZLE_128B K-major UMMA operand tile.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "f32x2.cuh"
#include "sm100_ptx.cuh"

namespace mst {
namespace f8 {

constexpr int kCh = MST_TCN_CH;
constexpr int kTaps = MST_TCN_K;
constexpr int kRowBytes = 512;
constexpr int kSubRows = 128;
constexpr int kTileRows = 256;
constexpr float kLoScale = 2048.f;   // 2^11

__device__ __forceinline__ float fp8_to_float(uint8_t v) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)v, __NV_E4M3);
  return __half2float(__half(h));
}
__device__ __forceinline__ uint8_t float_to_fp8(float v) {
  return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}
// v -> (fp16 hi, e4m3 of the scaled remainder, e4m3 of v)
__device__ __forceinline__ void encode3(float v, __half& hi, uint8_t& l8, uint8_t& h8) {
  hi = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  l8 = float_to_fp8((v - __half2float(hi)) * kLoScale);
  h8 = float_to_fp8(v);
}
__device__ __forceinline__ float decode2(__half hi, uint8_t l8) {
  return __half2float(hi) + fp8_to_float(l8) * (1.f / kLoScale);
}

// =====================================================================================================================
// weight packing
// =====================================================================================================================
__global__ void wmax_kernel(const float* __restrict__ w, const float* __restrict__ bn_w, const float* __restrict__ bn_var,
                            unsigned int* __restrict__ max_bits) {
  float m = 0.f;
  const int n = kCh * kCh * kTaps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i / (kCh * kTaps);
    m = fmaxf(m, fabsf(w[i] * (bn_w[co] / sqrtf(bn_var[co] + 1e-5f))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, __float_as_uint(m));
}

__device__ __forceinline__ float layer_scale(unsigned int max_bits) {
  const float m = __uint_as_float(max_bits);
  if (!(m > 0.f)) return 1.f;
  return exp2f(floorf(log2f(8.f / m)));   // max|W*S| in [4, 8)
}

// out (bytes): [tap][4 tiles][co 128][128 B];  inv_scale[0] = 1 / (S * 2^11)
__global__ void pack_kernel(const float* __restrict__ w, const float* __restrict__ bn_w, const float* __restrict__ bn_var,
                            const unsigned int* __restrict__ max_bits, uint8_t* __restrict__ out,
                            float* __restrict__ inv_scale) {
  const float S = layer_scale(*max_bits);
  const int n = kTaps * kCh * kCh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ci = i % kCh;
    const int co = (i / kCh) % kCh;
    const int tap = i / (kCh * kCh);
    const float s = bn_w[co] / sqrtf(bn_var[co] + 1e-5f);
    const float v = w[((size_t)co * kCh + ci) * kTaps + tap] * s * S;
    const float vm = v * kLoScale;
    const __half wm = __float2half_rn(vm);
    uint8_t* tap_base = out + (size_t)tap * 4 * kCh * 128;
    reinterpret_cast<__half*>(tap_base + (size_t)(ci >> 6) * kCh * 128 + (size_t)co * 128)[ci & 63] = wm;
    tap_base[(size_t)2 * kCh * 128 + (size_t)co * 128 + ci] = float_to_fp8(v);
    tap_base[(size_t)3 * kCh * 128 + (size_t)co * 128 + ci] = float_to_fp8(vm - __half2float(wm));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.f / (S * kLoScale);
}

// =====================================================================================================================
// format converters (block 0 of this format runs on the tensor cores: tcn_b0.cuh)
// =====================================================================================================================
__global__ void __launch_bounds__(256) act_pack_kernel(const float* __restrict__ x, uint8_t* __restrict__ act, int T, int Ts) {
  __shared__ float tile[kCh][33];
  const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = warp; c < kCh; c += 8) {
    const int t = t0 + lane;
    tile[c][lane] = t < T ? x[((size_t)b * kCh + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = warp; r < 32; r += 8) {
    const int t = t0 + r;
    if (t >= T) continue;
    uint8_t* row = act + ((size_t)b * Ts + t) * kRowBytes;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __half hi[2];
      uint8_t l8[2], h8[2];
      encode3(tile[64 * half + 2 * lane][r], hi[0], l8[0], h8[0]);
      encode3(tile[64 * half + 2 * lane + 1][r], hi[1], l8[1], h8[1]);
      reinterpret_cast<__half2*>(row + half * 128)[lane] = __halves2half2(hi[0], hi[1]);
      reinterpret_cast<uint16_t*>(row + 256 + half * 64)[lane] = (uint16_t)l8[0] | ((uint16_t)l8[1] << 8);
      reinterpret_cast<uint16_t*>(row + 384 + half * 64)[lane] = (uint16_t)h8[0] | ((uint16_t)h8[1] << 8);
    }
  }
}

__global__ void __launch_bounds__(256) act_unpack_kernel(const uint8_t* __restrict__ act, float* __restrict__ y, int T, int Ts) {
  __shared__ float tile[kCh][33];
  const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < 32; r += 8) {
    const int t = t0 + r;
    if (t >= T) continue;
    const uint8_t* row = act + ((size_t)b * Ts + t) * kRowBytes;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const __half2 h2 = reinterpret_cast<const __half2*>(row + half * 128)[lane];
      const uint16_t l2 = reinterpret_cast<const uint16_t*>(row + 256 + half * 64)[lane];
      tile[64 * half + 2 * lane][r] = decode2(__low2half(h2), (uint8_t)(l2 & 0xFF));
      tile[64 * half + 2 * lane + 1][r] = decode2(__high2half(h2), (uint8_t)(l2 >> 8));
    }
  }
  __syncthreads();
  for (int c = warp; c < kCh; c += 8) {
    const int t = t0 + lane;
    if (t < T) y[((size_t)b * kCh + c) * T + t] = tile[c][lane];
  }
}


}  // namespace f8

// ---- entry points used by tcn.cu ----
size_t tcn_f8_weight_bytes() { return (size_t)f8::kTaps * 4 * f8::kCh * 128; }   // 983,040 B per layer (same as bf16 x 2)

int tcn_f8_pack_layer(const float* conv_w, const float* bn_w, const float* bn_var, void* w_out, float* inv_scale,
                      unsigned int* scratch, cudaStream_t st) {
  MST_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(unsigned int), st));
  f8::wmax_kernel<<<128, 256, 0, st>>>(conv_w, bn_w, bn_var, scratch);
  if (launch_ok("tcn f8 wmax_kernel")) return 1;
  f8::pack_kernel<<<256, 256, 0, st>>>(conv_w, bn_w, bn_var, scratch, (uint8_t*)w_out, inv_scale);
  return launch_ok("tcn f8 pack_kernel");
}

int tcn_f8_act_pack(const float* x, void* act, int B, int T, cudaStream_t st) {
  f8::act_pack_kernel<<<dim3(cdiv(T, 32), B), 256, 0, st>>>(x, (uint8_t*)act, T, tcn_seg_rows(T));
  return launch_ok("tcn f8 act_pack_kernel");
}
int tcn_f8_act_unpack(const void* act, float* y, int B, int T, cudaStream_t st) {
  f8::act_unpack_kernel<<<dim3(cdiv(T, 32), B), 256, 0, st>>>((const uint8_t*)act, y, T, tcn_seg_rows(T));
  return launch_ok("tcn f8 act_unpack_kernel");
}

}  // namespace mst
