// pcm.cu -- WAV sample-format conversions on the device: the steps either side of the forward.
//
// Replaces (paths relative to /root/reference/):
//   load_wav_segment            mixing_style_transfer/data_loader/loader_utils.py:47-70   int16 / int32 interleaved PCM ->
//                               float (x / 2^15 or x / 2^31), de-interleaved to [channel][frame]
//   clamp of the loaded stems   mixing_style_transfer/data_loader/data_loader.py:589-590  (a no-op on PCM data, kept)
//   mono -> stereo duplication  inference/feature_extraction.py:87-89
//   remix + PCM_16 write        inference/style_transfer.py:165-177  sum of the per-instrument outputs (float32, in
//                               instrument order) and the PCM_16 quantisation of the written file
// Raw PCM crosses PCIe (2 bytes per sample instead of 4, and ONE int16 mixture back instead of n_stems fp32 waveforms);
// both kernels are pure HBM streams: 16-byte accesses, 4 frames per thread, grid-stride.
#include "common.cuh"

namespace mst {

template <typename S>
__device__ __forceinline__ float pcm_to_float(S v);
template <>
__device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }   // exact
template <>
__device__ __forceinline__ float pcm_to_float<int32_t>(int32_t v) { return (float)((double)v / 2147483648.0); }  // numpy: float64 divide, then .float()

template <typename S>
__global__ void __launch_bounds__(256)
pcm_decode_kernel(const S* __restrict__ pcm, int n_ch, long long n_frames, float* __restrict__ out, long long out_stride, int vec) {
  const long long n4 = n_frames / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float* o0 = out;
  float* o1 = out + out_stride;
  if (vec) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
      float a[4], c[4];
      const S* src = pcm + q * 4 * n_ch;
      S s[8];
      if (sizeof(S) == 2 && n_ch == 2) {
        *reinterpret_cast<uint4*>(s) = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = i < 4 * n_ch ? src[i] : S(0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = fminf(fmaxf(pcm_to_float<S>(s[i * n_ch]), -1.f), 1.f);
        c[i] = n_ch == 2 ? fminf(fmaxf(pcm_to_float<S>(s[i * n_ch + 1]), -1.f), 1.f) : a[i];
      }
      reinterpret_cast<float4*>(o0)[q] = make_float4(a[0], a[1], a[2], a[3]);
      reinterpret_cast<float4*>(o1)[q] = make_float4(c[0], c[1], c[2], c[3]);
    }
  }
  const long long t0 = vec ? n4 * 4 : 0;
  for (long long t = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_frames; t += stride) {
    const float a = fminf(fmaxf(pcm_to_float<S>(pcm[t * n_ch]), -1.f), 1.f);
    const float c = n_ch == 2 ? fminf(fmaxf(pcm_to_float<S>(pcm[t * n_ch + 1]), -1.f), 1.f) : a;
    o0[t] = a;
    o1[t] = c;
  }
}

__device__ __forceinline__ int16_t quantise_pcm16(float v) {
  // np.clip(np.rint(v * 32768.0), -32768, 32767): the product is exact in float32, rint = round half to even
  const int q = __float2int_rn(v * 32768.0f);
  return (int16_t)max(-32768, min(32767, q));
}

__global__ void __launch_bounds__(256)
pcm_encode_mix_kernel(const float* __restrict__ stems, int n_stems, long long stem_stride, long long n_frames,
                      int16_t* __restrict__ pcm, int vec) {
  const long long n4 = n_frames / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
      float4 l = __ldg(reinterpret_cast<const float4*>(stems) + q);
      float4 r = __ldg(reinterpret_cast<const float4*>(stems + stem_stride) + q);
      for (int s = 1; s < n_stems; ++s) {   // float32 adds in instrument order, like Python's sum() of the numpy arrays
        const float4 a = __ldg(reinterpret_cast<const float4*>(stems + (size_t)s * 2 * stem_stride) + q);
        const float4 c = __ldg(reinterpret_cast<const float4*>(stems + ((size_t)s * 2 + 1) * stem_stride) + q);
        l.x += a.x; l.y += a.y; l.z += a.z; l.w += a.w;
        r.x += c.x; r.y += c.y; r.z += c.z; r.w += c.w;
      }
      int16_t o[8] = {quantise_pcm16(l.x), quantise_pcm16(r.x), quantise_pcm16(l.y), quantise_pcm16(r.y),
                      quantise_pcm16(l.z), quantise_pcm16(r.z), quantise_pcm16(l.w), quantise_pcm16(r.w)};
      reinterpret_cast<uint4*>(pcm)[q] = *reinterpret_cast<const uint4*>(o);
    }
  }
  const long long t0 = vec ? n4 * 4 : 0;
  for (long long t = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_frames; t += stride) {
    float l = stems[t], r = stems[stem_stride + t];
    for (int s = 1; s < n_stems; ++s) {
      l += stems[(size_t)s * 2 * stem_stride + t];
      r += stems[((size_t)s * 2 + 1) * stem_stride + t];
    }
    pcm[2 * t] = quantise_pcm16(l);
    pcm[2 * t + 1] = quantise_pcm16(r);
  }
}

static int stream_grid(long long n_items) {
  const long long want = (n_items + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace mst

using namespace mst;

extern "C" {

int mst_pcm_decode(const void* pcm, int sample_bytes, int n_channels, long long n_frames, float* out,
                   long long out_stride, void* stream) {
  MST_CHECK(pcm && out, "pcm_decode: null pointer");
  MST_CHECK(sample_bytes == 2 || sample_bytes == 4, "ValueError: input audio's bit depth should be 16 or 32-bit (got %d bytes)",
            sample_bytes);
  MST_CHECK(n_channels == 1 || n_channels == 2, "pcm_decode: %d channels unsupported", n_channels);
  MST_CHECK(n_frames >= 0 && out_stride >= n_frames, "pcm_decode: bad sizes");
  if (n_frames == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (reinterpret_cast<uintptr_t>(pcm) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                  (out_stride % 4 == 0) ? 1 : 0;
  const int grid = stream_grid((n_frames + 3) / 4);
  if (sample_bytes == 2)
    pcm_decode_kernel<int16_t><<<grid, 256, 0, st>>>(reinterpret_cast<const int16_t*>(pcm), n_channels, n_frames, out, out_stride, vec);
  else
    pcm_decode_kernel<int32_t><<<grid, 256, 0, st>>>(reinterpret_cast<const int32_t*>(pcm), n_channels, n_frames, out, out_stride, vec);
  return launch_ok("pcm_decode_kernel");
}

int mst_pcm_encode_mix(const float* stems, int n_stems, long long stem_stride, long long n_frames, int16_t* pcm,
                       void* stream) {
  MST_CHECK(stems && pcm, "pcm_encode_mix: null pointer");
  MST_CHECK(n_stems >= 1 && n_frames >= 0 && stem_stride >= n_frames, "pcm_encode_mix: bad sizes");
  if (n_frames == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (reinterpret_cast<uintptr_t>(stems) % 16 == 0) && (reinterpret_cast<uintptr_t>(pcm) % 16 == 0) &&
                  (stem_stride % 4 == 0) ? 1 : 0;
  pcm_encode_mix_kernel<<<stream_grid((n_frames + 3) / 4), 256, 0, st>>>(stems, n_stems, stem_stride, n_frames, pcm, vec);
  return launch_ok("pcm_encode_mix_kernel");
}

}  // extern "C"
