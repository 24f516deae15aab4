// fxnorm.cu -- small stereo building blocks for the input FX normaliser and the remaining FXmanipulator
// processors (8f-4).  All HBM-bound streams: 16-byte accesses, grids sized to fill the 148 SMs.
//
// Replaces (paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/):
//   mst_stereo_stats     the reductions of normalize_imager / process_balance (normalization_imager.py:34-36, 94-99), of
//                        MidSideImager (common_audioeffects.py:968-971) and the peak of lufs_normalize (fx_utils.py:231)
//   mst_stereo_mix       every per-frame 2x2 map of that code: lr_to_ms / gains / ms_to_lr (normalization_imager.py:31-76,
//                        101-118), Panner.process (common_audioeffects.py:935), Gain, the loudness gain (fx_utils.py:229-232)
//   mst_block_energy     the gating-block mean squares of the BS.1770 meter behind fx_utils.lufs_normalize (:224,
//                        pyloudnorm.Meter.integrated_loudness -- third-party, restated in oracle/norm_oracle.py)
//   mst_haas             haas_process (common_audioeffects.py:767-787): y[ch] += feedback * roll(x[ch], delay)
#include "common.cuh"

namespace mst {
namespace fxn {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stats[b] = (sum L^2, sum R^2, sum L R, max |x|), float64 accumulation; out must be zeroed
__global__ void __launch_bounds__(kThreads)
stereo_stats_kernel(const float* __restrict__ x, long long L, int vec, double* __restrict__ out) {
  const int b = blockIdx.y;
  const float* l = x + (size_t)b * 2 * L;
  const float* r = l + L;
  double sll = 0.0, srr = 0.0, slr = 0.0;
  float mx = 0.f;
  const long long stride = (long long)gridDim.x * kThreads;
  if (vec) {
    const long long n4 = L / 4;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(l) + i), c = __ldg(reinterpret_cast<const float4*>(r) + i);
      const float ll = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;       // 4 frames in float32, then float64
      const float rr = c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
      const float lr = a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      sll += (double)ll; srr += (double)rr; slr += (double)lr;
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
    }
  } else {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < L; i += stride) {
      const float a = l[i], c = r[i];
      sll += (double)a * a; srr += (double)c * c; slr += (double)a * c;
      mx = fmaxf(mx, fmaxf(fabsf(a), fabsf(c)));
    }
  }
  __shared__ double red[kThreads / 32][4];
  sll = warp_sum(sll); srr = warp_sum(srr); slr = warp_sum(slr);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w][0] = sll; red[w][1] = srr; red[w][2] = slr; red[w][3] = (double)mx; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) a = threadIdx.x < 3 ? a + red[i][threadIdx.x] : fmax(a, red[i][3]);
    double* o = out + (size_t)b * 4 + threadIdx.x;
    if (threadIdx.x < 3) atomicAdd(o, a);
    else atomicMax(reinterpret_cast<unsigned long long*>(o), (unsigned long long)__double_as_longlong(a));  // a >= 0
  }
}

// y = M x per frame: y_L = m0 l + m1 r, y_R = m2 l + m3 r (one multiply-add each, like the numpy expressions it replaces)
__global__ void __launch_bounds__(kThreads)
stereo_mix_kernel(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ y, long long L, int vec) {
  const int b = blockIdx.y;
  const float m0 = m[b * 4 + 0], m1 = m[b * 4 + 1], m2 = m[b * 4 + 2], m3 = m[b * 4 + 3];
  const float* l = x + (size_t)b * 2 * L;
  const float* r = l + L;
  float* yl = y + (size_t)b * 2 * L;
  float* yr = yl + L;
  const long long stride = (long long)gridDim.x * kThreads;
  if (vec) {
    const long long n4 = L / 4;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(l) + i), c = __ldg(reinterpret_cast<const float4*>(r) + i);
      float4 ol, orr;
      ol.x = fmaf(m0, a.x, m1 * c.x); orr.x = fmaf(m2, a.x, m3 * c.x);
      ol.y = fmaf(m0, a.y, m1 * c.y); orr.y = fmaf(m2, a.y, m3 * c.y);
      ol.z = fmaf(m0, a.z, m1 * c.z); orr.z = fmaf(m2, a.z, m3 * c.z);
      ol.w = fmaf(m0, a.w, m1 * c.w); orr.w = fmaf(m2, a.w, m3 * c.w);
      reinterpret_cast<float4*>(yl)[i] = ol;
      reinterpret_cast<float4*>(yr)[i] = orr;
    }
  } else {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < L; i += stride) {
      const float a = l[i], c = r[i];
      yl[i] = fmaf(m0, a, m1 * c);
      yr[i] = fmaf(m2, a, m3 * c);
    }
  }
}

// z[c][j] = sum_{t in [lo[j], hi[j])} x[c][t]^2  (float64); one CTA per (block j, channel c)
__global__ void __launch_bounds__(kThreads)
block_energy_kernel(const float* __restrict__ x, long long T, const long long* __restrict__ lo, const long long* __restrict__ hi,
                    int nb, double* __restrict__ z) {
  const int j = blockIdx.x, c = blockIdx.y;
  long long a = lo[j], e = hi[j];
  a = a < 0 ? 0 : a;
  e = e > T ? T : e;                       // numpy slicing clips the same way
  const float* p = x + (size_t)c * T;
  double s = 0.0;
  for (long long t = a + threadIdx.x; t < e; t += kThreads) { const float v = __ldg(p + t); s += (double)v * v; }
  __shared__ double red[kThreads / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += red[i];
    z[(size_t)c * nb + j] = t;
  }
}

// y = x;  y[ch][t] += feedback * x[ch][(t - delay) mod L]   (np.roll wraps around)
__global__ void __launch_bounds__(kThreads)
haas_kernel(const float* __restrict__ x, float* __restrict__ y, long long L, const int* __restrict__ delay,
            const float* __restrict__ feedback, const int* __restrict__ channel) {
  const int b = blockIdx.y;
  const int ch = channel[b];
  const float fb = feedback[b];
  long long d = (long long)delay[b] % L;
  if (d < 0) d += L;
  const float* xb = x + (size_t)b * 2 * L;
  float* yb = y + (size_t)b * 2 * L;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < 2 * L; i += stride) {
    const int c = i >= L;
    const long long t = i - (c ? L : 0);
    float v = xb[i];
    if (c == ch) {
      long long s = t - d;
      if (s < 0) s += L;
      v += fb * xb[(size_t)c * L + s];
    }
    yb[i] = v;
  }
}

static int grid_x(long long L) {
  long long g = (L / 4 + kThreads - 1) / kThreads;
  const long long cap = 8LL * sm_count();
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fxn
}  // namespace mst

using namespace mst;

extern "C" {

int mst_stereo_stats(const float* x, int B, long long L, double* stats, void* stream) {
  MST_CHECK(x && stats && B > 0 && B <= 65535 && L > 0, "stereo_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  MST_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)B * 4 * sizeof(double), st));
  const int vec = (L % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  fxn::stereo_stats_kernel<<<dim3(fxn::grid_x(L), B), fxn::kThreads, 0, st>>>(x, L, vec, stats);
  return launch_ok("stereo_stats_kernel");
}

int mst_stereo_mix(const float* x, const float* matrices, float* y, int B, long long L, void* stream) {
  MST_CHECK(x && matrices && y && B > 0 && B <= 65535 && L > 0, "stereo_mix: bad arguments");
  const int vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0);
  fxn::stereo_mix_kernel<<<dim3(fxn::grid_x(L), B), fxn::kThreads, 0, (cudaStream_t)stream>>>(x, matrices, y, L, vec);
  return launch_ok("stereo_mix_kernel");
}

int mst_block_energy(const float* x, int n_channels, long long T, const long long* lo, const long long* hi, int n_blocks,
                     double* z, void* stream) {
  MST_CHECK(x && lo && hi && z && n_channels > 0 && n_channels <= 65535 && T > 0 && n_blocks > 0, "block_energy: bad arguments");
  fxn::block_energy_kernel<<<dim3(n_blocks, n_channels), fxn::kThreads, 0, (cudaStream_t)stream>>>(x, T, lo, hi, n_blocks, z);
  return launch_ok("block_energy_kernel");
}

int mst_haas(const float* x, float* y, int B, long long L, const int* delay, const float* feedback, const int* channel,
             void* stream) {
  MST_CHECK(x && y && x != y && delay && feedback && channel && B > 0 && B <= 65535 && L > 0, "haas: bad arguments (in-place is not supported)");
  fxn::haas_kernel<<<dim3(fxn::grid_x(2 * L), B), fxn::kThreads, 0, (cudaStream_t)stream>>>(x, y, L, delay, feedback, channel);
  return launch_ok("haas_kernel");
}

}  // extern "C"
