// spectral.cu -- the spectral pieces of the input FX normaliser's EQ matching for sm_100a:
// averaged STFT magnitude (65,536-point frames), zero-phase FIR filtering (filtfilt), per-row peaks.
//
// Replaces (paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/):
//   mst_stft_mag_mean   compute_stft + np.abs + np.mean of get_eq_matching (utils_data_normalization.py:74-79;
//                       common_miscellaneous.py:50-77: librosa.stft(center=False), frames of n_fft samples every hop samples,
//                       analysis window sqrt(hann), spectrum stored as complex64)
//   mst_fir_filtfilt    scipy.signal.filtfilt(taps, 1, x, padtype='odd', method='pad') of get_eq_matching (:100-102):
//                       odd extension by 3*n_taps samples, forward FIR, backward FIR, both started in the steady state of
//                       the first sample (lfilter_zi), float64 like scipy
//   mst_row_absmax      np.max(np.abs(x)) per channel (:69, fx_utils.py:231)
//   mst_fft_convolve    oaconvolve(x, h, mode='full', axes=0) + cut + dry / wet mix of ConvolutionalReverb.process
//                       (common_audioeffects.py:735-764), SURVEY.md 8f-4
//
// FFT: four-step decomposition N = 256 * N2 (N2 = 4 .. 256).  Two signals share one complex transform (real part = signal
// 2p, imaginary part = signal 2p + 1; the spectra are separated from Z[k] and conj(Z[N - k])).  Pass A: 16 columns per CTA,
// 256-point in-place radix-2 transforms in shared memory (bit-reversed on load), twiddle W_N^(n2 k1), coalesced 128-byte
// rows out.  Pass B: 16 rows per CTA, N2-point transforms, written so that bin k = k1 + 256 k2 lands at index k.  Pass C:
// one thread per bin, magnitudes summed over the frames of a batch in float64.  All three passes are HBM/L2 streams of
// 8 bytes per point; the transform arithmetic is float32 (the reference stores complex64).
//
// FIR: float64 on the B200's full-rate FP64 pipe (DFMA 1.67 warp-instructions / clk / SM, tools/ubench/fp_pipes.cu): 1001 taps
// x 2 passes over a 3-minute stem is 32 G DFMA = ~2 ms, so there is no reason to give up scipy's float64 arithmetic.  Thread =
// 8 consecutive outputs, a 16-sample register window slides over the shared-memory tile (4 LDS.128 of samples + 4 broadcast
// LDS.128 of taps per 64 DFMA); sample blocks of 8 are padded to 10 doubles so that the quarter-warp 16-byte loads hit 32
// distinct banks.
#include "common.cuh"

namespace mst {
namespace spec {

constexpr int kThreads = 256;
constexpr int kN1 = 256;          // column transform length
constexpr int kCols = 16;         // columns (pass A) / rows (pass B) per CTA
constexpr int kFrameBatch = 32;   // frames per workspace batch

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place radix-2 DIT over `n` points of `cnt` interleaved transforms: element (point p, transform j) at d[p * cnt + j];
// input already in bit-reversed order; tw[k] = exp(-2 pi i k / 256), k < 128
__device__ __forceinline__ void fft_inplace(float2* d, const float2* tw, int n, int logn, int cnt) {
  for (int s = 0; s < logn; ++s) {
    const int half = 1 << s;
    const int nb = (n >> 1) * cnt;                  // butterflies in this stage
    for (int q = threadIdx.x; q < nb; q += kThreads) {
      const int j = q % cnt, bf = q / cnt;
      const int grp = bf >> s, pos = bf & (half - 1);
      const int i0 = (grp << (s + 1)) + pos, i1 = i0 + half;
      const float2 w = tw[pos << (7 - s)];          // W_{2 half}^pos = W_256^(pos * 128 / half)
      const float2 a = d[i0 * cnt + j], b = cmul(d[i1 * cnt + j], w);
      d[i0 * cnt + j] = make_float2(a.x + b.x, a.y + b.y);
      d[i1 * cnt + j] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void load_twiddles(float2* tw) {
  if (threadIdx.x < 128) {
    double s, c;
    sincospi(-2.0 * (double)threadIdx.x / 256.0, &s, &c);
    tw[threadIdx.x] = make_float2((float)c, (float)s);
  }
}

// pass A: Y[pair][f][k1][n2] = W_N^(n2 k1) * sum_n1 w[n] x[f hop + n] W_256^(n1 k1),  n = n1 N2 + n2
__global__ void __launch_bounds__(kThreads)
fft_cols_kernel(const float* __restrict__ x, long long stride, int n_signals, long long f0, int hop, int N2, const float* __restrict__ win,
                float2* __restrict__ Y) {
  __shared__ float2 d[kN1 * kCols];
  __shared__ float2 tw[128];
  const int c0 = blockIdx.x * kCols, f = blockIdx.y, pair = blockIdx.z;
  const int N = kN1 * N2;
  const int cols = min(kCols, N2 - c0);              // N2 < 16: fewer columns
  load_twiddles(tw);
  const float* xa = x + (size_t)(2 * pair) * stride + (f0 + f) * (long long)hop;
  const bool has_b = 2 * pair + 1 < n_signals;
  const float* xb = xa + stride;
  for (int q = threadIdx.x; q < kN1 * cols; q += kThreads) {
    const int j = q % cols, n1 = q / cols;
    const int n = n1 * N2 + c0 + j;
    const float w = __ldg(win + n);
    const int r = __brev((unsigned)n1) >> 24;        // 8-bit reversal
    d[r * cols + j] = make_float2(w * __ldg(xa + n), has_b ? w * __ldg(xb + n) : 0.f);
  }
  __syncthreads();
  fft_inplace(d, tw, kN1, 8, cols);
  float2* out = Y + ((size_t)pair * gridDim.y + f) * N;
  for (int q = threadIdx.x; q < kN1 * cols; q += kThreads) {
    const int j = q % cols, k1 = q / cols;
    const int n2 = c0 + j;
    float s, c;
    sincospif(-2.0f * (float)((n2 * k1) & (N - 1)) / (float)N, &s, &c);   // exact argument: N is a power of two <= 2^16
    out[(size_t)k1 * N2 + n2] = cmul(d[k1 * cols + j], make_float2(c, s));
  }
}

// pass B: Z[pair][f][k1 + 256 k2] = sum_n2 Y[k1][n2] W_N2^(n2 k2)
__global__ void __launch_bounds__(kThreads)
fft_rows_kernel(const float2* __restrict__ Y, int N2, int logn2, float2* __restrict__ Z) {
  extern __shared__ float2 dsm[];                    // [N2][kCols] + 128 twiddles
  float2* d = dsm;
  float2* tw = dsm + (size_t)N2 * kCols;
  const int r0 = blockIdx.x * kCols, f = blockIdx.y, pair = blockIdx.z;
  const int N = kN1 * N2;
  load_twiddles(tw);
  const float2* in = Y + ((size_t)pair * gridDim.y + f) * N;
  for (int q = threadIdx.x; q < N2 * kCols; q += kThreads) {
    const int n2 = q % N2, j = q / N2;               // consecutive threads read consecutive n2 of one row
    const int r = (int)(__brev((unsigned)n2) >> (32 - logn2));
    d[r * kCols + j] = in[(size_t)(r0 + j) * N2 + n2];
  }
  __syncthreads();
  // the twiddle table is W_256: a transform of length N2 <= 256 uses every (256 / N2)-th entry -> fft_inplace's indexing
  // (pos << (7 - s)) is already expressed in W_256 units for a stage of half-length 2^s
  fft_inplace(d, tw, N2, logn2, kCols);
  float2* out = Z + ((size_t)pair * gridDim.y + f) * N;
  for (int q = threadIdx.x; q < N2 * kCols; q += kThreads) {
    const int j = q % kCols, k2 = q / kCols;
    out[(size_t)k2 * kN1 + r0 + j] = d[k2 * kCols + j];
  }
}

// pass C: acc[sig][k] += sum_f |X_sig,f[k]|,  X_a = (Z[k] + conj Z[N-k]) / 2,  X_b = (Z[k] - conj Z[N-k]) / (2i)
__global__ void __launch_bounds__(kThreads)
mag_accumulate_kernel(const float2* __restrict__ Z, int N, int n_frames, int n_signals, double* __restrict__ acc) {
  const int k = blockIdx.x * kThreads + threadIdx.x, pair = blockIdx.y;
  const int nb = N / 2 + 1;
  if (k >= nb) return;
  const int km = (N - k) & (N - 1);
  double sa = 0.0, sb = 0.0;
  for (int f = 0; f < n_frames; ++f) {
    const float2* z = Z + ((size_t)pair * n_frames + f) * N;
    const float2 p = z[k], q = z[km];
    const float ar = 0.5f * (p.x + q.x), ai = 0.5f * (p.y - q.y);     // (p + conj q) / 2
    const float br = 0.5f * (p.y + q.y), bi = 0.5f * (q.x - p.x);     // (p - conj q) / (2i)
    sa += (double)hypotf(ar, ai);
    sb += (double)hypotf(br, bi);
  }
  acc[(size_t)(2 * pair) * nb + k] += sa;
  if (2 * pair + 1 < n_signals) acc[(size_t)(2 * pair + 1) * nb + k] += sb;
}

__global__ void scale_kernel(double* __restrict__ a, size_t n, double s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] *= s;
}

// ---- FFT convolution (ConvolutionalReverb: scipy.signal.oaconvolve, common_audioeffects.py:748) ----------------------------------
// Partitioned overlap-add at frame length N, hop H = N / 2: x is cut into blocks of H samples, h into partitions of H samples,
// all zero-padded to N and transformed with the passes above (window = 1 on the first half, 0 on the second; the two channels
// of a stereo signal travel as real / imaginary part of one complex transform).  Output block j = sum over i + p = j of
// X_i H_p, built per bin from the separated channel spectra and re-packed as Y = Y_L + i Y_R; the inverse transform is the
// forward one on conj(Y); blocks overlap-add with hop H.
__device__ __forceinline__ void unpack_pair(float2 p, float2 q, float2& a, float2& b) {
  a = make_float2(0.5f * (p.x + q.x), 0.5f * (p.y - q.y));       // (p + conj q) / 2
  b = make_float2(0.5f * (p.y + q.y), 0.5f * (q.x - p.x));       // (p - conj q) / (2i)
}

// Yc[j][k] = conj( sum_p XL[j-p][k] HL[p][k]  +  i * sum_p XR[j-p][k] HR[p][k] )   (conjugated for the inverse pass)
__global__ void __launch_bounds__(kThreads)
conv_mac_kernel(const float2* __restrict__ Zx, int nbx, const float2* __restrict__ Zh, int P, int N, float2* __restrict__ Yc) {
  const int k = blockIdx.x * kThreads + threadIdx.x, j = blockIdx.y;
  if (k >= N) return;
  const int km = (N - k) & (N - 1);
  float2 yl = make_float2(0.f, 0.f), yr = yl;
  for (int p = 0; p < P; ++p) {
    const int i = j - p;
    if (i < 0 || i >= nbx) continue;
    float2 xl, xr, hl, hr;
    unpack_pair(Zx[(size_t)i * N + k], Zx[(size_t)i * N + km], xl, xr);
    unpack_pair(Zh[(size_t)p * N + k], Zh[(size_t)p * N + km], hl, hr);
    const float2 a = cmul(xl, hl), b = cmul(xr, hr);
    yl.x += a.x; yl.y += a.y; yr.x += b.x; yr.y += b.y;
  }
  // Y = yl + i yr = (yl.x - yr.y) + i (yl.y + yr.x);  store conj(Y)
  Yc[(size_t)j * N + k] = make_float2(yl.x - yr.y, -(yl.y + yr.x));
}

// pass A on complex input (one transform per frame): Y[f][k1][n2] = W_N^(n2 k1) * sum_n1 z[f][n1 N2 + n2] W_256^(n1 k1)
__global__ void __launch_bounds__(kThreads)
fft_cols_cplx_kernel(const float2* __restrict__ zin, int N2, float2* __restrict__ Y) {
  __shared__ float2 d[kN1 * kCols];
  __shared__ float2 tw[128];
  const int c0 = blockIdx.x * kCols, f = blockIdx.y;
  const int N = kN1 * N2;
  const int cols = min(kCols, N2 - c0);
  load_twiddles(tw);
  const float2* z = zin + (size_t)f * N;
  for (int q = threadIdx.x; q < kN1 * cols; q += kThreads) {
    const int j = q % cols, n1 = q / cols;
    d[(__brev((unsigned)n1) >> 24) * cols + j] = z[n1 * N2 + c0 + j];
  }
  __syncthreads();
  fft_inplace(d, tw, kN1, 8, cols);
  float2* out = Y + (size_t)f * N;
  for (int q = threadIdx.x; q < kN1 * cols; q += kThreads) {
    const int j = q % cols, k1 = q / cols;
    const int n2 = c0 + j;
    float sn, cs;
    sincospif(-2.0f * (float)((n2 * k1) & (N - 1)) / (float)N, &sn, &cs);
    out[(size_t)k1 * N2 + n2] = cmul(d[k1 * cols + j], make_float2(cs, sn));
  }
}

// y[c][t] = dry * x[c][t] + wet * full[c][t + offset],  full = overlap-add of the inverse blocks: the inverse transform of Y is
// conj(F) / N with F = FFT(conj Y), so full_L = Re(F) / N and full_R = -Im(F) / N
__global__ void __launch_bounds__(kThreads)
conv_ola_kernel(const float2* __restrict__ F, int nby, int N, const float* __restrict__ x, long long T, long long x_stride,
                long long offset, float dry, float wet, float* __restrict__ y, long long y_stride) {
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= T) return;
  const int H = N / 2;
  const long long u = t + offset;
  const long long j = u / H;
  const int n = (int)(u - j * H);
  float2 v = make_float2(0.f, 0.f);
  if (j < nby) { const float2 a = F[(size_t)j * N + n]; v.x += a.x; v.y += a.y; }
  if (j >= 1 && j - 1 < nby) { const float2 a = F[(size_t)(j - 1) * N + n + H]; v.x += a.x; v.y += a.y; }
  const float inv = 1.0f / (float)N;
  y[t] = dry * x[t] + wet * (v.x * inv);
  y[y_stride + t] = dry * x[x_stride + t] + wet * (-v.y * inv);
}

__global__ void __launch_bounds__(kThreads)
conv_pad_kernel(const float* __restrict__ src, long long n, long long src_stride, int mono, float* __restrict__ dst, long long n_pad) {
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_pad) return;
  const int c = blockIdx.y;
  dst[(size_t)c * n_pad + t] = t < n ? src[(size_t)(mono ? 0 : c) * src_stride + t] : 0.f;
}

__global__ void half_window_kernel(float* __restrict__ w, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) w[i] = i < N / 2 ? 1.f : 0.f;
}

// ---- per-row peak --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
row_absmax_kernel(const float* __restrict__ x, long long T, long long stride, double* __restrict__ out) {
  const float* p = x + (size_t)blockIdx.y * stride;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < T; i += (long long)gridDim.x * kThreads) m = fmaxf(m, fabsf(__ldg(p + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    atomicMax(reinterpret_cast<unsigned long long*>(out + blockIdx.y), (unsigned long long)__double_as_longlong((double)m));   // m >= 0
}

// ---- float64 FIR with steady-state start (one pass of filtfilt) ----------------------------------------------------------
constexpr int kFirPer = 8;                         // outputs per thread
constexpr int kFirTile = kThreads * kFirPer;       // 2048 outputs per CTA
constexpr int kFirMaxTaps = 2048;

// padded position of sample m inside the shared window (blocks of 8 doubles every 10)
__device__ __forceinline__ int fpad(int m) { return m + 2 * (m >> 3); }

// MODE 0: source = odd extension of x (float32) by P samples at both ends, index g in [0, T + 2P), constant before 0;
//         output y1[g] (float64).
// MODE 1: source r[i] = y1[n - 1 - i], constant before 0; outputs i in [P, P + T) -> y[T + P - 1 - i] = float(scale * z[i]).
template <int MODE>
__global__ void __launch_bounds__(kThreads)
fir64_kernel(const float* __restrict__ x, const double* __restrict__ y1_in, long long T, int P, const double* __restrict__ taps_all,
             int n_taps, const double* __restrict__ scale, double* __restrict__ y1_out, float* __restrict__ y, long long x_stride,
             long long y_stride) {
  extern __shared__ __align__(16) double fsm[];
  const int KP = (n_taps + 7) & ~7;
  double* tp = fsm;                                  // KP taps (zero-padded)
  double* sw = fsm + KP;                             // window, padded layout
  const int sig = blockIdx.y;
  const long long n_ext = T + 2LL * P;
  const long long o_first = (MODE == 0 ? 0 : (long long)P) + (long long)blockIdx.x * kFirTile;   // first output of this CTA
  const long long o_end = MODE == 0 ? n_ext : (long long)P + T;
  const float* xs = x + (size_t)sig * x_stride;
  const double* ys = y1_in + (size_t)sig * n_ext;
  const double* taps = taps_all + (size_t)sig * n_taps;
  for (int k = threadIdx.x; k < KP; k += kThreads) tp[k] = k < n_taps ? taps[k] : 0.0;
  const long long base = o_first - (KP - 1);         // source index of window sample 0
  const int n_win = kFirTile + KP - 1;
  for (int m = threadIdx.x; m < n_win; m += kThreads) {
    long long g = base + m;
    if (g < 0) g = 0;                                // steady state of the first sample (lfilter_zi * x[0])
    double v = 0.0;
    if (g < n_ext) {
      if (MODE == 0) {
        if (g < P) v = 2.0 * (double)__ldg(xs) - (double)__ldg(xs + (P - g));
        else if (g < P + T) v = (double)__ldg(xs + (g - P));
        else v = 2.0 * (double)__ldg(xs + (T - 1)) - (double)__ldg(xs + (2 * T + P - 2 - g));
      } else {
        v = ys[n_ext - 1 - g];
      }
    }
    sw[fpad(m)] = v;
  }
  __syncthreads();

  double acc[kFirPer];
#pragma unroll
  for (int j = 0; j < kFirPer; ++j) acc[j] = 0.0;
  double win[16];
  const int tl = threadIdx.x;
  {
    // hi block of the first step: samples 8 tl + KP .. + 7
    const double2* p = reinterpret_cast<const double2*>(sw + 10 * (tl + KP / 8));
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double2 v = p[i]; win[8 + 2 * i] = v.x; win[9 + 2 * i] = v.y; }
  }
  for (int k = 0; k < KP; k += 8) {
    const int c = KP - 8 - k;                        // window offset of the low block
    const double2* p = reinterpret_cast<const double2*>(sw + 10 * (tl + c / 8));
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double2 v = p[i]; win[2 * i] = v.x; win[2 * i + 1] = v.y; }
    const double2* tq = reinterpret_cast<const double2*>(tp + k);
    double b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double2 v = tq[i]; b[2 * i] = v.x; b[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < kFirPer; ++j) acc[j] = fma(b[i], win[7 + j - i], acc[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) win[8 + i] = win[i];
  }
  const long long o0 = o_first + (long long)tl * kFirPer;
  if (MODE == 0) {
    double* out = y1_out + (size_t)sig * n_ext;
#pragma unroll
    for (int j = 0; j < kFirPer; ++j)
      if (o0 + j < o_end) out[o0 + j] = acc[j];
  } else {
    const double s = scale != nullptr ? scale[sig] : 1.0;
    float* out = y + (size_t)sig * y_stride;
#pragma unroll
    for (int j = 0; j < kFirPer; ++j)
      if (o0 + j < o_end) out[T + P - 1 - (o0 + j)] = (float)(s * acc[j]);
  }
}

static size_t fir_smem_bytes(int n_taps) {
  const int KP = (n_taps + 7) & ~7;
  const int n_win = kFirTile + KP - 1;
  return ((size_t)KP + (size_t)(n_win + 2 * (n_win / 8) + 16)) * sizeof(double);
}

}  // namespace spec
}  // namespace mst

using namespace mst;

extern "C" {

int mst_row_absmax(const float* x, int n_rows, long long T, long long stride, double* out, void* stream) {
  MST_CHECK(x && out && n_rows > 0 && n_rows <= 65535 && T > 0 && stride >= T, "row_absmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  MST_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n_rows * sizeof(double), st));
  long long g = (T + spec::kThreads * 8 - 1) / (spec::kThreads * 8);
  const long long cap = 4LL * sm_count();
  g = g < 1 ? 1 : (g > cap ? cap : g);
  spec::row_absmax_kernel<<<dim3((unsigned)g, n_rows), spec::kThreads, 0, st>>>(x, T, stride, out);
  return launch_ok("row_absmax_kernel");
}

size_t mst_stft_workspace_bytes(int n_signals, int n_fft) {
  if (n_signals <= 0 || n_fft <= 0) return 0;
  const size_t pairs = (size_t)(n_signals + 1) / 2;
  return 2 * pairs * spec::kFrameBatch * (size_t)n_fft * sizeof(float2);
}

int mst_stft_mag_mean(const float* x, int n_signals, long long T, long long stride, int n_fft, int hop, const float* window,
                      double* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace spec;
  MST_CHECK(x && window && out && workspace, "stft_mag_mean: null pointer");
  MST_CHECK(n_signals > 0 && n_signals <= 2 * 65535 && hop > 0 && stride >= T, "stft_mag_mean: bad arguments");
  int logn = 0;
  while ((1 << logn) < n_fft) ++logn;
  MST_CHECK((1 << logn) == n_fft && n_fft >= 1024 && n_fft <= 65536, "stft_mag_mean: n_fft %d must be a power of two in [1024, 65536]", n_fft);
  MST_CHECK(T >= n_fft, "stft_mag_mean: signal (%lld samples) shorter than one frame (%d)", T, n_fft);
  MST_CHECK(workspace_bytes >= mst_stft_workspace_bytes(n_signals, n_fft), "stft_mag_mean: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int N2 = n_fft / kN1, logn2 = logn - 8;
  const int pairs = (n_signals + 1) / 2;
  const long long n_frames = 1 + (T - n_fft) / hop;            // common_miscellaneous.py:64
  const int nb = n_fft / 2 + 1;
  float2* Y = reinterpret_cast<float2*>(workspace);
  float2* Z = Y + (size_t)pairs * kFrameBatch * n_fft;
  MST_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n_signals * nb * sizeof(double), st));
  const size_t rows_smem = ((size_t)N2 * kCols + 128) * sizeof(float2);
  for (long long f0 = 0; f0 < n_frames; f0 += kFrameBatch) {
    const int nf = (int)((n_frames - f0) < kFrameBatch ? (n_frames - f0) : kFrameBatch);
    fft_cols_kernel<<<dim3(cdiv(N2, kCols), nf, pairs), kThreads, 0, st>>>(x, stride, n_signals, f0, hop, N2, window, Y);
    if (launch_ok("fft_cols_kernel")) return 1;
    fft_rows_kernel<<<dim3(kN1 / kCols, nf, pairs), kThreads, rows_smem, st>>>(Y, N2, logn2, Z);
    if (launch_ok("fft_rows_kernel")) return 1;
    mag_accumulate_kernel<<<dim3(cdiv(nb, kThreads), pairs), kThreads, 0, st>>>(Z, n_fft, nf, n_signals, out);
    if (launch_ok("mag_accumulate_kernel")) return 1;
  }
  const size_t n = (size_t)n_signals * nb;
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, 1.0 / (double)n_frames);
  return launch_ok("scale_kernel");
}

static int conv_frame_len(long long M) {
  int N = 1024;
  while (N < 65536 && (long long)N / 2 < M) N *= 2;       // one partition when the impulse response fits half a frame
  return N;
}

size_t mst_fft_convolve_workspace_bytes(long long T, long long M) {
  if (T <= 0 || M <= 0) return 0;
  const long long N = conv_frame_len(M), H = N / 2;
  const long long nbx = (T + H - 1) / H, P = (M + H - 1) / H, nby = nbx + P - 1;
  size_t b = 0;
  b += align_up((size_t)N * sizeof(float), 256);                               // half window
  b += align_up((size_t)2 * (nbx + 1) * H * sizeof(float), 256);               // padded x
  b += align_up((size_t)2 * (P + 1) * H * sizeof(float), 256);                 // padded h
  b += 2 * align_up((size_t)(nbx > nby ? nbx : nby) * N * sizeof(float2), 256); // pass A scratch + block spectra / inverse blocks
  b += align_up((size_t)nbx * N * sizeof(float2), 256);                        // Zx
  b += align_up((size_t)P * N * sizeof(float2), 256);                          // Zh
  return b;
}

int mst_fft_convolve(const float* x, long long T, long long x_stride, const float* h, long long M, long long h_stride,
                     int h_channels, long long offset, float dry, float wet, float* y, long long y_stride, void* workspace,
                     size_t workspace_bytes, void* stream) {
  using namespace spec;
  MST_CHECK(x && h && y && workspace, "fft_convolve: null pointer");
  MST_CHECK(T > 0 && M > 0 && x_stride >= T && y_stride >= T && h_stride >= M && (h_channels == 1 || h_channels == 2) && offset >= 0,
            "fft_convolve: bad arguments");
  MST_CHECK(workspace_bytes >= mst_fft_convolve_workspace_bytes(T, M), "fft_convolve: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = conv_frame_len(M), H = N / 2, N2 = N / kN1;
  int logn = 0;
  while ((1 << logn) < N) ++logn;
  const long long nbx = (T + H - 1) / H, P = (M + H - 1) / H, nby = nbx + P - 1;
  MST_CHECK(nby <= 65535, "fft_convolve: too many blocks (%lld)", nby);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) { uint8_t* r = w8; w8 += align_up(bytes, 256); return r; };
  float* win = reinterpret_cast<float*>(take((size_t)N * sizeof(float)));
  float* xp = reinterpret_cast<float*>(take((size_t)2 * (nbx + 1) * H * sizeof(float)));
  float* hp = reinterpret_cast<float*>(take((size_t)2 * (P + 1) * H * sizeof(float)));
  const size_t big = align_up((size_t)(nbx > nby ? nbx : nby) * N * sizeof(float2), 256);
  float2* S1 = reinterpret_cast<float2*>(take(big));
  float2* S2 = reinterpret_cast<float2*>(take(big));
  float2* Zx = reinterpret_cast<float2*>(take((size_t)nbx * N * sizeof(float2)));
  float2* Zh = reinterpret_cast<float2*>(take((size_t)P * N * sizeof(float2)));
  half_window_kernel<<<cdiv(N, 256), 256, 0, st>>>(win, N);
  const long long xpad = (nbx + 1) * H, hpad = (P + 1) * H;
  conv_pad_kernel<<<dim3((unsigned)((xpad + kThreads - 1) / kThreads), 2), kThreads, 0, st>>>(x, T, x_stride, 0, xp, xpad);
  conv_pad_kernel<<<dim3((unsigned)((hpad + kThreads - 1) / kThreads), 2), kThreads, 0, st>>>(h, M, h_stride, h_channels == 1, hp, hpad);
  if (launch_ok("conv_pad_kernel")) return 1;
  const size_t rows_smem = ((size_t)N2 * kCols + 128) * sizeof(float2);
  // forward transforms of the x blocks and the h partitions (pair-packed: L real, R imaginary)
  fft_cols_kernel<<<dim3(cdiv(N2, kCols), (unsigned)nbx, 1), kThreads, 0, st>>>(xp, xpad, 2, 0, H, N2, win, S1);
  fft_rows_kernel<<<dim3(kN1 / kCols, (unsigned)nbx, 1), kThreads, rows_smem, st>>>(S1, N2, logn - 8, Zx);
  fft_cols_kernel<<<dim3(cdiv(N2, kCols), (unsigned)P, 1), kThreads, 0, st>>>(hp, hpad, 2, 0, H, N2, win, S1);
  fft_rows_kernel<<<dim3(kN1 / kCols, (unsigned)P, 1), kThreads, rows_smem, st>>>(S1, N2, logn - 8, Zh);
  if (launch_ok("fft (convolution, forward)")) return 1;
  conv_mac_kernel<<<dim3(cdiv(N, kThreads), (unsigned)nby), kThreads, 0, st>>>(Zx, (int)nbx, Zh, (int)P, N, S2);
  if (launch_ok("conv_mac_kernel")) return 1;
  fft_cols_cplx_kernel<<<dim3(cdiv(N2, kCols), (unsigned)nby), kThreads, 0, st>>>(S2, N2, S1);
  fft_rows_kernel<<<dim3(kN1 / kCols, (unsigned)nby, 1), kThreads, rows_smem, st>>>(S1, N2, logn - 8, S2);
  if (launch_ok("fft (convolution, inverse)")) return 1;
  conv_ola_kernel<<<(unsigned)((T + kThreads - 1) / kThreads), kThreads, 0, st>>>(S2, (int)nby, N, x, T, x_stride, offset, dry, wet, y, y_stride);
  return launch_ok("conv_ola_kernel");
}

size_t mst_fir_filtfilt_workspace_bytes(int n_signals, long long T, int n_taps) {
  if (n_signals <= 0 || T <= 0 || n_taps <= 0) return 0;
  return align_up((size_t)n_signals * (size_t)(T + 6LL * n_taps) * sizeof(double), 256);
}

int mst_fir_filtfilt(const float* x, int n_signals, long long T, long long stride, const double* taps, int n_taps,
                     const double* scale, float* y, long long y_stride, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace spec;
  MST_CHECK(x && taps && y && workspace, "fir_filtfilt: null pointer");
  MST_CHECK(n_signals > 0 && n_signals <= 65535 && n_taps >= 1 && n_taps <= kFirMaxTaps && stride >= T && y_stride >= T,
            "fir_filtfilt: bad arguments (n_taps <= %d)", kFirMaxTaps);
  const int P = 3 * n_taps;                                     // scipy's default padlen for a = [1]
  MST_CHECK(T > P, "fir_filtfilt: the signal (%lld samples) must be longer than padlen = %d", T, P);   // scipy raises ValueError
  MST_CHECK(workspace_bytes >= mst_fir_filtfilt_workspace_bytes(n_signals, T, n_taps), "fir_filtfilt: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done[64] = {false};
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    MST_CUDA_OK(cudaFuncSetAttribute(fir64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fir_smem_bytes(kFirMaxTaps)));
    MST_CUDA_OK(cudaFuncSetAttribute(fir64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fir_smem_bytes(kFirMaxTaps)));
    attr_done[dev] = true;
  }
  double* y1 = reinterpret_cast<double*>(workspace);
  const long long n_ext = T + 2LL * P;
  const size_t smem = fir_smem_bytes(n_taps);
  fir64_kernel<0><<<dim3((unsigned)((n_ext + kFirTile - 1) / kFirTile), n_signals), kThreads, smem, st>>>(
      x, nullptr, T, P, taps, n_taps, nullptr, y1, nullptr, stride, 0);
  if (launch_ok("fir64_kernel<0>")) return 1;
  fir64_kernel<1><<<dim3((unsigned)((T + kFirTile - 1) / kFirTile), n_signals), kThreads, smem, st>>>(
      nullptr, y1, T, P, taps, n_taps, scale, nullptr, y, 0, y_stride);
  return launch_ok("fir64_kernel<1>");
}

}  // extern "C"
