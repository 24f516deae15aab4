// encoder.cu -- FXencoder forward for sm_100a.
//
// Replaces (reference paths relative to /root/reference/mixing_style_transfer/networks/):
//   Conv1d_layer  = ReflectionPad1d -> Conv1d(bias) -> BatchNorm1d(eval) -> ReLU   network_utils.py:28-34,47-51,74,79-89
//   Res_ConvBlock = conv1(x) + x ; conv2(.)                                        network_utils.py:116-119
//   FXencoder.forward = 12 blocks -> AdaptiveAvgPool1d(1).squeeze(-1)              architectures.py:62-70
//
// Data layout: activations are fp32 [B][C][T] (PyTorch contiguous, time fastest).  Weights are BN-folded once
// (mst_conv1d_fold_bn) and stored TRANSPOSED as w[ci][k][co] so a CTA's co-tile is contiguous.
//
// Kernel: im2col-free direct convolution on the fp32 CUDA cores.  One CTA computes a (CO_TILE x T_TILE) output tile of
// one segment; lanes run along time (coalesced loads/stores, conflict-free shared reads), warps along output channels.
// The input window is staged in shared memory DE-INTERLEAVED by stride phase (position p -> [p % S][p / S]) so that a
// strided convolution still reads consecutive words per warp; reflection padding is resolved while staging.
// Epilogue fuses bias(+BN) -> ReLU -> residual add.
#include "common.cuh"
#include "f32x2.cuh"

namespace mst {

template <int K, int S, int COT, int TT, int CI_TILE>
struct EncTile {
  static constexpr int kWarps = 8;
  static constexpr int CO_TILE = kWarps * COT;
  static constexpr int T_TILE = 32 * TT;
  static constexpr int XW = T_TILE + (K - 1) / S + 1;        // per-phase row length
  static constexpr int WIN = (T_TILE - 1) * S + K;           // input positions needed by the tile
  static constexpr int SMEM_FLOATS = CI_TILE * K * CO_TILE + CI_TILE * S * XW;
};

__device__ __forceinline__ int reflect_index(int p, int t_in) {
  // nn.ReflectionPad1d: mirror without repeating the edge sample; clamp covers the don't-care tail of the last tile
  if (p < 0) p = -p;
  if (p >= t_in) p = 2 * (t_in - 1) - p;
  return min(max(p, 0), t_in - 1);
}

template <int K, int S, int COT, int TT, int CI_TILE>
__global__ void __launch_bounds__(256)
enc_conv1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ residual, float* __restrict__ y, int c_in, int t_in, int c_out,
                  int t_out, int pad_left, int relu) {
  using Tile = EncTile<K, S, COT, TT, CI_TILE>;
  __shared__ __align__(16) float smem[Tile::SMEM_FLOATS];
  float* Ws = smem;                                        // [CI_TILE][K][CO_TILE]
  float* Xs = smem + CI_TILE * K * Tile::CO_TILE;          // [CI_TILE][S][XW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_base = blockIdx.x * Tile::T_TILE;
  const int co_base = blockIdx.y * Tile::CO_TILE;
  const int b = blockIdx.z;
  const float* xb = x + (size_t)b * c_in * t_in;
  const int p0 = t_base * S - pad_left;

  float acc[COT][TT];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[c][i] = 0.f;

  for (int ci0 = 0; ci0 < c_in; ci0 += CI_TILE) {
    // ---- stage weights: contiguous co per (ci, k) ----
    for (int idx = tid; idx < CI_TILE * K * Tile::CO_TILE; idx += 256) {
      const int co = idx % Tile::CO_TILE;
      const int r = idx / Tile::CO_TILE;  // ci * K + j
      const int ci = r / K, j = r - ci * K;
      float v = 0.f;
      if (ci0 + ci < c_in && co_base + co < c_out) v = __ldg(w + ((size_t)(ci0 + ci) * K + j) * c_out + co_base + co);
      Ws[idx] = v;
    }
    // ---- stage the input window, de-interleaved by stride phase, reflection resolved here ----
    for (int idx = tid; idx < CI_TILE * Tile::WIN; idx += 256) {
      const int ci = idx / Tile::WIN, m = idx - ci * Tile::WIN;
      float v = 0.f;
      if (ci0 + ci < c_in) v = __ldg(xb + (size_t)(ci0 + ci) * t_in + reflect_index(p0 + m, t_in));
      Xs[(ci * S + (m % S)) * Tile::XW + m / S] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CI_TILE; ++ci) {
      const float* wrow = Ws + ci * K * Tile::CO_TILE + warp * COT;
      const float* xrow = Xs + ci * S * Tile::XW + lane;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float wv[COT];
#pragma unroll
        for (int c4 = 0; c4 < COT / 4; ++c4) {
          const float4 q = *reinterpret_cast<const float4*>(wrow + j * Tile::CO_TILE + c4 * 4);  // warp-broadcast
          wv[c4 * 4 + 0] = q.x; wv[c4 * 4 + 1] = q.y; wv[c4 * 4 + 2] = q.z; wv[c4 * 4 + 3] = q.w;
        }
        float xv[TT];
#pragma unroll
        for (int i = 0; i < TT; ++i) xv[i] = xrow[(j % S) * Tile::XW + (j / S) + 32 * i];
#pragma unroll
        for (int c = 0; c < COT; ++c)
#pragma unroll
          for (int i = 0; i < TT; ++i) acc[c][i] = fmaf(wv[c], xv[i], acc[c][i]);
      }
    }
    __syncthreads();
  }

  // ---- epilogue: folded bias -> ReLU -> residual ----
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = co_base + warp * COT + c;
    if (co >= c_out) continue;
    const float bv = __ldg(bias + co);
    const size_t row = ((size_t)b * c_out + co) * t_out;
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      const int t = t_base + lane + 32 * i;
      if (t >= t_out) continue;
      float v = acc[c][i] + bv;
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v += __ldg(residual + row + t);
      y[row + t] = v;
    }
  }
}

// Narrow-channel variant for the first encoder blocks (C_out <= 64, long time axis).  A thread owns COT output channels
// x TT CONSECUTIVE time steps and keeps the input window of the current input channel in registers (one LDS.128 stream
// per stride phase), so every staged input value is reused by all K taps and all COT channels from registers: the inner
// loop is FMA-bound instead of LDS-bound.  CTA = GROUPS channel groups x (256 / GROUPS) time threads.
template <int K, int S, int COT, int TT, int GROUPS, int CI_TILE>
struct EncNarrow {
  static constexpr int TPG = 256 / GROUPS;             // threads along time
  static constexpr int T_TILE = TPG * TT;
  static constexpr int CO_TILE = GROUPS * COT;
  static constexpr int PH = TT + (K - 1) / S + 1;      // per-phase window length held in registers
  static constexpr int XW = ((T_TILE + (K - 1) / S + 1 + 3) / 4) * 4 + 4;   // per-phase smem row (16-byte multiple)
  static constexpr int WIN = (T_TILE - 1) * S + K;
  static constexpr int SMEM_FLOATS = CI_TILE * K * CO_TILE + CI_TILE * S * XW;
};

template <int K, int S, int COT, int TT, int GROUPS, int CI_TILE>
__global__ void __launch_bounds__(256)
enc_conv1d_narrow_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ residual, float* __restrict__ y, int c_in, int t_in, int c_out,
                         int t_out, int pad_left, int relu) {
  using Tile = EncNarrow<K, S, COT, TT, GROUPS, CI_TILE>;
  static_assert(TT % 4 == 0 && COT % 2 == 0, "vector widths");
  __shared__ __align__(16) float smem[Tile::SMEM_FLOATS];
  float* Ws = smem;                                        // [CI_TILE][K][CO_TILE]
  float* Xs = smem + CI_TILE * K * Tile::CO_TILE;          // [CI_TILE][S][XW]
  const int tid = threadIdx.x;
  const int g = tid / Tile::TPG, tl = tid - g * Tile::TPG;
  const int t_base = blockIdx.x * Tile::T_TILE, b = blockIdx.y;
  const float* xb = x + (size_t)b * c_in * t_in;
  const int p0 = t_base * S - pad_left;

  // channel PAIRS ride in packed fma.rn.f32x2 (each half rounds like the scalar FMA: bit-identical sums, half the FMA
  // instructions, 1.46x the FMA rate on B200); the staged weight pair is the 64-bit operand as loaded, the input sample the
  // broadcast one
  f2::u64 acc[COT / 2][TT];
#pragma unroll
  for (int c = 0; c < COT / 2; ++c)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[c][i] = 0ull;

  for (int ci0 = 0; ci0 < c_in; ci0 += CI_TILE) {
    for (int idx = tid; idx < CI_TILE * K * Tile::CO_TILE; idx += 256) {
      const int co = idx % Tile::CO_TILE;
      const int r = idx / Tile::CO_TILE;
      const int ci = r / K, j = r - ci * K;
      float v = 0.f;
      if (ci0 + ci < c_in && co < c_out) v = __ldg(w + ((size_t)(ci0 + ci) * K + j) * c_out + co);
      Ws[idx] = v;
    }
    for (int idx = tid; idx < CI_TILE * Tile::WIN; idx += 256) {
      const int ci = idx / Tile::WIN, m = idx - ci * Tile::WIN;
      float v = 0.f;
      if (ci0 + ci < c_in) v = __ldg(xb + (size_t)(ci0 + ci) * t_in + reflect_index(p0 + m, t_in));
      Xs[(ci * S + (m % S)) * Tile::XW + m / S] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CI_TILE; ++ci) {
      // register window: for every stride phase the TT + (K-1)/S + 1 values this thread's outputs touch
      float xw[S][Tile::PH + 3];
#pragma unroll
      for (int ph = 0; ph < S; ++ph) {
        const float* src = Xs + (ci * S + ph) * Tile::XW + tl * TT;
#pragma unroll
        for (int q = 0; q < (Tile::PH + 3) / 4; ++q) {
          const float4 v4 = *reinterpret_cast<const float4*>(src + 4 * q);
          xw[ph][4 * q] = v4.x; xw[ph][4 * q + 1] = v4.y; xw[ph][4 * q + 2] = v4.z; xw[ph][4 * q + 3] = v4.w;
        }
      }
      const float* wrow = Ws + ci * K * Tile::CO_TILE + g * COT;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        f2::u64 wv[COT / 2];
#pragma unroll
        for (int c2 = 0; c2 < COT / 2; ++c2) wv[c2] = *reinterpret_cast<const f2::u64*>(wrow + j * Tile::CO_TILE + c2 * 2);
        f2::u64 xv[TT];
#pragma unroll
        for (int i = 0; i < TT; ++i) xv[i] = f2::dup(xw[j % S][i + j / S]);
#pragma unroll
        for (int c2 = 0; c2 < COT / 2; ++c2)
#pragma unroll
          for (int i = 0; i < TT; ++i) acc[c2][i] = f2::fma2(wv[c2], xv[i], acc[c2][i]);
      }
    }
    __syncthreads();
  }

  const int t0 = t_base + tl * TT;
  const bool vec = (t_out % 4) == 0;
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = g * COT + c;
    if (co >= c_out) continue;
    const float bv = __ldg(bias + co);
    const size_t row = ((size_t)b * c_out + co) * t_out;
#pragma unroll
    for (int i4 = 0; i4 < TT / 4; ++i4) {
      const int t = t0 + 4 * i4;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = ((c & 1) ? f2::hi_of(acc[c / 2][4 * i4 + e]) : f2::lo_of(acc[c / 2][4 * i4 + e])) + bv;
        if (relu) v[e] = fmaxf(v[e], 0.f);
      }
      if (vec && t + 3 < t_out) {
        if (residual) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(residual + row + t));
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        *reinterpret_cast<float4*>(y + row + t) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (t + e < t_out) y[row + t + e] = v[e] + (residual ? __ldg(residual + row + t + e) : 0.f);
        }
      }
    }
  }
}

// generic (any k / stride) variant: same tiling, runtime tap loop.  Used only for configs outside configs.yaml.
template <int COT, int TT, int CI_TILE>
__global__ void __launch_bounds__(256)
enc_conv1d_generic_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                          const float* __restrict__ residual, float* __restrict__ y, int c_in, int t_in, int c_out,
                          int t_out, int pad_left, int relu, int K, int S) {
  constexpr int CO_TILE = 8 * COT, T_TILE = 32 * TT;
  extern __shared__ __align__(16) float dsm[];
  const int XW = T_TILE + (K - 1) / S + 1, WIN = (T_TILE - 1) * S + K;
  float* Ws = dsm;
  float* Xs = dsm + CI_TILE * K * CO_TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_base = blockIdx.x * T_TILE, co_base = blockIdx.y * CO_TILE, b = blockIdx.z;
  const float* xb = x + (size_t)b * c_in * t_in;
  const int p0 = t_base * S - pad_left;
  float acc[COT][TT];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[c][i] = 0.f;
  for (int ci0 = 0; ci0 < c_in; ci0 += CI_TILE) {
    for (int idx = tid; idx < CI_TILE * K * CO_TILE; idx += 256) {
      const int co = idx % CO_TILE, r = idx / CO_TILE, ci = r / K, j = r - ci * K;
      float v = 0.f;
      if (ci0 + ci < c_in && co_base + co < c_out) v = __ldg(w + ((size_t)(ci0 + ci) * K + j) * c_out + co_base + co);
      Ws[idx] = v;
    }
    for (int idx = tid; idx < CI_TILE * WIN; idx += 256) {
      const int ci = idx / WIN, m = idx - ci * WIN;
      float v = 0.f;
      if (ci0 + ci < c_in) v = __ldg(xb + (size_t)(ci0 + ci) * t_in + reflect_index(p0 + m, t_in));
      Xs[(ci * S + (m % S)) * XW + m / S] = v;
    }
    __syncthreads();
    for (int ci = 0; ci < CI_TILE; ++ci) {
      for (int j = 0; j < K; ++j) {
        const float* wrow = Ws + (ci * K + j) * CO_TILE + warp * COT;
        const float* xrow = Xs + (ci * S + (j % S)) * XW + (j / S) + lane;
#pragma unroll
        for (int c = 0; c < COT; ++c)
#pragma unroll
          for (int i = 0; i < TT; ++i) acc[c][i] = fmaf(wrow[c], xrow[32 * i], acc[c][i]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = co_base + warp * COT + c;
    if (co >= c_out) continue;
    const float bv = __ldg(bias + co);
    const size_t row = ((size_t)b * c_out + co) * t_out;
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      const int t = t_base + lane + 32 * i;
      if (t >= t_out) continue;
      float v = acc[c][i] + bv;
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v += __ldg(residual + row + t);
      y[row + t] = v;
    }
  }
}

__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ bn_w,
                               const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                               const float* __restrict__ bn_var, float eps, int c_out, int c_in, int k,
                               float* __restrict__ w_out, float* __restrict__ b_out) {
  const size_t n = (size_t)c_out * c_in * k;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    // output index order: [ci][j][co]
    const int co = idx % c_out;
    const size_t r = idx / c_out;
    const int j = r % k;
    const int ci = r / k;
    const float s = bn_w[co] / sqrtf(bn_var[co] + eps);
    w_out[idx] = w[((size_t)co * c_in + ci) * k + j] * s;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < c_out; co += gridDim.x * blockDim.x) {
    const float s = bn_w[co] / sqrtf(bn_var[co] + eps);
    b_out[co] = ((b ? b[co] : 0.f) - bn_mean[co]) * s + bn_b[co];
  }
}

__global__ void mean_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int T) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* p = x + (size_t)warp * T;
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += p[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[warp] = s / (float)T;
}

template <int K, int S, int COT, int TT, int CI_TILE>
static int launch_conv(const float* x, const float* w, const float* b, const float* res, float* y, int B, int c_in,
                       int t_in, int c_out, int t_out, int pad_left, int relu, cudaStream_t st) {
  using Tile = EncTile<K, S, COT, TT, CI_TILE>;
  static_assert(Tile::SMEM_FLOATS * 4 <= 48 * 1024, "static shared memory budget");
  dim3 grid(cdiv(t_out, Tile::T_TILE), cdiv(c_out, Tile::CO_TILE), B);
  enc_conv1d_kernel<K, S, COT, TT, CI_TILE><<<grid, 256, 0, st>>>(x, w, b, res, y, c_in, t_in, c_out, t_out, pad_left, relu);
  return launch_ok("enc_conv1d_kernel");
}

template <int K, int S, int CI_TILE>
static int launch_conv_by_shape(const float* x, const float* w, const float* b, const float* res, float* y, int B,
                                int c_in, int t_in, int c_out, int t_out, int pad_left, int relu, cudaStream_t st) {
  // long time axis: 8 co x 4 t per thread (64 x 128 tile); short time axis / many channels: 16 co x 2 t (128 x 64)
  if (t_out > 64 || c_out < 128)
    return launch_conv<K, S, 8, 4, CI_TILE>(x, w, b, res, y, B, c_in, t_in, c_out, t_out, pad_left, relu, st);
  constexpr int CI_WIDE = (K * CI_TILE * 128 > 10000) ? CI_TILE / 2 : CI_TILE;  // 128-wide co tile: keep Ws <= 40 KB
  return launch_conv<K, S, 16, 2, CI_WIDE>(x, w, b, res, y, B, c_in, t_in, c_out, t_out, pad_left, relu, st);
}

static int conv1d_dispatch(const float* x, const float* w, const float* b, const float* res, float* y, int B, int c_in,
                           int t_in, int c_out, int k, int stride, int relu, cudaStream_t st) {
  MST_CHECK(B > 0 && c_in > 0 && c_out > 0 && t_in > 0 && k > 0 && stride > 0, "enc_conv1d: bad shape");
  const int pad = k - 1, pad_left = pad / 2;  // network_utils.py:31-34 (dilation 1)
  MST_CHECK(t_in > pad - pad_left, "enc_conv1d: reflection padding (%d,%d) needs T_in > pad, got T_in=%d", pad_left,
            pad - pad_left, t_in);
  const int t_out = (t_in + pad - k) / stride + 1;  // == ceil(t_in / stride)
  // narrow-channel layers of blocks 0-2 (inference/configs.yaml): {K, S, COT, TT, GROUPS, CI_TILE} by (k, stride, c_out)
#define MST_NARROW_CASE(KK, SS, CO, COT, TT, GR, CI)                                                                  \
  if (k == KK && stride == SS && c_out == CO) {                                                                        \
    using Tile = EncNarrow<KK, SS, COT, TT, GR, CI>;                                                                   \
    static_assert(Tile::SMEM_FLOATS * 4 <= 48 * 1024, "static shared memory budget");                                  \
    dim3 grid(cdiv(t_out, Tile::T_TILE), B);                                                                           \
    enc_conv1d_narrow_kernel<KK, SS, COT, TT, GR, CI><<<grid, 256, 0, st>>>(x, w, b, res, y, c_in, t_in, c_out, t_out,  \
                                                                          pad_left, relu);                           \
    return launch_ok("enc_conv1d_narrow_kernel");                                                                     \
  }
  if (t_out >= 2048) {
    MST_NARROW_CASE(25, 1, 2, 2, 8, 1, 2)      // block 0 conv1:  2 ->  2
    MST_NARROW_CASE(25, 4, 16, 16, 4, 1, 2)    // block 0 conv2:  2 -> 16
    MST_NARROW_CASE(25, 1, 16, 16, 4, 1, 4)    // block 1 conv1: 16 -> 16
    MST_NARROW_CASE(25, 4, 32, 16, 4, 2, 4)    // block 1 conv2: 16 -> 32
    MST_NARROW_CASE(15, 1, 32, 16, 4, 2, 8)    // block 2 conv1: 32 -> 32
    MST_NARROW_CASE(15, 2, 64, 16, 4, 4, 8)    // block 2 conv2: 32 -> 64
  }
#undef MST_NARROW_CASE
#define MST_CONV_CASE(KK, SS, CI) \
  if (k == KK && stride == SS)    \
    return launch_conv_by_shape<KK, SS, CI>(x, w, b, res, y, B, c_in, t_in, c_out, t_out, pad_left, relu, st);
  MST_CONV_CASE(25, 4, 4)
  MST_CONV_CASE(25, 1, 4)
  MST_CONV_CASE(15, 2, 8)
  MST_CONV_CASE(15, 1, 8)
  MST_CONV_CASE(10, 2, 8)
  MST_CONV_CASE(10, 1, 8)
  MST_CONV_CASE(5, 2, 8)
  MST_CONV_CASE(5, 1, 8)
#undef MST_CONV_CASE
  {  // any other (k, stride): generic kernel, dynamic shared memory
    constexpr int COT = 8, TT = 4, CI_TILE = 2;
    const int XW = 32 * TT + (k - 1) / stride + 1;
    const size_t smem = (size_t)(CI_TILE * k * 8 * COT + CI_TILE * stride * XW) * sizeof(float);
    MST_CHECK(smem <= 200 * 1024, "enc_conv1d: kernel_size %d / stride %d too large for the generic kernel", k, stride);
    auto kern = enc_conv1d_generic_kernel<COT, TT, CI_TILE>;
    if (smem > 48 * 1024) MST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(t_out, 32 * TT), cdiv(c_out, 8 * COT), B);
    kern<<<grid, 256, smem, st>>>(x, w, b, res, y, c_in, t_in, c_out, t_out, pad_left, relu, k, stride);
    return launch_ok("enc_conv1d_generic_kernel");
  }
}

// Packed layout (byte offsets, 1024-aligned).  Blocks before `first_umma` run on the fp32 CUDA cores and store
// conv {w[ci][k][co] fp32, b[co]}; blocks from `first_umma` on run on tcgen05 (enc_umma.cu) and store
// conv {w[tap][kc][hi|lo][co][64] bf16, b[co]}.
struct EncOffsets {
  size_t w1[MST_MAX_ENC_BLOCKS], b1[MST_MAX_ENC_BLOCKS], w2[MST_MAX_ENC_BLOCKS], b2[MST_MAX_ENC_BLOCKS], total;
  int first_umma;  // == n_blocks when no block is eligible
};
static bool enc_force_fp32() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MST_ENC_FP32"); v = (e && atoi(e) != 0) ? 1 : 0; }
  return v == 1;
}
static int enc_offsets(const mst_enc_config* cfg, EncOffsets* o) {
  MST_CHECK(cfg && cfg->n_blocks > 0 && cfg->n_blocks <= MST_MAX_ENC_BLOCKS, "enc config: n_blocks out of range");
  // the tensor-core path starts at the first block from which EVERY later conv is eligible
  int first = cfg->n_blocks;
  for (int i = cfg->n_blocks - 1; i >= 0; --i) {
    const int ci = cfg->channels[i], co = cfg->channels[i + 1], k = cfg->kernels[i], s = cfg->strides[i];
    if (enc_umma_eligible(ci, ci, k, 1) && enc_umma_eligible(ci, co, k, s)) first = i; else break;
  }
  if (enc_force_fp32()) first = cfg->n_blocks;
  if (const char* e = getenv("MST_ENC_FIRST_UMMA")) { const int v = atoi(e); if (v > first && v <= cfg->n_blocks) first = v; }  // debug
  o->first_umma = first;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes, 1024); return r; };
  for (int i = 0; i < cfg->n_blocks; ++i) {
    const size_t ci = cfg->channels[i], co = cfg->channels[i + 1], k = cfg->kernels[i];
    MST_CHECK(ci > 0 && co > 0 && k > 0 && cfg->strides[i] > 0, "enc config: bad block %d", i);
    o->w1[i] = take(ci * ci * k * 4);
    o->b1[i] = take((ci + 2) * 4);   // + 1/weight-scale + scratch (tensor-core blocks)
    o->w2[i] = take(co * ci * k * 4);
    o->b2[i] = take((co + 2) * 4);
  }
  o->total = off;
  return 0;
}

static void conv_halo(int k, int* l, int* r) { *l = (k - 1) / 2; *r = (k - 1) - *l; }

// activation sizes of the two phases
static size_t enc_max_act_f32(const mst_enc_config* cfg, const EncOffsets& o, int B, int L) {
  size_t mx = 64;
  int t = L;
  for (int i = 0; i < o.first_umma; ++i) {
    const size_t a1 = (size_t)B * cfg->channels[i] * t;
    t = cdiv(t, cfg->strides[i]);
    const size_t a2 = (size_t)B * cfg->channels[i + 1] * t;
    mx = a1 > mx ? a1 : mx;
    mx = a2 > mx ? a2 : mx;
  }
  return align_up(mx * sizeof(float), 1024);
}
static size_t enc_max_act_split(const mst_enc_config* cfg, const EncOffsets& o, int B, int L) {
  size_t mx = 0;
  int t = L;
  for (int i = 0; i < cfg->n_blocks; ++i) {
    const int tin = t;
    t = cdiv(t, cfg->strides[i]);
    if (i < o.first_umma) continue;
    // every buffer of block i: input / c1 (T_in rows) and output (T_out rows), each with at most a 32-row halo
    const size_t a1 = (size_t)B * (tin + 64) * cfg->channels[i] * 4;
    const size_t a2 = (size_t)B * (t + 64) * cfg->channels[i + 1] * 4;
    mx = a1 > mx ? a1 : mx;
    mx = a2 > mx ? a2 : mx;
  }
  return align_up(mx, 1024);
}

}  // namespace mst

using namespace mst;

namespace mst {
// out[c] = scale * sum_r x[r][c], rows added in order (bit-reproducible): the mean over the reference segments' embeddings
// (inference/style_transfer.py:152-153) without an eager torch reduction on the path
__global__ void __launch_bounds__(256) rows_reduce_kernel(const float* __restrict__ x, int rows, int cols, float scale,
                                                           float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += __ldg(x + (size_t)r * cols + c);
  out[c] = s * scale;
}
}  // namespace mst

extern "C" {

int mst_conv1d_fold_bn(const float* w, const float* b, const float* bn_w, const float* bn_b, const float* bn_mean,
                       const float* bn_var, float eps, int c_out, int c_in, int k, float* w_out, float* b_out,
                       void* stream) {
  MST_CHECK(w && bn_w && bn_b && bn_mean && bn_var && w_out && b_out, "fold_bn: null pointer");
  const size_t n = (size_t)c_out * c_in * k;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  fold_bn_kernel<<<blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(w, b, bn_w, bn_b, bn_mean, bn_var, eps, c_out,
                                                                            c_in, k, w_out, b_out);
  return launch_ok("fold_bn_kernel");
}

int mst_enc_conv1d(const float* x, const float* w_folded, const float* b_folded, const float* residual, float* y, int B,
                   int c_in, int t_in, int c_out, int k, int stride, int relu, void* stream) {
  MST_CHECK(x && w_folded && b_folded && y, "enc_conv1d: null pointer");
  MST_CHECK(B <= 65535, "enc_conv1d: batch %d exceeds grid.z", B);
  return conv1d_dispatch(x, w_folded, b_folded, residual, y, B, c_in, t_in, c_out, k, stride, relu, (cudaStream_t)stream);
}

int mst_enc_mean_pool(const float* x, float* y, int B, int C, int T, void* stream) {
  MST_CHECK(x && y && B > 0 && C > 0 && T > 0, "mean_pool: bad arguments");
  const int rows = B * C;
  mean_pool_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, y, rows, T);
  return launch_ok("mean_pool_kernel");
}

int mst_rows_reduce(const float* x, int rows, int cols, float scale, float* out, void* stream) {
  MST_CHECK(x && out && rows > 0 && cols > 0, "rows_reduce: bad arguments");
  rows_reduce_kernel<<<cdiv(cols, 256), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, scale, out);
  return launch_ok("rows_reduce_kernel");
}

size_t mst_enc_packed_bytes(const mst_enc_config* cfg) {
  EncOffsets o;
  if (enc_offsets(cfg, &o)) return 0;
  return o.total;
}

int mst_enc_pack(const mst_enc_config* cfg, const float* const* raw, float* packed_f, void* stream) {
  EncOffsets o;
  if (enc_offsets(cfg, &o)) return 1;
  MST_CHECK(raw && packed_f, "enc_pack: null pointer");
  MST_CHECK((reinterpret_cast<uintptr_t>(packed_f) & 1023) == 0, "enc_pack: packed buffer must be 1024-byte aligned");
  uint8_t* packed = reinterpret_cast<uint8_t*>(packed_f);
  for (int i = 0; i < cfg->n_blocks; ++i) {
    const float* const* r = raw + 12 * i;
    const int ci = cfg->channels[i], co = cfg->channels[i + 1], k = cfg->kernels[i];
    if (i < o.first_umma) {
      if (mst_conv1d_fold_bn(r[0], r[1], r[2], r[3], r[4], r[5], 1e-5f, ci, ci, k, (float*)(packed + o.w1[i]),
                             (float*)(packed + o.b1[i]), stream)) return 1;
      if (mst_conv1d_fold_bn(r[6], r[7], r[8], r[9], r[10], r[11], 1e-5f, co, ci, k, (float*)(packed + o.w2[i]),
                             (float*)(packed + o.b2[i]), stream)) return 1;
    } else {
      if (enc_umma_pack(r[0], r[1], r[2], r[3], r[4], r[5], ci, ci, k, packed + o.w1[i], (float*)(packed + o.b1[i]),
                        (cudaStream_t)stream)) return 1;
      if (enc_umma_pack(r[6], r[7], r[8], r[9], r[10], r[11], co, ci, k, packed + o.w2[i], (float*)(packed + o.b2[i]),
                        (cudaStream_t)stream)) return 1;
    }
  }
  return 0;
}

size_t mst_enc_workspace_bytes(const mst_enc_config* cfg, int B, int L) {
  EncOffsets o;
  if (!cfg || B <= 0 || L <= 0 || enc_offsets(cfg, &o)) return 0;
  return 3 * enc_max_act_f32(cfg, o, B, L) + 3 * enc_max_act_split(cfg, o, B, L) + 1024;
}

int mst_enc_forward(const mst_enc_config* cfg, const float* packed_f, const float* x, int B, int L, float* emb,
                    void* workspace, size_t workspace_bytes, void* stream) {
  EncOffsets o;
  if (enc_offsets(cfg, &o)) return 1;
  MST_CHECK(packed_f && x && emb && workspace, "enc_forward: null pointer");
  MST_CHECK(B > 0 && L > 0 && B <= 65535, "enc_forward: bad shape B=%d L=%d", B, L);
  MST_CHECK(workspace_bytes >= mst_enc_workspace_bytes(cfg, B, L), "enc_forward: workspace too small (%zu < %zu)",
            workspace_bytes, mst_enc_workspace_bytes(cfg, B, L));
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(packed_f);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const size_t fstride = enc_max_act_f32(cfg, o, B, L), sstride = enc_max_act_split(cfg, o, B, L);
  float* fbuf[3] = {(float*)ws, (float*)(ws + fstride), (float*)(ws + 2 * fstride)};
  uint8_t* sbuf[3] = {ws + 3 * fstride, ws + 3 * fstride + sstride, ws + 3 * fstride + 2 * sstride};

  // ---- phase 1: small-channel blocks on the fp32 CUDA cores, [B][C][T] ----
  const float* cur = x;
  int cur_buf = -1, t = L;
  for (int i = 0; i < o.first_umma; ++i) {
    const int ci = cfg->channels[i], co = cfg->channels[i + 1], k = cfg->kernels[i], s = cfg->strides[i];
    float* c1 = fbuf[(cur_buf + 1) % 3];
    float* c2 = fbuf[(cur_buf + 2) % 3];
    // c1 = relu(bn(conv1(x))) + x          (Res_ConvBlock, network_utils.py:117)
    if (mst_enc_conv1d(cur, (const float*)(packed + o.w1[i]), (const float*)(packed + o.b1[i]), cur, c1, B, ci, t, ci, k,
                       1, 1, stream)) return 1;
    // c2 = relu(bn(conv2(c1)))             (:118)
    if (mst_enc_conv1d(c1, (const float*)(packed + o.w2[i]), (const float*)(packed + o.b2[i]), nullptr, c2, B, ci, t, co,
                       k, s, 1, stream)) return 1;
    t = cdiv(t, s);
    cur = c2;
    cur_buf = (cur_buf + 2) % 3;
  }
  if (o.first_umma == cfg->n_blocks) return mst_enc_mean_pool(cur, emb, B, cfg->channels[cfg->n_blocks], t, stream);

  // ---- hand-over: fp32 [B][C][T] -> split rows carrying the first tensor-core block's reflection halo ----
  int l, r;
  conv_halo(cfg->kernels[o.first_umma], &l, &r);
  MST_CHECK(t > r, "enc_forward: length %d too short for reflection padding (%d,%d) at block %d", t, l, r, o.first_umma);
  if (enc_split_from_f32(cur, sbuf[0], B, cfg->channels[o.first_umma], t, l, r, st)) return 1;

  // ---- phase 2: tcgen05 blocks on split rows ----
  int sb = 0;
  for (int i = o.first_umma; i < cfg->n_blocks; ++i) {
    const int ci = cfg->channels[i], co = cfg->channels[i + 1], k = cfg->kernels[i], s = cfg->strides[i];
    conv_halo(k, &l, &r);
    int nl = 0, nr = 0;
    if (i + 1 < cfg->n_blocks) conv_halo(cfg->kernels[i + 1], &nl, &nr);
    uint8_t* xin = sbuf[sb];
    uint8_t* c1 = sbuf[(sb + 1) % 3];
    uint8_t* c2 = sbuf[(sb + 2) % 3];
    if (enc_umma_conv(xin, packed + o.w1[i], (const float*)(packed + o.b1[i]), c1, B, ci, t, ci, k, 1, true, l, r, st))
      return 1;
    if (enc_umma_conv(c1, packed + o.w2[i], (const float*)(packed + o.b2[i]), c2, B, ci, t, co, k, s, false, nl, nr, st))
      return 1;
    t = cdiv(t, s);
    sb = (sb + 2) % 3;
    if (i + 1 == cfg->n_blocks)
      return enc_pool_split(c2, emb, B, co, t, 0, t, st);
  }
  return 0;
}

}  // extern "C"
