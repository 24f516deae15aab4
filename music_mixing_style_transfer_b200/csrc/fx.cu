// fx.cu -- batched FX chain (parametric EQ -> compressor -> mid/side imager -> gain) for sm_100a.
//
// Replaces (reference paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/):
//   AugmentationChain.__call__ / apply_processor   common_audioeffects.py:156-192, 115-148 (RMS re-normalisation :142-145)
//   Equaliser.process (5 cascaded RBJ biquads, float64, state reset per band)        :501-525, filters :438-462
//   compressor_process / Compressor.process (log-domain gain computer + 1-pole attack/release smoother)   :529-587, :624-652
//   MidSideImager.process                           :965-992
//   Gain.process                                    :1038-1051
//
// Layout: x, y fp32 [B][2][L] (channel-major); params fp32 [B][20]; stats double [B][16].
// EQ: one CTA (256 threads) per (segment, channel).  Compressor: one CTA per segment, 512 threads (threads 0-255 own
// channel 0, 256-511 channel 1; it needs the cross-channel sum L*R).  A CTA streams its data in tiles of 4096 frames;
// inside a tile every thread owns 16 consecutive samples of its channel.
//
// The two recurrences are made time-parallel without changing their result:
//  * EQ: each biquad is a 2-state linear system.  Per tile and biquad: (1) every thread runs its 16 samples from zero
//    state (float64, DF-II-T like scipy.signal.lfilter), (2) the per-chunk end states are combined with a block-wide
//    scan of the constant 2x2 chunk-transition matrix (warp shuffles + one shared-memory hop), (3) each thread adds the
//    natural response to its true incoming state.  The tile's end state carries to the next tile.
//  * Compressor smoother  y[i] = a_i*y[i-1] + (1-a_i)*x_l[i],  a_i = (x_l[i] > y[i-1]) ? alpha_att : alpha_rel
//    is a max of two increasing affine maps per step (alpha_att < alpha_rel), so the exact solution is the fixed point
//    of "guess the attack/release pattern -> solve the now-linear recurrence by an affine block scan -> re-derive the
//    pattern".  Iterating from the tile's exact incoming state the correct prefix grows every round, so the loop
//    terminates with the sequential result (up to float64 re-association).
// Each of EQ / compressor / imager is followed by a whole-segment RMS re-normalisation, which forces three passes
// (EQ | compressor | imager+gain); the imager's own energies and its RMS factor follow analytically from the
// compressor-pass sums (sum L^2, sum R^2, sum L*R), so imager + gain are a single element-wise pass.
#include "common.cuh"

#include <cstdlib>

namespace mst {

constexpr int kFxThreads = 512;
constexpr int kFxHalf = 256;          // threads per channel
constexpr int kFxChunk = 16;          // samples per thread per tile
constexpr int kFxTile = kFxHalf * kFxChunk;  // 4096 frames
constexpr int kFxStats = 16;          // doubles per segment

// stats slots
enum { S_X2_0 = 0, S_X2_1, S_Y1_0, S_Y1_1, S_U2_0, S_U2_1, S_Y2_0, S_Y2_1, S_LR, S_ROUNDS };  // S_ROUNDS: smoother iterations (diagnostic)

struct Biquad { double b0, b1, b2, a1, a2; };

__device__ Biquad rbj(double G, double Q, double fc, double rate, int type) {  // 0 low shelf, 1 peaking, 2 high shelf
  const double A = pow(10.0, G / 40.0);
  const double w0 = 2.0 * 3.14159265358979323846 * (fc / rate);
  const double alpha = sin(w0) / (2.0 * Q), c = cos(w0), s = 2.0 * sqrt(A) * alpha;
  double b0, b1, b2, a0, a1, a2;
  if (type == 1) {
    b0 = 1.0 + alpha * A; b1 = -2.0 * c; b2 = 1.0 - alpha * A;
    a0 = 1.0 + alpha / A; a1 = -2.0 * c; a2 = 1.0 - alpha / A;
  } else if (type == 0) {
    b0 = A * ((A + 1) - (A - 1) * c + s); b1 = 2 * A * ((A - 1) - (A + 1) * c); b2 = A * ((A + 1) - (A - 1) * c - s);
    a0 = (A + 1) + (A - 1) * c + s; a1 = -2 * ((A - 1) + (A + 1) * c); a2 = (A + 1) + (A - 1) * c - s;
  } else {
    b0 = A * ((A + 1) + (A - 1) * c + s); b1 = -2 * A * ((A - 1) + (A + 1) * c); b2 = A * ((A + 1) + (A - 1) * c - s);
    a0 = (A + 1) - (A - 1) * c + s; a1 = 2 * ((A - 1) - (A + 1) * c); a2 = (A + 1) - (A - 1) * c - s;
  }
  Biquad q;
  q.b0 = b0 / a0; q.b1 = b1 / a0; q.b2 = b2 / a0; q.a1 = a1 / a0; q.a2 = a2 / a0;
  return q;
}

// tile loads / stores of 16 consecutive samples per thread (float4 when the row is 16-byte aligned)
__device__ __forceinline__ void load_chunk(const float* __restrict__ row, int L, int s0, bool vec, float (&v)[kFxChunk]) {
  if (vec && s0 + kFxChunk <= L) {
#pragma unroll
    for (int i = 0; i < kFxChunk / 4; ++i) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(row + s0) + i);
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i) v[i] = (s0 + i < L) ? __ldg(row + s0 + i) : 0.f;
  }
}
__device__ __forceinline__ void store_chunk(float* __restrict__ row, int L, int s0, bool vec, const float (&v)[kFxChunk]) {
  if (vec && s0 + kFxChunk <= L) {
#pragma unroll
    for (int i = 0; i < kFxChunk / 4; ++i)
      reinterpret_cast<float4*>(row + s0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i)
      if (s0 + i < L) row[s0 + i] = v[i];
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// =====================================================================================================================
// pass A: EQ
// Precision split: the recurrences have long memory (pole radius up to 0.997), so everything that is CARRIED -- the
// per-chunk end states, the block scan of transition matrices, the tile carry -- is float64; everything LOCAL to a
// 16-sample chunk (zero-state response, natural-response correction) is float32, where rounding cannot accumulate for
// more than 16 steps (the double-integrator noise growth of a low-frequency section over 16 float32 steps is
// ~16^2/2 * 2^-23 ~ 1.5e-5 relative).  Measured against the float64 scipy cascade + reference compressor/imager:
// 2e-6 ... 1.1e-5 RMS absolute on signals of RMS 0.1-0.2 (budget 1e-4), tests/test_gpu_fx.py.
// =====================================================================================================================
struct EqShared {
  double mpow[5][33][4];     // (16-sample chunk transition)^l, l = 0..32, row-major 2x2
  float2 h[5][kFxChunk];     // output natural response to unit initial state (s1, s2)
  double2 hd[2][kFxChunk];   // float64 copy for the two low-frequency sections
  double coefd[2][5];        // float64 b0 b1 b2 a1 a2 of the two low-frequency sections
  float coef[5][8];          // b0 b1 b2 a1 a2 as float
  double carry[2][5][2];     // [tile parity][biquad]: state entering the tile
  double wtot[2][5][8][2];   // [tile parity][biquad][warp]: warp totals of the scan
  double red[8][2];
};

// one CTA (256 threads) per (segment, channel): the EQ needs no cross-channel term
__global__ void __launch_bounds__(kFxHalf)
fx_eq_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ y,
             double* __restrict__ stats, int L, float sample_rate, int enable) {
  __shared__ EqShared sh;
  const int b = blockIdx.x >> 1, ch = blockIdx.x & 1, tid = threadIdx.x;
  const int ct = tid, lane = tid & 31, wic = tid >> 5;
  const float* p = params + (size_t)b * MST_FX_NPARAMS;

  if (tid < 5) {
    // dict order low_shelf, first, second, third, high_shelf (:391); shelves Q = 0.707 (:454)
    const int gi[5] = {0, 2, 5, 8, 11}, fi[5] = {1, 3, 6, 9, 12}, qi[5] = {-1, 4, 7, 10, -1}, ty[5] = {0, 1, 1, 1, 2};
    const double Q = qi[tid] < 0 ? 0.707 : (double)p[qi[tid]];
    const Biquad q = rbj((double)p[gi[tid]], Q, (double)p[fi[tid]], (double)sample_rate, ty[tid]);
    sh.coef[tid][0] = (float)q.b0; sh.coef[tid][1] = (float)q.b1; sh.coef[tid][2] = (float)q.b2;
    sh.coef[tid][3] = (float)q.a1; sh.coef[tid][4] = (float)q.a2;
    if (tid < 2) { sh.coefd[tid][0] = q.b0; sh.coefd[tid][1] = q.b1; sh.coefd[tid][2] = q.b2; sh.coefd[tid][3] = q.a1; sh.coefd[tid][4] = q.a2; }
    // natural response + chunk transition from the two unit states (float64)
    double M[4];
    for (int u = 0; u < 2; ++u) {
      double s1 = u == 0 ? 1.0 : 0.0, s2 = u == 0 ? 0.0 : 1.0;
      for (int i = 0; i < kFxChunk; ++i) {
        const double yy = s1;
        if (u == 0) sh.h[tid][i].x = (float)yy; else sh.h[tid][i].y = (float)yy;
        if (tid < 2) { if (u == 0) sh.hd[tid][i].x = yy; else sh.hd[tid][i].y = yy; }
        s1 = s2 - q.a1 * yy;
        s2 = -q.a2 * yy;
      }
      M[0 + u] = s1;  // column u of the transition
      M[2 + u] = s2;
    }
    double P[4] = {1.0, 0.0, 0.0, 1.0};
    for (int l = 0; l <= 32; ++l) {
      for (int e = 0; e < 4; ++e) sh.mpow[tid][l][e] = P[e];
      const double n0 = M[0] * P[0] + M[1] * P[2], n1 = M[0] * P[1] + M[1] * P[3];
      const double n2 = M[2] * P[0] + M[3] * P[2], n3 = M[2] * P[1] + M[3] * P[3];
      P[0] = n0; P[1] = n1; P[2] = n2; P[3] = n3;
    }
    sh.carry[0][tid][0] = 0.0; sh.carry[0][tid][1] = 0.0;  // state reset (:512)
  }
  __syncthreads();

  const float* xrow = x + ((size_t)b * 2 + ch) * L;
  float* yrow = y + ((size_t)b * 2 + ch) * L;
  const bool vec = (L % 4) == 0;
  double sum_x2 = 0.0, sum_y2 = 0.0;
  int par = 0;

  for (int tile0 = 0; tile0 < L; tile0 += kFxTile, par ^= 1) {
    const int s0 = tile0 + ct * kFxChunk;
    float v[kFxChunk];
    load_chunk(xrow, L, s0, vec, v);
    float sx = 0.f;
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i) sx = fmaf(v[i], v[i], sx);
    sum_x2 += (double)sx;
    if (enable) {
#pragma unroll 1
      for (int k = 0; k < 5; ++k) {
        // Sections 0-1 (low shelf 30-200 Hz, first band 200-1000 Hz) act as double integrators over a 16-sample chunk:
        // float32 rounding noise grows ~n^2/2 there (up to 5e-5 absolute after the imager).  They run in float64; the
        // three sections above 1 kHz are well conditioned over 16 steps and run in float32.
        const bool dbl = k < 2;
        double yd[kFxChunk];
        double z1, z2;
        if (dbl) {
          const double b0 = sh.coefd[k][0], b1 = sh.coefd[k][1], b2 = sh.coefd[k][2], a1 = sh.coefd[k][3], a2 = sh.coefd[k][4];
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int i = 0; i < kFxChunk; ++i) {
            const double xi = (double)v[i];
            const double yi = fma(b0, xi, s1);
            s1 = fma(b1, xi, fma(-a1, yi, s2));
            s2 = fma(b2, xi, -a2 * yi);
            yd[i] = yi;
          }
          z1 = s1; z2 = s2;
        } else {
          const float b0 = sh.coef[k][0], b1 = sh.coef[k][1], b2 = sh.coef[k][2], a1 = sh.coef[k][3], a2 = sh.coef[k][4];
          // (1) zero-state response of this chunk (DF-II transposed like scipy.signal.lfilter; float32, local)
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < kFxChunk; ++i) {
            const float xi = v[i];
            const float yi = fmaf(b0, xi, s1);
            s1 = fmaf(b1, xi, fmaf(-a1, yi, s2));
            s2 = fmaf(b2, xi, -a2 * yi);
            v[i] = yi;
          }
          z1 = (double)s1; z2 = (double)s2;
        }
        // (2) inclusive scan over the warp (float64): I_l = sum_{i<=l} M^(l-i) z_i
        double i1 = z1, i2 = z2;
#pragma unroll
        for (int st = 0; st < 5; ++st) {
          const int off = 1 << st;
          const double o1 = __shfl_up_sync(0xffffffffu, i1, off), o2 = __shfl_up_sync(0xffffffffu, i2, off);
          if (lane >= off) {
            const double* Mp = sh.mpow[k][off];
            i1 += Mp[0] * o1 + Mp[1] * o2;
            i2 += Mp[2] * o1 + Mp[3] * o2;
          }
        }
        if (lane == 31) { sh.wtot[par][k][wic][0] = i1; sh.wtot[par][k][wic][1] = i2; }
        double e1 = __shfl_up_sync(0xffffffffu, i1, 1), e2 = __shfl_up_sync(0xffffffffu, i2, 1);
        if (lane == 0) { e1 = 0.0; e2 = 0.0; }
        __syncthreads();
        // state entering this warp: Q_0 = carry, Q_{w+1} = M^32 Q_w + W_w
        double q1 = sh.carry[par][k][0], q2 = sh.carry[par][k][1];
        const double* M32 = sh.mpow[k][32];
        for (int w = 0; w < wic; ++w) {
          const double n1 = M32[0] * q1 + M32[1] * q2 + sh.wtot[par][k][w][0];
          const double n2 = M32[2] * q1 + M32[3] * q2 + sh.wtot[par][k][w][1];
          q1 = n1; q2 = n2;
        }
        const double* Ml = sh.mpow[k][lane];
        const double in1 = Ml[0] * q1 + Ml[1] * q2 + e1, in2 = Ml[2] * q1 + Ml[3] * q2 + e2;
        // (3) add the natural response to the true incoming state
        if (dbl) {
#pragma unroll
          for (int i = 0; i < kFxChunk; ++i) {
            const double2 hh = sh.hd[k][i];
            v[i] = (float)(yd[i] + fma(hh.x, in1, hh.y * in2));
          }
        } else {
          const float f1 = (float)in1, f2 = (float)in2;
#pragma unroll
          for (int i = 0; i < kFxChunk; ++i) {
            const float2 hh = sh.h[k][i];
            v[i] = fmaf(hh.x, f1, fmaf(hh.y, f2, v[i]));
          }
        }
        if (ct == 255) {   // state leaving the tile = M * in + z of the last chunk; read by the NEXT tile (other parity)
          const double* M1 = sh.mpow[k][1];
          sh.carry[par ^ 1][k][0] = M1[0] * in1 + M1[1] * in2 + z1;
          sh.carry[par ^ 1][k][1] = M1[2] * in1 + M1[3] * in2 + z2;
        }
      }
    }
    float sy = 0.f;
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i)
      if (s0 + i < L) sy = fmaf(v[i], v[i], sy);
    sum_y2 += (double)sy;
    store_chunk(yrow, L, s0, vec, v);
  }

  sum_x2 = warp_sum(sum_x2);
  sum_y2 = warp_sum(sum_y2);
  __syncthreads();
  if (lane == 0) { sh.red[wic][0] = sum_x2; sh.red[wic][1] = sum_y2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh.red[w][0]; c += sh.red[w][1]; }
    stats[(size_t)b * kFxStats + S_X2_0 + ch] = a;
    stats[(size_t)b * kFxStats + S_Y1_0 + ch] = c;
  }
}

// =====================================================================================================================
// pass B: compressor (in place on y)
// Same precision split: the smoother state that is carried from chunk to chunk and tile to tile is float64 (chunk maps
// y_out = A*y_in + B with A = alpha_att^na * alpha_rel^nr taken from float64 power tables); the 16 local steps, the
// gain computer and the gain itself are float32 (the reference evaluates log10 in float32 as well).
// =====================================================================================================================
struct CompShared {
  double wa[2][2][8], wb[2][2][8];   // [round parity][channel][warp]: warp-total affine maps
  double carry[2][2];                // [tile parity][channel]: smoother state entering the tile
  double pa[kFxChunk + 1], pr[kFxChunk + 1];   // alpha_att^k, alpha_rel^k
  alignas(16) float xch[kFxTile];    // channel-1 output of the current tile (for sum L*R)
  double red[16][4];
  unsigned long long rounds;
};

__global__ void __launch_bounds__(kFxThreads)
fx_comp_kernel(const float* __restrict__ params, float* __restrict__ y, double* __restrict__ stats, int L,
               float sample_rate, int enable, int rms_norm) {
  __shared__ CompShared sh;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ch = tid >> 8, ct = tid & 255, lane = tid & 31, wic = ct >> 5;
  const float* p = params + (size_t)b * MST_FX_NPARAMS;
  double* st = stats + (size_t)b * kFxStats;

  // RMS re-normalisation of the EQ stage (common_audioeffects.py:142-145), float32 like numpy's `y *= scale`
  float scale1 = 1.f;
  if (rms_norm) {
    const double n = 2.0 * (double)L;
    const double mx = (st[S_X2_0] + st[S_X2_1]) / n, my = (st[S_Y1_0] + st[S_Y1_1]) / n;
    scale1 = (float)sqrt(mx / fmax(1e-7, my));
  }
  const double thr_d = (double)p[13], att_ms = (double)p[14], rel_ms = (double)p[15], ratio_d = (double)p[16];
  const double a_att_d = exp(-1.0 / (0.001 * (double)sample_rate * att_ms));   // :555
  const double a_rel_d = exp(-1.0 / (0.001 * (double)sample_rate * rel_ms));   // :556
  const bool active = enable && !(thr_d == 0.0 && ratio_d == 1.0);              // :635
  const float thr = (float)thr_d, ratio = (float)ratio_d, inv_ratio = 1.f / ratio;
  const float a_att = (float)a_att_d, a_rel = (float)a_rel_d;
  const float c_att = (float)(1.0 - a_att_d), c_rel = (float)(1.0 - a_rel_d);

  if (tid < 2) sh.carry[0][tid] = 0.0;  // yL_prev = 0 at every call (:553)
  if (tid == 0) {
    double pa = 1.0, pr = 1.0;
    for (int k = 0; k <= kFxChunk; ++k) { sh.pa[k] = pa; sh.pr[k] = pr; pa *= a_att_d; pr *= a_rel_d; }
    sh.rounds = 0ull;
  }
  __syncthreads();

  float* yrow = y + ((size_t)b * 2 + ch) * L;
  const bool vec = (L % 4) == 0;
  double sum_u2 = 0.0, sum_y2 = 0.0, sum_lr = 0.0;
  int tpar = 0, rpar = 0;
  unsigned rounds_local = 0;

  for (int tile0 = 0; tile0 < L; tile0 += kFxTile, tpar ^= 1) {
    const int s0 = tile0 + ct * kFxChunk;
    float u[kFxChunk];
    load_chunk(yrow, L, s0, vec, u);
    float su = 0.f;
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i) {
      u[i] = u[i] * scale1;
      if (s0 + i < L) su = fmaf(u[i], u[i], su);
    }
    sum_u2 += (double)su;
    float out[kFxChunk];
    if (active) {
      // gain computer (:559-575): x_g in dB, static curve, x_l = x_g - y_g   (float32)
      float xl[kFxChunk];
#pragma unroll
      for (int i = 0; i < kFxChunk; ++i) {
        const float ax = fabsf(u[i]);
        const float xg = ax < 0.000001f ? -120.f : 20.f * log10f(ax);
        float yg;
        if (ratio > 1.f) yg = xg >= thr ? thr + (xg - thr) * inv_ratio : xg;
        else if (ratio < 1.f) yg = xg <= thr ? thr + (xg - thr) * ratio : xg;   // (x_g - thr) / (1 / ratio)
        else yg = 0.f;
        xl[i] = xg - yg;
      }
      // smoother (:577-583): attack/release pattern fixed-point iteration + affine block scan
      float yl[kFxChunk];
      const double tile_in = sh.carry[tpar][ch];
      double g_in = tile_in;
      unsigned prev_mask = 0xFFFFFFFFu;  // impossible 16-bit pattern -> first round always "changed"
      double end_state = 0.0;
      for (int round = 0; round < kFxHalf + 2; ++round, rpar ^= 1) {
        float yy = (float)g_in, Bc = 0.f;
        unsigned mask = 0;
#pragma unroll
        for (int i = 0; i < kFxChunk; ++i) {
          const bool at = xl[i] > yy;
          const float al = at ? a_att : a_rel;
          const float c = (at ? c_att : c_rel) * xl[i];
          yy = fmaf(al, yy, c);
          Bc = fmaf(al, Bc, c);
          mask |= (at ? 1u : 0u) << i;
          yl[i] = yy;
        }
        const int na = __popc(mask);
        const double A = sh.pa[na] * sh.pr[kFxChunk - na];
        end_state = A * g_in + (double)Bc;
        const int my_changed = mask != prev_mask;
        prev_mask = mask;
        // inclusive affine scan over the warp: (A,B)_l <- map_l o ... o map_0
        double sa = A, sb = (double)Bc;
#pragma unroll
        for (int stp = 0; stp < 5; ++stp) {
          const int off = 1 << stp;
          const double oa = __shfl_up_sync(0xffffffffu, sa, off), ob = __shfl_up_sync(0xffffffffu, sb, off);
          if (lane >= off) { sb = sa * ob + sb; sa = sa * oa; }
        }
        if (lane == 31) { sh.wa[rpar][ch][wic] = sa; sh.wb[rpar][ch][wic] = sb; }
        double ea = __shfl_up_sync(0xffffffffu, sa, 1), eb = __shfl_up_sync(0xffffffffu, sb, 1);
        if (lane == 0) { ea = 1.0; eb = 0.0; }
        const int any = __syncthreads_or(my_changed);
        ++rounds_local;
        if (!any) break;          // pattern reproduced itself -> yl[] is the sequential solution
        double qv = tile_in;      // state entering this warp
        for (int w = 0; w < wic; ++w) qv = sh.wa[rpar][ch][w] * qv + sh.wb[rpar][ch][w];
        g_in = ea * qv + eb;      // state entering this thread's chunk under the current pattern
      }
      if (ct == 255) sh.carry[tpar ^ 1][ch] = end_state;   // read by the next tile (other parity)
#pragma unroll
      for (int i = 0; i < kFxChunk; ++i) {
        const float c = exp2f(yl[i] * -0.16609640474436813f);   // 10^((0 - y_l)/20), makeup 0 (:582, :646)
        out[i] = u[i] * c;                                       // (:585, :638)
      }
    } else {
#pragma unroll
      for (int i = 0; i < kFxChunk; ++i) out[i] = u[i];
    }
    // sums for the compressor RMS factor and the imager energies
    if (ch == 1) {
#pragma unroll
      for (int i = 0; i < kFxChunk / 4; ++i)
        reinterpret_cast<float4*>(sh.xch + ct * kFxChunk)[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    }
    __syncthreads();
    float sy = 0.f, slr = 0.f;
#pragma unroll
    for (int i = 0; i < kFxChunk; ++i) {
      if (s0 + i < L) {
        sy = fmaf(out[i], out[i], sy);
        if (ch == 0) slr = fmaf(out[i], sh.xch[ct * kFxChunk + i], slr);
      }
    }
    sum_y2 += (double)sy;
    sum_lr += (double)slr;
    store_chunk(yrow, L, s0, vec, out);
    __syncthreads();  // xch reuse
  }

  sum_u2 = warp_sum(sum_u2);
  sum_y2 = warp_sum(sum_y2);
  sum_lr = warp_sum(sum_lr);
  if (lane == 0) { sh.red[tid >> 5][0] = sum_u2; sh.red[tid >> 5][1] = sum_y2; sh.red[tid >> 5][2] = sum_lr; }
  __syncthreads();
  if (tid < 2) {
    double a = 0.0, c = 0.0, e = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh.red[tid * 8 + w][0]; c += sh.red[tid * 8 + w][1]; e += sh.red[tid * 8 + w][2]; }
    st[S_U2_0 + tid] = a;
    st[S_Y2_0 + tid] = c;
    if (tid == 0) { st[S_LR] = e; st[S_ROUNDS] = (double)rounds_local; }
  }
}

// =====================================================================================================================
// pass C: compressor RMS factor -> imager (+ its RMS factor, analytic) -> gain, element-wise in place
// =====================================================================================================================
__global__ void __launch_bounds__(256)
fx_final_kernel(const float* __restrict__ params, float* __restrict__ y, const double* __restrict__ stats, int L,
                int stages) {
  const int b = blockIdx.y;
  const float* p = params + (size_t)b * MST_FX_NPARAMS;
  const double* st = stats + (size_t)b * kFxStats;
  const bool rms = stages & MST_FX_RMSNORM;
  const double n = 2.0 * (double)L;
  // compressor stage RMS factor
  float scale2 = 1.f;
  if (rms && (stages & MST_FX_COMP)) {
    const double mu = (st[S_U2_0] + st[S_U2_1]) / n, my = (st[S_Y2_0] + st[S_Y2_1]) / n;
    scale2 = (float)sqrt(mu / fmax(1e-7, my));
  }
  float mg = 1.f, sg = 1.f, scale3 = 1.f;
  if (stages & MST_FX_IMAGER) {
    // energies of v = y2*scale2 :  mid = L+R, side = L-R   (:968-971)
    const double s2 = (double)scale2 * (double)scale2;
    const double ll = st[S_Y2_0] * s2, rr = st[S_Y2_1] * s2, lr = st[S_LR] * s2;
    const float mid_e = (float)(ll + rr + 2.0 * lr), side_e = (float)fmax(ll + rr - 2.0 * lr, 0.0);
    const float total_e = mid_e + side_e;
    const float max_side = sqrtf(total_e / (side_e + 1e-3f));                 // :973
    const double bal = rint((double)p[17] * 1000.0) / 1000.0;                  // round(bal, 3) (:975)
    sg = bal <= 1.0 ? (float)bal : max_side * (float)(bal - 1.0);              // :976
    const float new_side_e = side_e * (sg * sg);
    const float left_mid_e = total_e - new_side_e;
    mg = sqrtf(left_mid_e / (mid_e + 1e-3f));                                  // :981
    if (rms) {
      // mean(y3^2) with y3 = ((m' + s')/2, (m' - s')/2):  sum = (mg^2 mid_e + sg^2 side_e) / 2
      const double in_ms = (ll + rr) / n;
      const double out_ms = 0.5 * ((double)mg * mg * mid_e + (double)sg * sg * side_e) / n;
      scale3 = (float)sqrt(in_ms / fmax(1e-7, out_ms));
    }
  }
  float g = 1.f;
  if (stages & MST_FX_GAIN) {
    g = (float)pow(10.0, (double)p[18] / 20.0);                                // :1048
    if (p[19] >= 0.5f) g = -g;                                                 // :1049-1050
  }
  float* l = y + ((size_t)b * 2) * L;
  float* r = l + L;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
    const float vl = l[t] * scale2, vr = r[t] * scale2;
    float ol = vl, orr = vr;
    if (stages & MST_FX_IMAGER) {
      const float m = (vl + vr) * mg, s = (vl - vr) * sg;
      ol = (m + s) * 0.5f * scale3;
      orr = (m - s) * 0.5f * scale3;
    }
    l[t] = g * ol;
    r[t] = g * orr;
  }
}

}  // namespace mst

using namespace mst;

namespace mst {
int fx2_chain_forward(const float* x, const float* params, float* y, int B, int L, float sample_rate, int stages, double* stats,
                      cudaStream_t st);   // fx2.cu
}

// MST_FX_IMPL=v1 selects the first-generation kernels of this file (kept for A/B timing); default = fx2.cu
static bool fx_use_v1() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("MST_FX_IMPL");
    cached = (e && strcmp(e, "v1") == 0) ? 1 : 0;
  }
  return cached == 1;
}

extern "C" {

size_t mst_fx_workspace_bytes(int B, int L) {
  if (B <= 0 || L <= 0) return 0;
  return align_up((size_t)B * kFxStats * sizeof(double), 256);
}

int mst_fx_chain_forward(const float* x, const float* params, float* y, int B, int L, float sample_rate, int stages,
                         void* workspace, size_t workspace_bytes, void* stream) {
  MST_CHECK(x && params && y && workspace, "fx_chain_forward: null pointer");
  MST_CHECK(B > 0 && L > 0, "fx_chain_forward: bad shape B=%d L=%d", B, L);
  MST_CHECK(workspace_bytes >= mst_fx_workspace_bytes(B, L), "fx_chain_forward: workspace too small");
  MST_CHECK(sample_rate > 0.f, "fx_chain_forward: bad sample rate");
  cudaStream_t st = (cudaStream_t)stream;
  double* stats = reinterpret_cast<double*>(workspace);
  if (!fx_use_v1()) return fx2_chain_forward(x, params, y, B, L, sample_rate, stages, stats, st);
  const int rms = (stages & MST_FX_RMSNORM) ? 1 : 0;
  fx_eq_kernel<<<2 * B, kFxHalf, 0, st>>>(x, params, y, stats, L, sample_rate, (stages & MST_FX_EQ) ? 1 : 0);
  if (launch_ok("fx_eq_kernel")) return 1;
  fx_comp_kernel<<<B, kFxThreads, 0, st>>>(params, y, stats, L, sample_rate, (stages & MST_FX_COMP) ? 1 : 0,
                                           rms && (stages & MST_FX_EQ) ? 1 : 0);
  if (launch_ok("fx_comp_kernel")) return 1;
  dim3 grid(cdiv(L, 256 * 8) < 64 ? cdiv(L, 256 * 8) : 64, B);
  fx_final_kernel<<<grid, 256, 0, st>>>(params, y, stats, L, stages);
  return launch_ok("fx_final_kernel");
}

}  // extern "C"
