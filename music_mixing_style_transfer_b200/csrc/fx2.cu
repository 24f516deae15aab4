// fx2.cu -- FX chain kernels (parametric EQ -> compressor -> mid/side imager -> gain) for sm_100a and their C-ABI entry.
//
// Replaces (paths relative to /root/reference/mixing_style_transfer/mixing_manipulator/):
//   AugmentationChain.__call__ / apply_processor   common_audioeffects.py:156-192, 115-148 (RMS re-normalisation :142-145)
//   Equaliser.process                               :501-525 (filters :438-462)
//   compressor_process / Compressor.process         :529-587, :624-652
//   MidSideImager.process                           :965-992
//   Gain.process                                    :1038-1051
//
// "Second generation": ncu on the first version (profiles/r01f_summary.md, r01f_fx_v1_ncu_full.csv; removed in round 2)
// showed 124 thread-instructions per sample in each of the EQ and compressor kernels, 22-25 % of the warp slots occupied and
// 0.7 eligible warps per scheduler cycle -- instruction- and latency-bound at ~1 TB/s.  This file keeps the time-parallel
// algorithms (chunked recurrences + block scans, compressor pattern fixed point) with this arithmetic and geometry:
//   * One CTA (256 threads) per SEGMENT; a thread owns the SAME frames of both channels and works on (L, R) pairs with
//     packed fma.rn.f32x2 (FFMA2: measured 1.46x the FMA rate of scalar FFMA on B200, tools/ubench/fp_pipes.cu) -- the two
//     channels share every coefficient, and the compressor's cross-channel sum L*R becomes thread-local.
//   * EQ biquads run in the state coordinates (u, beta) = (s1, s1 + s2) of the DF-II-transposed states:
//         y = b0 x + u;   u' = -(1 + a1) y + (b0 + b1) x + beta;   beta' = beta + (b0+b1+b2) x - (1+a1+a2) y
//     which is algebraically the same filter (5 FMAs per sample) but keeps the ill-conditioned "slope" direction of a
//     low-frequency section (poles near z = 1) in its own small-magnitude state, so plain float32 is accurate to
//     ~1e-7 RMS even for a 30 Hz shelf (numpy emulation + tests) -- no float64 in the signal path at all: the per-chunk end
//     states, their scan and the tile carry are float32 too (packed over the channel pair); only the per-segment sums of
//     squares are accumulated in float64.
//   * Every thread runs TWO independent 16-frame chunks (ILP 2) -> 32 frames x 2 channels per thread, 8192-frame tiles,
//     so the scan and the five block barriers are amortised over 4x more samples per thread than in the first version.
//   * Compressor: gain computer on MUFU lg2 (x_l = max(0, (x_g - T)(1 - 1/R)) in one FMA + max), smoother step
//     y' = max(y + c_att (x - y), y + c_rel (x - y)) (the attack/release choice is the max of two affine maps), float32
//     maps and scans (all-float32 smoothing is within 1e-6 relative of the float64 reference), exp2 on MUFU.
//   * Global memory is touched only by fully coalesced 16-byte accesses: tiles are transposed through a padded,
//     warp-private shared-memory staging buffer (conflict-free LDS.128 on both sides).
//   * The final pass is a 2x2 matrix per frame whose coefficients are produced once per segment by the compressor CTA.
//
// Layout: x, y fp32 [B][2][L]; params fp32 [B][20]; stats double [B][16] (workspace).
#include "common.cuh"
#include "f32x2.cuh"

#include <cstdlib>

namespace mst {
namespace fx2 {

using namespace ::mst::f2;   // packed float32 x 2 helpers: lo = left channel, hi = right channel

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kChunk = 16;                    // frames per recurrence chunk
constexpr int kEqLane = 32;                   // frames per thread per tile in the EQ kernel (two chunks)
constexpr int kEqTile = kThreads * kEqLane;   // 8192 frames
constexpr int kEqRow = kEqLane + 4;           // padded staging row (floats): 36 -> lanes 0..7 start in distinct 4-bank groups
constexpr int kCpLane = 16;                   // frames per thread per tile in the compressor kernel
constexpr int kCpTile = kThreads * kCpLane;   // 4096 frames
constexpr int kCpRow = kCpLane + 4;           // 20
constexpr int kFxStats = 16;

// stats slots (doubles per segment); 10..13 = final-pass matrix (out_L = m0 l + m1 r, out_R = m2 l + m3 r)
enum { S_X2_0 = 0, S_X2_1, S_Y1_0, S_Y1_1, S_U2_0, S_U2_1, S_Y2_0, S_Y2_1, S_LR, S_ROUNDS, S_M0, S_M1, S_M2, S_M3 };

__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 16 smoother steps of one thread (both channels packed):  d = x_l - y;  y' = d > 0 ? y + c_att d : y + c_rel d  (:577-580).
// The choice is a max of the two candidates when c_att >= c_rel (attack faster than release, the usual case) and a min
// otherwise -- one FMNMX instead of compare + select.  rel[ch] collects the sign bits of d, newest in bit 0 (funnel
// shift, one instruction per sample): bit set <=> release step (d < 0; d == 0 leaves y unchanged under either label).
template <bool kMax>
__device__ __forceinline__ void smoother_steps(const u64 (&w)[kChunk], u64 (&yl)[kChunk], u64& yy, unsigned (&rel)[2], u64 c_att2,
                                               u64 c_rel2, u64 m_one2) {
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const u64 d = fma2(yy, m_one2, w[i]);            // x_l - y
    const u64 ya = fma2(c_att2, d, yy), yr = fma2(c_rel2, d, yy);
    const float n0 = kMax ? fmaxf(lo_of(ya), lo_of(yr)) : fminf(lo_of(ya), lo_of(yr));
    const float n1 = kMax ? fmaxf(hi_of(ya), hi_of(yr)) : fminf(hi_of(ya), hi_of(yr));
    rel[0] = __funnelshift_l(__float_as_uint(lo_of(d)), rel[0], 1);
    rel[1] = __funnelshift_l(__float_as_uint(hi_of(d)), rel[1], 1);
    yy = pk(n0, n1);
    yl[i] = yy;
  }
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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

// =====================================================================================================================
// Tile staging: a warp moves NF frames per lane (32 * NF consecutive frames per channel) between global memory and
// registers through its private shared-memory buffer stg[2][32 * ROW] (ROW = NF + 4 floats).
//   global side : 16-byte accesses, lane l of access i covers frames 4 * (32 i + l) .. +3  -> 512 contiguous bytes
//   register side: lane l owns frames NF * l .. NF * l + NF - 1 (row l of the buffer)
// Frame f lives at stg[ch][(f / NF) * ROW + f % NF].  `fast` = whole tile inside the row and 16-byte aligned; otherwise
// a scalar, bounds-checked path zero-fills / skips frames >= L.
// =====================================================================================================================
template <int NF, int ROW>
__device__ __forceinline__ void stage_in(const float* __restrict__ g0, const float* __restrict__ g1, int wf0, int L, bool fast,
                                         float* stg, int lane) {
  constexpr int N4 = NF / 4;   // 16-byte accesses per lane per channel
  if (fast) {
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* g = (ch == 0 ? g0 : g1) + wf0;
      float4 q[N4];
#pragma unroll
      for (int i = 0; i < N4; ++i) q[i] = __ldg(reinterpret_cast<const float4*>(g) + 32 * i + lane);
#pragma unroll
      for (int i = 0; i < N4; ++i) {
        const int f = 4 * (32 * i + lane);
        *reinterpret_cast<float4*>(stg + ch * 32 * ROW + (f / NF) * ROW + (f % NF)) = q[i];
      }
    }
  } else {
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      const float* g = (ch == 0 ? g0 : g1);
#pragma unroll 4
      for (int j = 0; j < NF; ++j) {
        const int f = 32 * j + lane;
        stg[ch * 32 * ROW + (f / NF) * ROW + (f % NF)] = (wf0 + f < L) ? __ldg(g + wf0 + f) : 0.f;
      }
    }
  }
  __syncwarp();
}

template <int NF, int ROW>
__device__ __forceinline__ void stage_out(float* __restrict__ g0, float* __restrict__ g1, int wf0, int L, bool fast,
                                          const float* stg, int lane) {
  constexpr int N4 = NF / 4;
  __syncwarp();
  if (fast) {
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      float* g = (ch == 0 ? g0 : g1) + wf0;
#pragma unroll
      for (int i = 0; i < N4; ++i) {
        const int f = 4 * (32 * i + lane);
        const float4 q = *reinterpret_cast<const float4*>(stg + ch * 32 * ROW + (f / NF) * ROW + (f % NF));
        reinterpret_cast<float4*>(g)[32 * i + lane] = q;
      }
    }
  } else {
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float* g = (ch == 0 ? g0 : g1);
#pragma unroll 4
      for (int j = 0; j < NF; ++j) {
        const int f = 32 * j + lane;
        if (wf0 + f < L) g[wf0 + f] = stg[ch * 32 * ROW + (f / NF) * ROW + (f % NF)];
      }
    }
  }
  __syncwarp();
}

// Asynchronous variant of the fast path of stage_in: 16-byte cp.async copies straight into the staging layout (no
// registers, the data lands while the warp works on the previous tile).  Complete with stage_wait().
template <int NF, int ROW>
__device__ __forceinline__ void stage_in_async(const float* __restrict__ g0, const float* __restrict__ g1, int wf0, float* stg,
                                               int lane) {
  constexpr int N4 = NF / 4;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const float* g = (ch == 0 ? g0 : g1) + wf0;
#pragma unroll
    for (int i = 0; i < N4; ++i) {
      const int f = 4 * (32 * i + lane);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(stg + ch * 32 * ROW + (f / NF) * ROW + (f % NF));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + f) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// =====================================================================================================================
// pass A: EQ (5 cascaded biquads), x -> y, per-channel sums of x^2 and y^2
//
// Per section and tile: every thread runs its two 16-frame chunks from zero state (packed over the channel pair), the
// chunk end states are combined by a warp scan + one block barrier + an 8-entry prefix over the warp totals (3-step shuffle
// scan, identical work in every warp), and the natural response to the true incoming state is added.  All of it is
// float32: in (u, beta) coordinates the float32 scan + carry is as accurate as the float64 one (numpy emulation of this
// very schedule: 6.6e-8 .. 2.4e-7 RMS either way, 30 Hz shelf and Q = 0.1 included).
// Negative result kept out of the code: replacing the barrier by a warp-to-warp hand-over through shared-memory flags
// (each warp taking its 1024-frame tile through all five sections on its own) measured 614-738 us against 499 us --
// the spinning lanes eat issue slots and the chain moves at the pace of the slowest warp just like the barrier.
// =====================================================================================================================
struct EqShared {
  float4 mpow[5][33];         // (32-frame lane transition)^l, l = 0..32, row-major 2x2 (m0 m1; m2 m3), (u, beta) coordinates
  float4 m16[5];              // 16-frame chunk transition
  float2 nat[5][kChunk];      // natural response of y to unit (u, beta): (Nu, Nb)
  float cf[5][8];             // b0, B = b0+b1+b2, -delta = -(1+a1+a2), cy = -(1+a1), cx = b0+b1
  float4 mw[5][kWarps + 1];   // (warp transition = M32^32)^j, j = 0..8
  float4 wtot[2][5][kWarps];  // [tile parity][biquad][warp]: warp totals (u_L, u_R, beta_L, beta_R)
  float4 carryb[2][5];        // [tile parity][biquad]: state entering the tile
  double red[kWarps][4];
};

__device__ __forceinline__ u64 shfl_up2(u64 v, int off) { return __shfl_up_sync(0xffffffffu, v, off); }
__global__ void __launch_bounds__(kThreads, 2)
eq_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ y, double* __restrict__ stats,
          int L, float sample_rate, int enable, int vec, const double* __restrict__ coef, int n_coef_sections) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EqShared& sh = *reinterpret_cast<EqShared*>(smem_raw);
  float* stg_all = reinterpret_cast<float*>(smem_raw + ((sizeof(EqShared) + 15) / 16) * 16);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* stg = stg_all + wid * (2 * 32 * kEqRow);
  const float* p = params != nullptr ? params + (size_t)b * MST_FX_NPARAMS : nullptr;

  if (tid < 5) {
    // dict order low_shelf, first, second, third, high_shelf (:391); shelves Q = 0.707 (:454)
    const int gi[5] = {0, 2, 5, 8, 11}, fi[5] = {1, 3, 6, 9, 12}, qi[5] = {-1, 4, 7, 10, -1}, ty[5] = {0, 1, 1, 1, 2};
    Biquad q;
    if (coef != nullptr) {
      // explicit sections (mst_biquad_cascade): (b0, b1, b2, a1, a2) per section, a0 = 1; the unused ones are identities
      const double* c5 = coef + ((size_t)b * n_coef_sections + tid) * 5;
      if (tid < n_coef_sections) { q.b0 = c5[0]; q.b1 = c5[1]; q.b2 = c5[2]; q.a1 = c5[3]; q.a2 = c5[4]; }
      else { q.b0 = 1.0; q.b1 = 0.0; q.b2 = 0.0; q.a1 = 0.0; q.a2 = 0.0; }
    } else {
      const double Q = qi[tid] < 0 ? 0.707 : (double)p[qi[tid]];
      q = rbj((double)p[gi[tid]], Q, (double)p[fi[tid]], (double)sample_rate, ty[tid]);
    }
    const double Bc = q.b0 + q.b1 + q.b2, dl = 1.0 + q.a1 + q.a2, cy = -(1.0 + q.a1), cx = q.b0 + q.b1;
    sh.cf[tid][0] = (float)q.b0;
    sh.cf[tid][1] = (float)Bc;
    sh.cf[tid][2] = (float)-dl;
    sh.cf[tid][3] = (float)cy;
    sh.cf[tid][4] = (float)cx;
    // natural response and 16-frame transition from the two unit states (float64, stored as float32)
    double M[4];
    for (int s = 0; s < 2; ++s) {
      double u = s == 0 ? 1.0 : 0.0, be = s == 0 ? 0.0 : 1.0;
      for (int i = 0; i < kChunk; ++i) {
        const double yy = u;
        if (s == 0) sh.nat[tid][i].x = (float)yy;
        else        sh.nat[tid][i].y = (float)yy;
        const double un = fma(cy, yy, be);
        be = fma(-dl, yy, be);
        u = un;
      }
      M[0 + s] = u;   // column s of the transition
      M[2 + s] = be;
    }
    sh.m16[tid] = make_float4((float)M[0], (float)M[1], (float)M[2], (float)M[3]);
    // lane transition = M16^2, and its powers 0..32
    double M32[4];
    M32[0] = M[0] * M[0] + M[1] * M[2]; M32[1] = M[0] * M[1] + M[1] * M[3];
    M32[2] = M[2] * M[0] + M[3] * M[2]; M32[3] = M[2] * M[1] + M[3] * M[3];
    double P[4] = {1.0, 0.0, 0.0, 1.0};
    for (int l = 0; l <= 32; ++l) {
      sh.mpow[tid][l] = make_float4((float)P[0], (float)P[1], (float)P[2], (float)P[3]);
      const double n0 = M32[0] * P[0] + M32[1] * P[2], n1 = M32[0] * P[1] + M32[1] * P[3];
      const double n2 = M32[2] * P[0] + M32[3] * P[2], n3 = M32[2] * P[1] + M32[3] * P[3];
      P[0] = n0; P[1] = n1; P[2] = n2; P[3] = n3;
    }
    {
      // P now holds M32^33; M32^32 was stored at l = 32.  Powers of the warp transition W = M32^32 in float64:
      double Wm[4], R[4] = {1.0, 0.0, 0.0, 1.0};
      double T[4] = {1.0, 0.0, 0.0, 1.0};
      for (int l = 0; l < 32; ++l) {
        const double n0 = M32[0] * T[0] + M32[1] * T[2], n1 = M32[0] * T[1] + M32[1] * T[3];
        const double n2 = M32[2] * T[0] + M32[3] * T[2], n3 = M32[2] * T[1] + M32[3] * T[3];
        T[0] = n0; T[1] = n1; T[2] = n2; T[3] = n3;
      }
      for (int e = 0; e < 4; ++e) Wm[e] = T[e];
      for (int j = 0; j <= kWarps; ++j) {
        sh.mw[tid][j] = make_float4((float)R[0], (float)R[1], (float)R[2], (float)R[3]);
        const double n0 = Wm[0] * R[0] + Wm[1] * R[2], n1 = Wm[0] * R[1] + Wm[1] * R[3];
        const double n2 = Wm[2] * R[0] + Wm[3] * R[2], n3 = Wm[2] * R[1] + Wm[3] * R[3];
        R[0] = n0; R[1] = n1; R[2] = n2; R[3] = n3;
      }
    }
    sh.carryb[0][tid] = make_float4(0.f, 0.f, 0.f, 0.f);   // state reset (:512)
  }
  __syncthreads();

  const float* x0 = x + ((size_t)b * 2) * L;
  const float* x1 = x0 + L;
  float* y0 = y + ((size_t)b * 2) * L;
  float* y1 = y0 + L;
  double sum_x2[2] = {0.0, 0.0}, sum_y2[2] = {0.0, 0.0};
  int par = 0;                                             // tile parity (double-buffered warp totals / carry)

#pragma unroll 1
  for (int tile0 = 0; tile0 < L; tile0 += kEqTile, par ^= 1) {
    const int wf0 = tile0 + wid * (32 * kEqLane);          // first frame of this warp
    const bool fast = vec && (tile0 + kEqTile <= L);
    stage_in<kEqLane, kEqRow>(x0, x1, wf0, L, fast, stg, lane);
    if (tile0 + 2 * kEqTile <= L) {
      // pull this warp's part of the NEXT tile into L2 now (one 128-byte line per lane and channel): the staging buffer is
      // busy until the end of the tile, so the next stage_in cannot be issued early, but it can find its lines in L2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(x0 + wf0 + kEqTile + lane * kEqLane));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(x1 + wf0 + kEqTile + lane * kEqLane));
    }
    u64 v[2][kChunk];                                      // [chunk][i] = (L, R)
    {
      const float* r0 = stg + lane * kEqRow;
      const float* r1 = stg + 32 * kEqRow + lane * kEqRow;
#pragma unroll
      for (int i = 0; i < kEqLane / 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * i);
        const float4 c = *reinterpret_cast<const float4*>(r1 + 4 * i);
        const int cc = (4 * i) / kChunk, ii = (4 * i) % kChunk;
        v[cc][ii] = pk(a.x, c.x); v[cc][ii + 1] = pk(a.y, c.y); v[cc][ii + 2] = pk(a.z, c.z); v[cc][ii + 3] = pk(a.w, c.w);
      }
    }
    {
      u64 sx = 0ull;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < kChunk; ++i) sx = fma2(v[c][i], v[c][i], sx);
      sum_x2[0] += (double)lo_of(sx);
      sum_x2[1] += (double)hi_of(sx);
    }
    if (enable) {
#pragma unroll 1
      for (int k = 0; k < 5; ++k) {
        // (1) zero-state response of both chunks, packed over the channels (float32, 5 FMAs per frame and chunk)
        u64 uA = 0ull, bA = 0ull, uB = 0ull, bB = 0ull;
        {
          // scalar coefficients, duplicated at use: ptxas turns dup() into the FFMA2 ".F32" broadcast operand form
          const u64 c_b0 = dup(sh.cf[k][0]), c_B = dup(sh.cf[k][1]), c_nd = dup(sh.cf[k][2]);
          const u64 c_cy = dup(sh.cf[k][3]), c_cx = dup(sh.cf[k][4]);
#pragma unroll
          for (int i = 0; i < kChunk; ++i) {
            const u64 xa = v[0][i], xb = v[1][i];
            const u64 ya = fma2(c_b0, xa, uA), yb = fma2(c_b0, xb, uB);
            const u64 tba = fma2(c_B, xa, bA), tbb = fma2(c_B, xb, bB);
            const u64 tua = fma2(c_cx, xa, bA), tub = fma2(c_cx, xb, bB);
            bA = fma2(c_nd, ya, tba); bB = fma2(c_nd, yb, tbb);
            uA = fma2(c_cy, ya, tua); uB = fma2(c_cy, yb, tub);
            v[0][i] = ya; v[1][i] = yb;
          }
        }
        // (2) lane total Z = M16 zA + zB, inclusive warp scan I_l = sum_{i<=l} M32^(l-i) Z_i
        const float4 m16 = sh.m16[k];
        const u64 h0 = dup(m16.x), h1 = dup(m16.y), h2 = dup(m16.z), h3 = dup(m16.w);
        u64 IU = fma2(h0, uA, fma2(h1, bA, uB));
        u64 IB = fma2(h2, uA, fma2(h3, bA, bB));
#pragma unroll
        for (int st = 0; st < 5; ++st) {
          const int off = 1 << st;
          const float4 mp = sh.mpow[k][off];
          const u64 oU = shfl_up2(IU, off), oB = shfl_up2(IB, off);
          if (lane >= off) {
            const u64 nU = fma2(dup(mp.x), oU, fma2(dup(mp.y), oB, IU));
            const u64 nB = fma2(dup(mp.z), oU, fma2(dup(mp.w), oB, IB));
            IU = nU; IB = nB;
          }
        }
        u64 EU = shfl_up2(IU, 1), EB = shfl_up2(IB, 1);    // state after the lanes before this one, from zero
        if (lane == 0) { EU = 0ull; EB = 0ull; }
        // (3) state entering this warp
        u64 QU, QB;
        {
          // warp totals -> shared memory, one block barrier, then every warp computes the prefix over the
          // (at most 8) warp totals with a 3-step shuffle scan: P_j = state after warps 0..j from zero
          if (lane == 31) sh.wtot[par][k][wid] = make_float4(lo_of(IU), hi_of(IU), lo_of(IB), hi_of(IB));
          __syncthreads();
          const float4 cq = sh.carryb[par][k];
          const u64 CU = pk(cq.x, cq.y), CB = pk(cq.z, cq.w);
          u64 PU = 0ull, PB = 0ull;
          if (lane < kWarps) {
            const float4 w4 = sh.wtot[par][k][lane];
            PU = pk(w4.x, w4.y); PB = pk(w4.z, w4.w);
          }
#pragma unroll
          for (int off = 1; off < kWarps; off <<= 1) {
            const float4 mp = sh.mw[k][off];
            const u64 oU = shfl_up2(PU, off), oB = shfl_up2(PB, off);
            if (lane >= off) {
              const u64 nU = fma2(dup(mp.x), oU, fma2(dup(mp.y), oB, PU));
              const u64 nB = fma2(dup(mp.z), oU, fma2(dup(mp.w), oB, PB));
              PU = nU; PB = nB;
            }
          }
          const int src = (wid + 31) & 31;
          u64 SU = __shfl_sync(0xffffffffu, PU, src), SB = __shfl_sync(0xffffffffu, PB, src);
          if (wid == 0) { SU = 0ull; SB = 0ull; }
          const float4 mq = sh.mw[k][wid];                  // Q_w = W^w carry + P_{w-1}
          QU = fma2(dup(mq.x), CU, fma2(dup(mq.y), CB, SU));
          QB = fma2(dup(mq.z), CU, fma2(dup(mq.w), CB, SB));
          if (tid == kThreads - 1) {                         // state leaving the tile = W Q_7 + W_7 (next tile, other parity)
            const float4 mt = sh.mw[k][1];
            const u64 nU = fma2(dup(mt.x), QU, fma2(dup(mt.y), QB, IU));
            const u64 nB = fma2(dup(mt.z), QU, fma2(dup(mt.w), QB, IB));
            sh.carryb[par ^ 1][k] = make_float4(lo_of(nU), hi_of(nU), lo_of(nB), hi_of(nB));
          }
        }
        // in_A = M32^lane Q + E;  in_B = M16 in_A + zA
        const float4 ml = sh.mpow[k][lane];
        const u64 inUA = fma2(dup(ml.x), QU, fma2(dup(ml.y), QB, EU));
        const u64 inBA = fma2(dup(ml.z), QU, fma2(dup(ml.w), QB, EB));
        const u64 inUB = fma2(h0, inUA, fma2(h1, inBA, uA));
        const u64 inBB = fma2(h2, inUA, fma2(h3, inBA, bA));
        // (4) add the natural response to the true incoming state
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
          const float2 n = sh.nat[k][i];
          const u64 nu = dup(n.x), nb = dup(n.y);
          v[0][i] = fma2(nu, inUA, fma2(nb, inBA, v[0][i]));
          v[1][i] = fma2(nu, inUB, fma2(nb, inBB, v[1][i]));
        }
      }
    }
    // sums of y^2 (frames >= L excluded: the filters ring into the zero padding of a partial tile) and store
    {
      const int f0 = wf0 + lane * kEqLane;
      u64 sy = 0ull;
      if (fast || f0 + kEqLane <= L) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < kChunk; ++i) sy = fma2(v[c][i], v[c][i], sy);
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < kChunk; ++i)
            if (f0 + c * kChunk + i < L) sy = fma2(v[c][i], v[c][i], sy);
      }
      sum_y2[0] += (double)lo_of(sy);
      sum_y2[1] += (double)hi_of(sy);
      float* r0 = stg + lane * kEqRow;
      float* r1 = stg + 32 * kEqRow + lane * kEqRow;
#pragma unroll
      for (int i = 0; i < kEqLane / 4; ++i) {
        const int cc = (4 * i) / kChunk, ii = (4 * i) % kChunk;
        *reinterpret_cast<float4*>(r0 + 4 * i) = make_float4(lo_of(v[cc][ii]), lo_of(v[cc][ii + 1]), lo_of(v[cc][ii + 2]), lo_of(v[cc][ii + 3]));
        *reinterpret_cast<float4*>(r1 + 4 * i) = make_float4(hi_of(v[cc][ii]), hi_of(v[cc][ii + 1]), hi_of(v[cc][ii + 2]), hi_of(v[cc][ii + 3]));
      }
    }
    stage_out<kEqLane, kEqRow>(y0, y1, wf0, L, fast, stg, lane);
  }

#pragma unroll
  for (int ch = 0; ch < 2; ++ch) { sum_x2[ch] = warp_sum(sum_x2[ch]); sum_y2[ch] = warp_sum(sum_y2[ch]); }
  __syncthreads();
  if (lane == 0) { sh.red[wid][0] = sum_x2[0]; sh.red[wid][1] = sum_x2[1]; sh.red[wid][2] = sum_y2[0]; sh.red[wid][3] = sum_y2[1]; }
  __syncthreads();
  if (tid < 4) {
    double a = 0.0;
    for (int w = 0; w < kWarps; ++w) a += sh.red[w][tid];
    const int slot[4] = {S_X2_0, S_X2_1, S_Y1_0, S_Y1_1};
    stats[(size_t)b * kFxStats + slot[tid]] = a;
  }
}

// =====================================================================================================================
// pass B: compressor, in place on y; per-channel sums, sum L*R, and the final-pass matrix of the segment
// =====================================================================================================================
struct CompShared {
  float4 wab[2][kWarps];                      // [round parity][warp]: warp-total affine maps (A_L, B_L, A_R, B_R)
  float cend[2][2];                           // [round parity][channel]: smoother state leaving the tile (thread 255)
  float pa[kChunk + 1], pr[kChunk + 1];       // alpha_att^k, alpha_rel^k
  double red[kWarps][5];
};

__device__ void final_matrix(const float* p, double* st, int L, int stages);

__global__ void __launch_bounds__(kThreads, 2)
comp_kernel(const float* __restrict__ params, float* __restrict__ y, double* __restrict__ stats, int L, float sample_rate,
            int stages, int vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CompShared& sh = *reinterpret_cast<CompShared*>(smem_raw);
  float* stg_all = reinterpret_cast<float*>(smem_raw + ((sizeof(CompShared) + 15) / 16) * 16);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* stg_w = stg_all + wid * (2 * 2 * 32 * kCpRow);   // two tile buffers per warp (the next tile is prefetched)
  const float* p = params + (size_t)b * MST_FX_NPARAMS;
  double* st = stats + (size_t)b * kFxStats;
  const bool enable = (stages & MST_FX_COMP) != 0;
  const bool rms_in = (stages & MST_FX_RMSNORM) && (stages & MST_FX_EQ);

  // RMS re-normalisation of the EQ stage (common_audioeffects.py:142-145), float32 like numpy's `y *= scale`
  float scale1 = 1.f;
  if (rms_in) {
    const double n = 2.0 * (double)L;
    const double mx = (st[S_X2_0] + st[S_X2_1]) / n, my = (st[S_Y1_0] + st[S_Y1_1]) / n;
    scale1 = (float)sqrt(mx / fmax(1e-7, my));
  }
  const double thr_d = (double)p[13], att_ms = (double)p[14], rel_ms = (double)p[15], ratio_d = (double)p[16];
  const double a_att_d = exp(-1.0 / (0.001 * (double)sample_rate * att_ms));   // :555
  const double a_rel_d = exp(-1.0 / (0.001 * (double)sample_rate * rel_ms));   // :556
  const bool active = enable && !(thr_d == 0.0 && ratio_d == 1.0);              // :635
  // static curve (:564-573) folded into x_l = x_g - y_g:
  //   ratio > 1: x_l = max(0, (x_g - T)(1 - 1/R));  ratio < 1: x_l = min(0, (x_g - T)(1 - R));  ratio == 1: x_l = x_g
  //   i.e. x_l = clamp(x_g * cs + co, c_lo, c_hi) with the bounds below (branch-free in the sample loop)
  float cs, co, c_lo, c_hi;
  const float kInf = __int_as_float(0x7f800000);
  if (ratio_d > 1.0)      { cs = (float)(1.0 - 1.0 / ratio_d); co = (float)(-thr_d * (1.0 - 1.0 / ratio_d)); c_lo = 0.f; c_hi = kInf; }
  else if (ratio_d < 1.0) { cs = (float)(1.0 - ratio_d);       co = (float)(-thr_d * (1.0 - ratio_d));       c_lo = -kInf; c_hi = 0.f; }
  else                    { cs = 1.f; co = 0.f; c_lo = -kInf; c_hi = kInf; }
  const float cs_db = cs * 6.0205999132796239f;            // x_g = 20 log10 |u| = 6.0206 log2 |u|
  const u64 c_att2 = dup((float)(1.0 - a_att_d)), c_rel2 = dup((float)(1.0 - a_rel_d)), m_one2 = dup(-1.f);
  const bool use_max = (float)(1.0 - a_att_d) >= (float)(1.0 - a_rel_d);
  const u64 scale1_2 = dup(scale1);

  if (tid == 0) {
    double pa = 1.0, pr = 1.0;
    for (int k = 0; k <= kChunk; ++k) { sh.pa[k] = (float)pa; sh.pr[k] = (float)pr; pa *= a_att_d; pr *= a_rel_d; }
  }
  __syncthreads();

  float* y0 = y + ((size_t)b * 2) * L;
  float* y1 = y0 + L;
  double sum_u2[2] = {0.0, 0.0}, sum_y2[2] = {0.0, 0.0}, sum_lr = 0.0;
  int rpar = 0, bufi = 0;
  unsigned rounds_local = 0;
  float tile_in[2] = {0.f, 0.f};        // smoother state entering the tile; yL_prev = 0 at every call (:553)

#pragma unroll 1
  for (int tile0 = 0; tile0 < L; tile0 += kCpTile) {
    const int wf0 = tile0 + wid * (32 * kCpLane);
    const bool fast = vec && (tile0 + kCpTile <= L);
    float* stg = stg_w + bufi * (2 * 32 * kCpRow);
    if (fast) {
      if (tile0 == 0) stage_in_async<kCpLane, kCpRow>(y0, y1, wf0, stg, lane);
      stage_wait();
    } else {
      stage_in<kCpLane, kCpRow>(y0, y1, wf0, L, false, stg, lane);
    }
    if (vec && tile0 + 2 * kCpTile <= L)   // prefetch the next tile into the other buffer
      stage_in_async<kCpLane, kCpRow>(y0, y1, wf0 + kCpTile, stg_w + (bufi ^ 1) * (2 * 32 * kCpRow), lane);
    bufi ^= 1;
    float* r0 = stg + lane * kCpRow;
    float* r1 = stg + 32 * kCpRow + lane * kCpRow;
    u64 w[kChunk];   // x_l (dB over the static curve) while the smoother runs, then the output
    {
      u64 su = 0ull;
#pragma unroll
      for (int i = 0; i < kCpLane / 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * i);
        const float4 c = *reinterpret_cast<const float4*>(r1 + 4 * i);
        w[4 * i] = mul2(pk(a.x, c.x), scale1_2); w[4 * i + 1] = mul2(pk(a.y, c.y), scale1_2);
        w[4 * i + 2] = mul2(pk(a.z, c.z), scale1_2); w[4 * i + 3] = mul2(pk(a.w, c.w), scale1_2);
      }
#pragma unroll
      for (int i = 0; i < kChunk; ++i) su = fma2(w[i], w[i], su);
      sum_u2[0] += (double)lo_of(su);
      sum_u2[1] += (double)hi_of(su);
    }
    if (active) {
      // gain computer (:559-575): x_g = 20 log10 |u| (-120 below 1e-6), static curve, x_l = x_g - y_g      (float32)
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        float xl[2];
        xl[0] = lo_of(w[i]); xl[1] = hi_of(w[i]);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          xl[ch] = fminf(fmaxf(fmaf(lg2_approx(fmaxf(fabsf(xl[ch]), 0.000001f)), cs_db, co), c_lo), c_hi);
        }
        w[i] = pk(xl[0], xl[1]);
      }
      // smoother (:577-583): attack/release pattern fixed-point iteration + affine block scan (float32)
      u64 yl[kChunk];
      float g_in[2] = {tile_in[0], tile_in[1]};
      unsigned prev_mask[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // impossible 16-bit pattern -> first round always "changed"
      float y_end[2] = {0.f, 0.f};
      int round = 0;
#pragma unroll 1
      while (true) {
        u64 yy = pk(g_in[0], g_in[1]);
        unsigned mask[2] = {0u, 0u};                       // release pattern of this round (16 bits per channel)
        if (use_max) smoother_steps<true>(w, yl, yy, mask, c_att2, c_rel2, m_one2);
        else         smoother_steps<false>(w, yl, yy, mask, c_att2, c_rel2, m_one2);
        y_end[0] = lo_of(yy); y_end[1] = hi_of(yy);
        int changed = 0;
        float sa[2], sb[2], ea[2], eb[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int nr = __popc(mask[ch]);
          sa[ch] = sh.pa[kChunk - nr] * sh.pr[nr];         // chunk map under this pattern: y_out = A y_in + B
          sb[ch] = fmaf(-sa[ch], g_in[ch], y_end[ch]);
          changed |= (mask[ch] != prev_mask[ch]);
          prev_mask[ch] = mask[ch];
        }
        // inclusive affine scan over the warp: (A,B)_l <- map_l o ... o map_0
#pragma unroll
        for (int stp = 0; stp < 5; ++stp) {
          const int off = 1 << stp;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const float oa = __shfl_up_sync(0xffffffffu, sa[ch], off), ob = __shfl_up_sync(0xffffffffu, sb[ch], off);
            if (lane >= off) { sb[ch] = fmaf(sa[ch], ob, sb[ch]); sa[ch] = sa[ch] * oa; }
          }
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          if (tid == kThreads - 1) sh.cend[rpar][ch] = y_end[ch];
          ea[ch] = __shfl_up_sync(0xffffffffu, sa[ch], 1);
          eb[ch] = __shfl_up_sync(0xffffffffu, sb[ch], 1);
          if (lane == 0) { ea[ch] = 1.f; eb[ch] = 0.f; }
        }
        if (lane == 31) sh.wab[rpar][wid] = make_float4(sa[0], sb[0], sa[1], sb[1]);
        const int any = __syncthreads_or(changed);
        ++rounds_local;
        ++round;
        if (!any || round >= kThreads + 2) {   // pattern reproduced itself -> yl[] is the sequential solution
          tile_in[0] = sh.cend[rpar][0];       // state leaving this tile = state entering the next one
          tile_in[1] = sh.cend[rpar][1];
          rpar ^= 1;                           // the next tile's first round must not overwrite what is being read here
          break;
        }
        {
          // state entering this warp: prefix of the warp-total maps, by a 3-step scan over lanes 0..7 (every warp does the
          // same work, so nobody arrives late at the next barrier), then lane wid-1 holds the map of warps 0..wid-1
          float4 m = make_float4(1.f, 0.f, 1.f, 0.f);
          if (lane < kWarps) m = sh.wab[rpar][lane];
#pragma unroll
          for (int off = 1; off < kWarps; off <<= 1) {
            const float oa0 = __shfl_up_sync(0xffffffffu, m.x, off), ob0 = __shfl_up_sync(0xffffffffu, m.y, off);
            const float oa1 = __shfl_up_sync(0xffffffffu, m.z, off), ob1 = __shfl_up_sync(0xffffffffu, m.w, off);
            if (lane >= off) {
              m.y = fmaf(m.x, ob0, m.y); m.x = m.x * oa0;
              m.w = fmaf(m.z, ob1, m.w); m.z = m.z * oa1;
            }
          }
          const int src = (wid + 31) & 31;
          float pa0 = __shfl_sync(0xffffffffu, m.x, src), pb0 = __shfl_sync(0xffffffffu, m.y, src);
          float pa1 = __shfl_sync(0xffffffffu, m.z, src), pb1 = __shfl_sync(0xffffffffu, m.w, src);
          if (wid == 0) { pa0 = 1.f; pb0 = 0.f; pa1 = 1.f; pb1 = 0.f; }
          const float q0 = fmaf(pa0, tile_in[0], pb0), q1 = fmaf(pa1, tile_in[1], pb1);
          g_in[0] = fmaf(ea[0], q0, eb[0]);   // state entering this thread's chunk under the current pattern
          g_in[1] = fmaf(ea[1], q1, eb[1]);
        }
        rpar ^= 1;
      }
      // gain: c = 10^((0 - y_l)/20), makeup 0 (:582, :646);  y = u * c (:585, :638).  u is re-read from the staging rows.
#pragma unroll
      for (int i = 0; i < kCpLane / 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * i);
        const float4 c = *reinterpret_cast<const float4*>(r1 + 4 * i);
        const float ua[4] = {a.x, a.y, a.z, a.w}, uc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float g0 = ex2_approx(lo_of(yl[4 * i + j]) * -0.16609640474436813f);
          const float g1 = ex2_approx(hi_of(yl[4 * i + j]) * -0.16609640474436813f);
          w[4 * i + j] = mul2(mul2(pk(ua[j], uc[j]), scale1_2), pk(g0, g1));
        }
      }
    }
    // sums for the compressor RMS factor and the imager energies; frames >= L hold u = 0 -> out = 0
    {
      u64 sy = 0ull;
      float slr = 0.f;
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        sy = fma2(w[i], w[i], sy);
        slr = fmaf(lo_of(w[i]), hi_of(w[i]), slr);
      }
      sum_y2[0] += (double)lo_of(sy);
      sum_y2[1] += (double)hi_of(sy);
      sum_lr += (double)slr;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kCpLane / 4; ++i) {
      *reinterpret_cast<float4*>(r0 + 4 * i) = make_float4(lo_of(w[4 * i]), lo_of(w[4 * i + 1]), lo_of(w[4 * i + 2]), lo_of(w[4 * i + 3]));
      *reinterpret_cast<float4*>(r1 + 4 * i) = make_float4(hi_of(w[4 * i]), hi_of(w[4 * i + 1]), hi_of(w[4 * i + 2]), hi_of(w[4 * i + 3]));
    }
    stage_out<kCpLane, kCpRow>(y0, y1, wf0, L, fast, stg, lane);
  }

  double r5[5] = {sum_u2[0], sum_u2[1], sum_y2[0], sum_y2[1], sum_lr};
#pragma unroll
  for (int j = 0; j < 5; ++j) r5[j] = warp_sum(r5[j]);
  __syncthreads();
  if (lane == 0)
    for (int j = 0; j < 5; ++j) sh.red[wid][j] = r5[j];
  __syncthreads();
  if (tid == 0) {
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int ww = 0; ww < kWarps; ++ww)
      for (int j = 0; j < 5; ++j) a[j] += sh.red[ww][j];
    st[S_U2_0] = a[0]; st[S_U2_1] = a[1]; st[S_Y2_0] = a[2]; st[S_Y2_1] = a[3]; st[S_LR] = a[4];
    st[S_ROUNDS] = (double)rounds_local;
    final_matrix(p, st, L, stages);
  }
}

// compressor RMS factor -> imager gains (+ its RMS factor, analytic) -> gain, folded into one 2x2 matrix per segment
__device__ void final_matrix(const float* p, double* st, int L, int stages) {
  const bool rms = stages & MST_FX_RMSNORM;
  const double n = 2.0 * (double)L;
  float scale2 = 1.f;
  if (rms && (stages & MST_FX_COMP)) {
    const double mu = (st[S_U2_0] + st[S_U2_1]) / n, my = (st[S_Y2_0] + st[S_Y2_1]) / n;
    scale2 = (float)sqrt(mu / fmax(1e-7, my));
  }
  float mg = 1.f, sg = 1.f, scale3 = 1.f;
  const bool imager = stages & MST_FX_IMAGER;
  if (imager) {
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
  double g = 1.0;
  if (stages & MST_FX_GAIN) {
    g = (double)(float)pow(10.0, (double)p[18] / 20.0);                        // :1048
    if (p[19] >= 0.5f) g = -g;                                                 // :1049-1050
  }
  // out_L = g scale3 ((vl + vr) mg + (vl - vr) sg) / 2,  out_R = g scale3 ((vl + vr) mg - (vl - vr) sg) / 2,  v = y2 scale2
  const double k = g * (double)scale3 * (double)scale2;
  if (imager) {
    const double a = 0.5 * k * ((double)mg + (double)sg), c = 0.5 * k * ((double)mg - (double)sg);
    st[S_M0] = a; st[S_M1] = c; st[S_M2] = c; st[S_M3] = a;
  } else {
    st[S_M0] = k; st[S_M1] = 0.0; st[S_M2] = 0.0; st[S_M3] = k;
  }
}

// =====================================================================================================================
// pass C: out = M * (l, r) per frame, in place
// =====================================================================================================================
constexpr int kFinThreads = 256;
constexpr int kFinFrames = kFinThreads * 4 * 4;   // frames per CTA: 4 float4 per thread and channel

__global__ void __launch_bounds__(kFinThreads)
final_kernel(float* __restrict__ y, const double* __restrict__ stats, int L, int vec) {
  const int b = blockIdx.y;
  const double* st = stats + (size_t)b * kFxStats;
  const float m0 = (float)st[S_M0], m1 = (float)st[S_M1], m2 = (float)st[S_M2], m3 = (float)st[S_M3];
  float* l = y + ((size_t)b * 2) * L;
  float* r = l + L;
  const int f0 = blockIdx.x * kFinFrames;
  if (vec && f0 + kFinFrames <= L) {
    float4 a[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = *(reinterpret_cast<const float4*>(l + f0) + i * kFinThreads + threadIdx.x);
      c[i] = *(reinterpret_cast<const float4*>(r + f0) + i * kFinThreads + threadIdx.x);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 ol, orr;
      ol.x = fmaf(m0, a[i].x, m1 * c[i].x); orr.x = fmaf(m2, a[i].x, m3 * c[i].x);
      ol.y = fmaf(m0, a[i].y, m1 * c[i].y); orr.y = fmaf(m2, a[i].y, m3 * c[i].y);
      ol.z = fmaf(m0, a[i].z, m1 * c[i].z); orr.z = fmaf(m2, a[i].z, m3 * c[i].z);
      ol.w = fmaf(m0, a[i].w, m1 * c[i].w); orr.w = fmaf(m2, a[i].w, m3 * c[i].w);
      *(reinterpret_cast<float4*>(l + f0) + i * kFinThreads + threadIdx.x) = ol;
      *(reinterpret_cast<float4*>(r + f0) + i * kFinThreads + threadIdx.x) = orr;
    }
  } else {
    const int f1 = f0 + kFinFrames < L ? f0 + kFinFrames : L;
    for (int t = f0 + threadIdx.x; t < f1; t += kFinThreads) {
      const float vl = l[t], vr = r[t];
      l[t] = fmaf(m0, vl, m1 * vr);
      r[t] = fmaf(m2, vl, m3 * vr);
    }
  }
}

static size_t eq_smem_bytes() { return align_up(sizeof(EqShared), 16) + (size_t)kWarps * 2 * 32 * kEqRow * sizeof(float); }
static size_t comp_smem_bytes() { return align_up(sizeof(CompShared), 16) + (size_t)kWarps * 2 * 2 * 32 * kCpRow * sizeof(float); }

}  // namespace fx2

static int fx2_set_attributes() {
  using namespace fx2;
  static bool attr_done[64] = {false};
  int dev = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    MST_CUDA_OK(cudaFuncSetAttribute(eq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eq_smem_bytes()));
    MST_CUDA_OK(cudaFuncSetAttribute(comp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)comp_smem_bytes()));
    attr_done[dev] = true;
  }
  return 0;
}

static int fx2_chain_forward(const float* x, const float* params, float* y, int B, int L, float sample_rate, int stages, double* stats,
                             cudaStream_t st) {
  using namespace fx2;
  if (fx2_set_attributes()) return 1;
  const int vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0) ? 1 : 0;
  eq_kernel<<<B, kThreads, eq_smem_bytes(), st>>>(x, params, y, stats, L, sample_rate, (stages & MST_FX_EQ) ? 1 : 0, vec, nullptr, 0);
  if (launch_ok("fx2::eq_kernel")) return 1;
  comp_kernel<<<B, kThreads, comp_smem_bytes(), st>>>(params, y, stats, L, sample_rate, stages, vec);
  if (launch_ok("fx2::comp_kernel")) return 1;
  dim3 grid(cdiv(L, kFinFrames), B);
  final_kernel<<<grid, kFinThreads, 0, st>>>(y, stats, L, vec);
  return launch_ok("fx2::final_kernel");
}

}  // namespace mst

using namespace mst;

extern "C" {

size_t mst_fx_workspace_bytes(int B, int L) {
  if (B <= 0 || L <= 0) return 0;
  return align_up((size_t)B * fx2::kFxStats * sizeof(double), 256);
}

int mst_fx_chain_forward(const float* x, const float* params, float* y, int B, int L, float sample_rate, int stages,
                         void* workspace, size_t workspace_bytes, void* stream) {
  MST_CHECK(x && params && y && workspace, "fx_chain_forward: null pointer");
  MST_CHECK(B > 0 && L > 0, "fx_chain_forward: bad shape B=%d L=%d", B, L);
  MST_CHECK(workspace_bytes >= mst_fx_workspace_bytes(B, L), "fx_chain_forward: workspace too small");
  MST_CHECK(sample_rate > 0.f, "fx_chain_forward: bad sample rate");
  return fx2_chain_forward(x, params, y, B, L, sample_rate, stages, reinterpret_cast<double*>(workspace), (cudaStream_t)stream);
}

int mst_biquad_cascade(const float* x, const double* coef, int n_sections, float* y, int B, int L, void* workspace,
                       size_t workspace_bytes, void* stream) {
  MST_CHECK(x && coef && y && workspace, "biquad_cascade: null pointer");
  MST_CHECK(B > 0 && L > 0 && n_sections >= 1 && n_sections <= 5, "biquad_cascade: bad shape B=%d L=%d sections=%d", B, L, n_sections);
  MST_CHECK(workspace_bytes >= mst_fx_workspace_bytes(B, L), "biquad_cascade: workspace too small");
  if (fx2_set_attributes()) return 1;
  const int vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0) ? 1 : 0;
  fx2::eq_kernel<<<B, fx2::kThreads, fx2::eq_smem_bytes(), (cudaStream_t)stream>>>(x, nullptr, y, reinterpret_cast<double*>(workspace), L,
                                                                                  44100.f, 1, vec, coef, n_sections);
  return launch_ok("fx2::eq_kernel (biquad cascade)");
}

}  // extern "C"
