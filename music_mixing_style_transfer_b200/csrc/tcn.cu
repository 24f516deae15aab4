// tcn.cu -- MixFXcloner (FiLM-conditioned TCN) forward for sm_100a.
//
// Replaces (reference paths relative to /root/reference/mixing_style_transfer/networks/):
//   TCNModel.forward   architectures.py:135-147   14 blocks, dilation 2^n, then clamp(Conv1d(128->2,k=1)(x), -1, 1)
//   TCNBlock.forward   architectures.py:222-234   h = LeakyReLU(BN(conv1_dilated(x))); h = gamma*h + beta; h += res(x)
//   FiLM.forward       network_utils.py:180-182   [gamma|beta] = Linear(2048 -> 256)(cond)
//
// ---- data layout in HBM --------------------------------------------------------------------------------------------
// Between blocks the 128-channel activation is kept TIME-MAJOR (channels-last) and pre-split into two bf16 terms,
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi)): 16 mantissa bits in 4 bytes/element -- the same HBM bytes as fp32.
// One time step is one 512-byte row of 4 "planes" of 64 bf16:  [ch 0-63 hi | ch 0-63 lo | ch 64-127 hi | ch 64-127 lo]
//   act[b][t][plane][64]   (b = segment).  A plane row is exactly one 128-byte swizzle row, so a TMA box of
// {64 ch, 128 t} lands in shared memory as a canonical K-major SWIZZLE_128B UMMA operand tile.  Zero padding of the
// dilated convolution = TMA out-of-bounds zero fill along t (segments are a separate tensor dimension, so taps never
// leak across segments).  Weights are BN-folded, split the same way and stored tap-major:
//   w[tap][kc][hi|lo][co 128][ci 64]   (kc = input-channel half).
//
// ---- the hot kernel (blocks 1..13, 98 % of all FLOPs) ---------------------------------------------------------------
// tcn_block_umma_kernel: im2col-free dilated implicit GEMM on the 5th-gen tensor cores.
//   D[t, co] = sum_{tap, ci} X[t + (tap-7)*d, ci] * W[tap][co][ci]        M = 128 time rows, N = 128 co, K = 15*128
// fp32-grade accuracy from the tensor cores via split operands: by default fp16(X)*fp16(W) + two e4m3 correction products
// (FMT 1 = MST_TCN_F16F8, the format of tcn_f8.cu: 2 bf16-MMA equivalents per algorithmic MMA; activations must stay within
// +-448, which the epilogue checks and reports through `range_flag`), or the 3-product bf16 split Xhi*Whi + Xlo*Whi + Xhi*Wlo
// (FMT 0 = MST_TCN_BF16X3, fp32 range); fp32 accumulate in TMEM.  The precision is a per-call argument.  Persistent, warp-specialised,
// one CTA per SM:
//   warp 0   TMA producer (one lane): streams 32 KB slots (W tap-group, X sub-tile) through a 6-deep mbarrier ring
//   warp 1   MMA issuer: the whole warp runs the uniform loop, the tcgen05.mma / commit instructions are predicated on the
//            elected lane -> straight UTCHMMA sequences from uniform registers (8 or 12 MMAs per slot pair)
//   warp 2   TMEM allocator (512 columns = 2 tiles x 2 sub-tiles x 128 fp32 columns -> double-buffered accumulators)
//   warps 4-7 epilogue: tcgen05.ld -> +BN bias -> LeakyReLU -> FiLM -> + res*x_in (x_in tile TMA-loaded, requested one piece
//            ahead) -> re-split -> swizzled shared tile -> TMA store;
//            the LAST block instead fuses Conv1d(128->2,k=1)+clamp and writes the fp32 [B,2,L] output directly (the
//            128-channel tensor of block 13 never touches HBM).
// Taps whose shifted tile lies entirely in the zero padding are skipped by producer and issuer alike.
// For the f16f8 format two CTAs on one TPC run as a PAIR (cluster of 2, tcgen05 cta_group::2): one M = 256 MMA stream issued by
// the leader over two consecutive work items, every weight tile staged half in each CTA (template parameter CG, see the
// kernel).  Block 0 (C_in = 2) of that format is its own small tensor-core kernel (tcn_b0.cuh); the bf16 x 3 format keeps the
// CUDA-core block 0 below.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "f32x2.cuh"
#include "sm100_ptx.cuh"

// Development-only ablation switches (tools/tcn_ablate.sh builds side libraries with -DMST_TCN_ABLATE=<mask>; the shipped
// library is always built with 0): 1 = issue no MMAs (barrier traffic and commits only), 2 = epilogue drains TMEM but does no
// math / residual load / store, 4 = producer moves no activation bytes (weights only).
#ifndef MST_TCN_ABLATE
#define MST_TCN_ABLATE 0
#endif
// CTA-pair variant of the dilated blocks (tcgen05 cta_group::2, see tcn_block_umma_kernel): 2 = on for the f16f8 format
#ifndef MST_TCN_CG
#define MST_TCN_CG 2
#endif

namespace mst {

using namespace f2;

constexpr int kCh = MST_TCN_CH;      // 128
constexpr int kTaps = MST_TCN_K;     // 15
constexpr int kRowBytes = 512;       // one time step: 4 planes x 64 bf16
constexpr int kSubRows = 128;        // UMMA M
constexpr int kTileRows = 256;       // two sub-tiles share every weight slot
constexpr int kStageBytes = 32768;   // epilogue staging: 64-channel hi tile (128 x 128 B) + lo tile
constexpr int kRingBytes = 196608;   // operand ring: 6 x 32 KB (64-channel slots)
constexpr int kMaxSlots = 6;
constexpr size_t kWBytesPerLayer = (size_t)kTaps * kCh * kCh * 2 * 2;  // hi + lo bf16 = 983,040 B

// Both operand formats of blocks >= 1 are packed (983,040 B per layer each), so the precision is a per-call choice and a
// forward that leaves the f16f8 range can be repeated in bf16 x 3 without touching the weights again.
struct TcnPacked {
  size_t w0, wumma, wbf16, bn_bias, res, film_w, film_b, out_w, out_b, f8_scale, total;
};

static int tcn_layout(const mst_tcn_config* c, TcnPacked* o) {
  MST_CHECK(c, "tcn config is null");
  MST_CHECK(c->channels == kCh && c->kernel_size == kTaps,
            "tcn config: only channel_width=128 / kernel_size=15 has a CUDA path (got %d / %d)", c->channels,
            c->kernel_size);
  MST_CHECK(c->n_blocks >= 2 && c->n_blocks <= 32, "tcn config: n_blocks %d out of range [2,32]", c->n_blocks);
  MST_CHECK(c->n_inputs >= 1 && c->n_inputs <= 2 && c->n_outputs >= 1 && c->n_outputs <= 2,
            "tcn config: n_inputs/n_outputs must be 1 or 2");
  MST_CHECK(c->cond_dim > 0 && c->cond_dim % 32 == 0 && c->cond_dim <= 4096, "tcn config: cond_dim %d unsupported", c->cond_dim);
  MST_CHECK(c->dilation_growth >= 1 && c->stack_size >= 1, "tcn config: bad dilation");
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes, 1024); return r; };
  o->w0 = take((size_t)kCh * c->n_inputs * kTaps * 4);
  o->wumma = take((size_t)(c->n_blocks - 1) * kWBytesPerLayer);   // f16f8: [tap][fp16 ci 0-63 | fp16 ci 64-127 | e4m3 | e4m3][co][128 B]
  o->wbf16 = take((size_t)(c->n_blocks - 1) * kWBytesPerLayer);   // bf16 x 3: [tap][kc][hi|lo][co][ci 64]
  o->bn_bias = take((size_t)c->n_blocks * kCh * 4);
  o->res = take((size_t)c->n_blocks * kCh * 4);
  o->film_w = take((size_t)c->n_blocks * 2 * kCh * c->cond_dim * 4);
  o->film_b = take((size_t)c->n_blocks * 2 * kCh * 4);
  o->out_w = take((size_t)c->n_outputs * kCh * 4);
  o->out_b = take((size_t)c->n_outputs * 4);
  o->f8_scale = take((size_t)c->n_blocks * 2 * 4);   // per block: 1/(S*2^11) and a scratch word (f16f8 mode)
  o->total = off;
  return 0;
}

static long long block_dilation(const mst_tcn_config* c, int n) {
  long long d = 1;
  for (int i = 0; i < n % c->stack_size; ++i) d *= c->dilation_growth;  // architectures.py:122
  return d;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// =====================================================================================================================
// weight packing
// =====================================================================================================================
__global__ void tcn_pack_block0_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                       const float* __restrict__ bn_var, int n_in, float* __restrict__ w0) {
  const int n = kCh * n_in * kTaps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i / (n_in * kTaps);
    w0[i] = w[i] * (bn_w[co] / sqrtf(bn_var[co] + 1e-5f));
  }
}

// w: [co 128][ci 128][tap 15] fp32  ->  out[tap][kc][split][co][ci KCH] bf16 (KCH = 64 or 32 input channels per
// pipeline slot), BN scale folded before the split
__global__ void tcn_pack_umma_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                     const float* __restrict__ bn_var, __nv_bfloat16* __restrict__ out, int kch) {
  const int n = kTaps * kCh * kCh;
  const int nkc = kCh / kch;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cil = i % kch;
    const int co = (i / kch) % kCh;
    const int kc = (i / (kch * kCh)) % nkc;
    const int tap = i / (kch * kCh * nkc);
    const float s = bn_w[co] / sqrtf(bn_var[co] + 1e-5f);
    const float v = w[((size_t)co * kCh + kc * kch + cil) * kTaps + tap] * s;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const size_t base = ((size_t)(tap * nkc + kc) * 2) * kCh * kch;
    out[base + (size_t)co * kch + cil] = hi;
    out[base + (size_t)kCh * kch + (size_t)co * kch + cil] = lo;
  }
}

__global__ void tcn_pack_vec_kernel(const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                    const float* __restrict__ bn_mean, const float* __restrict__ bn_var,
                                    const float* __restrict__ res_w, float* __restrict__ bn_bias,
                                    float* __restrict__ res) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kCh) return;
  const float s = bn_w[c] / sqrtf(bn_var[c] + 1e-5f);
  bn_bias[c] = bn_b[c] - bn_mean[c] * s;
  res[c] = res_w[c];
}

// =====================================================================================================================
// FiLM precompute: one warp per (block, output row of Linear(2048 -> 256)); weight row kept in registers and reused for
// every conditioning row.  Emits (bn_bias, gamma, beta, res_scale) per (block, cond row, channel), interleaved over channel
// PAIRS -- film[block][cond row][pair][component][channel & 1], two float4 per pair -- so that the epilogues, which work on
// two channels at a time with packed fma.rn.f32x2, load their operands as ready-made 64-bit register pairs.
// =====================================================================================================================
template <int MAX_PER_LANE>
__global__ void __launch_bounds__(256)
tcn_film_kernel(const float* __restrict__ film_w, const float* __restrict__ film_b, const float* __restrict__ bn_bias,
                const float* __restrict__ res, const float* __restrict__ cond, int n_blocks, int n_cond, int cond_dim,
                float* __restrict__ out) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_blocks * 2 * kCh) return;
  const int n = gw / (2 * kCh), row = gw % (2 * kCh);
  const float* wr = film_w + (size_t)gw * cond_dim;
  float wreg[MAX_PER_LANE];
  const int per_lane = cond_dim / 32;
#pragma unroll
  for (int i = 0; i < MAX_PER_LANE; ++i) wreg[i] = (i < per_lane) ? __ldg(wr + lane + 32 * i) : 0.f;
  const float bias = film_b[gw];
  const int c = row & (kCh - 1);
  const int comp = row < kCh ? 1 : 2;  // gamma rows first, then beta (torch.split, network_utils.py:181)
  for (int bc = 0; bc < n_cond; ++bc) {
    const float* cr = cond + (size_t)bc * cond_dim;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_PER_LANE; ++i)
      if (i < per_lane) s = fmaf(wreg[i], __ldg(cr + lane + 32 * i), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float* q = out + (((size_t)n * n_cond + bc) * kCh + (c & ~1)) * 4 + (c & 1);     // pair base + channel parity
      q[2 * comp] = s + bias;
      if (comp == 1) {
        q[0] = bn_bias[n * kCh + c];
        q[6] = res[n * kCh + c];
      }
    }
  }
}

// =====================================================================================================================
// block 0 (C_in = 1 or 2, K = 15 or 30): CUDA cores, fp32 in -> split-bf16 activation rows out.  HBM-write bound.
// Lane l owns channels {2l, 2l+1, 64+2l, 65+2l} so each warp store is one full 128-byte plane row.
// =====================================================================================================================
template <int NIN>
__global__ void __launch_bounds__(256, 2)
tcn_block0_kernel(const float* __restrict__ x, const float* __restrict__ w0, const float4* __restrict__ film, int n_cond,
                  uint8_t* __restrict__ act, int T, int Ts) {
  constexpr int ROWS = 256, HALO = 7;
  __shared__ float xs[NIN][ROWS + 2 * HALO + 4];
  const int b = blockIdx.y, t0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NIN * (ROWS + 2 * HALO); i += 256) {
    const int ci = i / (ROWS + 2 * HALO), m = i % (ROWS + 2 * HALO);
    const int t = t0 - HALO + m;  // dilation 1, zero padding 7 (architectures.py:199-206)
    xs[ci][m] = (t >= 0 && t < T) ? __ldg(x + ((size_t)b * NIN + ci) * T + t) : 0.f;
  }
  // warp = (row group of 64 rows, channel half): lane owns channels {2l, 2l+1} + 64*half -> every warp store is one
  // full 128-byte plane row; 60 weights per lane keep the kernel at 2 CTAs / SM
  const int half = warp & 1, rgrp = warp >> 1;
  const int ch[2] = {64 * half + 2 * lane, 64 * half + 2 * lane + 1};
  u64 wr[NIN * kTaps];      // (channel 2l, channel 2l + 1) packed: one fma.rn.f32x2 per tap and row, bit-identical to two scalar FMAs
  float4 P[2];
#pragma unroll
  for (int i = 0; i < NIN * kTaps; ++i) wr[i] = pk(__ldg(w0 + ch[0] * NIN * kTaps + i), __ldg(w0 + ch[1] * NIN * kTaps + i));
  {
    // (bn_bias, gamma, beta, res) of the channel pair {ch[0], ch[1]}: two float4 of the pair-interleaved table
    const float4* fp = film + (size_t)(n_cond > 1 ? b : 0) * kCh + ch[0];
    const float4 A = __ldg(fp), Bq = __ldg(fp + 1);
    P[0] = make_float4(A.x, A.z, Bq.x, Bq.z);
    P[1] = make_float4(A.y, A.w, Bq.y, Bq.w);
  }
  // res = Conv1d(in, 128, k=1, groups=in): out channel c reads input channel c / (128/in)  (architectures.py:216-220)
  const int res_ci = ch[0] / (kCh / NIN);
  __syncthreads();
  // four consecutive rows per iteration share their input window: 18 broadcast reads per input channel feed 240 FMAs
  constexpr int RB = 4;
  for (int r = rgrp * 64; r < rgrp * 64 + 64; r += RB) {
    if (t0 + r >= T) break;
    u64 acc[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) acc[u] = 0ull;
#pragma unroll
    for (int ci = 0; ci < NIN; ++ci) {
#pragma unroll
      for (int m = 0; m < kTaps + RB - 1; ++m) {
        const u64 xv = dup(xs[ci][r + m]);
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int j = m - u;   // tap of row r+u that touches window position m
          if (j >= 0 && j < kTaps) acc[u] = fma2(wr[ci * kTaps + j], xv, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int t = t0 + r + u;
      if (t >= T) break;
      __nv_bfloat16 hi[2], lo[2];
      const float xin = xs[res_ci][r + u + HALO];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v = (q == 0 ? lo_of(acc[u]) : hi_of(acc[u])) + P[q].x;
        v = v > 0.f ? v : 0.01f * v;
        v = fmaf(P[q].y, v, P[q].z) + P[q].w * xin;
        split_bf16(v, hi[q], lo[q]);
      }
      uint32_t* row = reinterpret_cast<uint32_t*>(act + ((size_t)b * Ts + t) * kRowBytes);
      row[(2 * half) * 32 + lane] = pack_bf16(hi[0], hi[1]);
      row[(2 * half + 1) * 32 + lane] = pack_bf16(lo[0], lo[1]);
    }
  }
}

// fp32 [B][128][T]  <->  split-bf16 activation rows (module-level TCNBlock surface and per-block parity tests)
__global__ void __launch_bounds__(256) tcn_act_pack_kernel(const float* __restrict__ x, uint8_t* __restrict__ act, int T, int Ts) {
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
    __nv_bfloat16 hi[4], lo[4];
    const int ch[4] = {2 * lane, 2 * lane + 1, 64 + 2 * lane, 65 + 2 * lane};
#pragma unroll
    for (int q = 0; q < 4; ++q) split_bf16(tile[ch[q]][r], hi[q], lo[q]);
    uint32_t* row = reinterpret_cast<uint32_t*>(act + ((size_t)b * Ts + t) * kRowBytes);
    row[0 * 32 + lane] = pack_bf16(hi[0], hi[1]);
    row[1 * 32 + lane] = pack_bf16(lo[0], lo[1]);
    row[2 * 32 + lane] = pack_bf16(hi[2], hi[3]);
    row[3 * 32 + lane] = pack_bf16(lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(256) tcn_act_unpack_kernel(const uint8_t* __restrict__ act, float* __restrict__ y, int T, int Ts) {
  __shared__ float tile[kCh][33];
  const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < 32; r += 8) {
    const int t = t0 + r;
    if (t >= T) continue;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(act + ((size_t)b * Ts + t) * kRowBytes);
    const uint32_t h0 = row[lane], l0 = row[32 + lane], h1 = row[64 + lane], l1 = row[96 + lane];
    tile[2 * lane][r] = bf16_lo_f(h0) + bf16_lo_f(l0);
    tile[2 * lane + 1][r] = bf16_hi_f(h0) + bf16_hi_f(l0);
    tile[64 + 2 * lane][r] = bf16_lo_f(h1) + bf16_lo_f(l1);
    tile[65 + 2 * lane][r] = bf16_hi_f(h1) + bf16_hi_f(l1);
  }
  __syncthreads();
  for (int c = warp; c < kCh; c += 8) {
    const int t = t0 + lane;
    if (t < T) y[((size_t)b * kCh + c) * T + t] = tile[c][lane];
  }
}

// =====================================================================================================================
// the tcgen05 kernel
// =====================================================================================================================
struct TcnLayerArgs {
  int B, T, dilation, tiles_per_seg, n_tiles, n_cond;
  int pair_m;           // mode 1: sub-tiles per half block = dilation / 128
  int pad_rows;         // T is not a multiple of the 256-row segment stride unit: rows [T, tcn_seg_rows(T)) exist and must stay zero
  const float* inv_scale;   // FMT 1 (f16f8): 1 / (S * 2^11) of this layer's packed weights (tcn_f8.cu)
  unsigned int* range_flag; // FMT 1: receives max |activation| (float bits, atomicMax) if it exceeds the e4m3 range; may be null
  const float4* film;   // this block's [n_cond][128] (bn_bias, gamma, beta, res)
  int n_out;
  const float* out_w;   // [n_out][128]
  const float* out_b;   // [n_out]
  float* out;
};

struct __align__(8) TcnBarriers {
  uint64_t full[kMaxSlots], empty[kMaxSlots];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t stage_full;
  uint32_t tmem_base;
};

constexpr size_t kTcnSmemBytes = 1024 /*align slack*/ + (size_t)kRingBytes + kStageBytes + 256;

// Work-item geometry.  MODE selects how the two 128-row sub-tiles of a work item sit in time:
//   0  one 256-row tile, sub-tile 1 directly after sub-tile 0 (any dilation, any length).
//   1  "far pairing", dilation d a multiple of 128: sub-tile 1 sits d rows after sub-tile 0.  The time axis is cut into blocks
//      of 2d rows; item i of a block pairs rows [i*128, +128) of its first half with the same rows of its second half.
//   2  "interleaved pairing", d < 128 dividing 128, T a multiple of 2d: the work item is still one 256-row span, but sub-tile 0
//      takes the FIRST d rows of each of its 128/d blocks of 2d rows and sub-tile 1 the second d rows -- 128 rows gathered by
//      one 5-D TMA box {128 B, d rows, 1 half, 128/d blocks} over the activation viewed as [b][block][half][row][512 B].
//   In modes 1 and 2 sub-tile 1 is sub-tile 0 shifted by exactly d rows, so tap j of sub-tile 1 reads the rows tap j+1 of
//   sub-tile 0 reads: one activation slot serves two MMA groups and the activation bytes entering the SM halve (see the
//   producer / issuer loops).  `ts` below is always the first row of a (shifted) sub-tile; in mode 2 it is a multiple of d.
struct TcnTile { int b; long long r0, r1; bool sub0, sub1; };
template <int MODE>
__device__ __forceinline__ TcnTile tcn_tile(int tile, const TcnLayerArgs& a) {
  TcnTile c;
  c.b = tile / a.tiles_per_seg;
  const int p = tile - c.b * a.tiles_per_seg;
  if (MODE == 1) {
    // block index fastest: consecutive work items (= the CTAs of one wave) are 2d rows apart and share 14 of their 16 tap
    // tiles, so a wave's working set stays in L2 also for the large dilations (with i fastest the 148 x 16 tiles of a wave are
    // all distinct at d >= 2048: 150 MB, DRAM reads up to 3x the algorithmic bytes)
    const int nblk = a.tiles_per_seg / a.pair_m;
    const int i = p / nblk, blk = p - i * nblk;
    c.r0 = (long long)blk * 2 * a.dilation + (long long)i * kSubRows;
    c.r1 = c.r0 + a.dilation;
  } else if (MODE == 2) {
    c.r0 = (long long)p * kTileRows;
    c.r1 = c.r0 + a.dilation;
  } else {
    c.r0 = (long long)p * kTileRows;
    c.r1 = c.r0 + kSubRows;
  }
  c.sub0 = c.r0 < a.T;
  c.sub1 = c.r1 < a.T;
  return c;
}
// does the (shifted) sub-tile starting at row ts touch the real signal [0, T)?  (otherwise it is all zero padding).  Mode 2
// sub-tiles span 256 - d rows; the looser bound only costs a few all-zero taps at the segment edges.
template <int MODE>
__device__ __forceinline__ bool tap_live(long long ts, int T) { return ts < (long long)T && ts + (MODE == 2 ? kTileRows : kSubRows) > 0; }
// time row of sub-tile row `rl` (modes 0 / 1: consecutive rows; mode 2: d rows out of every 2d)
template <int MODE>
__device__ __forceinline__ int tile_row(int ts, int rl, int d) { return MODE == 2 ? ts + (rl / d) * 2 * d + rl % d : ts + rl; }

// FMT 0: bf16 hi/lo operands, three products per K-step (the default).
// FMT 1: the "2 tensor units" split of tcn_f8.cu -- rows [fp16 ch 0-63 | fp16 ch 64-127 | e4m3 (x - hi) 2^11, ch 0-127 | e4m3 x, ch
//        0-127], weights per tap [fp16 W S 2^11 ci 0-63 | ci 64-127 | e4m3 W S | e4m3 lo]; operand group 0 = the two fp16 tiles
//        (kind::f16), group 1 = the two e4m3 tiles (kind::f8f6f4, tile i of X times tile i of W), all into ONE accumulator; the
//        epilogue multiplies by 1 / (S 2^11).  Tensor maps are byte-typed in this mode (coordinates in bytes).
// FUSE:  the last block -- Conv1d(128 -> n_out, k=1) + clamp in the epilogue, no activation store.
// CG:    1 = one CTA per work item; 2 = CTA pair (cluster of 2 on one TPC, tcgen05 cta_group::2): the pair takes two consecutive
//        work items, one per CTA, as ONE M = 256 MMA stream issued by the leader; every weight tile is staged half in each CTA
//        (64 of its 128 output-channel rows), so the weight bytes entering an SM halve.  Both CTAs walk the same (operand group,
//        tap) steps -- a step runs when it is live for either work item; a CTA whose own item (or tap) lies outside the signal
//        loads zero-filled rows and stores nothing.
template <int KCH, int MODE, int FMT, bool FUSE, int CG>
__global__ void __launch_bounds__(256, 1)
tcn_block_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_xs, const __grid_constant__ CUtensorMap tm_y,
                      const __grid_constant__ CUtensorMap tm_l8, const __grid_constant__ CUtensorMap tm_y8,
                      const TcnLayerArgs a) {
  // tm_x / tm_w: operand boxes {KCH ch, 128 rows}, swizzle = 2*KCH bytes;  tm_xs / tm_y: epilogue boxes {64 ch, 128 rows};
  // tm_l8 / tm_y8 (FMT 1 only): epilogue boxes {64 bytes, 128 rows}, SWIZZLE_64B, for the e4m3 half planes
  static_assert(FMT == 0 || KCH == 64, "the f16f8 format uses 128-byte operand rows");
  static_assert(CG == 1 || (KCH == 64 && (MODE != 0 || FMT == 1)), "the CTA pair uses the operand-group-outermost schedule");
  constexpr bool PAIRED = MODE != 0;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
  const int n_workers = CG == 2 ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int worker = CG == 2 ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int n_items = CG == 2 ? (a.n_tiles + 1) / 2 : a.n_tiles;        // work items of a worker step (pairs of tiles for CG 2)
  // activation boxes: 3-D {columns, 128 rows, segment} in modes 0 / 1, 5-D {columns, d rows, half, 128/d blocks, segment} in mode 2
  auto act_load = [&](const CUtensorMap* m, uint64_t* bar, void* dst, int col, int ts, int b) {
    if (MODE == 2) { const int hh = ts / a.dilation; ptx::tma_load_5d(m, bar, dst, col, 0, hh & 1, hh >> 1, b); }
    else ptx::tma_load_3d(m, bar, dst, col, ts, b);
  };
  // operand loads of the CTA pair: completion is signalled on the LEADER's barrier (cluster address)
  auto act_load_cg2 = [&](const CUtensorMap* m, uint32_t bar_caddr, void* dst, int col, int ts, int b) {
    if (MODE == 2) { const int hh = ts / a.dilation; ptx::tma_load_5d_cg2(m, bar_caddr, dst, col, 0, hh & 1, hh >> 1, b); }
    else ptx::tma_load_3d_cg2(m, bar_caddr, dst, col, ts, b);
  };
  // the two work items of a worker step: `c` = this CTA's, `o` = the partner's (CG 1: o == c)
  auto item_tiles = [&](int q, TcnTile& c, TcnTile& o) {
    const int t_own = CG == 2 ? 2 * q + (int)cta_rank : q, t_oth = CG == 2 ? 2 * q + 1 - (int)cta_rank : q;
    c = tcn_tile<MODE>(t_own, a);
    o = tcn_tile<MODE>(t_oth, a);
    if (t_own >= a.n_tiles) { c.sub0 = false; c.sub1 = false; }
    if (t_oth >= a.n_tiles) { o.sub0 = false; o.sub1 = false; }
  };
  auto act_store = [&](const CUtensorMap* m, const void* src, int col, int ts, int b) {
    if (MODE == 2) { const int hh = ts / a.dilation; ptx::tma_store_5d(m, src, col, 0, hh & 1, hh >> 1, b); }
    else ptx::tma_store_3d(m, src, col, ts, b);
  };
  constexpr int kCoord = FMT == 1 ? 2 : 1;             // byte-typed maps: column coordinates in bytes instead of bf16 elements
  constexpr int kKcPerTap = kCh / KCH;                 // 2 or 4 input-channel chunks per tap
  constexpr int kHalf = kSubRows * KCH * 2;            // bytes of one hi (or lo) operand tile: 16 KB / 8 KB
  constexpr int kSlotBytes = 2 * kHalf;
  constexpr int kNumSlots = kRingBytes / kSlotBytes;   // 6
  constexpr int kK16 = KCH / 16;                       // MMA K-steps per slot
  constexpr int kSwz = KCH * 2;                        // swizzle span in bytes (128 / 64)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET into the shared array (not integer pointer arithmetic): the compiler keeps the shared
  // state space and emits LDS / STS instead of generic LD / ST for the epilogue's staging accesses
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                  // kNumSlots x kSlotBytes, 1024-aligned
  uint8_t* staging = smem + kRingBytes;                  // 32 KB epilogue tile (hi 16 KB | lo 16 KB), SWIZZLE_128B
  TcnBarriers* bars = reinterpret_cast<TcnBarriers*>(staging + kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_w);
    ptx::prefetch_tensormap(&tm_xs);
    ptx::prefetch_tensormap(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNumSlots; ++i) {
      ptx::mbar_init(&bars->full[i], CG);            // one producer arrival per CTA of the pair
      ptx::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->tmem_full[i], 1);
      ptx::mbar_init(&bars->tmem_empty[i], 128 * CG);  // the epilogue threads of both CTAs release the leader's accumulators
    }
    ptx::mbar_init(&bars->stage_full, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 2) {
    if (CG == 2) { ptx::tmem_alloc_cg2(&bars->tmem_base, 512); ptx::tmem_relinquish_cg2(); }
    else { ptx::tmem_alloc(&bars->tmem_base, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();                  // the partner's barriers exist before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const long long d = a.dilation;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // one lane only: this warp spends most of its time polling `empty` barriers, and 32 polling lanes slow the issuer's
    // barrier traffic down (measured: tensor-pipe activity 76 -> 68 % with a warp-convergent producer)
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      auto next = [&]() { if (++slot == kNumSlots) { slot = 0; phase ^= 1; } };
      auto load_w = [&](int j, int kc) {
        // weights: rows ((j*kKcPerTap+kc)*2 + split)*128 .. : hi tile then lo tile
        ptx::mbar_wait(&bars->empty[slot], phase ^ 1);
        uint8_t* dst = ring + (size_t)slot * kSlotBytes;
        const int wrow = ((j * kKcPerTap + kc) * 2) * kCh;
        if (CG == 2) {
          // this CTA's 64 output-channel rows of both tiles (tm_w boxes are 64 rows high), same slot offsets as the full tiles
          const uint32_t fb = ptx::mapa(ptx::smem_u32(&bars->full[slot]), 0);
          ptx::mbar_expect_tx_cluster(fb, kSlotBytes / 2);
          ptx::tma_load_2d_cg2(&tm_w, fb, dst, 0, wrow + 64 * (int)cta_rank);
          ptx::tma_load_2d_cg2(&tm_w, fb, dst + kHalf, 0, wrow + kCh + 64 * (int)cta_rank);
        } else {
          ptx::mbar_expect_tx(&bars->full[slot], kSlotBytes);
          ptx::tma_load_2d(&tm_w, &bars->full[slot], dst, 0, wrow);
          ptx::tma_load_2d(&tm_w, &bars->full[slot], dst + kHalf, 0, wrow + kCh);
        }
        next();
      };
      auto load_x = [&](int kc, long long ts, int b) {
        // activation columns of this chunk: plane (hi / lo) of channel half kc*KCH/64, offset inside the plane
        const int c_hi = (((kc * KCH) / 64) * 128 + (kc * KCH) % 64) * kCoord, c_lo = c_hi + 64 * kCoord;
        ptx::mbar_wait(&bars->empty[slot], phase ^ 1);
        if (MST_TCN_ABLATE & 4) { ptx::mbar_arrive(&bars->full[slot]); next(); return; }
        uint8_t* dst = ring + (size_t)slot * kSlotBytes;
        if (CG == 2) {
          const uint32_t fb = ptx::mapa(ptx::smem_u32(&bars->full[slot]), 0);
          ptx::mbar_expect_tx_cluster(fb, kSlotBytes);
          act_load_cg2(&tm_x, fb, dst, c_hi, (int)ts, b);
          act_load_cg2(&tm_x, fb, dst + kHalf, c_lo, (int)ts, b);
        } else {
          ptx::mbar_expect_tx(&bars->full[slot], kSlotBytes);
          act_load(&tm_x, &bars->full[slot], dst, c_hi, (int)ts, b);
          act_load(&tm_x, &bars->full[slot], dst + kHalf, c_lo, (int)ts, b);
        }
        next();
      };
      for (int q = worker; q < n_items; q += n_workers) {
        TcnTile c, o;
        item_tiles(q, c, o);
        if (!c.sub0 && !o.sub0) continue;
        if (PAIRED || FMT == 1) {
          // operand group (channel half / MMA kind) outermost.  Slot order per executed step (kc, j): W, [rows of sub-tile 0's
          // tap j unless step j-1 loaded them as sub-tile 1's tap j-1 (PAIRED only)], [rows of sub-tile 1's tap j].  Liveness is
          // the union over the work items of the step (CG 2); this CTA always loads ITS item's rows (zero-filled outside).
          for (int kc = 0; kc < kKcPerTap; ++kc) {
            bool prev_live1 = false;
            for (int j = 0; j < kTaps; ++j) {
              const long long ts0 = c.r0 + (long long)(j - 7) * d, ts1 = c.r1 + (long long)(j - 7) * d;
              const long long os0 = o.r0 + (long long)(j - 7) * d, os1 = o.r1 + (long long)(j - 7) * d;
              const bool live0 = (c.sub0 && tap_live<MODE>(ts0, a.T)) || (o.sub0 && tap_live<MODE>(os0, a.T));
              const bool live1 = (c.sub1 && tap_live<MODE>(ts1, a.T)) || (o.sub1 && tap_live<MODE>(os1, a.T));
              if (!live0 && !live1) { prev_live1 = false; continue; }
              load_w(j, kc);
              const bool resident = PAIRED && prev_live1 && live0;   // ts0 == rows of sub-tile 1 at tap j-1, still in its slot
              if (live0 && !resident) load_x(kc, ts0, c.b);
              if (live1) load_x(kc, ts1, c.b);
              prev_live1 = live1;
            }
          }
        } else {
          for (int j = 0; j < kTaps; ++j) {
            const long long ts0 = c.r0 + (long long)(j - 7) * d, ts1 = c.r1 + (long long)(j - 7) * d;
            const bool live0 = tap_live<MODE>(ts0, a.T), live1 = c.sub1 && tap_live<MODE>(ts1, a.T);
            if (!live0 && !live1) continue;
            for (int kc = 0; kc < kKcPerTap; ++kc) {
              load_w(j, kc);
              if (live0) load_x(kc, ts0, c.b);
              if (live1) load_x(kc, ts1, c.b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // all 32 lanes run this (warp-uniform) loop; only the tcgen05 instructions are predicated on the elected lane.
    // CTA pair: the leader CTA issues for both (M = 256); the partner's warp 1 has nothing to do.
    if (CG == 1 || cta_rank == 0) {
      const uint32_t leader = ptx::elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = FMT == 1 ? ptx::umma_idesc_f16_f32(kSubRows * CG, kCh)     // format code 0 = F16 (kind::f16) = E4M3 (kind::f8f6f4)
                                          : ptx::umma_idesc_bf16_f32(kSubRows * CG, kCh);
      auto mma_f16 = [&](uint32_t dt, uint64_t ad, uint64_t bd, uint32_t acc) {
        if (CG == 2) ptx::umma_mma_f16kind_elect_cg2(dt, ad, bd, idesc, acc, leader);
        else ptx::umma_mma_f16kind_elect(dt, ad, bd, idesc, acc, leader);
      };
      auto mma_f8 = [&](uint32_t dt, uint64_t ad, uint64_t bd, uint32_t acc) {
        if (CG == 2) ptx::umma_mma_f8kind_elect_cg2(dt, ad, bd, idesc, acc, leader);
        else ptx::umma_mma_f8kind_elect(dt, ad, bd, idesc, acc, leader);
      };
      auto commit = [&](uint64_t* bar) {
        if (CG == 2) ptx::umma_commit_elect_cg2(bar, leader);
        else ptx::umma_commit_elect(bar, leader);
      };
      uint32_t slot = 0, phase = 0;
      bool f8_group = false;     // FMT 1: the slots being consumed hold the e4m3 tiles (operand group 1)
      auto next = [&]() { if (++slot == kNumSlots) { slot = 0; phase ^= 1; } };
      // 3-product split: (Xhi, Whi) + (Xlo, Whi) + (Xhi, Wlo), KCH/16 K16 steps per slot
      auto issue_group = [&](uint32_t x_addr, uint32_t w_addr, uint32_t d_tmem, bool first) {
        if (MST_TCN_ABLATE & 1) return;
        const uint64_t xh = ptx::umma_desc_kmajor<kSwz>(x_addr), xl = ptx::umma_desc_kmajor<kSwz>(x_addr + kHalf);
        const uint64_t wh = ptx::umma_desc_kmajor<kSwz>(w_addr), wl = ptx::umma_desc_kmajor<kSwz>(w_addr + kHalf);
        if (FMT == 1) {
          // tile 0 of X times tile 0 of W, tile 1 times tile 1; four 32-byte K-steps each (K16 fp16 / K32 e4m3)
          if (f8_group) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              mma_f8(d_tmem, xh + adv, wh + adv, (first && k == 0) ? 0u : 1u);
              mma_f8(d_tmem, xl + adv, wl + adv, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);
              mma_f16(d_tmem, xh + adv, wh + adv, (first && k == 0) ? 0u : 1u);
              mma_f16(d_tmem, xl + adv, wl + adv, 1u);
            }
          }
          return;
        }
#pragma unroll
        for (int k = 0; k < kK16; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);  // +32 bytes along K inside the 128-byte swizzle row
          mma_f16(d_tmem, xh + adv, wh + adv, (first && k == 0) ? 0u : 1u);
          mma_f16(d_tmem, xl + adv, wh + adv, 1u);
          mma_f16(d_tmem, xh + adv, wl + adv, 1u);
        }
      };
      int it = 0;
      for (int q = worker; q < n_items; q += n_workers) {
        TcnTile c, o;
        item_tiles(q, c, o);
        if (!c.sub0 && !o.sub0) continue;
        const int buf = it & 1;
        ptx::mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * 2 + 0) * kCh, acc1 = tmem_base + (uint32_t)(buf * 2 + 1) * kCh;
        bool first0 = true, first1 = true;
        if (PAIRED || FMT == 1) {
          for (int kc = 0; kc < kKcPerTap; ++kc) {
            f8_group = FMT == 1 && kc == 1;
            int carried = -1;      // slot holding the rows sub-tile 1 used at the previous tap = rows of sub-tile 0 at this tap
            for (int j = 0; j < kTaps; ++j) {
              const long long ts0 = c.r0 + (long long)(j - 7) * d, ts1 = c.r1 + (long long)(j - 7) * d;
              const long long os0 = o.r0 + (long long)(j - 7) * d, os1 = o.r1 + (long long)(j - 7) * d;
              const bool live0 = (c.sub0 && tap_live<MODE>(ts0, a.T)) || (o.sub0 && tap_live<MODE>(os0, a.T));
              const bool live1 = (c.sub1 && tap_live<MODE>(ts1, a.T)) || (o.sub1 && tap_live<MODE>(os1, a.T));
              if (!live0 && !live1) { carried = -1; continue; }
              const uint32_t wslot = slot;
              ptx::mbar_wait(&bars->full[wslot], phase);
              const uint32_t w_addr = ptx::smem_u32(ring + (size_t)wslot * kSlotBytes);
              next();
              if (live0) {
                uint32_t xs;
                if (carried >= 0) {
                  xs = (uint32_t)carried;
                } else {
                  xs = slot;
                  ptx::mbar_wait(&bars->full[xs], phase);
                  next();
                }
                ptx::tc_fence_after();
                issue_group(ptx::smem_u32(ring + (size_t)xs * kSlotBytes), w_addr, acc0, first0);
                first0 = false;
                commit(&bars->empty[xs]);
              }
              carried = -1;
              if (live1) {
                const uint32_t xs = slot;
                ptx::mbar_wait(&bars->full[xs], phase);
                next();
                ptx::tc_fence_after();
                issue_group(ptx::smem_u32(ring + (size_t)xs * kSlotBytes), w_addr, acc1, first1);
                first1 = false;
                // the same rows are sub-tile 0's operand at tap j+1 (ts1 = r0 + (j-6) d): keep the slot if that tap runs
                if (PAIRED && j + 1 < kTaps) carried = (int)xs;
                else commit(&bars->empty[xs]);
              }
              commit(&bars->empty[wslot]);
            }
          }
        } else {
          for (int j = 0; j < kTaps; ++j) {
            const long long ts0 = c.r0 + (long long)(j - 7) * d, ts1 = c.r1 + (long long)(j - 7) * d;
            const bool live0 = tap_live<MODE>(ts0, a.T), live1 = c.sub1 && tap_live<MODE>(ts1, a.T);
            if (!live0 && !live1) continue;
            for (int kc = 0; kc < kKcPerTap; ++kc) {
              const uint32_t wslot = slot;
              ptx::mbar_wait(&bars->full[wslot], phase);
              const uint32_t w_addr = ptx::smem_u32(ring + (size_t)wslot * kSlotBytes);
              next();
              if (live0) {
                const uint32_t xs = slot;
                ptx::mbar_wait(&bars->full[xs], phase);
                next();
                ptx::tc_fence_after();
                issue_group(ptx::smem_u32(ring + (size_t)xs * kSlotBytes), w_addr, acc0, first0);
                first0 = false;
                commit(&bars->empty[xs]);
              }
              if (live1) {
                const uint32_t xs = slot;
                ptx::mbar_wait(&bars->full[xs], phase);
                next();
                ptx::tc_fence_after();
                issue_group(ptx::smem_u32(ring + (size_t)xs * kSlotBytes), w_addr, acc1, first1);
                first1 = false;
                commit(&bars->empty[xs]);
              }
              commit(&bars->empty[wslot]);
            }
          }
        }
        commit(&bars->tmem_full[buf]);
        ++it;
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue (128 threads, thread <-> one time row) ==============================
    // Pieces = (sub-tile, channel half).  The residual tile of the NEXT piece is requested as soon as the TMA store of the
    // current one has read the staging buffer (for the first piece of the next tile: long before its accumulator is ready),
    // by the same thread that issued the store.  What was tried and measured on top (profiles/r02_tcn_ablation.md): a second
    // epilogue group with its own staging buffer (sub-tile g drained by group g, 5-slot ring) -- more cycles per launch at a
    // higher clock, the same milliseconds: the kernel runs AT the 1000 W power cap, only energy per tile moves its time.
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int et = threadIdx.x - 128;       // 0..127
    const int rl = q * 32 + lane;           // row inside the sub-tile == TMEM lane
    uint64_t* stage_full = &bars->stage_full;
    uint32_t stage_phase = 0;
    __half2 vmax2 = __float2half2_rn(0.f);  // FMT 1: max |activation| this thread re-split (operand-range guard), per pair half
    // residual x_in of channels 64h .. 64h+63, rows ts .. ts+127 of segment b -> staging (issued by thread 0 of the group)
    auto request_residual = [&](int h, int ts, int b) {
      if (FMT == 1) {
        // the fp16 plane h (16 KB, SWIZZLE_128B) and half of the e4m3 remainder plane (64 bytes per row, 8 KB, SWIZZLE_64B)
        ptx::mbar_expect_tx(stage_full, 16384 + 8192);
        act_load(&tm_xs, stage_full, staging, 128 * h, ts, b);
        act_load(&tm_l8, stage_full, staging + 16384, 256 + 64 * h, ts, b);
      } else {
        ptx::mbar_expect_tx(stage_full, kStageBytes);
        act_load(&tm_xs, stage_full, staging, (2 * h) * 64, ts, b);              // x_in hi, ch 64h..
        act_load(&tm_xs, stage_full, staging + 16384, (2 * h + 1) * 64, ts, b);  // x_in lo
      }
    };
    // requests the h = 0 residual tile of the first live sub-tile of THIS CTA at or after (item q, sub) in processing order
    // (thread 0 only)
    auto request_next = [&](int q, int sub) {
      if (MST_TCN_ABLATE & 2) return;
      for (; q < n_items; q += n_workers, sub = 0) {
        TcnTile c, o;
        item_tiles(q, c, o);
        if (!c.sub0) continue;
        for (; sub < 2; ++sub) {
          const long long ts = sub == 0 ? c.r0 : c.r1;
          if (ts < a.T) { request_residual(0, (int)ts, c.b); return; }
        }
      }
    };
    if (et == 0) request_next(worker, 0);
    const uint32_t tmem_empty_lead[2] = {CG == 2 ? ptx::mapa(ptx::smem_u32(&bars->tmem_empty[0]), 0) : 0u,
                                         CG == 2 ? ptx::mapa(ptx::smem_u32(&bars->tmem_empty[1]), 0) : 0u};
    int it = 0;
    for (int q = worker; q < n_items; q += n_workers) {
      TcnTile c, o;
      item_tiles(q, c, o);
      if (!c.sub0 && !o.sub0) continue;
      const int b = c.b;
      const int buf = it & 1;
      const float4* film = a.film + (size_t)(a.n_cond > 1 && c.sub0 ? b : 0) * kCh;
      ptx::mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      for (int sub = 0; sub < 2; ++sub) {
        const int ts = (int)(sub == 0 ? c.r0 : c.r1);
        if (!c.sub0 || ts >= a.T || (MST_TCN_ABLATE & 2)) break;     // !c.sub0: the partner's item only (CTA pair)
        float o0 = 0.f, o1 = 0.f;
        // interleaved sub-tiles run over the padded segment: a row at or beyond T is stored as zeros (it is zero padding for
        // the next block's taps)
        const bool dead_row = MODE == 2 && a.pad_rows && tile_row<MODE>(ts, rl, a.dilation) >= a.T;
        for (int h = 0; h < 2; ++h) {
          uint32_t acc[64];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + sub) * kCh + h * 64);
          ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
          ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          ptx::tmem_ld_wait();
          ptx::mbar_wait(stage_full, stage_phase);
          stage_phase ^= 1;
          uint8_t* rowp = staging + rl * 128;
          if constexpr (FMT == 1) {
            // f16f8 rows: fp16 hi at staging (128-byte rows, SWIZZLE_128B), e4m3 remainder at +16 KB and e4m3 copy at +24 KB
            // (64-byte rows, SWIZZLE_64B: 16-byte chunk index XOR ((row / 2) mod 4))
            const u64 inv_scale2 = dup(__ldg(a.inv_scale));
            uint8_t* lrow = staging + 16384 + rl * 64;
            uint8_t* hrow8 = staging + 24576 + rl * 64;
            const int sw64 = (rl >> 1) & 3;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {     // 8 channels per iteration
              const int off = ((c8 ^ (rl & 7)) << 4);
              const int off8 = (((c8 >> 1) ^ sw64) << 4) + (c8 & 1) * 8;
              const uint4 xh = *reinterpret_cast<const uint4*>(rowp + off);
              const uint2 xl = *reinterpret_cast<const uint2*>(lrow + off8);
              const uint32_t xhw[4] = {xh.x, xh.y, xh.z, xh.w};
              const uint32_t xlw[2] = {xl.x, xl.y};
              uint32_t oh[4], ol[2] = {0, 0}, oh8[2] = {0, 0};
#pragma unroll
              for (int pr = 0; pr < 4; ++pr) {
                const int cl = c8 * 8 + 2 * pr;
                const int ch = h * 64 + cl;
                // two adjacent channels at a time in packed fma.rn.f32x2 (each half rounds like the scalar op): the pair-
                // interleaved table delivers (bn_bias, gamma) and (beta, res) of the pair as 64-bit operands
                const ulonglong2 Pa = __ldg(reinterpret_cast<const ulonglong2*>(film + ch));
                const ulonglong2 Pb = __ldg(reinterpret_cast<const ulonglong2*>(film + ch) + 1);
                // x_in = fp16 hi + e4m3 lo * 2^-11
                const float2 hif = __half22float2(*reinterpret_cast<const __half2*>(&xhw[pr]));
                const unsigned short l8pair = (unsigned short)((xlw[pr >> 1] >> (16 * (pr & 1))) & 0xFFFFu);
                const __half2_raw lraw = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)l8pair, __NV_E4M3);
                const float2 lof = __half22float2(__half2(lraw));
                const u64 xin = fma2(pk(lof.x, lof.y), dup(1.f / 2048.f), pk(hif.x, hif.y));
                u64 u = fma2(pk(__uint_as_float(acc[cl]), __uint_as_float(acc[cl + 1])), inv_scale2, Pa.x);
                const u64 ul = mul2(u, dup(0.01f));                   // LeakyReLU(0.01): max(u, 0.01 u)
                u = pk(fmaxf(lo_of(u), lo_of(ul)), fmaxf(hi_of(u), hi_of(ul)));
                u = fma2(Pb.y, xin, fma2(Pa.y, u, Pb.x));             // gamma u + beta + res x_in
                const float u0 = lo_of(u), u1 = hi_of(u);
                if (FUSE) {
                  o0 = fmaf(u0, __ldg(a.out_w + ch), o0);
                  o0 = fmaf(u1, __ldg(a.out_w + ch + 1), o0);
                  if (a.n_out > 1) {
                    o1 = fmaf(u0, __ldg(a.out_w + kCh + ch), o1);
                    o1 = fmaf(u1, __ldg(a.out_w + kCh + ch + 1), o1);
                  }
                } else {
                  const uint32_t hbits = ptx::cvt_f16x2_satfinite(u0, u1);       // fp16 pair, clamped to +-65504
                  vmax2 = __hmax2(vmax2, __habs2(*reinterpret_cast<const __half2*>(&hbits)));   // range guard on the packed pair
                  const float2 hb = __half22float2(*reinterpret_cast<const __half2*>(&hbits));
                  oh[pr] = hbits;
                  const u64 rem = mul2(fma2(pk(hb.x, hb.y), dup(-1.f), u), dup(2048.f));     // (u - fp16 u) 2^11, exact subtraction
                  const uint32_t l2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo_of(rem), hi_of(rem)), __NV_SATFINITE, __NV_E4M3);
                  const uint32_t h2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(u0, u1), __NV_SATFINITE, __NV_E4M3);
                  ol[pr >> 1] |= l2 << (16 * (pr & 1));
                  oh8[pr >> 1] |= h2 << (16 * (pr & 1));
                }
              }
              if (!FUSE) {
                if (MODE == 2 && dead_row) { oh[0] = oh[1] = oh[2] = oh[3] = 0u; ol[0] = ol[1] = 0u; oh8[0] = oh8[1] = 0u; }
                *reinterpret_cast<uint4*>(rowp + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                *reinterpret_cast<uint2*>(lrow + off8) = make_uint2(ol[0], ol[1]);
                *reinterpret_cast<uint2*>(hrow8 + off8) = make_uint2(oh8[0], oh8[1]);
              }
            }
          } else {
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const int off = ((c8 ^ (rl & 7)) << 4);  // 128-byte swizzle: 16-byte chunk index XOR (row mod 8)
              const uint4 xh = *reinterpret_cast<const uint4*>(rowp + off);
              const uint4 xl = *reinterpret_cast<const uint4*>(rowp + 16384 + off);
              const uint32_t xhw[4] = {xh.x, xh.y, xh.z, xh.w}, xlw[4] = {xl.x, xl.y, xl.z, xl.w};
              uint32_t oh[4], ol[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float v[2];
#pragma unroll
                for (int sidx = 0; sidx < 2; ++sidx) {
                  const int cl = c8 * 8 + e * 2 + sidx;
                  const int ch = h * 64 + cl;
                  const float* Pp = reinterpret_cast<const float*>(film + (ch & ~1)) + (ch & 1);   // pair-interleaved table
                  const float4 P = make_float4(__ldg(Pp), __ldg(Pp + 2), __ldg(Pp + 4), __ldg(Pp + 6));
                  const float xin = sidx == 0 ? bf16_lo_f(xhw[e]) + bf16_lo_f(xlw[e]) : bf16_hi_f(xhw[e]) + bf16_hi_f(xlw[e]);
                  float u = __uint_as_float(acc[cl]) + P.x;
                  u = fmaxf(u, 0.01f * u);
                  u = fmaf(P.y, u, P.z) + P.w * xin;
                  v[sidx] = u;
                  if (FUSE) {
                    o0 = fmaf(u, __ldg(a.out_w + ch), o0);
                    if (a.n_out > 1) o1 = fmaf(u, __ldg(a.out_w + kCh + ch), o1);
                  }
                }
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v[0], h0, l0);
                split_bf16(v[1], h1, l1);
                oh[e] = pack_bf16(h0, h1);
                ol[e] = pack_bf16(l0, l1);
              }
              if (!FUSE) {
                if (MODE == 2 && dead_row) { oh[0] = oh[1] = oh[2] = oh[3] = 0u; ol[0] = ol[1] = ol[2] = ol[3] = 0u; }
                *reinterpret_cast<uint4*>(rowp + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                *reinterpret_cast<uint4*>(rowp + 16384 + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
              }
            }
          }
          // every thread of the group is done with the staging tile (reads, and the in-place rewrite when it is stored)
          if (!FUSE) ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1, 128);
          if (et == 0) {
            if (!FUSE) {
              if (FMT == 1) {
                act_store(&tm_y, staging, 128 * h, ts, b);
                act_store(&tm_y8, staging + 16384, 256 + 64 * h, ts, b);
                act_store(&tm_y8, staging + 24576, 384 + 64 * h, ts, b);
              } else {
                act_store(&tm_y, staging, (2 * h) * 64, ts, b);
                act_store(&tm_y, staging + 16384, (2 * h + 1) * 64, ts, b);
              }
              ptx::tma_store_commit();
              ptx::tma_store_wait_read0();       // the staging tile may be overwritten once the store has read it
            }
            if (h == 0) request_residual(1, ts, b);
            else request_next(q, sub + 1);
          }
        }
        if (FUSE) {
          const int t = tile_row<MODE>(ts, rl, a.dilation);
          if (t < a.T) {
            // clamp(output(x), -1, 1)   architectures.py:143-145
            a.out[((size_t)b * a.n_out + 0) * a.T + t] = fminf(fmaxf(o0 + __ldg(a.out_b), -1.f), 1.f);
            if (a.n_out > 1) a.out[((size_t)b * a.n_out + 1) * a.T + t] = fminf(fmaxf(o1 + __ldg(a.out_b + 1), -1.f), 1.f);
          }
        }
      }
      ptx::tc_fence_before();
      if (CG == 2) ptx::mbar_arrive_cluster(tmem_empty_lead[buf]);
      else ptx::mbar_arrive(&bars->tmem_empty[buf]);
      ++it;
    }
    if (et == 0) ptx::tma_store_wait_all();
    if (FMT == 1 && a.range_flag != nullptr) {
      // |x| > 448 saturates the e4m3 planes (x itself and (x - fp16 x) 2^11 <= |x|): the next block would then run at single-pass
      // fp16 accuracy.  Report it (one atomic per warp, only when it happened) so the caller can repeat the forward in bf16 x 3.
      float vmax = fmaxf(__low2float(vmax2), __high2float(vmax2));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (lane == 0 && vmax > MST_TCN_F16F8_RANGE) atomicMax(a.range_flag, __float_as_uint(vmax));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();       // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
static int encode_act_map(CUtensorMap* m, const void* base, int B, int T, int box_ch) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  // dims (fastest first): 256 bf16 per time row (4 planes x 64), T rows, B segments; box = box_ch channels x 128 rows,
  // swizzle span = the box's row bytes (64 ch -> SWIZZLE_128B, 32 ch -> SWIZZLE_64B)
  cuuint64_t dims[3] = {256, (cuuint64_t)T, (cuuint64_t)B};      // rows >= T: zero on load, clipped on store
  cuuint64_t strides[2] = {(cuuint64_t)kRowBytes, (cuuint64_t)tcn_seg_rows(T) * kRowBytes};
  cuuint32_t box[3] = {(cuuint32_t)box_ch, (cuuint32_t)kSubRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation B=%d T=%d box=%d) failed: CUresult %d", B, T, box_ch, (int)r);
  return 0;
}

static int encode_w_map(CUtensorMap* m, const void* base, int kch) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)kch, (cuuint64_t)(kWBytesPerLayer / (kch * 2))};
  cuuint64_t strides[1] = {(cuuint64_t)kch * 2};
  cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)kCh};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
  return 0;
}

// byte-typed maps of the f16f8 format (FMT 1): activation rows of 512 bytes, box = box_bytes x 128 rows (128 -> SWIZZLE_128B operand /
// fp16 epilogue tiles, 64 -> SWIZZLE_64B e4m3 half planes); weights: rows of 128 bytes, 4 tiles of 128 rows per tap
static int encode_act_map_bytes(CUtensorMap* m, const void* base, int B, int T, int box_bytes) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t dims[3] = {(cuuint64_t)kRowBytes, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)kRowBytes, (cuuint64_t)tcn_seg_rows(T) * kRowBytes};
  cuuint32_t box[3] = {(cuuint32_t)box_bytes, (cuuint32_t)kSubRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f16f8 activation B=%d T=%d box=%d) failed: CUresult %d", B, T, box_bytes, (int)r);
  return 0;
}
static int encode_w_map_bytes(CUtensorMap* m, const void* base, int box_rows = kCh) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t dims[2] = {128, (cuuint64_t)kTaps * 4 * kCh};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};      // 64 rows: the half of a weight tile one CTA of a pair stages
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f16f8 weights) failed: CUresult %d", (int)r);
  return 0;
}

// mode-2 activation maps: the time axis viewed as [block of 2d rows][half][d rows], box = {box columns, d rows, 1 half, 128/d blocks}
// = the 128 rows of an interleaved sub-tile (tcn_tile<2>).  The segment stride tcn_seg_rows(T) is a multiple of 2d.  `bytes` = byte-typed map (f16f8) or bf16 elements.
static int encode_act_map5(CUtensorMap* m, const void* base, int B, int T, int d, int box_cols, bool bytes) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  const int Ts = tcn_seg_rows(T);        // the view runs over the padded segment: the pad rows are in bounds and hold zeros
  cuuint64_t dims[5] = {(cuuint64_t)(bytes ? kRowBytes : 256), (cuuint64_t)d, 2, (cuuint64_t)(Ts / (2 * d)), (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)kRowBytes, (cuuint64_t)d * kRowBytes, (cuuint64_t)2 * d * kRowBytes, (cuuint64_t)Ts * kRowBytes};
  cuuint32_t box[5] = {(cuuint32_t)box_cols, (cuuint32_t)d, 1, (cuuint32_t)(kSubRows / d), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const int row_bytes = bytes ? box_cols : box_cols * 2;
  CUresult r = enc(m, bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(interleaved activation B=%d T=%d d=%d box=%d) failed: CUresult %d", B, T, d, box_cols, (int)r);
  return 0;
}

// rows [T, tcn_seg_rows(T)) of every segment of an activation buffer are zero padding for the interleaved tensor maps: cleared
// before every launch that writes the buffer (the 3-D stores clip at T, the interleaved ones write zeros there themselves)
static int zero_pad_rows(uint8_t* act, int B, int T, cudaStream_t st) {
  const int Ts = tcn_seg_rows(T);
  if (Ts == T) return 0;
  MST_CUDA_OK(cudaMemset2DAsync(act + (size_t)T * kRowBytes, (size_t)Ts * kRowBytes, 0, (size_t)(Ts - T) * kRowBytes, (size_t)B, st));
  return 0;
}

// one dilated block (n >= 1): act_in -> act_out, or -> fp32 `out` when fuse_out
static int launch_umma_block(const mst_tcn_config* cfg, const uint8_t* packed, const TcnPacked& L, int n,
                             const uint8_t* act_in, uint8_t* act_out, const float* film, int n_cond, int B, int T,
                             bool fuse_out, float* out, int precision, unsigned int* range_flag, cudaStream_t st) {
  const long long d = block_dilation(cfg, n);
  MST_CHECK(7 * d + kTileRows < (1ll << 31) - T, "tcn: dilation %lld too large", d);
  const bool f8 = precision == MST_TCN_F16F8;
  // work-item geometry (tcn_tile): far pairing for dilations that are a multiple of 128, interleaved pairing for the small
  // ones when the length allows it, plain 256-row tiles otherwise
  const int mode = (d >= kSubRows && d % kSubRows == 0) ? 1 : ((d < kSubRows && kSubRows % d == 0) ? 2 : 0);
  // CTA pairs for the f16f8 format in the paired geometries (plain 256-row tiles need three slots per step: the half-filled
  // weight slot then costs ring depth, measured 8 % slower at L = 82,412)
  const int cg = (f8 && MST_TCN_CG == 2 && mode != 0 && sm_count() % 2 == 0) ? 2 : 1;
  CUtensorMap tm_x, tm_w, tm_xs, tm_y, tm_l8, tm_y8;
  const uint8_t* w_layer = packed + (f8 ? L.wumma : L.wbf16) + (size_t)(n - 1) * kWBytesPerLayer;
  const void* dst_act = fuse_out ? (const void*)act_in : (const void*)act_out;
  if (mode == 2) {
    if (encode_act_map5(&tm_x, act_in, B, T, (int)d, f8 ? 128 : 64, f8)) return 1;
    tm_xs = tm_x;
    if (encode_act_map5(&tm_y, dst_act, B, T, (int)d, f8 ? 128 : 64, f8)) return 1;
    if (f8) {
      if (encode_act_map5(&tm_l8, act_in, B, T, (int)d, 64, true)) return 1;
      if (encode_act_map5(&tm_y8, dst_act, B, T, (int)d, 64, true)) return 1;
      if (encode_w_map_bytes(&tm_w, w_layer, cg == 2 ? 64 : kCh)) return 1;
    } else {
      tm_l8 = tm_xs;
      tm_y8 = tm_y;
      if (encode_w_map(&tm_w, w_layer, 64)) return 1;
    }
  } else if (f8) {
    if (encode_act_map_bytes(&tm_x, act_in, B, T, 128)) return 1;
    tm_xs = tm_x;
    if (encode_act_map_bytes(&tm_y, fuse_out ? act_in : act_out, B, T, 128)) return 1;
    if (encode_act_map_bytes(&tm_l8, act_in, B, T, 64)) return 1;
    if (encode_act_map_bytes(&tm_y8, fuse_out ? act_in : act_out, B, T, 64)) return 1;
    if (encode_w_map_bytes(&tm_w, w_layer, cg == 2 ? 64 : kCh)) return 1;
  } else {
    if (encode_act_map(&tm_x, act_in, B, T, 64)) return 1;
    tm_xs = tm_x;
    if (encode_act_map(&tm_y, fuse_out ? act_in : act_out, B, T, 64)) return 1;
    if (encode_w_map(&tm_w, w_layer, 64)) return 1;
    tm_l8 = tm_xs;      // unused in this format
    tm_y8 = tm_y;
  }
  TcnLayerArgs a;
  a.inv_scale = reinterpret_cast<const float*>(packed + L.f8_scale) + 2 * n;
  a.range_flag = f8 ? range_flag : nullptr;
  a.B = B; a.T = T; a.dilation = (int)d;
  a.pair_m = mode == 1 ? (int)(d / kSubRows) : 1;
  a.pad_rows = tcn_seg_rows(T) != T ? 1 : 0;
  if (!fuse_out && zero_pad_rows(act_out, B, T, st)) return 1;
  a.tiles_per_seg = mode == 1 ? (int)(((T + 2 * d - 1) / (2 * d)) * a.pair_m) : cdiv(T, kTileRows);
  a.n_tiles = B * a.tiles_per_seg;
  a.n_cond = n_cond;
  a.film = reinterpret_cast<const float4*>(film) + (size_t)n * n_cond * kCh;
  a.n_out = cfg->n_outputs;
  a.out_w = reinterpret_cast<const float*>(packed + L.out_w);
  a.out_b = reinterpret_cast<const float*>(packed + L.out_b);
  a.out = out;
  int grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  if (cg == 2) {
    const int pairs = (a.n_tiles + 1) / 2;
    grid = 2 * (pairs < sm_count() / 2 ? pairs : sm_count() / 2);
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(256);
  lc.dynamicSmemBytes = kTcnSmemBytes;
  lc.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = 2;
  lattr[0].val.clusterDim.y = 1;
  lattr[0].val.clusterDim.z = 1;
  lc.attrs = lattr;
  lc.numAttrs = 1;
#define MST_TCN_LAUNCH(MODE_, FMT_, FUSE_)                                                                                        \
  do {                                                                                                                           \
    if (FMT_ == 1 && cg == 2) {                                                                                                  \
      auto kern = tcn_block_umma_kernel<64, MODE_, 1, FUSE_, 2>;                                                                 \
      MST_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcnSmemBytes));                  \
      MST_CUDA_OK(cudaLaunchKernelEx(&lc, kern, tm_x, tm_w, tm_xs, tm_y, tm_l8, tm_y8, a));                                      \
    } else {                                                                                                                     \
      MST_CUDA_OK(cudaFuncSetAttribute(tcn_block_umma_kernel<64, MODE_, FMT_, FUSE_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)kTcnSmemBytes));                                                                     \
      tcn_block_umma_kernel<64, MODE_, FMT_, FUSE_, 1><<<grid, 256, kTcnSmemBytes, st>>>(tm_x, tm_w, tm_xs, tm_y, tm_l8, tm_y8, a); \
    }                                                                                                                            \
  } while (0)
#define MST_TCN_LAUNCH2(MODE_, FMT_) do { if (fuse_out) MST_TCN_LAUNCH(MODE_, FMT_, true); else MST_TCN_LAUNCH(MODE_, FMT_, false); } while (0)
#define MST_TCN_LAUNCH3(FMT_) do { if (mode == 1) MST_TCN_LAUNCH2(1, FMT_); else if (mode == 2) MST_TCN_LAUNCH2(2, FMT_); else MST_TCN_LAUNCH2(0, FMT_); } while (0)
  if (f8) MST_TCN_LAUNCH3(1);
  else MST_TCN_LAUNCH3(0);
#undef MST_TCN_LAUNCH3
#undef MST_TCN_LAUNCH2
#undef MST_TCN_LAUNCH2
#undef MST_TCN_LAUNCH
  return launch_ok("tcn_block_umma_kernel");
}

}  // namespace mst
#include "tcn_b0.cuh"
namespace mst {

static int launch_block0(const mst_tcn_config* cfg, const uint8_t* packed, const TcnPacked& L, const float* x,
                         const float* film, int n_cond, uint8_t* act, int B, int T, int precision,
                         unsigned int* range_flag, cudaStream_t st) {
  dim3 grid(cdiv(T, 256), B);
  const float* w0 = reinterpret_cast<const float*>(packed + L.w0);
  if (zero_pad_rows(act, B, T, st)) return 1;
  if (precision == MST_TCN_F16F8) {
    // block 0 on the tensor cores (tcn_b0.cuh): persistent, one CTA per SM, 128-row tiles
    CUtensorMap tm_y, tm_y8;
    if (encode_act_map_bytes(&tm_y, act, B, T, 128)) return 1;
    if (encode_act_map_bytes(&tm_y8, act, B, T, 64)) return 1;
    b0::Args a;
    a.x = x; a.w0 = w0; a.film = reinterpret_cast<const float4*>(film); a.range_flag = range_flag;
    a.n_cond = n_cond; a.B = B; a.T = T; a.nin = cfg->n_inputs;
    a.tiles_per_seg = cdiv(T, b0::kRows);
    a.n_tiles = B * a.tiles_per_seg;
    MST_CUDA_OK(cudaFuncSetAttribute(b0::block0_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b0::kSmemBytes));
    const int g = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
    b0::block0_umma_kernel<<<g, b0::kThreads, b0::kSmemBytes, st>>>(tm_y, tm_y8, a);
    return launch_ok("block0_umma_kernel");
  }
  const float4* f = reinterpret_cast<const float4*>(film);
  if (cfg->n_inputs == 2) tcn_block0_kernel<2><<<grid, 256, 0, st>>>(x, w0, f, n_cond, act, T, tcn_seg_rows(T));
  else tcn_block0_kernel<1><<<grid, 256, 0, st>>>(x, w0, f, n_cond, act, T, tcn_seg_rows(T));
  return launch_ok("tcn_block0_kernel");
}

static size_t act_bytes(int B, int L) { return align_up((size_t)B * tcn_seg_rows(L) * kRowBytes, 1024); }

static int check_precision(int precision) {
  MST_CHECK(precision == MST_TCN_F16F8 || precision == MST_TCN_BF16X3, "tcn: precision must be MST_TCN_F16F8 (0) or MST_TCN_BF16X3 (1), got %d",
            precision);
  return 0;
}

}  // namespace mst

using namespace mst;

extern "C" {

size_t mst_tcn_packed_bytes(const mst_tcn_config* cfg) {
  TcnPacked L;
  if (tcn_layout(cfg, &L)) return 0;
  return L.total;
}

int mst_tcn_pack(const mst_tcn_config* cfg, const void* const* raw, void* packed_v, void* stream) {
  TcnPacked L;
  if (tcn_layout(cfg, &L)) return 1;
  MST_CHECK(raw && packed_v, "tcn_pack: null pointer");
  MST_CHECK((reinterpret_cast<uintptr_t>(packed_v) & 1023) == 0, "tcn_pack: packed buffer must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* packed = reinterpret_cast<uint8_t*>(packed_v);
  for (int n = 0; n < cfg->n_blocks; ++n) {
    const float* conv_w = (const float*)raw[8 * n + 0];
    const float* bn_w = (const float*)raw[8 * n + 1];
    const float* bn_b = (const float*)raw[8 * n + 2];
    const float* bn_m = (const float*)raw[8 * n + 3];
    const float* bn_v = (const float*)raw[8 * n + 4];
    const float* res_w = (const float*)raw[8 * n + 5];
    const float* film_w = (const float*)raw[8 * n + 6];
    const float* film_b = (const float*)raw[8 * n + 7];
    MST_CHECK(conv_w && bn_w && bn_b && bn_m && bn_v && res_w && film_w && film_b, "tcn_pack: null weight in block %d", n);
    if (n == 0) {
      tcn_pack_block0_kernel<<<16, 256, 0, st>>>(conv_w, bn_w, bn_v, cfg->n_inputs, (float*)(packed + L.w0));
    } else {
      float* sc = reinterpret_cast<float*>(packed + L.f8_scale) + 2 * n;
      if (tcn_f8_pack_layer(conv_w, bn_w, bn_v, packed + L.wumma + (size_t)(n - 1) * kWBytesPerLayer, sc,
                            reinterpret_cast<unsigned int*>(sc + 1), st)) return 1;
      tcn_pack_umma_kernel<<<256, 256, 0, st>>>(
          conv_w, bn_w, bn_v, (__nv_bfloat16*)(packed + L.wbf16 + (size_t)(n - 1) * kWBytesPerLayer), 64);
    }
    tcn_pack_vec_kernel<<<1, 128, 0, st>>>(bn_w, bn_b, bn_m, bn_v, res_w, (float*)(packed + L.bn_bias) + n * kCh,
                                           (float*)(packed + L.res) + n * kCh);
    MST_CUDA_OK(cudaMemcpyAsync(packed + L.film_w + (size_t)n * 2 * kCh * cfg->cond_dim * 4, film_w,
                                (size_t)2 * kCh * cfg->cond_dim * 4, cudaMemcpyDeviceToDevice, st));
    MST_CUDA_OK(cudaMemcpyAsync(packed + L.film_b + (size_t)n * 2 * kCh * 4, film_b, (size_t)2 * kCh * 4,
                                cudaMemcpyDeviceToDevice, st));
  }
  const void* out_w = raw[8 * cfg->n_blocks], *out_b = raw[8 * cfg->n_blocks + 1];
  MST_CHECK(out_w && out_b, "tcn_pack: null output weight");
  MST_CUDA_OK(cudaMemcpyAsync(packed + L.out_w, out_w, (size_t)cfg->n_outputs * kCh * 4, cudaMemcpyDeviceToDevice, st));
  MST_CUDA_OK(cudaMemcpyAsync(packed + L.out_b, out_b, (size_t)cfg->n_outputs * 4, cudaMemcpyDeviceToDevice, st));
  return launch_ok("tcn_pack kernels");
}

int mst_tcn_film_precompute(const mst_tcn_config* cfg, const void* packed_v, const float* cond, int n_cond,
                            float* film_out, void* stream) {
  TcnPacked L;
  if (tcn_layout(cfg, &L)) return 1;
  MST_CHECK(packed_v && cond && film_out && n_cond >= 1, "tcn_film_precompute: bad arguments");
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(packed_v);
  const int warps = cfg->n_blocks * 2 * kCh;
  tcn_film_kernel<128><<<cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(
      (const float*)(packed + L.film_w), (const float*)(packed + L.film_b), (const float*)(packed + L.bn_bias),
      (const float*)(packed + L.res), cond, cfg->n_blocks, n_cond, cfg->cond_dim, film_out);
  return launch_ok("tcn_film_kernel");
}

size_t mst_tcn_workspace_bytes(const mst_tcn_config* cfg, int B, int L) {
  if (!cfg || B <= 0 || L <= 0) return 0;
  return 2 * act_bytes(B, L);
}

int mst_tcn_forward(const mst_tcn_config* cfg, const void* packed_v, const float* x, const float* film, int n_cond,
                    float* y, int B, int L, void* workspace, size_t workspace_bytes, int precision,
                    unsigned int* range_flag, void* stream) {
  TcnPacked P;
  if (tcn_layout(cfg, &P) || check_precision(precision)) return 1;
  MST_CHECK(packed_v && x && film && y && workspace, "tcn_forward: null pointer");
  MST_CHECK(B > 0 && L > 0 && B <= 65535, "tcn_forward: bad shape B=%d L=%d", B, L);
  MST_CHECK(n_cond == 1 || n_cond == B, "tcn_forward: n_cond must be 1 or B (got %d, B=%d)", n_cond, B);
  MST_CHECK(workspace_bytes >= mst_tcn_workspace_bytes(cfg, B, L), "tcn_forward: workspace too small (%zu < %zu)",
            workspace_bytes, mst_tcn_workspace_bytes(cfg, B, L));
  MST_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "tcn_forward: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(packed_v);
  uint8_t* act[2] = {(uint8_t*)workspace, (uint8_t*)workspace + act_bytes(B, L)};
  if (range_flag) MST_CUDA_OK(cudaMemsetAsync(range_flag, 0, sizeof(unsigned int), st));
  if (launch_block0(cfg, packed, P, x, film, n_cond, act[0], B, L, precision, range_flag, st)) return 1;
  int cur = 0;
  for (int n = 1; n < cfg->n_blocks; ++n) {
    const bool last = n == cfg->n_blocks - 1;
    if (launch_umma_block(cfg, packed, P, n, act[cur], act[cur ^ 1], film, n_cond, B, L, last, y, precision, range_flag, st))
      return 1;
    cur ^= 1;
  }
  return 0;
}

int mst_tcn_block0_forward(const mst_tcn_config* cfg, const void* packed_v, const float* x, const float* film, int n_cond,
                           void* act_out, int B, int L, int precision, unsigned int* range_flag, void* stream) {
  TcnPacked P;
  if (tcn_layout(cfg, &P) || check_precision(precision)) return 1;
  MST_CHECK(packed_v && x && film && act_out, "tcn_block0_forward: null pointer");
  MST_CHECK(B > 0 && L > 0 && B <= 65535 && (n_cond == 1 || n_cond == B), "tcn_block0_forward: bad shape");
  return launch_block0(cfg, reinterpret_cast<const uint8_t*>(packed_v), P, x, film, n_cond, (uint8_t*)act_out, B, L,
                       precision, range_flag, (cudaStream_t)stream);
}

int mst_tcn_layer_forward(const mst_tcn_config* cfg, const void* packed_v, int block, const void* act_in, void* act_out,
                          const float* film, int n_cond, int B, int L, int fuse_out, float* y, int precision,
                          unsigned int* range_flag, void* stream) {
  TcnPacked P;
  if (tcn_layout(cfg, &P) || check_precision(precision)) return 1;
  MST_CHECK(packed_v && act_in && film, "tcn_layer_forward: null pointer");
  MST_CHECK(block >= 1 && block < cfg->n_blocks, "tcn_layer_forward: block %d out of range [1,%d)", block, cfg->n_blocks);
  MST_CHECK(B > 0 && L > 0 && (n_cond == 1 || n_cond == B), "tcn_layer_forward: bad shape");
  MST_CHECK(fuse_out ? (y != nullptr) : (act_out != nullptr), "tcn_layer_forward: missing output buffer");
  MST_CHECK((reinterpret_cast<uintptr_t>(act_in) & 1023) == 0 && (reinterpret_cast<uintptr_t>(act_out) & 1023) == 0,
            "tcn_layer_forward: activation buffers must be 1024-byte aligned");
  return launch_umma_block(cfg, reinterpret_cast<const uint8_t*>(packed_v), P, block, (const uint8_t*)act_in,
                           (uint8_t*)act_out, film, n_cond, B, L, fuse_out != 0, y, precision, range_flag,
                           (cudaStream_t)stream);
}

int mst_tcn_block_forward(const mst_tcn_config* cfg, const void* packed_v, int block, const float* x, const float* film,
                          int n_cond, float* y, int B, int L, void* workspace, size_t workspace_bytes, int precision,
                          void* stream) {
  TcnPacked P;
  if (tcn_layout(cfg, &P) || check_precision(precision)) return 1;
  MST_CHECK(packed_v && x && film && y && workspace, "tcn_block_forward: null pointer");
  MST_CHECK(block >= 0 && block < cfg->n_blocks, "tcn_block_forward: block %d out of range", block);
  MST_CHECK(B > 0 && L > 0 && B <= 65535, "tcn_block_forward: bad shape B=%d L=%d", B, L);
  MST_CHECK(n_cond == 1 || n_cond == B, "tcn_block_forward: n_cond must be 1 or B");
  MST_CHECK(workspace_bytes >= mst_tcn_workspace_bytes(cfg, B, L), "tcn_block_forward: workspace too small");
  MST_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "tcn_block_forward: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* packed = reinterpret_cast<const uint8_t*>(packed_v);
  uint8_t* act[2] = {(uint8_t*)workspace, (uint8_t*)workspace + act_bytes(B, L)};
  const bool f8 = precision == MST_TCN_F16F8;
  dim3 grid(cdiv(L, 32), B);
  if (block == 0) {
    // film for block 0 sits at the start of the table
    if (launch_block0(cfg, packed, P, x, film, n_cond, act[1], B, L, precision, nullptr, st)) return 1;
  } else {
    if (zero_pad_rows(act[0], B, L, st)) return 1;
    if (f8) {
      if (tcn_f8_act_pack(x, act[0], B, L, st)) return 1;
    } else {
      tcn_act_pack_kernel<<<grid, 256, 0, st>>>(x, act[0], L, tcn_seg_rows(L));
      if (launch_ok("tcn_act_pack_kernel")) return 1;
    }
    if (launch_umma_block(cfg, packed, P, block, act[0], act[1], film, n_cond, B, L, false, nullptr, precision, nullptr, st))
      return 1;
  }
  if (f8) return tcn_f8_act_unpack(act[1], y, B, L, st);
  tcn_act_unpack_kernel<<<grid, 256, 0, st>>>(act[1], y, L, tcn_seg_rows(L));
  return launch_ok("tcn_act_unpack_kernel");
}

}  // extern "C"
