// enc_umma.cu -- FXencoder blocks with >= 64 input channels on the tcgen05 tensor cores.
//
// Same reference semantics as encoder.cu (Conv1d_layer: ReflectionPad1d -> Conv1d(bias) -> BatchNorm1d(eval) -> ReLU,
// Res_ConvBlock: conv1(x) + x -> conv2; mst/networks/network_utils.py:28-34,47-51,74,79-89,116-119), different engine:
// an im2col-free implicit GEMM   D[(b,t), co] = sum_{tap, ci} X[b, s*t + tap - l, ci] * W[tap][co][ci]
// with the same split-bf16 3-product scheme and warp-specialised TMA / tcgen05 / TMEM pipeline as csrc/tcn.cu.
// Accuracy note (measured, round 1): with the whole K loop chained in one TMEM accumulator the embedding came out 1.4e-4
// relative off the fp32 CPU forward -- and identically so with FP16 operand pairs (22 mantissa bits) instead of BF16 pairs
// (16 bits): the error is not operand rounding but the tensor core's FP32 accumulation, which truncates when aligning
// addends (a bias of ~2^-24 of the running sum per tcgen05.mma, over 45 ... 480 chained MMAs per output at K up to
// 10240, amplified ~10x by the following conv -> BN -> ReLU layers).  The kernel therefore accumulates at most
// kFlushSteps K-steps (24 MMAs) in TMEM and adds these partial sums in fp32 registers (round-to-nearest) in the epilogue
// warps, double-buffered against the next group.  BF16 pairs are kept (no range hazard).
//
// Activation format ("split rows"): channels-last, one time step = C*4 bytes = C/64 groups of [64 hi bf16 | 64 lo bf16],
// x = hi + lo.  Every buffer carries the reflection halo of its CONSUMER: rows [0, l) and [l+T, l+T+r) hold the mirrored
// samples (nn.ReflectionPad1d), written by the producing kernel's epilogue, so the consumer's taps are plain shifted TMA
// boxes:  padded row = s*t + tap.  Stride-2 convolutions use the tensor map's element stride along time.
// M tile = 128 rows = NB segments x TT time steps (TT = 128 ... 8), so late layers (T = 64) still fill the MMA;
// N tile = 128 output channels (64 for the one 64-channel layer).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace mst {

constexpr int kEncSlotBytes = 32768;
constexpr int kEncSlots = 6;
constexpr int kFlushSteps = 2;   // K64-steps (12 MMAs each) accumulated in TMEM before the fp32 register flush
constexpr size_t kEncSmemBytes = 1024 + (size_t)kEncSlots * kEncSlotBytes + kEncSlotBytes + 256;

struct ConvUmmaArgs {
  int B, T_out, TT, NB, ntt, n_sub, n_mtiles, n_ntiles, n_items;
  int taps, kchunks, stride, c_out, nt;
  int res_row_off;          // padded input row of unpadded t = 0 for the residual read (conv1: its own l); -1 = none
  int out_halo_l, out_halo_r;
  const float* bias;        // [c_out], BN folded
  uint8_t* out;             // output buffer (for the mirrored halo rows)
  long long out_seg_bytes;  // (T_out + halo_l + halo_r) * c_out * 4
  int out_row_bytes;        // c_out * 4
};

struct __align__(8) EncBarriers {
  uint64_t full[kEncSlots], empty[kEncSlots];
  uint64_t tmem_full[2], tmem_empty[2];
  uint64_t stage_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void enc_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t enc_pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ float enc_lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float enc_hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// tm_x: operand boxes {64 ch, TT (x stride), NB} on the padded input;  tm_w: {64 ci, nt co};
// tm_r: residual boxes {64, TT, NB} on the padded input (stride 1);  tm_y: store boxes {64, TT, NB} on the output,
// whose time extent is clipped to halo_l + T_out so partial tiles never spill into the right halo.
__global__ void __launch_bounds__(256, 1)
enc_conv_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_y,
                     const ConvUmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET into the shared array (not integer pointer arithmetic): the compiler keeps the shared
  // state space and emits LDS / STS instead of generic LD / ST for the epilogue's staging accesses
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* staging = smem + (size_t)kEncSlots * kEncSlotBytes;
  EncBarriers* bars = reinterpret_cast<EncBarriers*>(staging + kEncSlotBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_w);
    ptx::prefetch_tensormap(&tm_r);
    ptx::prefetch_tensormap(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kEncSlots; ++i) {
      ptx::mbar_init(&bars->full[i], 1);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->tmem_full[i], 1);
      ptx::mbar_init(&bars->tmem_empty[i], 128);
    }
    ptx::mbar_init(&bars->stage_full, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&bars->tmem_base, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t w_bytes = 2u * (uint32_t)a.nt * 128u;

  // Work item = (n tile, 128-row sub-tile); consecutive CTAs share a weight tile (the big operand of the late layers).
  // The K loop (taps x 64-channel chunks) is cut into GROUPS of kFlushSteps steps (= 12*kFlushSteps chained MMAs): each
  // group accumulates from zero into one of two TMEM buffers and is then added into fp32 REGISTERS by the epilogue
  // warps while the next group runs.  This bounds the tensor core's truncating accumulation chain (see file header).
  const int n_steps = a.taps * a.kchunks;
  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      auto next = [&]() { if (++slot == kEncSlots) { slot = 0; phase ^= 1; } };
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int ntile = item / a.n_sub, su = item - ntile * a.n_sub;
        const int n0 = ntile * a.nt;
        const int bg = su / a.ntt, t0 = (su - bg * a.ntt) * a.TT;
        for (int j = 0; j < a.taps; ++j) {
          for (int kc = 0; kc < a.kchunks; ++kc) {
            ptx::mbar_wait(&bars->empty[slot], phase ^ 1);
            ptx::mbar_expect_tx(&bars->full[slot], w_bytes);
            uint8_t* dst = ring + (size_t)slot * kEncSlotBytes;
            const int wrow = ((j * a.kchunks + kc) * 2) * a.c_out + n0;
            ptx::tma_load_2d(&tm_w, &bars->full[slot], dst, 0, wrow);
            ptx::tma_load_2d(&tm_w, &bars->full[slot], dst + 16384, 0, wrow + a.c_out);
            next();
            ptx::mbar_wait(&bars->empty[slot], phase ^ 1);
            ptx::mbar_expect_tx(&bars->full[slot], kEncSlotBytes);
            dst = ring + (size_t)slot * kEncSlotBytes;
            ptx::tma_load_3d(&tm_x, &bars->full[slot], dst, kc * 128, a.stride * t0 + j, bg * a.NB);
            ptx::tma_load_3d(&tm_x, &bars->full[slot], dst + 16384, kc * 128 + 64, a.stride * t0 + j, bg * a.NB);
            next();
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-convergent MMA issue (see umma_mma_f16kind_elect in sm100_ptx.cuh): all lanes run the uniform loop, the tcgen05
    // instructions are predicated on the elected lane, so ptxas emits straight UTCHMMA sequences
    {
      const uint32_t leader = ptx::elect_one() ? 1u : 0u;
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, a.nt);
      uint32_t slot = 0, phase = 0;
      auto next = [&]() { if (++slot == kEncSlots) { slot = 0; phase ^= 1; } };
      auto issue_group = [&](uint32_t x_addr, uint32_t w_addr, uint32_t d_tmem, bool first) {
        const uint64_t xh = ptx::umma_desc_kmajor<128>(x_addr), xl = ptx::umma_desc_kmajor<128>(x_addr + 16384);
        const uint64_t wh = ptx::umma_desc_kmajor<128>(w_addr), wl = ptx::umma_desc_kmajor<128>(w_addr + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);
          ptx::umma_mma_f16kind_elect(d_tmem, xh + adv, wh + adv, idesc, (first && k == 0) ? 0u : 1u, leader);
          ptx::umma_mma_f16kind_elect(d_tmem, xl + adv, wh + adv, idesc, 1u, leader);
          ptx::umma_mma_f16kind_elect(d_tmem, xh + adv, wl + adv, idesc, 1u, leader);
        }
      };
      uint32_t gtotal = 0;   // accumulation groups issued so far (buffer = gtotal & 1)
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        for (int s = 0; s < n_steps; ++s) {
          const uint32_t buf = gtotal & 1u;
          const bool group_start = (s % kFlushSteps) == 0;
          if (group_start) {
            ptx::mbar_wait(&bars->tmem_empty[buf], ((gtotal >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
          }
          const uint32_t wslot = slot;
          ptx::mbar_wait(&bars->full[wslot], phase);
          const uint32_t w_addr = ptx::smem_u32(ring + (size_t)wslot * kEncSlotBytes);
          next();
          ptx::mbar_wait(&bars->full[slot], phase);
          ptx::tc_fence_after();
          issue_group(ptx::smem_u32(ring + (size_t)slot * kEncSlotBytes), w_addr, tmem_base + buf * 128u, group_start);
          ptx::umma_commit_elect(&bars->empty[slot], leader);
          next();
          ptx::umma_commit_elect(&bars->empty[wslot], leader);
          if ((s % kFlushSteps) == kFlushSteps - 1 || s == n_steps - 1) {
            ptx::umma_commit_elect(&bars->tmem_full[buf], leader);
            ++gtotal;
          }
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int et = threadIdx.x - 128;
    const int rl = q * 32 + lane;              // row of the sub-tile == TMEM lane
    const int nb = rl / a.TT, tt = rl - nb * a.TT;
    uint32_t stage_phase = 0;
    uint32_t gtotal = 0;
    const int halves = a.nt / 64;
    const int n_groups = (n_steps + kFlushSteps - 1) / kFlushSteps;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int ntile = item / a.n_sub, su = item - ntile * a.n_sub;
      const int n0 = ntile * a.nt;
      const int bg = su / a.ntt, t0 = (su - bg * a.ntt) * a.TT;
      const int b = bg * a.NB + nb, t = t0 + tt;
      const bool row_ok = b < a.B && t < a.T_out;
      // ---- fp32 register accumulation of the K groups ----
      float sum[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) sum[i] = 0.f;
      for (int g = 0; g < n_groups; ++g, ++gtotal) {
        const uint32_t buf = gtotal & 1u;
        ptx::mbar_wait(&bars->tmem_full[buf], (gtotal >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128u;
#pragma unroll
        for (int c32 = 0; c32 < 4; ++c32) {
          if (c32 * 32 < a.nt) {
            uint32_t part[32];
            ptx::tmem_ld_32x32(taddr + c32 * 32, part);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[c32 * 32 + i] += __uint_as_float(part[i]);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bars->tmem_empty[buf]);
      }
      // mirrored destination rows of this thread's time step, if it falls into the consumer's reflection halo
      // (a short segment can need BOTH: t = 2 of T = 5 with halo (2,2) is mirrored to the left and to the right)
      long long mirror_off[2] = {-1, -1};
      if (row_ok) {
        const long long seg = (long long)b * a.out_seg_bytes;
        if (t >= 1 && t <= a.out_halo_l) mirror_off[0] = seg + (long long)(a.out_halo_l - t) * a.out_row_bytes;
        if (t <= a.T_out - 2 && t >= a.T_out - 1 - a.out_halo_r)
          mirror_off[1] = seg + (long long)(a.out_halo_l + 2 * (a.T_out - 1) - t) * a.out_row_bytes;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h < halves) {
          const int cg = n0 / 64 + h;   // 64-channel group of the output
          if (et == 0) ptx::tma_store_wait_read0();
          ptx::named_bar_sync(1, 128);
          if (a.res_row_off >= 0) {
            if (et == 0) {
              ptx::mbar_expect_tx(&bars->stage_full, kEncSlotBytes);
              ptx::tma_load_3d(&tm_r, &bars->stage_full, staging, cg * 128, a.res_row_off + t0, bg * a.NB);
              ptx::tma_load_3d(&tm_r, &bars->stage_full, staging + 16384, cg * 128 + 64, a.res_row_off + t0, bg * a.NB);
            }
            ptx::mbar_wait(&bars->stage_full, stage_phase);
            stage_phase ^= 1;
          }
          uint8_t* rowp = staging + rl * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int off = ((c ^ (rl & 7)) << 4);
            uint32_t xhw[4] = {0, 0, 0, 0}, xlw[4] = {0, 0, 0, 0};
            if (a.res_row_off >= 0) {
              const uint4 xh = *reinterpret_cast<const uint4*>(rowp + off);
              const uint4 xl = *reinterpret_cast<const uint4*>(rowp + 16384 + off);
              xhw[0] = xh.x; xhw[1] = xh.y; xhw[2] = xh.z; xhw[3] = xh.w;
              xlw[0] = xl.x; xlw[1] = xl.y; xlw[2] = xl.z; xlw[3] = xl.w;
            }
            uint32_t oh[4], ol[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v[2];
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) {
                const int cl = c * 8 + e * 2 + s2;
                float u = sum[h * 64 + cl] + __ldg(a.bias + n0 + h * 64 + cl);
                u = fmaxf(u, 0.f);                                       // ReLU (network_utils.py:79-80)
                const float xin = s2 == 0 ? enc_lo_f(xhw[e]) + enc_lo_f(xlw[e]) : enc_hi_f(xhw[e]) + enc_hi_f(xlw[e]);
                v[s2] = u + xin;                                         // conv1(x) + x (network_utils.py:117); xin = 0 for conv2
              }
              __nv_bfloat16 h0, l0, h1, l1;
              enc_split(v[0], h0, l0);
              enc_split(v[1], h1, l1);
              oh[e] = enc_pack2(h0, h1);
              ol[e] = enc_pack2(l0, l1);
            }
            *reinterpret_cast<uint4*>(rowp + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
            *reinterpret_cast<uint4*>(rowp + 16384 + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
              if (mirror_off[mi] >= 0) {
                uint8_t* m = a.out + mirror_off[mi] + (size_t)cg * 256 + c * 16;
                *reinterpret_cast<uint4*>(m) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                *reinterpret_cast<uint4*>(m + 128) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
              }
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(2, 128);
          if (et == 0) {
            ptx::tma_store_3d(&tm_y, staging, cg * 128, a.out_halo_l + t0, bg * a.NB);
            ptx::tma_store_3d(&tm_y, staging + 16384, cg * 128 + 64, a.out_halo_l + t0, bg * a.NB);
            ptx::tma_store_commit();
          }
        }
      }
    }
    if (et == 0) ptx::tma_store_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight packing: w[co][ci][k] fp32 (+BN) -> out[tap][kc][split][co][64] bf16 pairs, bias folded
__global__ void enc_pack_umma_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ bn_w,
                                     const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                     const float* __restrict__ bn_var, float eps, int c_out, int c_in, int k,
                                     __nv_bfloat16* __restrict__ w_out, float* __restrict__ b_out) {
  const size_t n = (size_t)c_out * c_in * k;
  const int kchunks = c_in / 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cil = i & 63;
    size_t r = i >> 6;
    const int co = r % c_out; r /= c_out;
    const int kc = r % kchunks;
    const int tap = r / kchunks;
    const float s = bn_w[co] / sqrtf(bn_var[co] + eps);
    const float v = w[((size_t)co * c_in + kc * 64 + cil) * k + tap] * s;
    __nv_bfloat16 hi, lo;
    enc_split(v, hi, lo);
    const size_t base = ((size_t)(tap * kchunks + kc) * 2) * c_out * 64;
    w_out[base + (size_t)co * 64 + cil] = hi;
    w_out[base + (size_t)c_out * 64 + (size_t)co * 64 + cil] = lo;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < c_out; co += gridDim.x * blockDim.x) {
    const float s = bn_w[co] / sqrtf(bn_var[co] + eps);
    b_out[co] = ((b ? b[co] : 0.f) - bn_mean[co]) * s + bn_b[co];
  }
}

// fp32 [B][C][T] -> split rows with reflection halo (the hand-over from the CUDA-core blocks to the tensor-core blocks)
__global__ void __launch_bounds__(256)
enc_split_from_f32_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int C, int T, int halo_l, int halo_r) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, cg = blockIdx.y, t0 = blockIdx.x * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = warp; c < 64; c += 8) {
    const int t = t0 + lane;
    tile[c][lane] = t < T ? x[((size_t)b * C + cg * 64 + c) * T + t] : 0.f;
  }
  __syncthreads();
  const int t_pad = T + halo_l + halo_r;
  const size_t row_bytes = (size_t)C * 4;
  for (int r = warp; r < 32; r += 8) {
    const int t = t0 + r;
    if (t >= T) continue;
    __nv_bfloat16 h0, l0, h1, l1;
    enc_split(tile[2 * lane][r], h0, l0);
    enc_split(tile[2 * lane + 1][r], h1, l1);
    const uint32_t hi = enc_pack2(h0, h1), lo = enc_pack2(l0, l1);
    int rows[3] = {halo_l + t, -1, -1};
    if (t >= 1 && t <= halo_l) rows[1] = halo_l - t;
    if (t <= T - 2 && t >= T - 1 - halo_r) rows[2] = halo_l + 2 * (T - 1) - t;
    for (int m = 0; m < 3; ++m) {
      if (rows[m] < 0) continue;
      uint32_t* row = reinterpret_cast<uint32_t*>(y + ((size_t)b * t_pad + rows[m]) * row_bytes + (size_t)cg * 256);
      row[lane] = hi;
      row[32 + lane] = lo;
    }
  }
}

// AdaptiveAvgPool1d(1).squeeze(-1) on split rows (architectures.py:62,67)
__global__ void __launch_bounds__(256)
enc_pool_split_kernel(const uint8_t* __restrict__ x, float* __restrict__ emb, int B, int C, int T, int halo_l, int t_pad) {
  // one warp per (b, 64-channel group): lane owns a channel pair
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int groups = C / 64;
  const int b = gw / groups, cg = gw - b * groups;
  if (b >= B) return;
  float s0 = 0.f, s1 = 0.f;
  const size_t row_bytes = (size_t)C * 4;
  for (int t = 0; t < T; ++t) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(x + ((size_t)b * t_pad + halo_l + t) * row_bytes + (size_t)cg * 256);
    const uint32_t hi = row[lane], lo = row[32 + lane];
    s0 += enc_lo_f(hi) + enc_lo_f(lo);
    s1 += enc_hi_f(hi) + enc_hi_f(lo);
  }
  emb[(size_t)b * C + cg * 64 + 2 * lane] = s0 / (float)T;
  emb[(size_t)b * C + cg * 64 + 2 * lane + 1] = s1 / (float)T;
}

// ---------------------------------------------------------------------------------------------------------------------
static int encode_rows_map(CUtensorMap* m, const void* base, int B, int t_extent, int t_pad, int C, int TT, int NB,
                           int stride) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t dims[3] = {(cuuint64_t)2 * C, (cuuint64_t)t_extent, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)t_pad * C * 4};
  cuuint32_t box[3] = {64, (cuuint32_t)(TT * stride), (cuuint32_t)NB};
  cuuint32_t estr[3] = {1, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(enc rows C=%d T=%d TT=%d NB=%d s=%d) failed: CUresult %d", C,
            t_extent, TT, NB, stride, (int)r);
  return 0;
}

static int encode_encw_map(CUtensorMap* m, const void* base, int rows, int nt) {
  PFN_encodeTiled enc = tensor_map_encoder();
  if (!enc) return 1;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)nt};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(enc weights rows=%d) failed: CUresult %d", rows, (int)r);
  return 0;
}

size_t enc_umma_weight_bytes(int c_in, int c_out, int k) { return align_up((size_t)c_in * c_out * k * 4, 1024); }

bool enc_umma_eligible(int c_in, int c_out, int k, int stride) {
  return c_in >= 64 && c_in % 64 == 0 && c_out % 64 == 0 && (c_out == 64 || c_out % 128 == 0) && stride >= 1 &&
         stride <= 2 && k >= 1 && k <= 64;
}

int enc_umma_pack(const float* w, const float* b, const float* bn_w, const float* bn_b, const float* bn_mean,
                  const float* bn_var, int c_out, int c_in, int k, void* w_out, float* b_out, cudaStream_t st) {
  const size_t n = (size_t)c_out * c_in * k;
  const int blocks = (int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192);
  enc_pack_umma_kernel<<<blocks, 256, 0, st>>>(w, b, bn_w, bn_b, bn_mean, bn_var, 1e-5f, c_out, c_in, k,
                                               (__nv_bfloat16*)w_out, b_out);
  return launch_ok("enc_pack_umma_kernel");
}

static void pick_tile(int t_out, int* TT, int* NB) {
  int tt = 128;
  while (tt > 8 && tt / 2 >= t_out) tt /= 2;
  *TT = tt;
  *NB = 128 / tt;
}

// x: split rows [B][t_in + l + r][c_in*4 B] with this conv's own reflection halo; y: split rows with (out_l, out_r)
int enc_umma_conv(const void* x, const void* w_packed, const float* bias, void* y, int B, int c_in, int t_in, int c_out,
                  int k, int stride, bool residual, int out_l, int out_r, cudaStream_t st) {
  const int pad = k - 1, l = pad / 2, r = pad - l;
  MST_CHECK(t_in > r, "enc_conv(umma): reflection padding (%d,%d) needs T_in > pad, got T_in=%d", l, r, t_in);
  const int t_out = (t_in + pad - k) / stride + 1;
  MST_CHECK(!residual || (stride == 1 && c_in == c_out), "enc_conv(umma): residual needs stride 1 and c_in == c_out");
  MST_CHECK(t_out > out_r, "enc_conv(umma): output length %d too short for the next layer's reflection halo", t_out);
  ConvUmmaArgs a;
  a.B = B; a.T_out = t_out;
  pick_tile(t_out, &a.TT, &a.NB);
  a.ntt = cdiv(t_out, a.TT);
  a.n_sub = cdiv(B, a.NB) * a.ntt;
  a.n_mtiles = a.n_sub;
  a.nt = c_out >= 128 ? 128 : 64;
  a.n_ntiles = c_out / a.nt;
  a.n_items = a.n_sub * a.n_ntiles;
  a.taps = k; a.kchunks = c_in / 64; a.stride = stride; a.c_out = c_out;
  a.res_row_off = residual ? l : -1;
  a.out_halo_l = out_l; a.out_halo_r = out_r;
  a.bias = bias;
  a.out = (uint8_t*)y;
  const int t_pad_in = t_in + l + r, t_pad_out = t_out + out_l + out_r;
  a.out_seg_bytes = (long long)t_pad_out * c_out * 4;
  a.out_row_bytes = c_out * 4;
  CUtensorMap tm_x, tm_w, tm_r, tm_y;
  if (encode_rows_map(&tm_x, x, B, t_pad_in, t_pad_in, c_in, a.TT, a.NB, stride)) return 1;
  if (encode_rows_map(&tm_r, x, B, t_pad_in, t_pad_in, c_in, a.TT, a.NB, 1)) return 1;
  if (encode_rows_map(&tm_y, y, B, out_l + t_out, t_pad_out, c_out, a.TT, a.NB, 1)) return 1;
  if (encode_encw_map(&tm_w, w_packed, k * a.kchunks * 2 * c_out, a.nt)) return 1;
  MST_CUDA_OK(cudaFuncSetAttribute(enc_conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSmemBytes));
  const int grid = a.n_items < sm_count() ? a.n_items : sm_count();
  enc_conv_umma_kernel<<<grid, 256, kEncSmemBytes, st>>>(tm_x, tm_w, tm_r, tm_y, a);
  return launch_ok("enc_conv_umma_kernel");
}

int enc_split_from_f32(const float* x, void* y, int B, int C, int T, int halo_l, int halo_r, cudaStream_t st) {
  MST_CHECK(C % 64 == 0 && B <= 65535, "enc_split_from_f32: bad shape");
  dim3 grid(cdiv(T, 32), C / 64, B);
  enc_split_from_f32_kernel<<<grid, 256, 0, st>>>(x, (uint8_t*)y, C, T, halo_l, halo_r);
  return launch_ok("enc_split_from_f32_kernel");
}

int enc_pool_split(const void* x, float* emb, int B, int C, int T, int halo_l, int t_pad, cudaStream_t st) {
  const int warps = B * (C / 64);
  enc_pool_split_kernel<<<cdiv(warps, 8), 256, 0, st>>>((const uint8_t*)x, emb, B, C, T, halo_l, t_pad);
  return launch_ok("enc_pool_split_kernel");
}

}  // namespace mst
