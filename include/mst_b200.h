/*
 * mst_b200.h -- C ABI of libmst_b200.so: the sm_100a (B200) compute library behind the segment-batched
 * style-transfer forward path of jhtonyKoo/music_mixing_style_transfer.
 *
 * The reference is 100 % Python and has no FFI of its own: the drop-in boundary is its Python module
 * surface (networks.FXencoder / networks.TCNModel / mixing_manipulator chain).  This header is what OUR modules
 * bind (ctypes, music_mixing_style_transfer_b200/_cabi.py); INTEGRATION.md shows the binding a reference maintainer
 * would add.  Every entry point cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host"
 *   - no allocation inside: outputs and workspaces are caller-allocated (`*_bytes` queries)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - return 0 on success, non-zero on error; mst_last_error() (thread-local) holds the message
 *   - fp32 tensors are PyTorch-contiguous [B, C, T] (time fastest) exactly as the reference modules take them
 */
#ifndef MST_B200_H_
#define MST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MST_MAX_ENC_BLOCKS 16
#define MST_TCN_CH 128          /* channel_width the tcgen05 path is specialised for (inference/configs.yaml:27) */
#define MST_TCN_K 15            /* kernel_size (configs.yaml:26) */
#define MST_FX_NPARAMS 20       /* 13 EQ + 4 compressor + 1 imager + 2 gain, order in oracle/fx_oracle.py */

/* ---- library ---------------------------------------------------------------------------------------------- */
const char* mst_last_error(void);
int mst_version(void);
/* 0 if `device` is an sm_100 part with a driver that can encode TMA tensor maps */
int mst_device_check(int device);

/* ---- FXencoder: replaces FXencoder.forward (mixing_style_transfer/networks/architectures.py:65-70) and the
 *      Conv1d_layer / Res_ConvBlock stack under it (networks/network_utils.py:15-89, 96-119) ------------------ */
typedef struct {
  int n_blocks;                              /* 12 (inference/configs.yaml:8-10) */
  int channels[MST_MAX_ENC_BLOCKS + 1];      /* 2,16,32,...,2048: channels[i] -> channels[i+1] */
  int kernels[MST_MAX_ENC_BLOCKS];
  int strides[MST_MAX_ENC_BLOCKS];
} mst_enc_config;

/* Fold eval-mode BatchNorm1d into one Conv1d (network_utils.py:50,74): w_out[ci][k][co] = w[co][ci][k]*s[co]
 * (transposed, co fastest), b_out[co] = (b[co]-mean[co])*s[co] + bn_b[co], s = bn_w/sqrt(var+eps).  */
int mst_conv1d_fold_bn(const float* w, const float* b, const float* bn_w, const float* bn_b, const float* bn_mean,
                       const float* bn_var, float eps, int c_out, int c_in, int k, float* w_out, float* b_out,
                       void* stream);

/* One Conv1d_layer (mode "conv", padding "SAME"): y = relu(conv_stride(reflect_pad(x)) + b) [+ residual]
 * (network_utils.py:28-34,47-51,74,79-80; the `+ residual` is Res_ConvBlock's `conv1(x) + x`, :117).
 * w_folded/b_folded come from mst_conv1d_fold_bn.  T_out = ceil(T_in/stride).  residual may be NULL. */
int mst_enc_conv1d(const float* x, const float* w_folded, const float* b_folded, const float* residual, float* y,
                   int B, int c_in, int t_in, int c_out, int k, int stride, int relu, void* stream);

/* AdaptiveAvgPool1d(1).squeeze(-1) (architectures.py:62,67): y[B,C] = mean_t x[B,C,T] */
int mst_enc_mean_pool(const float* x, float* y, int B, int C, int T, void* stream);

/* out[c] = scale * sum over rows of x[rows, cols] (rows added in order): `stack -> reshape -> mean(0)` over the reference
 * segments' embeddings (inference/style_transfer.py:152-153) with scale = 1/rows, or a rank's partial sum with scale = 1 */
int mst_rows_reduce(const float* x, int rows, int cols, float scale, float* out, void* stream);

size_t mst_enc_packed_bytes(const mst_enc_config* cfg);
/* raw: host array of 12*n_blocks device pointers, per block: conv1 {w,b,bn_w,bn_b,bn_mean,bn_var}, conv2 {...} */
int mst_enc_pack(const mst_enc_config* cfg, const float* const* raw, float* packed, void* stream);
size_t mst_enc_workspace_bytes(const mst_enc_config* cfg, int B, int L);
/* whole encoder: x[B,2,L] -> emb[B,channels[n_blocks]] */
int mst_enc_forward(const mst_enc_config* cfg, const float* packed, const float* x, int B, int L, float* emb,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- MixFXcloner TCN: replaces TCNModel.forward (architectures.py:135-147), TCNBlock.forward (:222-234) and
 *      FiLM.forward (network_utils.py:180-182) ------------------------------------------------------------- */
typedef struct {
  int n_blocks;          /* 14 */
  int n_inputs;          /* 2  */
  int n_outputs;         /* 2  */
  int channels;          /* 128 (only value with a CUDA path) */
  int kernel_size;       /* 15 (only value with a CUDA path) */
  int dilation_growth;   /* 2  */
  int stack_size;        /* 15 */
  int cond_dim;          /* 2048 */
} mst_tcn_config;

/* Operand format of the dilated blocks (per call; both weight packs are always built):
 *   MST_TCN_F16F8   fp16(X) fp16(W) + two e4m3 correction products into one fp32 accumulator: 2 tensor units per algorithmic
 *                   MMA, fp32-grade accuracy (1e-5 RMS at full length) while every inter-block activation stays inside
 *                   +-MST_TCN_F16F8_RANGE (the e4m3 planes saturate above it).  `range_flag` reports a violation.
 *   MST_TCN_BF16X3  bf16 hi/lo pairs, three products: 3 tensor units, fp32 dynamic range -- the fall-back the Python
 *                   module repeats a forward in when range_flag fired.  The reference's FiLM gamma is unbounded
 *                   (networks/network_utils.py:180-182), hence the guard. */
#define MST_TCN_F16F8 0
#define MST_TCN_BF16X3 1
#define MST_TCN_F16F8_RANGE 448.0f

size_t mst_tcn_packed_bytes(const mst_tcn_config* cfg);
/* raw: host array of device pointers, per block n: {conv1.weight, bn.weight, bn.bias, bn.running_mean,
 * bn.running_var, res.weight, film.film_fc.weight, film.film_fc.bias}, then {output.weight, output.bias}.
 * Packs: BN folded into conv1 (block 0 fp32; blocks >=1 in BOTH operand formats, tap-major, K-major tiles),
 * per-channel BN bias and residual scale, FiLM weights, output projection. */
int mst_tcn_pack(const mst_tcn_config* cfg, const void* const* raw, void* packed, void* stream);

/* FiLM for all blocks at once (network_utils.py:180-181): film[n][bc][c] = (bn_bias, gamma, beta, res_scale) as
 * float4, gamma|beta = Linear_n(cond[bc]).  cond: [n_cond, cond_dim]; n_cond is 1 (broadcast) or B.
 * film_out: float[n_blocks * n_cond * channels * 4]. */
int mst_tcn_film_precompute(const mst_tcn_config* cfg, const void* packed, const float* cond, int n_cond,
                            float* film_out, void* stream);

size_t mst_tcn_workspace_bytes(const mst_tcn_config* cfg, int B, int L);
/* whole TCN: x[B,n_inputs,L] fp32 -> y[B,n_outputs,L] fp32 = clamp(output(blocks(x)), -1, 1).
 * `film` from mst_tcn_film_precompute with the same n_cond.  precision: MST_TCN_F16F8 / MST_TCN_BF16X3.
 * range_flag (device uint32, may be NULL; F16F8 only): zeroed at the start of the call, afterwards 0 if every inter-block
 * activation stayed inside the f16f8 range, else the float bits of the largest |activation| seen -> y is then only
 * single-pass-fp16 accurate and the caller should repeat the call with MST_TCN_BF16X3. */
int mst_tcn_forward(const mst_tcn_config* cfg, const void* packed, const float* x, const float* film, int n_cond,
                    float* y, int B, int L, void* workspace, size_t workspace_bytes, int precision,
                    unsigned int* range_flag, void* stream);

/* Layer-granular launches on the INTERNAL activation format (DESIGN.md "data layout": act[b][t][4 planes of 128 B],
 * 512 bytes per time step; a segment occupies ceil(L / 256) * 256 rows, i.e. a buffer is mst_tcn_workspace_bytes(cfg, B, L) / 2
 * bytes, 1024-byte aligned; the rows beyond L of a segment are zero padding the library maintains).  mst_tcn_forward is exactly
 * block0 + layer(1) ... layer(n_blocks-1, fuse_out=1); these entry points exist so a host can time or interleave
 * individual launches (bench.py's roofline leg).
 *   block0: x fp32 [B,n_inputs,L] -> act_out.   layer n>=1: act_in -> act_out, or, when fuse_out != 0 (last block),
 *   -> y fp32 [B,n_outputs,L] = clamp(Conv1d(128->n_out,k=1)(block(act_in)), -1, 1) and act_out is not written. */
int mst_tcn_block0_forward(const mst_tcn_config* cfg, const void* packed, const float* x, const float* film, int n_cond,
                           void* act_out, int B, int L, int precision, unsigned int* range_flag, void* stream);
int mst_tcn_layer_forward(const mst_tcn_config* cfg, const void* packed, int block, const void* act_in, void* act_out,
                          const float* film, int n_cond, int B, int L, int fuse_out, float* y, int precision,
                          unsigned int* range_flag, void* stream);

/* single TCNBlock n on fp32 [B,C,L] tensors (module-level surface + per-dilation parity tests):
 * y[B,128,L] = film(leaky_relu(bn(conv1(x)))) + res(x).  workspace >= mst_tcn_workspace_bytes(cfg,B,L). */
int mst_tcn_block_forward(const mst_tcn_config* cfg, const void* packed, int block, const float* x, const float* film,
                          int n_cond, float* y, int B, int L, void* workspace, size_t workspace_bytes, int precision,
                          void* stream);

/* ---- FX chain: replaces AugmentationChain.__call__ over Equaliser/Compressor/MidSideImager/Gain
 *      (mixing_manipulator/common_audioeffects.py:156-192, 501-525, 529-652, 965-992, 1038-1051) -------------- */
#define MST_FX_EQ 1
#define MST_FX_COMP 2
#define MST_FX_IMAGER 4
#define MST_FX_GAIN 8
#define MST_FX_RMSNORM 16   /* apply the chain's RMS re-normalisation after EQ / comp / imager (:142-145) */
size_t mst_fx_workspace_bytes(int B, int L);
/* x,y: fp32 [B,2,L] (channel-major; the reference's arrays are [L,2]); params: fp32 [B,20]; stages: MST_FX_* mask */
int mst_fx_chain_forward(const float* x, const float* params, float* y, int B, int L, float sample_rate, int stages,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- stereo building blocks of the input FX normaliser and the remaining processors ----------
 * Reference paths relative to mixing_style_transfer/mixing_manipulator/.
 * mst_biquad_cascade: up to 5 cascaded biquads with explicit coefficients coef[B][n_sections][5] = (b0, b1, b2, a1, a2), a0 = 1,
 *   device float64, zero initial state, on x[B,2,L] -> y (the EQ kernel's time-parallel scan).  Replaces the K-weighting
 *   `IIRfilter.apply_filter` pair of the BS.1770 meter behind fx_utils.lufs_normalize (fx_utils.py:220-238; pyloudnorm).
 * mst_stereo_stats: stats[B][4] = (sum L^2, sum R^2, sum L*R, max|x|) in float64: the reductions of normalize_imager /
 *   process_balance (normalization_imager.py:34-36,94-99) and the peak of lufs_normalize (fx_utils.py:231).
 * mst_stereo_mix: y = M x per frame, matrices[B][4] = (m0 m1; m2 m3) device float32: mid/side gains and L/R balance of
 *   normalize_imager (normalization_imager.py:31-76), Panner.process (common_audioeffects.py:935), the loudness gain.
 * mst_block_energy: z[c][j] = sum of squares of x[c][lo[j]:hi[j]] (float64): the gating blocks of the loudness meter.
 * mst_haas: y = x, y[ch] += feedback * roll(x[ch], delay) per segment (haas_process, common_audioeffects.py:767-787); x != y. */
int mst_biquad_cascade(const float* x, const double* coef, int n_sections, float* y, int B, int L, void* workspace,
                       size_t workspace_bytes, void* stream);
int mst_stereo_stats(const float* x, int B, long long L, double* stats, void* stream);
int mst_stereo_mix(const float* x, const float* matrices, float* y, int B, long long L, void* stream);
int mst_block_energy(const float* x, int n_channels, long long T, const long long* lo, const long long* hi, int n_blocks,
                     double* z, void* stream);
int mst_haas(const float* x, float* y, int B, long long L, const int* delay, const float* feedback, const int* channel,
             void* stream);

/* ---- spectral pieces of the input FX normaliser's EQ matching (csrc/spectral.cu) -------------------------
 * Reference paths relative to mixing_style_transfer/mixing_manipulator/.
 * mst_row_absmax: out[r] = max |x[r][0..T)| (float64, rows `stride` floats apart): np.max(np.abs(.)) per channel
 *   (utils_data_normalization.py:69, fx_utils.py:231).
 * mst_stft_mag_mean: out[s][k] = mean over frames of |rfft(window * x_s[f*hop : f*hop + n_fft])[k]|, k = 0..n_fft/2, float64:
 *   compute_stft + np.abs + np.mean of get_eq_matching (utils_data_normalization.py:74-79; common_miscellaneous.py:50-77,
 *   librosa.stft(center=False), n_frames = 1 + (T - n_fft) / hop, spectra in complex64).  n_fft: power of two in
 *   [1024, 65536]; window: device float32 [n_fft]; signals `stride` floats apart.
 * mst_fir_filtfilt: y_s = float32(scale[s] * filtfilt(taps_s, 1, x_s)) with scipy's defaults (padtype='odd', padlen = 3*n_taps,
 *   method='pad': utils_data_normalization.py:100-102), float64 arithmetic; taps: device float64 [n_signals][n_taps];
 *   scale: device float64 [n_signals] or NULL; needs T > 3*n_taps like scipy. */
int mst_row_absmax(const float* x, int n_rows, long long T, long long stride, double* out, void* stream);
size_t mst_stft_workspace_bytes(int n_signals, int n_fft);
int mst_stft_mag_mean(const float* x, int n_signals, long long T, long long stride, int n_fft, int hop, const float* window,
                      double* out, void* workspace, size_t workspace_bytes, void* stream);
size_t mst_fir_filtfilt_workspace_bytes(int n_signals, long long T, int n_taps);
int mst_fir_filtfilt(const float* x, int n_signals, long long T, long long stride, const double* taps, int n_taps,
                     const double* scale, float* y, long long y_stride, void* workspace, size_t workspace_bytes, void* stream);

/* ---- remaining FXmanipulator processors -------------------------------------------------------------------
 * mst_fft_convolve: ConvolutionalReverb.process (common_audioeffects.py:735-764): y[c][t] = dry * x[c][t] + wet * full[c][t + offset],
 *   full = x[c] convolved with h[c] (mode='full'; scipy.signal.oaconvolve in the reference), as a partitioned overlap-add on the
 *   device FFT of csrc/spectral.cu.  x, y: fp32 [2][T] (rows x_stride / y_stride apart); h: fp32 [h_channels][M], h_channels 1
 *   (the mono response serves both channels, :738-739) or 2; offset = the reference's cut index (:752-757).
 * mst_algo_reverb: AlgorithmicReverb.process (:1446-1509): per channel four damped feedback comb filters (the reference
 *   overwrites the sum of combs 1-4 with comb 5, :1474-1488 -- reproduced) followed by four all-pass sections, then the wet /
 *   dry / width mix.  The comb / all-pass arithmetic is pymixconsole's (third-party, absent): restated, PARITY UNPINNED.
 *   x, y: fp32 [B][2][L]; params: fp32 [B][5] = (room_size, damping, dry_mix, wet_mix, width). */
size_t mst_fft_convolve_workspace_bytes(long long T, long long M);
int mst_fft_convolve(const float* x, long long T, long long x_stride, const float* h, long long M, long long h_stride,
                     int h_channels, long long offset, float dry, float wet, float* y, long long y_stride, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t mst_algo_reverb_workspace_bytes(int B, int L);
int mst_algo_reverb(const float* x, const float* params, float* y, int B, int L, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- WAV sample formats on the device: the steps either side of the forward -------------------
 * mst_pcm_decode replaces load_wav_segment's int -> float conversion and de-interleave
 * (mixing_style_transfer/data_loader/loader_utils.py:54-70) plus the stem clamp (data_loader/data_loader.py:589-590) and
 * the mono duplication of inference/feature_extraction.py:87-89:
 *   pcm: interleaved [n_frames][n_channels] int16 (sample_bytes 2, x / 2^15) or int32 (sample_bytes 4, x / 2^31)
 *   out: fp32 planar, channel c at out + c * out_stride (always two planes; mono input fills both)
 * mst_pcm_encode_mix replaces the remix `sum(inst_outputs)` and the PCM_16 file quantisation
 * (inference/style_transfer.py:165-177): stems fp32 [n_stems][2][stem_stride] -> int16 interleaved [n_frames][2],
 * float32 adds in stem order, then clip(rint(x * 32768), -32768, 32767). */
int mst_pcm_decode(const void* pcm, int sample_bytes, int n_channels, long long n_frames, float* out,
                   long long out_stride, void* stream);
int mst_pcm_encode_mix(const float* stems, int n_stems, long long stem_stride, long long n_frames, int16_t* pcm,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_B200_H_ */
