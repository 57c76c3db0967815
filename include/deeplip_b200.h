/*
 * deeplip_b200 -- C ABI of the B200 (sm_100a) kernels behind the DeepLip audio-visual
 * embedding-extraction + trial-scoring hot path.
 *
 * The reference (DanielMengLiu/DeepLip) has no FFI of its own: the boundary it exposes is the
 * Python module interface of models/audio_models, models/video_models and models/fusion_models
 * (SURVEY.md 8(b)).  Every entry point below replaces the PyTorch/NumPy/sklearn library call(s)
 * made at the cited reference lines; the drop-in nn.Modules in deeplip_b200/ bind them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions (all entry points):
 *   - plain device pointers + sizes; no torch types; caller owns every buffer (no allocation here);
 *   - work is ENQUEUED on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - returns DL_OK (0) or a negative DL_ERR_* code; dl_last_error() gives the message (thread-local);
 *   - never throws; there is no CPU fallback: without a CUDA device every compute call fails.
 *   - "bf16" buffers are raw 16-bit bfloat16; activations are channels-last.
 */
#ifndef DEEPLIP_B200_H
#define DEEPLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DL_OK 0
#define DL_ERR_INVALID (-1)     /* bad argument / unsupported shape                       */
#define DL_ERR_CUDA (-2)        /* CUDA runtime / driver error (launch, tensor-map encode) */
#define DL_ERR_UNSUPPORTED (-3) /* device is not sm_100                                    */

int dl_version(void);
const char* dl_last_error(void);
/* Tuning switches for A/B measurements: "pair" (CTA-pair igemm kernels, default 1), "pair_resident"
 * (smem-resident weight half in the pair kernel, default 1), "tap_share" (one operand-A box per filter row in
 * the guarded-linear pair kernel, default 1), "frontend" (2 = register-resident FFT front end, 1 = first generation),
 * "prepass" (2 | 1, stem pre-pass generation), "stem" (2 = channels-on-lanes stem kernel where the shape allows it:
 * W = 88 and H % 8 == 0, 1 = first-generation kernel for every shape; the two agree to fp32 summation order),
 * "small_linear" (1 = fc layers on linear_small_kernel, 0 = igemm),
 * "statpool_mlp" (4 | 8 loads in flight), "staged_epilogue" (1 = the CTA-pair kernels send their output tiles through
 * shared-memory slabs and TMA stores, 0 = per-lane 16-byte stores; bit-identical), "stft_pad" (0 reflect | 1 zeros: a
 * convention, not a tuning switch), "pool_fuse" (1 = dl_conv_desc.avgpool is taken in the pair kernel's epilogue where
 * the shape allows it, 0 = always conv + pooling kernel; bit-identical).  Results agree to fp32 summation order either way.
 * "dbg" (default 0) is a measurement aid only: bits 1/2/4 drop the residual / stores / whole epilogue of the
 * pair kernel, 8 issues one MMA in four, 16/32 idle the stem's builders / epilogue, 4096/8192 widen the stem's MMAs,
 * 16384 drops its strip loads, bits 16..19 set its strip ring depth (tools/entry_ablate.py, tools/stem_try.py); in the
 * second-generation stem 2/32 drop the bulk stores / staging writes, 8 issues one MMA per stage, 16 skips the builders'
 * strip reads, 64 the epilogue arithmetic, 128 the strip loads, 256/512 substitute the operand data, 1024 multiplies
 * the padded 8th window row too (tools/stem2_abl.py) -- any non-zero value produces WRONG results by design. */
int dl_set_option(const char* name, int value);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long dl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  audio front end.  Replaces python_speech_features.{mfcc,fbank,logfbank} / librosa.stft + magphase +
 *     log1p, and `_normalize` (models/fusion_models/datasets.py:227-246, 214-215).  wav: (B, nsamp) f32 ->
 *     feat_bf16: (B, T, ld_bf16) channels-last bf16 (zero padded to ld_bf16 channels; may be NULL),
 *     feat_f32 : (B, F, T) f32, the layout the reference hands to the model (required: it doubles
 *                as the pre-CMVN scratch, this library never allocates).
 *     kind: 0 = mfcc(numcep=F, nfilt=26), 1 = fbank(nfilt=F), 2 = logfbank(nfilt=F),
 *           3 = stft(n_fft=512, hop 160, periodic Hann 400 centred in 512, center=True): F = 257 rows of
 *               log1p|S|; the centre padding is librosa's pre-0.10 default `reflect` unless
 *               dl_set_option("stft_pad", 1) selects zeros (librosa >= 0.10).
 *     delta: 0, or the order of `_delta` (datasets.py:217-225; the reference always uses 2): python_speech_features.
 *       delta(feat, N) for N = 1 .. delta of the (normalised) features is appended, feat_f32 is then
 *       (B, F (1 + delta), T) = [feat | delta N=1 | delta N=2] and feat_bf16 carries the same F (1 + delta) channels.
 *     lengths: per-utterance valid sample counts (device int32, may be NULL = nsamp for all);
 *     T must equal 1 + ceil((nsamp - 400) / 160) for the padded length nsamp (16 kHz, 25/10 ms), or
 *     1 + nsamp / 160 for kind 3.
 *     dl_set_option("frontend", 1) selects the first-generation kernels (radix-2 FFT in shared memory;
 *     kinds 0-2 only), kept for A/B measurements; the default (2) holds the FFT in registers.
 */
int dl_frontend_features(const float* wav, const int32_t* lengths, int B, int nsamp, int kind, int F,
                         int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T, void* stream);

/* The same with wav as (B, nsamp) int16 PCM -- the sample format of the corpus' wav files; soundfile.read
 * (models/fusion_models/datasets.py:70-76, 327-331) hands the reference value / 32768, which is what the kernel's load
 * computes (exact in f32: both entry points give the same bits).  Halves the audio's share of the host-to-device
 * ingest (96 KB instead of 192 KB per 3 s utterance). */
int dl_frontend_features_pcm16(const int16_t* wav, const int32_t* lengths, int B, int nsamp, int kind, int F,
                               int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T, void* stream);

/* (B, C, T) f32 -> (B, T, ldc) bf16 channels-last, zero padded: the layout change in front of the
 * TDNN for callers that bring their own features (models/audio_models/tdnn.py:89 input). */
int dl_nct_to_ntc_bf16(const float* x, int B, int C, int T, void* y, int ldc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  video stem.  Replaces Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3)) + BatchNorm3d + PReLU +
 *     MaxPool3d((1,3,3),(1,2,2),(0,1,1)) and the NCTHW -> (N*T)CHW copy
 *     (models/video_models/model.py:81-85, 9-13).
 *     Input either f32 frames (B, T, H, W) already normalised (u8 == 0) or raw u8 crops
 *     (B, T, Hraw, Wraw) with the reference preprocessing fused into the load (u8 == 1):
 *     x/255 -> centre crop to HxW -> (x-mean)/std  (models/video_models/dataloaders.py:19-24).
 *     w_packed: (64, 320) bf16 = [cout][kt][kh*8+kw] zero padded (see deeplip_b200/packing.py).
 *     y: (B*T, out_img_rows, W/4, 64) bf16 channels-last; out_img_rows >= H/4 is the row pitch of one
 *     frame (0 = H/4).  Rows H/4 .. out_img_rows-1 are not written ("stacked rows" layout of
 *     dl_conv3x3_c64_halo_bf16: the caller keeps them zero).
 *     lengths: valid frames per clip (device int32[B], may be NULL): frames t >= lengths[b] are taken as
 *     all-zero NORMALISED frames, the reference's pad-after-preprocessing convention for ragged batches
 *     (pad_packed_collate, models/video_models/dataset.py:123-139) -- needed for raw u8 input, where no
 *     padding byte normalises to zero.
 *     workspace: caller-owned scratch of dl_stem_workspace_bytes(B, T, H, W) bytes (the normalised,
 *     zero-bordered bf16 frames the TMA unit streams from; this library never allocates).
 *     dl_set_option("prepass", 1) selects the first-generation pre-pass kernel (A/B measurements; no lengths).
 *     Two kernels: W == 88 (the corpus crop) with H % 8 == 0 runs on stem2_conv3d_kernel (channels on the
 *     accumulator lanes), every other shape -- and every shape after dl_set_option("stem", 1) --
 *     on stem_conv3d_kernel.  Same contract, results equal to fp32 summation order.
 */
long long dl_stem_workspace_bytes(int B, int T, int H, int W);
/* The two halves of dl_stem_conv3d_bn_prelu_pool as calls of their own, for callers that run the pre-pass (x ->
 * workspace: HBM-bound, no shared memory) on another stream under kernels that leave room for it -- the audio branch
 * of deeplip_b200.pipeline -- and the main kernel (workspace -> y) afterwards.  Same arithmetic, same bits. */
int dl_stem_prepass(const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw, float mean, float std,
                    const int32_t* lengths, void* workspace, void* stream);
int dl_stem_conv3d_prepassed(int B, int T, int H, int W, const void* w_packed, const float* scale, const float* shift,
                             const float* slope, void* y, int out_img_rows, void* workspace, void* stream);
int dl_stem_conv3d_bn_prelu_pool(const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw,
                                 float mean, float std, const void* w_packed, const float* scale,
                                 const float* shift, const float* slope, void* y, int out_img_rows,
                                 const int32_t* lengths, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3/K5/K7  implicit-GEMM convolution on tcgen05 tensor cores (TMA im2col operand A, TMA tiled
 *     operand B, fp32 accumulation in TMEM) with a fused epilogue:
 *         v = acc * scale[c] + shift[c] (+ residual);  y = v > 0 ? v : v * slope[c]
 *     Replaces nn.Conv2d + BatchNorm2d + PReLU (+ residual add) in BasicBlock
 *     (models/video_models/resnet.py:56-69), the 1x1 stride-2 downsample (resnet.py:13-17),
 *     nn.Conv1d + BatchNorm1d + LeakyReLU in TDNN_Block (models/audio_models/tdnn.py:35-43, H = R = 1)
 *     and nn.Linear + BatchNorm1d + LeakyReLU in the heads (tdnn.py:93-100, model_fusion.py:19-24;
 *     H = W = R = S = 1).
 *     x        : (N, H, W, ldx) bf16 channels-last, C logical channels (ldx % 8 == 0)
 *     w_packed : (Cout, R*S*ceil64(C)) bf16, K index = (r*S + s) * ceil64(C) + c
 *     y        : (N*P*Q, ldy) bf16 or NULL;  residual: same layout as y or NULL
 *     y_f32    : (N*P*Q, ldf) f32 or NULL = lrelu(acc * scale2[c] + shift2[c], f32_slope)  (side output,
 *                e.g. x_a of extract_embedding, tdnn.py:93; scale2/shift2 NULL = raw accumulator)
 *     Cout % 8 == 0.
 *
 *     Guarded layouts.  TMA im2col delivers ~5 cycles per pixel row on B200, which bounds every stride-1
 *     layer below the tensor pipe; plain tiled TMA boxes are several times faster.  A stride-1 convolution
 *     can use them when the zero padding is materialised in the tensor itself:
 *       lin = 1 : x, y and residual all are (N, H, W, ld) where H and W INCLUDE the guard rows / columns
 *                 (>= pad_h rows below and >= pad_w columns right of every image, kept zero by the caller).
 *                 Output pixel (n, h, w) sits at the same (n, h, w) as the input pixel under the filter
 *                 anchor + pad; it is stored only for h < valid_h and w < valid_w, so the guards stay zero.
 *                 Tap (r, s) of a 128-pixel tile is the tile's row range shifted by
 *                 (r*dil_h - pad_h) * W + (s*dil_w - pad_w) rows.  stride must be 1.  Conv1d over
 *                 (B, T, C) needs no guards at all: H = 1, W = T, valid_w = T - (S-1)*dil (tdnn.py:35-43).
 *       out_img_rows / out_img_cols : (lin = 0) the im2col kernel writes y / reads residual at
 *                 ((n * out_img_rows + p) * out_img_cols + q) instead of the dense (n*P + p)*Q + q, i.e.
 *                 straight into a guarded tensor;  img_rows / img_cols: it reads x from one.
 */
typedef struct dl_conv_desc {
  int N, H, W, C, ldx;
  int Cout, R, S;
  int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int ldy, ldf;
  float f32_slope;   /* leaky slope applied to the f32 side output (1.0f = none) */
  int img_rows;      /* row pitch of one input image (>= H; 0 = H): input may be in a stacked / guarded layout */
  int img_cols;      /* column pitch of one input row (>= W; 0 = W) */
  int lin;           /* 1 = guarded-linear operand A (tiled TMA), see above */
  int valid_h, valid_w;            /* lin = 1: extents of the stored outputs */
  int out_img_rows, out_img_cols;  /* lin = 0: pitches of y / residual (0 = dense P, Q) */
  /* Sibling convs as one launch (0 = off): output channels c >= split_channel are stored to y_split[row * ldy +
   * c - split_channel] instead of y (both with pitch ldy >= max(split_channel, Cout - split_channel)); with
   * split_center_only = 1 the caller declares that those channels' weights are zero off the filter's centre tap
   * (a 1x1 stride-s pad-0 conv riding a 3x3 stride-s pad-1 one, resnet.py:13-17 + :56-59): the CTA-pair kernel then
   * skips their other K blocks.  Needs split_channel % 256 == 0, no residual, no f32 output, lin = 0. */
  int split_channel, split_center_only;
  void* y_split;
  /* 0 = off.  Declares that output channels c >= center_only_from have zero weights off the filter's centre tap,
   * WITHOUT a split output (layer2's entry block: conv1 and the 1x1 skip as one 64 -> 256 conv whose halves share one
   * pitch-256 buffer).  With Cout = 256 and center_only_from = 128 the CTA-pair kernel multiplies the conv1 half with
   * N = 128 MMAs over all taps and the skip half with N = 128 MMAs on the centre tap only (5/9 of the tensor work of
   * the N = 256 form); every other kernel ignores the hint and multiplies the zeros: same bits. */
  int center_only_from;
  /* K4 in the conv's epilogue (0 = off): avgpool = 1 asks for the global average pool of every output image
   * (AdaptiveAvgPool2d(1), models/video_models/resnet.py:125-126) as avgpool_out[n * Cout + c] f32, computed from the
   * bf16-rounded outputs in the order dl_frame_pool_temporal_mean uses (same bits).  Where the shape allows it -- the
   * CTA-pair kernel with the staged epilogue, P*Q <= 128, Cout % 64 == 0 -- the means are taken from the staged output
   * tiles (m tiles then step by the largest multiple of P*Q rows <= 128, so that no image straddles two tiles) and,
   * with avgpool_keep_y = 0, the bf16 tensor is not written at all; otherwise the conv runs as usual and a pooling
   * kernel follows.  y must be a valid (N, P, Q, Cout) buffer either way (scratch when avgpool_keep_y = 0).
   * Needs a dense output (no split, no guarded / sliced y), no f32 side output. */
  int avgpool, avgpool_keep_y;
  float* avgpool_out;
} dl_conv_desc;

int dl_conv_igemm_bf16(const void* x, const void* w_packed, const float* scale, const float* shift,
                       const float* slope, const void* residual, void* y, float* y_f32, const float* scale2,
                       const float* shift2, const dl_conv_desc* desc, void* stream);

/* 3x3 / stride 1 / pad 1, 64 -> 64 channel convolution with shared-memory operand reuse (one halo patch
 * feeds all nine taps, weights resident): the ResNet layer1 BasicBlock convs (resnet.py:56-69) at a ninth
 * of the L2 traffic of the generic kernel.  Same fused epilogue (scale, shift, residual, slope).
 * x, y, residual: (N, img_rows, W, 64) bf16 "stacked rows": img_rows >= H + 1 and rows H .. img_rows-1 of
 * every image are zero in x (they act as vertical padding); the kernel writes zeros to those rows of y.
 * w_packed: (64, 576) bf16 as for dl_conv_igemm_bf16.  W >= 8.
 */
int dl_conv3x3_c64_halo_bf16(const void* x, const void* w_packed, const float* scale, const float* shift,
                             const float* slope, const void* residual, void* y, int N, int H, int W,
                             int img_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  per-frame global average pool + masked temporal mean.  Replaces AdaptiveAvgPool2d(1)
 *     (models/video_models/resnet.py:125-126), `_average_batch` (model.py:16-17) and
 *     torch.mean(..., dim=0) over one clip (train_fusion.py:400).
 *     x: (B*T, HW, C) bf16 -> frame_feats (B, T, C) f32 (may be NULL), utt_mean (B, C) f32 (may be
 *     NULL) = mean over the first lengths[b] frames (lengths NULL = T).
 */
int dl_frame_pool_temporal_mean(const void* x, int B, int T, int HW, int C, const int32_t* lengths,
                                float* frame_feats, float* utt_mean, void* stream);
/* The temporal half of K4 on frame features that already exist (dl_conv_desc.avgpool): frame_feats (B, T, C) f32 ->
 * utt_mean (B, C) f32 = mean over the first lengths[b] frames (NULL = T), summed in dl_frame_pool_temporal_mean's
 * order (same bits as the one-kernel form). */
int dl_temporal_mean_f32(const float* frame_feats, int B, int T, int C, const int32_t* lengths, float* utt_mean,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6  statistics pooling.  Replaces MeanStdPooling (models/audio_models/pooling.py:18-26):
 *     mean || unbiased std over time.  x: (B, T, ldx) bf16 channels-last, C channels ->
 *     out_f32 (B, 2C) f32 (may be NULL) and out_bf16 (B, ld_out) bf16 (may be NULL).
 *     lengths: valid frames per utterance (NULL = T) for ragged batches.
 */
int dl_stat_pool(const void* x, int B, int T, int C, int ldx, const int32_t* lengths, float* out_f32,
                 void* out_bf16, int ld_out, void* stream);

/* Attentive statistics pooling tail (models/audio_models/pooling.py:96-106): given the attention
 * logits e (B, T) f32 (= v . relu(W x + b) + k, produced with dl_conv_igemm_bf16 + dl_attn_logits),
 * alpha = softmax_T(e); mean = sum alpha x; std = sqrt(sum alpha x^2 - mean^2).  Same outputs as
 * dl_stat_pool. */
int dl_attn_stat_pool(const void* x, const float* logits, int B, int T, int C, int ldx,
                      const int32_t* lengths, float* out_f32, void* out_bf16, int ld_out, void* stream);
/* e[b,t] = sum_h v[h] * relu(h[b,t,h]) + k   (pooling.py:97-99); h: (B*T, ldh) f32 = W x + b. */
int dl_attn_logits(const float* h, int rows, int Hd, int ldh, const float* v, float k, float* e,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * K8  fusion.
 *  dl_znorm_concat: Trainer.feature_normalize x2 + torch.cat([audio, video], 1)
 *     (train_fusion.py:233-238, 405-410; unbiased std) or, with biased != 0 and video_first != 0,
 *     the NumPy variant feature_normalize + hstack((video, audio)) (models/fusion_models/utils.py:
 *     465-471, 524-527).  Optionally L2-normalises the fused row (what cosine scoring needs).
 *     a: (B, Da) f32, v: (B, Dv) f32 -> out (B, Da+Dv) f32, out_bf16 optional (B, Da+Dv) bf16.
 *  dl_lowfer: LowFER.forward's live path cat([e1, sigmoid(e2), e1*sigmoid(e2)], 1)
 *     (models/fusion_models/LBP.py:46-50).
 *  dl_l2_normalize: F.normalize(xv) (train_audio.py:430) / the row normalisation inside
 *     sklearn cosine_similarity (zero norm -> divide by 1).
 */
int dl_znorm_concat(const float* a, int Da, const float* v, int Dv, int B, int biased, int video_first,
                    int l2norm, float* out, void* out_bf16, void* stream);
int dl_lowfer(const float* e1, const float* e2, int B, int D, float* out, void* stream);
int dl_l2_normalize(const float* x, int B, int D, float* out, void* out_bf16, void* stream);
/* y = lrelu(x * scale[c] + shift[c], slope) on f32 (rows, C): bf16 copy (rows, ldc, zero padded) and/or f32
 * copy (rows, C).  BatchNorm1d + LeakyReLU after the heads (tdnn.py:103-111) and the f32 -> bf16 hand-off. */
int dl_affine_act(const float* x, int rows, int C, const float* scale, const float* shift, float slope,
                  void* y_bf16, int ldc, float* y_f32, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9  trial scoring.  Replaces the per-trial sklearn cosine_similarity loop
 *     (models/fusion_models/utils.py:272-279): scores[i] = cos(emb[enrol[i]], emb[test[i]]).
 *     emb: (N_utt, D) f32 (need not be normalised); enrol/test: int32[n_trials] rows into emb.
 *     dl_score_fusion adds the 0.5/0.5 score fusion of utils.py:343-377 (second embedding table,
 *     eps = 1e-8 clamp on the video norms like F.cosine_similarity).
 */
int dl_cosine_score_trials(const float* emb, int n_utt, int D, const int32_t* enrol, const int32_t* test,
                           int n_trials, float* scores, void* stream);
int dl_score_fusion_trials(const float* emb_a, int Da, const float* emb_v, int Dv, int n_utt,
                           const int32_t* enrol, const int32_t* test, int n_trials, float* scores,
                           void* stream);
/* Dense formulation: gather scores[i] = S[enrol_row[i], test_col[i]] from a score matrix S (ld) f32
 * produced by dl_conv_igemm_bf16 on L2-normalised bf16 embeddings (enrol x test^T). */
int dl_gather_scores(const float* S, int ld, const int32_t* rows, const int32_t* cols, int n_trials,
                     float* scores, void* stream);

/* ------------------------------------------------------------------------------------------------
 * S5  PLDA trial scoring.  Replaces the per-trial body of eer_plda_grid / eer_plda_lomgrid
 *     (models/audio_models/utils.py:285-329): `model.transform(em, 'D', 'U_model')` +
 *     `model.calc_same_diff_log_likelihood_ratio(U_0, U_1)` of the third-party `plda` package.
 *     dl_plda_transform: every utterance once, u[row, r] = bias[r] + sum_d emb[row, d] M[r, d]  (M: R x D f32, the
 *       fitted PCA + A^-1 maps restricted to the R <= 32 relevant dimensions, folded by deeplip_b200/plda.py).
 *     dl_plda_llr_trials: scores[t] = c0 + sum_r k1[r] (a_r + b_r)^2 - k2[r] (a_r^2 + b_r^2) with a = u[enrol[t]],
 *       b = u[test[t]] (k1 = psi / (2 (2 psi + 1)), k2 = psi / (2 (psi + 1)), c0 = sum_r log(psi + 1) - log(2 psi + 1) / 2);
 *       out-of-range indices give NaN, like dl_cosine_score_trials.
 */
int dl_plda_transform(const float* emb, int n_utt, int D, const float* M, const float* bias, int R, float* u,
                      void* stream);
int dl_plda_llr_trials(const float* u, int n_utt, int R, const float* k1, const float* k2, float c0,
                       const int32_t* enrol, const int32_t* test, int n_trials, float* scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPLIP_B200_H */
