/* dinounet_b200 — C-ABI of the B200-native Dino U-Net forward path.
 *
 * Plain pointers and sizes only (no torch types).  Every function launches hand-written sm_100a kernels on the
 * given CUDA stream, never allocates, never synchronises, and returns 0 on success or a negative error code
 * (message via b2u_last_error()).  All device tensors are channels-last ("token-major"): [rows, C] with C contiguous.
 *
 * Reference interface each entry replaces (paths relative to the reference repo root):
 *   - the reference's ONLY native FFI is the pybind11 module `MultiScaleDeformableAttention`
 *     (dinounet/dinov3/eval/segmentation/models/utils/ops/src/vision.cpp:18-21, ms_deform_attn.h:26-67);
 *     b2u_msda_forward is its drop-in (same tensor meaning, fused softmax/location prologue available).
 *   - every other entry replaces an ATen/cuDNN/cuBLAS library call that the reference's Python forward makes
 *     (SURVEY.md section 2.3); the call site is cited per function.
 */
#ifndef DINOUNET_B200_H_
#define DINOUNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2u_stream_t; /* cudaStream_t */

enum { B2U_F16 = 0, B2U_BF16 = 1 };
enum { B2U_ACT_NONE = 0, B2U_ACT_GELU = 1, B2U_ACT_RELU = 2, B2U_ACT_LRELU = 3,
       /* act1 = B2U_ACT_SWIGLU (SwiGLUFFN.forward, ffn_layers.py:73-77): Wp rows (and bias) are packed in 32-row blocks
        * [w1 rows 32j..32j+31 | w2 rows 32j..32j+31], N = 2*hidden; out[:, j] = silu(x.w1_j + b1_j) * (x.w2_j + b2_j),
        * 16-bit [M, N/2] (ldc/col_off in output columns). */
       B2U_ACT_SWIGLU = 4 };
enum { B2U_CONV_NONE = 0, B2U_CONV3X3_S1 = 1, B2U_CONV3X3_S2 = 2 };

/* Epilogue applied to each fp32 accumulator element acc[m, n], in this order:
 *   v = acc + bias[n]; if (round16) v = round_to_dtype(v); v = act1(v); v = v * scale[n] + shift[n]; v = act2(v);
 *   v += residual[orow, ocol] (fp32) ; v += add16[orow, ocol] (16-bit) ; out[orow, ocol] = v
 * (orow, ocol) = output addressing: identity, batch row-remap, or 2x2 pixel-shuffle (ConvTranspose2d k2 s2). */
typedef struct b2u_epilogue {
  void* out;            /* 16-bit (dtype) or fp32 */
  int32_t out_fp32;     /* 1: out is float */
  int64_t ldc;          /* output row stride, elements */
  int32_t col_off;      /* output column offset (concat buffers) */
  int32_t rows_in;      /* row remap: orow = (m / rows_in) * rows_out + row_off + m % rows_in ; 0 = identity */
  int32_t rows_out;
  int32_t row_off;
  int32_t ps_cout;      /* >0: pixel-shuffle store, n = q*ps_cout + co, q = 2*a + b, input pixel grid ps_h x ps_w */
  int32_t ps_h;
  int32_t ps_w;
  const float* bias;    /* [N] or NULL */
  const float* scale;   /* [N] or NULL */
  const float* shift;   /* [N] or NULL */
  int32_t act1;
  int32_t act2;
  int32_t round16;
  const float* residual; /* fp32 [*, ldres], may alias out */
  int64_t ldres;
  const void* add16;     /* 16-bit [*, ldadd] */
  int64_t ldadd;
} b2u_epilogue;

/* tcgen05 tensor-core GEMM / implicit-GEMM 3x3 convolution:  acc[M, N] = A[M, K] * Wp[N, K]^T.
 * conv == NONE : A is [M, K] row-major (row stride lda).  Replaces F.linear / 1x1 Conv2d / ConvTranspose2d(k2,s2)
 *                (attention.py:88-90, ffn_layers.py:43-49, ms_deform_attn.py:183-215, dinov3_adapter.py:274-277,467,
 *                dinounet_training.py:419-441,255-264,613).
 * conv != NONE : A is an NHWC image [B, Hin, Win, C]; the kernel walks the 9 taps with TMA halo tiles (zero fill),
 *                no im2col buffer.  Wp is [N, 9*Cpad] with k = tap*Cpad + c, Cpad = ceil(C/64)*64.  Output pixel
 *                grid is Hin/stride x Win/stride.  Replaces Conv2d(3x3,pad 1) (dinov3_adapter.py:239-273,
 *                dynamic_network_architectures ConvDropoutNormReLU). */
typedef struct b2u_gemm_params {
  int32_t M, N, K;
  const void* A;
  int64_t lda;
  const void* Wp;       /* packed weights [N, ldw], 16-bit, zero padded */
  int64_t ldw;
  int32_t dtype;        /* B2U_F16 / B2U_BF16 for A, Wp and 16-bit outputs */
  int32_t conv;
  int32_t B, Hin, Win, C;
  b2u_epilogue epi;
} b2u_gemm_params;

int b2u_gemm(const b2u_gemm_params* p, b2u_stream_t stream);

/* QKV projection with masked bias + RoPE + head split fused in the epilogue (attention.py:30-40,66-92):
 *   qkv = A[M=B*ntok, D] * Wp[3D, D]^T + bias (already multiplied by bias_mask); rounded to dtype; q,k rows with
 *   token index >= prefix rotated with sin/cos[(tok - prefix), hd] (fp32); written as q,k,v [B, heads, ntok, hd],
 *   hd = D / heads = 64 (ViT-S/B/L) or 128 (ViT-7B). */
typedef struct b2u_qkv_params {
  int32_t B, ntok, D, heads, prefix;
  const void* A;
  int64_t lda;
  const void* Wp;
  int64_t ldw;
  const float* bias;    /* [3D] or NULL */
  const float* rope_sin;
  const float* rope_cos;
  void* q;
  void* k;
  void* v;
  int32_t dtype;
  int32_t v_transposed; /* 1: write V as V^T [B, heads, 64, npad] (keys contiguous) for b2u_attention_tc */
  int32_t npad;         /* row length of V^T: multiple of 8, >= ntok; columns >= ntok are never written (keep them 0) */
  int32_t rope_w;       /* > 0: the rope tables describe an (ntok-prefix)/rope_w x rope_w patch grid with separable angles
                           (rope_position_encoding.py:99-104: first 16 of 32 angles depend on the row, last 16 on the column,
                           tiled twice) -> the kernel keeps compact per-row / per-column tables in shared memory */
} b2u_qkv_params;

int b2u_qkv_rope(const b2u_qkv_params* p, b2u_stream_t stream);

/* tcgen05 / TMEM flash attention, non-causal softmax attention (attention.py:106-118 -> F.scaled_dot_product_attention):
 * q,k [B, heads, ntok, 64], vt = V^T [B, heads, 64, npad]
 * with zero padding columns (see b2u_qkv_params.v_transposed) -> out [B, ntok, heads*64].  Only the query rows
 * [q_begin, ntok) are produced (128-row tiles, two per work item); keys always span [0, ntok).  With q_begin = the
 * number of cls/storage tokens the ViT's 1024 patch rows fill whole tiles and b2u_attention_rows does the prefix. */
int b2u_attention_tc(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads, int32_t ntok,
                     int32_t npad, int32_t q_begin, float scale, int32_t dtype, b2u_stream_t stream);

/* Same kernel for head_dim 64 or 128 (ViT-7B: q,k [B, heads, ntok, 128], vt [B, heads, 128, npad]); all query rows. */
int b2u_attention_tc_hd(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads, int32_t ntok,
                        int32_t npad, int32_t head_dim, float scale, int32_t dtype, b2u_stream_t stream);

/* Few-row attention (same layouts as b2u_attention_tc): query rows [row_begin, row_begin + nrows), one warp per row,
 * fp32 softmax.  Meant for the handful of prefix-token rows. */
int b2u_attention_rows(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads, int32_t ntok,
                       int32_t npad, int32_t row_begin, int32_t nrows, float scale, int32_t dtype, b2u_stream_t stream);

/* LayerNorm over the last dim (block.py:193-194, vision_transformer.py:300, dinov3_adapter.py:142-148):
 *   in fp32 [rows_sel, D] -> out (16-bit or fp32).  Row selection: for output row r,
 *   input row = (r / rows_out) * rows_in + row_off + r % rows_out  (rows_out == 0: identity). */
int b2u_layernorm(const float* in, void* out, const float* gamma, const float* beta, int32_t rows, int32_t D,
                  float eps, int32_t rows_in, int32_t rows_out, int32_t row_off, int32_t out_fp32, int32_t dtype,
                  b2u_stream_t stream);
/* Same, for a 16-bit input stream `in` (element type = dtype). */
int b2u_layernorm16(const void* in, void* out, const float* gamma, const float* beta, int32_t rows, int32_t D,
                  float eps, int32_t rows_in, int32_t rows_out, int32_t row_off, int32_t out_fp32, int32_t dtype,
                  b2u_stream_t stream);

/* fp32 -> 16-bit cast of a [rows, D] matrix with the same row selection as b2u_layernorm. */
int b2u_cast_rows(const float* in, void* out, int32_t rows, int32_t D, int32_t rows_in, int32_t rows_out,
                  int32_t row_off, int32_t dtype, b2u_stream_t stream);
/* 16-bit -> 16-bit row gather with the same row selection (no conversion). */
int b2u_copy_rows16(const void* in, void* out, int32_t rows, int32_t D, int32_t rows_in, int32_t rows_out,
                    int32_t row_off, b2u_stream_t stream);

/* Patch-embed operand: x fp32 NCHW [B,3,S,S] -> [B*(S/16)^2, 768] 16-bit with k = c*256 + ky*16 + kx
 * (non-overlapping patches: a permutation + cast, no duplication)  (patch_embed.py:64-76). */
int b2u_patchify(const float* x, void* out, int32_t B, int32_t S, int32_t dtype, b2u_stream_t stream);

/* cls/storage prefix rows of the token stream: X[b, t, :] = prefix[t, :] for t < n_prefix (vision_transformer.py:207-214). */
int b2u_write_prefix(float* X, const float* prefix, int32_t B, int32_t ntok, int32_t n_prefix, int32_t D,
                     b2u_stream_t stream);

/* SPM stem conv0: Conv2d(3,64,3,s2,p1,no bias)+BN(eval)+ReLU reading fp32 NCHW, writing NHWC 16-bit
 * (dinov3_adapter.py:241-243).  w is [64,3,3,3] fp32; scale/shift = folded BN. */
int b2u_stem_conv0(const float* x, const float* w, const float* scale, const float* shift, void* out, int32_t B,
                   int32_t S, int32_t dtype, b2u_stream_t stream);

/* MaxPool2d(3, s2, p1) on NHWC 16-bit (dinov3_adapter.py:250). */
int b2u_maxpool3x3s2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t dtype,
                     b2u_stream_t stream);

/* Depthwise 3x3 (pad 1) + bias (+GELU) on NHWC 16-bit.  `planes` > 1 processes the ConvFFN token stream
 * [B, 21n, C] as three planes (2h x 2w, h x w, h/2 x w/2) with shared weights (dinov3_adapter.py:99-109);
 * planes == 1 is a plain [B,H,W,C] image (dinounet_training.py:241-243).  w is tap-major [9, C] fp32. */
int b2u_dwconv3x3(const void* in, void* out, const float* w, const float* bias, int32_t B, int32_t H, int32_t W,
                  int32_t C, int32_t planes, int32_t act, int32_t dtype, b2u_stream_t stream);

/* MultiScaleDeformableAttention forward, single level, n_heads x n_points, fused prologue
 * (ms_deform_attn.py:185-213 + ms_deform_im2col_cuda.cuh:242-304):
 *   value  [B, Hv*Wv, heads, dh] 16-bit;  offaw fp32 [B*Lq, heads*points*3] = (offsets(x,y) | attention logits);
 *   reference point of query q = cell centre of q in the pyramid of (2Hv x 2Wv, Hv x Wv, Hv/2 x Wv/2) grids;
 *   loc = ref + off / (Wv, Hv); weights = softmax over points; bilinear, zero padding, align_corners=False.
 *   out [B*Lq, heads*dh] 16-bit. */
int b2u_msda_forward(const void* value, const float* offaw, void* out, int32_t B, int32_t Hv, int32_t Wv,
                     int32_t heads, int32_t dh, int32_t points, int32_t dtype, b2u_stream_t stream);

/* Drop-in for the reference pybind op `ms_deform_attn_forward` (ops/src/vision.cpp:18, ms_deform_attn.h:26-45):
 * value fp32 [B, S, heads, dh]; spatial_shapes int64 [levels, 2] (H, W) and level_start_index int64 [levels] on the
 * DEVICE (as in the reference); sampling locations fp32 [B, Lq, heads, levels, points, 2]; attention weights fp32
 * [B, Lq, heads, levels, points]; out fp32 [B, Lq, heads*dh].  (im2col_step is a launch-chunking detail of the
 * reference kernel and has no equivalent here.) */
int b2u_msda_forward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                         const float* loc, const float* attw, float* out, int32_t B, int32_t S, int32_t Lq,
                         int32_t heads, int32_t dh, int32_t levels, int32_t points, b2u_stream_t stream);

/* Drop-in for the reference pybind op `ms_deform_attn_backward` (ops/src/vision.cpp:19, ms_deform_attn.h:48-67; the one
 * native kernel the reference's training step runs, ms_deform_attn.py:58-66).  Same inputs as the forward plus
 * grad_out fp32 [B, Lq, heads*dh]; writes grad_value [B, S, heads, dh] (zeroed here, then accumulated),
 * grad_loc [B, Lq, heads, levels, points, 2] and grad_attw [B, Lq, heads, levels, points], all fp32, caller-owned.
 * grad_value is accumulated with fp32 atomics (as in the reference kernel): summation order is not fixed. */
int b2u_msda_backward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const float* loc, const float* attw, const float* grad_out, float* grad_value,
                          float* grad_loc, float* grad_attw, int32_t B, int32_t S, int32_t Lq, int32_t heads,
                          int32_t dh, int32_t levels, int32_t points, b2u_stream_t stream);

/* ---- Sliding-window prediction (inference/predict_from_raw_data.py:537-621, sliding_window_prediction.py:11-56).
 * tile_desc: DEVICE int32 [n_entries][4] = (slice d, y0, x0, flip bits: 1 = rows / tensor dim 2, 2 = cols / dim 3).
 * b2u_sw_gather_tiles: volume fp32 [Cin, D, H, W] -> batch fp32 [n_entries, 3, th, tw]; crops, mirrors and applies the
 *   1/2/>3 -> 3 channel rule of DINOv3EncoderAdapter.forward (dinounet_training.py:491-497).
 * b2u_sw_accumulate: one tile whose mirror variants are logits[first .. first+nvar) (fp32 [n, C, th, tw], variant 0
 *   unflipped, nvar in {1,2,4}): un-mirror, mean over variants, * gaussian (fp16 [th, tw], NULL = 1), accumulate into
 *   acc fp16 [C, D, H, W] and npred fp16 [D, H, W] with the reference's fp16 rounding sequence (:545-551, :598-599).
 *   Launches for overlapping tiles must be stream-ordered.
 * b2u_sw_finalize: acc /= npred (:601); *inf_flag |= 1 if any result is inf (:603-606 raises). */
int b2u_sw_gather_tiles(const float* volume, float* batch, const int32_t* tile_desc, int32_t n_entries, int32_t Cin,
                        int32_t D, int32_t H, int32_t W, int32_t th, int32_t tw, b2u_stream_t stream);
int b2u_sw_accumulate(const float* logits, const int32_t* tile_desc, int32_t first, int32_t nvar, const void* gaussian,
                      void* acc, void* npred, int32_t C, int32_t D, int32_t H, int32_t W, int32_t th, int32_t tw,
                      b2u_stream_t stream);
int b2u_sw_finalize(void* acc, const void* npred, int32_t C, int64_t plane, int32_t* inf_flag, b2u_stream_t stream);

/* ---- Dice + cross-entropy loss and online-validation statistics on the logits of the forward path
 * (training/loss/compound_losses.py:31-56 `DC_and_CE_loss`, dice.py:72-119 `MemoryEfficientSoftDiceLoss`,
 * robust_ce_loss.py:12-16, built at nnUNetTrainer.py:363-365; consumer nnUNetTrainer.py:961-1005 `validation_step`).
 * logits fp32 [B, C, plane] (plane = H*W); target [B, plane] labels, target_kind 0 = uint8, 1 = int32, 2 = int64,
 * 3 = float32 (nnU-Net hands float label maps).  work: caller-owned, b2u_dice_ce_work_doubles(B, C, plane) doubles.
 * forward: out3[0] = weight_ce*CE + weight_dice*dice, out3[1] = CE (mean over all pixels), out3[2] = dice term
 *   (-mean_c (2*sum(p*y)+smooth)/max(sum(y)+sum(p)+smooth, 1e-8); over the batch if batch_dice, background class skipped
 *   unless do_bg);  tp_fp_fn (optional) int64 [3][C] = hard tp / fp / fn of argmax(logits) vs target, summed over the
 *   batch (get_tp_fp_fn_tn with axes (0,2,3), nnUNetTrainer.py:971-991);  *bad_label is set to 1 if a label is outside
 *   [0, C) (torch raises a device assert there).  Deterministic (fixed reduction order, fp64 partials).
 * backward: grad_logits fp32 [B, C, plane] = grad_scale * dLoss/dlogits, from the sums the forward left in `work`.
 * 2 <= C <= 32.  ignore_label / region (BCE) training are not covered. */
int64_t b2u_dice_ce_work_doubles(int32_t B, int32_t C, int64_t plane);
int b2u_dice_ce_forward(const float* logits, const void* target, int32_t target_kind, double* work, float* out3,
                        int64_t* tp_fp_fn, int32_t* bad_label, int32_t B, int32_t C, int64_t plane, float weight_ce,
                        float weight_dice, int32_t batch_dice, int32_t do_bg, float smooth, b2u_stream_t stream);
int b2u_dice_ce_backward(const float* logits, const void* target, int32_t target_kind, const double* work,
                         float* grad_logits, int32_t B, int32_t C, int64_t plane, float weight_ce, float weight_dice,
                         int32_t batch_dice, int32_t do_bg, float smooth, float grad_scale, b2u_stream_t stream);

/* Adapter tail (dinov3_adapter.py:467-482): out[b,y,x,:] = BN_eval( base[b,y,x,:] + bilinear(tap[b,:,:,:] -> HxW) ).
 * base: fp32 token stream slice or 16-bit image (base_fp32), tap fp32 [B, Ht*Wt, D] (token-major), align_corners=False. */
int b2u_tail_fuse(const void* base, int32_t base_fp32, int64_t base_batch_stride, const float* tap, void* out,
                  const float* scale, const float* shift, int32_t B, int32_t H, int32_t W, int32_t Ht, int32_t Wt,
                  int32_t D, int32_t dtype, b2u_stream_t stream);

/* InstanceNorm statistics: sums[b, c, 0..1] = (sum x, sum (x - mean)^2) over the rows of image b (accumulated as shifted
 * sums around the image's first pixel: no cancellation when |mean| >> std).  x 16-bit [B*rows, C] with row
 * stride ldx (dinounet_training.py:400-401; ConvDropoutNormReLU norm).  Deterministic two-level reduction (no float
 * atomics): `work` holds >= b2u_in_stats_work_floats(B, rows, C) floats whose first B words are ticket counters that
 * must be ZERO before the first call (the kernel leaves them zero); a workspace may be shared by consecutive calls. */
int64_t b2u_in_stats_work_floats(int32_t B, int32_t rows, int32_t C);
int b2u_in_stats(const void* x, int64_t ldx, float* sums, float* work, int32_t B, int32_t rows, int32_t C, int32_t dtype,
                 b2u_stream_t stream);

/* InstanceNorm apply + LeakyReLU(0.01): y = lrelu((x - mean) * rstd * gamma + beta), biased variance, eps. */
int b2u_in_apply(const void* x, int64_t ldx, void* y, int64_t ldy, const float* sums, const float* gamma,
                 const float* beta, int32_t B, int32_t rows, int32_t C, float eps, int32_t dtype, b2u_stream_t stream);

/* FiLM: z[r, j] = gb[r, j] * zz[r, zoff + j] + gb[r, R + j]  (dinounet_training.py:427-429). */
int b2u_film(const void* gb, const void* zz, int64_t ldzz, int32_t zoff, void* z, int32_t rows, int32_t R,
             int32_t dtype, b2u_stream_t stream);

/* Squeeze-excitation gate + shortcut (dinounet_training.py:222-225,437-439):
 *   pooled[b,c] = mean_rows t ; g = sigmoid(W2 relu(W1 pooled + b1) + b2) ; out = t * g + sc. */
int b2u_se_gate(const float* sums, const float* w1, const float* b1, const float* w2, const float* b2, float* gate,
                int32_t B, int32_t C, int32_t Cr, int32_t rows, b2u_stream_t stream);
int b2u_se_apply(const void* t, const void* sc, int64_t ldsc, const float* gate, void* out, int32_t B, int32_t rows,
                 int32_t C, int32_t dtype, b2u_stream_t stream);

/* Final InstanceNorm + LeakyReLU + 1x1 segmentation head + argmax (dinounet_training.py:597,619; nnUNetTrainer.py:977):
 *   x 16-bit [B*rows, C]; logits NCHW fp32 [B, ncls, rows]; labels uint8 [B, rows] (may be NULL). */
int b2u_seg_head(const void* x, const float* sums, const float* gamma, const float* beta, float eps, const float* w,
                 const float* b, float* logits, uint8_t* labels, int32_t B, int32_t rows, int32_t C, int32_t ncls,
                 int32_t dtype, b2u_stream_t stream);

/* cudaMemsetAsync(ptr, 0, bytes) on the stream (statistics accumulators must be zeroed every forward). */
int b2u_zero(void* ptr, int64_t bytes, b2u_stream_t stream);

/* Tuning / A-B switches.  key 0 (B2U_OPT_GEMM_IMPL): 0 = persistent 128x256-tile tcgen05 kernel (default),
 * 1 = the first-generation one-tile-per-CTA kernel. */
/* key 1 (B2U_OPT_MSDA_IMPL): 0 = shared-memory value-slab gather (default), 1 = first-generation warp-per-query kernel. */
/* key 2 (B2U_OPT_CONV_HALO): 0 = 3x3 convs with <= 64 in/out channels and >= 128-px rows use the halo-reuse mode
 * (default), 1 = always the per-tap TMA walk. */
/* key 3 (B2U_OPT_GEMM_PAIR): 0 = plain GEMMs with 256-wide tiles run on CTA pairs (tcgen05 cta_group::2, 256 x 256
 * tile per pair, each CTA stages half of the weight tile) (default), 1 = one CTA per tile. */
/* key 4 (B2U_OPT_ATTN_SPLIT): 0 = two softmax warps per (query tile, TMEM lane quarter), each owning half of the key
 * columns / head dims (default), 1 = one warp. */
enum { B2U_OPT_GEMM_IMPL = 0, B2U_OPT_MSDA_IMPL = 1, B2U_OPT_CONV_HALO = 2, B2U_OPT_GEMM_PAIR = 3, B2U_OPT_ATTN_SPLIT = 4 };
/* ---------------------------------------------------------------------------------------------------------------
 * fp32 parity tier (csrc/fp32_tier.cu): the same forward in IEEE fp32 with plain SIMT kernels, selected by
 * `DinoUNet.precision = "fp32"`.  It is the path held to north_star's 1e-5 against the reference's fp32 (CPU) forward;
 * it is not the benchmarked path.  All tensors fp32, same layouts as above (tokens [rows, C], images NHWC).
 * Epilogue of b2u_f32_gemm (same meaning as b2u_epilogue): v = acc + bias[n]; v = act1(v); v = v*scale[n] + shift[n];
 * v = act2(v); v += residual[orow, ocol]; out[orow, ocol] = v.  conv != 0: A is an NHWC image, W is [N, 9*Cpad].
 * A-row remap (a_rows_in > 0): arow = (m / a_rows_in) * a_rows_out + a_row_off + m % a_rows_in. */
typedef struct b2u_f32_gemm_params {
  int64_t M;
  int32_t N, K;
  const float* A;
  int64_t lda;
  int32_t a_rows_in, a_rows_out, a_row_off;
  const float* W;
  int64_t ldw;
  int32_t conv, Hin, Win, C, Cpad;
  float* out;
  int64_t ldc;
  int32_t col_off;
  int32_t rows_in, rows_out, row_off;
  int32_t ps_cout, ps_h, ps_w;
  const float* bias;
  const float* scale;
  const float* shift;
  int32_t act1, act2;
  const float* residual;
  int64_t ldres;
  /* backward-pass addressing (csrc/train_bwd.cu callers): a_trans: A'(m, k) = A[k][m]; w_mode 1: W'(n, k) = W[k][n];
   * w_mode 2 (with conv != 0): 3x3 data gradient, W'(c, (tap', n)) = W[n][(8 - tap') * w_cpad + c];
   * w_mode 3 (with conv != 0, a_trans): 3x3 weight gradient, W is the layer's NHWC input, W'((tap, c), pixel) = its window;
   * ksplit > 1: the K range is split over gridDim.z and the raw products are atomically ADDED to out (zero it first);
   * accumulate: out += instead of out =. */
  int32_t a_trans, w_mode, w_cpad, ksplit, accumulate;
} b2u_f32_gemm_params;
int b2u_f32_gemm(const b2u_f32_gemm_params* p, b2u_stream_t stream);   /* F.linear / Conv2d 1x1, 3x3 / ConvTranspose2d k2 s2 */
/* Same parameter block, addressing modes and epilogue on the tensor cores (csrc/gemm_tf32.cu): operands rounded to TF32
 * (round-to-nearest, 10 mantissa bits = the fp16 autocast the reference trains under, nnUNetTrainer.py:899-929, with fp32's
 * exponent range), fp32 accumulation in tensor memory (tcgen05.mma kind::tf32).  The fast tier of the TRAINING step's
 * matrix products (forward, data gradient, weight gradient of F.linear / Conv2d / ConvTranspose2d); not bit-comparable
 * with b2u_f32_gemm (about 1e-3 relative per product). */
int b2u_tf32_gemm(const b2u_f32_gemm_params* p, b2u_stream_t stream);
/* F.layer_norm; input row = (r / out_per_b) * in_per_b + in_off + r % out_per_b when in_per_b > 0 (ViT taps drop the prefix) */
int b2u_f32_layernorm(const float* in, float* out, const float* w, const float* b, int64_t rows, int32_t D, float eps,
                      int32_t in_per_b, int32_t out_per_b, int32_t in_off, b2u_stream_t stream);
int b2u_f32_patchify(const float* x_nchw, float* rows768, int32_t B, int32_t S, b2u_stream_t stream);   /* patch_embed.py:64-76 */
int b2u_f32_nchw_to_nhwc(const float* x, float* out, int32_t B, int32_t C, int64_t HW, b2u_stream_t stream);
int b2u_f32_seg_out(const float* lin, float* logits_nchw, uint8_t* labels, int32_t B, int64_t HW, int32_t ncls, b2u_stream_t stream);
int b2u_f32_maxpool3x3s2(const float* in, float* out, int32_t B, int32_t H, int32_t W, int32_t C, b2u_stream_t stream);
int b2u_f32_dwconv3x3(const float* in, float* out, const float* w9, const float* bias, int32_t B, int32_t H, int32_t W,
                      int32_t C, int32_t planes, int32_t act, b2u_stream_t stream);              /* dinov3_adapter.py:99-109 */
/* attention.py:66-118 on the fused qkv rows [B*N, 3D]: RoPE on tokens >= prefix, exact softmax, fp32 */
int b2u_f32_attention(const float* qkv, const float* rope_sin, const float* rope_cos, float* out, int32_t B, int32_t N,
                      int32_t heads, int32_t head_dim, int32_t prefix, float scale, b2u_stream_t stream);
int b2u_f32_msda(const float* value, const float* offaw, float* out, int32_t B, int32_t Hv, int32_t Wv, int32_t heads,
                 int32_t dh, b2u_stream_t stream);                                                 /* ms_deform_attn.py:158-216 */
/* InstanceNorm2d(affine) (+ LeakyReLU 0.01): work = 2*B*C doubles (scratch), stats = [B][C][2] floats (mean, rstd) out */
int b2u_f32_instnorm(const float* in, int64_t ld_in, float* out, int64_t ld_out, const float* w, const float* b, double* work,
                     float* stats, int32_t B, int64_t HW, int32_t C, float eps, int32_t lrelu, b2u_stream_t stream);
int b2u_f32_se(const float* t, const float* shortcut, int64_t ld_shortcut, float* pooled_work, const float* w1, const float* b1,
               const float* w2, const float* b2, float* out, int32_t B, int64_t HW, int32_t C, int32_t hidden, b2u_stream_t stream);
int b2u_f32_film(const float* gamma_beta, const float* zs_zp, float* z, int64_t px, int32_t R, b2u_stream_t stream);
int b2u_f32_tail(const float* c, int64_t c_rows_per_b, int64_t c_off, const float* tap, float* out, const float* bn_scale,
                 const float* bn_shift, int32_t B, int32_t r, int32_t h, int32_t D, b2u_stream_t stream);   /* dinov3_adapter.py:468-482 */

/* ---------------------------------------------------------------------------------------------------------------
 * Backward of the trainable part + optimizer (csrc/train_bwd.cu; SURVEY.md section 8f rank 2, BASELINE config 3).
 * fp32, same layouts as the fp32 tier.  Matrix-product gradients use b2u_f32_gemm (a_trans / w_mode / ksplit).  Every
 * parameter-gradient output is ACCUMULATED with atomics: zero it first.  Replaces ATen autograd kernels of
 * nnUNetTrainer.train_step (nnUNetTrainer.py:899-929); the one native backward op of the reference is b2u_msda_backward_f32. */
int b2u_f32_add(const float* a, const float* b, float* out, int64_t n, b2u_stream_t stream);
int b2u_f32_act(const float* x, float* y, int64_t n, int32_t act, b2u_stream_t stream);
int b2u_f32_act_bwd(const float* x_pre, const float* dy, float* dx, int64_t n, int32_t act, b2u_stream_t stream);
int b2u_f32_colsum(const float* in, int64_t ld, int64_t rows, int32_t C, float* out_accum, b2u_stream_t stream);
int b2u_f32_layernorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int64_t rows,
                          int32_t D, float eps, b2u_stream_t stream);
int b2u_f32_instnorm_bwd(const float* x, int64_t ldx, const float* dy, int64_t lddy, const float* w, const float* b,
                         const float* stats, double* work, float* dx, int64_t lddx, float* dw, float* db, int32_t B, int64_t HW,
                         int32_t C, int32_t lrelu, b2u_stream_t stream);
/* eval-mode (Sync)BatchNorm + optional ReLU, forward and backward (the gradient oracle's semantics) */
int b2u_f32_bn_act(const float* x, float* y, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float eps, int64_t rows, int32_t C, int32_t act, b2u_stream_t stream);
int b2u_f32_bn_act_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                       int32_t act, b2u_stream_t stream);
int b2u_f32_dwconv_wgrad(const float* x, const float* dy, float* dw9, float* db, int32_t B, int32_t H, int32_t W, int32_t C,
                         int32_t planes, b2u_stream_t stream);
int b2u_f32_maxpool3x3s2_bwd(const float* x, const float* dy, float* dx_zeroed, int32_t B, int32_t H, int32_t W, int32_t C,
                             b2u_stream_t stream);
int b2u_f32_film_bwd(const float* gamma_beta, const float* zs_zp, const float* dz, float* dgb, float* dzz, int64_t px, int32_t R,
                     b2u_stream_t stream);
int b2u_f32_se_bwd(const float* t, const float* dy, const float* pooled, const float* w1, const float* b1, const float* w2,
                   const float* b2, float* work_3BC, float* dt, float* dw1, float* db1, float* dw2, float* db2, int32_t B,
                   int64_t HW, int32_t C, int32_t hidden, b2u_stream_t stream);
int b2u_f32_msda_prep(const float* offaw, float* loc, float* attw, int32_t B, int32_t Hv, int32_t Wv, int32_t heads, b2u_stream_t stream);
int b2u_f32_msda_prep_bwd(const float* attw, const float* dloc, const float* dattw, float* doffaw, int32_t B, int32_t Hv, int32_t Wv,
                          int32_t heads, b2u_stream_t stream);
int b2u_f32_unshuffle(const float* dy_image, int64_t ld, int32_t col_off, float* rows_4cout, int32_t B, int32_t h, int32_t w,
                      int32_t Cout, b2u_stream_t stream);
int b2u_f32_conv3x3_dgrad(const float* dy, const float* W, float* dx, int32_t B, int32_t H, int32_t Wd, int32_t C, int32_t Cpad,
                          int32_t N, int32_t stride, b2u_stream_t stream);
int b2u_f32_conv3x3_wgrad(const float* x, const float* dy, float* dW_accum, int32_t B, int32_t H, int32_t Wd, int32_t C, int32_t Cpad,
                          int32_t N, int32_t stride, b2u_stream_t stream);
int b2u_f32_sqsum(const float* g, int64_t n, double* out_accum, b2u_stream_t stream);
int b2u_f32_sgd_nesterov(float* p, const float* g, float* momentum_buf, int64_t n, float lr, float momentum, float weight_decay,
                         const double* grad_sqsum, float max_norm, int32_t first_step, b2u_stream_t stream);

int b2u_set_option(int32_t key, int32_t value);

const char* b2u_last_error(void);
int b2u_version(void);
/* Number of kernel launches issued through this library since load (bench.py's gpu_launches evidence). */
int64_t b2u_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DINOUNET_B200_H_ */
