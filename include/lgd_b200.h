/* liblgd_b200 -- C ABI of the B200-native LGD distillation hot path.
 *
 * Every entry point replaces a piece of the reference's PyTorch hot path (paths relative to
 * megvii-research/LGD; see SURVEY.md section 8(a) for the full map):
 *   models/customized_detectors/dynamic_teacher/label_encoder.py      (a1, a2)
 *   models/customized_detectors/dynamic_teacher/spatial_transformer.py (a2)
 *   models/customized_detectors/dynamic_teacher/utils.py:53-89         (a4 get_inside_gt_mask)
 *   models/customized_detectors/dynamic_teacher/dynamic_teacher.py     (a3, a5..a9)
 *   models/adapters/sequential_convs.py:7-15                            (a10)
 *   models/base_distillator.py:34-64                                    (a11)
 *
 * Conventions
 *   - all pointers are DEVICE pointers to fp32 / int32 unless the name ends in _host;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing, keeps no global mutable state and returns 0 on success or a negative LGD_E* code;
 *     lgd_last_error() gives the message for the calling thread;
 *   - "pyramid buffer": one fp32 buffer holding all FPN levels, level l at element offset
 *     256*B*sum_{j<l} h_j*w_j, laid out [B][h_l][w_l][256] (NHWC). Shapes travel in lgd_pyramid_t;
 *   - box table: (T,4) clamped XYXY boxes of all images back to back, CSR offsets img_start[B+1];
 *   - ranges: int32 (F,T,4) = {x_lo, x_hi, y_lo, y_hi} half-open pixel intervals per level and box,
 *     the exact solution set of the reference's fp32 membership test (utils.py:67-88).
 */
#ifndef LGD_B200_H_
#define LGD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGD_MAX_LEVELS 8
#define LGD_CHANNELS 256
#define LGD_DESC_DIM 84 /* 4 box coords + 80 one-hot classes, label_encoder.py:25-28 */

#define LGD_OK 0
#define LGD_EINVAL (-1)  /* bad argument / shape */
#define LGD_ECUDA (-2)   /* CUDA runtime or driver error */
#define LGD_ENOSUP (-3)  /* unsupported device (needs sm_100) */

typedef struct lgd_pyramid {
  int32_t num_levels;
  int32_t batch;
  int32_t h[LGD_MAX_LEVELS];
  int32_t w[LGD_MAX_LEVELS];
} lgd_pyramid_t;

int lgd_version(void);
const char* lgd_last_error(void);
/* number of CUDA kernels this library has launched in this process so far (diagnostics only) */
int64_t lgd_launch_count(void);
/* number of fp32 elements of one pyramid buffer */
int64_t lgd_pyramid_elems(const lgd_pyramid_t* pyr);

/* ---- a1: box_descriptor_encode (label_encoder.py:88-112): boxes/W,H | one-hot -> 2x-1 ---- */
/* labels[t] in [0,80) or -1 for the context / dummy row (all-zero one-hot). desc: (T,84). */
int lgd_encode_descriptors(const float* boxes, const int32_t* labels, int T, int img_h, int img_w, float* desc,
                           void* stream);

/* ---- a2/a3/a6/a7 building blocks for the per-instance ("small-T") network ---- */
/* The three products of one linear layer. workspace (optional, may be NULL): scratch for the deterministic split-K
 * path that keeps all SMs busy although M = T is only ~10^2 rows; lgd_linear_workspace() bytes are always enough. */
size_t lgd_linear_workspace(int M, int N, int K);
/* y[M,N] = x[M,K] * w[N,K]^T + bias[N]   (nn.Linear / Conv1d(k=1), label_encoder.py:149-155) */
int lgd_linear_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy, int M, int N,
                   int K, void* workspace, size_t workspace_bytes, void* stream);
/* gx[M,K] (+)= gy[M,N] * w[N,K] */
int lgd_linear_bwd_input(const float* gy, int ldgy, const float* w, int ldw, float* gx, int ldgx, int M, int N, int K,
                         int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* gw[N,K] (+)= gy[M,N]^T * x[M,K];  gb[N] (+)= sum_m gy[m,n] */
int lgd_linear_bwd_weight(const float* gy, int ldgy, const float* x, int ldx, float* gw, int ldgw, float* gb, int M,
                          int N, int K, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* LayerNorm over the last dim (no affine, eps 1e-5, biased var) + optional ReLU; saves mean/rstd (M each). */
int lgd_layernorm_fwd(const float* x, float* y, float* mean, float* rstd, int M, int N, int relu, void* stream);
/* backward of y = relu?(LN(x)) from the saved input x and row statistics. gx may alias gy. */
int lgd_layernorm_bwd(const float* gy, const float* x, const float* mean, const float* rstd, float* gx, int M, int N,
                      int relu, void* stream);
/* y[t,j] = sum_i x[t,i] * mats[t,i,j]   (torch.bmm with the STN transform, label_encoder.py:241,248) */
int lgd_rowvec_matmul_fwd(const float* x, const float* mats, float* y, int T, int k, void* stream);
/* gx[t,i] = sum_j gy[t,j]*mats[t,i,j];  gmats[t,i,j] = x[t,i]*gy[t,j] */
int lgd_rowvec_matmul_bwd(const float* gy, const float* x, const float* mats, float* gx, float* gmats, int T, int k,
                          void* stream);
/* per-image max over instances (hier_pool, label_encoder.py:195-213) and its broadcast-concat:
 * out[t, 0:c_local] = local[t], out[t, c_local:] = max_{s in image(t)} x[s]; argmax saved (B,C). */
int lgd_segmax_concat_fwd(const float* local, int c_local, const float* x, int C, const int32_t* img_start, int B,
                          float* out, int32_t* argmax, void* stream);
int lgd_segmax_concat_bwd(const float* gout, int c_local, int C, const int32_t* img_start, int B,
                          const int32_t* argmax, float* glocal, float* gx, void* stream);
/* block-diagonal multi-head attention core (nn.MultiheadAttention, dynamic_teacher.py:255-275).
 * q: (nsets_q*T, E) already projected, k/v: (nsets_kv*T, E); set l of the output uses q set
 * (nsets_q>1 ? l : 0) and kv set (nsets_kv>1 ? l : 0). Row t attends to the rows of its own image only.
 * out: (F*T, E). probs: (F, heads, T, max_n) saved for backward. */
int lgd_attention_fwd(const float* q, int nsets_q, const float* k, const float* v, int nsets_kv, int F, int T,
                      int heads, int E, const int32_t* img_of, const int32_t* img_start, int max_n, float* out,
                      float* probs, void* stream);
/* gq: (F,T,E) (when nsets_q == 1 the sum over levels is returned in its first T*E floats); gk, gv: (nsets_kv*T, E);
 * gs_scratch: same size as probs. */
int lgd_attention_bwd(const float* gout, const float* q, int nsets_q, const float* k, const float* v, int nsets_kv,
                      int F, int T, int heads, int E, const int32_t* img_of, const int32_t* img_start, int max_n,
                      const float* probs, float* gs_scratch, float* gq, float* gk, float* gv, void* stream);

/* ---- a4: get_inside_gt_mask (utils.py:53-89), bit exact ---- */
int lgd_box_ranges(const float* boxes, int T, int img_h, int img_w, const lgd_pyramid_t* pyr, int32_t* ranges,
                   void* stream);
/* materialise the reference's float masks: level l at offset T*sum_{j<l} h_j*w_j, row t = (h_l*w_l) floats */
int lgd_masks_from_ranges(const int32_t* ranges, int T, const lgd_pyramid_t* pyr, float* masks, void* stream);

/* ---- layout movers between detectron2's NCHW maps and the NHWC pyramid buffer ---- */
/* src_levels_host: host array of num_levels device pointers to contiguous (B,256,h,w) fp32 */
/* dst_half (optional): fp16 pyramid copy of the same values, the operand of the fp16 forward convolutions */
int lgd_nchw_to_pyramid(const float* const* src_levels_host, const lgd_pyramid_t* pyr, float* dst, int round_tf32,
                        void* dst_half, void* stream);
int lgd_pyramid_to_nchw(const float* src, const lgd_pyramid_t* pyr, float* const* dst_levels_host, int accumulate,
                        void* stream);
/* channels_last boundary (retinanet.py:52-59 with the FPN in channels_last memory format): the per-level maps are
 * already NHWC in memory, (B,h,w,256) contiguous fp32 each; no transposition, one streaming pass that gathers the
 * levels into the pyramid buffer as fp32 (dst, optional) and / or as the fp16 conv operand (dst_half, optional). */
int lgd_nhwc_to_pyramid(const float* const* src_levels_host, const lgd_pyramid_t* pyr, float* dst, void* dst_half,
                        void* stream);

/* ---- K1: 3x3 / stride 1 / pad 1 / 256->256 convolution on tcgen05 CTA pairs (TF32 operands, fp32 accumulate) ----
 * weights: mode 0 (forward)  packed[tap][co][ci] = tf32(w[co][ci][ky][kx]), tap = ky*3+kx
 *          mode 1 (dgrad)    packed[tap][ci][co] = tf32(w[co][ci][2-ky][2-kx])
 *          mode 2 / 3        the same layouts holding the residual tf32(w - tf32(w)) (split-operand forward)   */
int lgd_pack_conv_weight(const float* w, float* packed, int mode, void* stream);
int lgd_unpack_conv_wgrad(const float* packed_grad, float* gw, int accumulate, void* stream);
int lgd_conv3x3_num_tiles(const lgd_pyramid_t* pyr);
/* out = conv(in) + bias[(l*bias_level_stride + b*bias_image_stride) + c]; optional ReLU; optional
 * relu_mask (same layout as out): out = mask>0 ? out : 0; optional TF32 rounding of the stored value;
 * tile_stats (num_tiles,2) receives per-tile sum / sum of squares of the un-rounded, pre-ReLU output.
 * Optional by-products (either may be NULL; need lgd_conv3x3_fwd_workspace() bytes of workspace): chan_sums (F,B,256) =
 * per-(level,image) channel sums of the un-rounded stored values, chan_total (256) = their sum -- with mode-1 weights
 * and a relu_mask this is the bias gradient of the convolution in front (dgrad through a ReLU). */
size_t lgd_conv3x3_fwd_workspace(const lgd_pyramid_t* pyr);
int lgd_conv3x3_fwd(const lgd_pyramid_t* pyr, const float* in, const float* packed_w, const float* bias,
                    int bias_level_stride, int bias_image_stride, float* out, int relu, int round_out,
                    const float* relu_mask, float* tile_stats, float* chan_sums, float* chan_total, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Split-operand ("tf32x3", fp32-accurate) convolution: out = conv(in) + addend, then bias / ReLU / relu_mask /
 * statistics / channel sums exactly as lgd_conv3x3_fwd. With x = x_hi + x_lo from lgd_tf32_split and mode-0 / mode-2
 * weights (mode 1 / 3 for a dgrad), three chained launches
 *   t = conv(x_lo, w_hi);  t = conv(x_hi, w_lo) + t;  out = conv(x_hi, w_hi) + t + bias
 * reproduce the fp32 convolution to ~4e-6 (the first through lgd_conv3x3_fwd). addend has the layout of out and may
 * alias it. Parity-verification mode: it removes the ReLU-mask flips that any 10-bit-mantissa forward shows against
 * an fp32 reference at 3x the tensor work. */
int lgd_conv3x3_fwd_addend(const lgd_pyramid_t* pyr, const float* in, const float* packed_w, const float* addend,
                           const float* bias, int bias_level_stride, int bias_image_stride, float* out, int relu,
                           int round_out, const float* relu_mask, float* tile_stats, float* chan_sums,
                           float* chan_total, void* workspace, size_t workspace_bytes, void* stream);
/* Forward convolution with fp16 operands (fp32 accumulate): same 10-bit mantissa as TF32 at twice the MMA rate and half
 * the operand bytes. in_half: pyramid buffer of __half (same element offsets as the fp32 layout); packed_w_half from
 * lgd_pack_conv_weight_f16 ([tap][co][ci] __half). out: fp32 pyramid (optionally TF32-rounded); out_half (optional):
 * fp16 copy of the stored values for the next forward convolution. Used for the forward direction only: activations
 * are O(1) after the norms, gradients (unbounded dynamic range) stay on the TF32 path. */
/* mode 0: forward layout [tap][co][ci]; mode 1: dgrad layout [tap][ci][co], taps flipped. gain (optional device
 * scalar; needs 9*256 floats of workspace): sum over the taps of ||W_tap||_F, an upper bound of the l2 operator norm
 * of the convolution and of its transpose -- it bounds the norm of an fp16 dgrad's output before it is computed. */
int lgd_pack_conv_weight_f16(const float* w, void* packed_half, int mode, float* gain, void* workspace,
                             size_t workspace_bytes, void* stream);
/* The same for up to 8 convolutions in one launch (+ one for the gains): w_host / fwd_host / dgrad_host are host arrays
 * of n device pointers (fwd_host, dgrad_host or single entries may be NULL = layout not needed); gains (optional,
 * n floats; needs n*9*64 floats of workspace) as above. */
int lgd_pack_conv_weights_f16_multi(const float* const* w_host, int n, void* const* fwd_host, void* const* dgrad_host,
                                    float* gains, void* workspace, size_t workspace_bytes, void* stream);
int lgd_conv3x3_fwd_f16(const lgd_pyramid_t* pyr, const void* in_half, const void* packed_w_half, const float* bias,
                        int bias_level_stride, int bias_image_stride, float* out, void* out_half, int relu,
                        int round_out, float* tile_stats, void* stream);
/* Input gradient on fp16 operands: gout_half = fp16(gout * s) with a power-of-two s chosen by the producer of gout so
 * that nothing overflows (||gout||_2 * s <= 2^14, see lgd_grad_scale); acc_scale = device scalar 1/s applied to the
 * fp32 accumulator. relu_mask / tile_stats / chan_sums / chan_total / round_out as lgd_conv3x3_fwd. out_half
 * (optional): fp16(stored value * half_scale[0]), saturated -- the operand of the next fp16 dgrad. */
int lgd_conv3x3_dgrad_f16(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                          const float* acc_scale, float* out /* optional when out_half is given */, int round_out,
                          const float* relu_mask, const void* relu_mask_half /* the mask as the activation's fp16 copy */,
                          void* out_half, const float* half_scale, float* tile_stats, float* chan_sums,
                          float* chan_total, void* workspace, size_t workspace_bytes, void* stream);
/* packed_grad[tap][co][ci] = sum_pixels gout[p][co] * in[p+tap][ci]; gbias[co] = sum gout.
 * workspace: lgd_conv3x3_wgrad_workspace() bytes. */
size_t lgd_conv3x3_wgrad_workspace(const lgd_pyramid_t* pyr);
int lgd_conv3x3_wgrad(const lgd_pyramid_t* pyr, const float* in, const float* gout, float* packed_grad, float* gbias,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The same on fp16 operands: in_half = the fp16 copy of the convolution's input that the forward producers wrote,
 * gout_half = fp16(gout * s) from the gradient's producer, inv_scale = device scalar 1/s (applied in the deterministic
 * second-stage reduction). */
int lgd_conv3x3_wgrad_f16(const lgd_pyramid_t* pyr, const void* in_half, const void* gout_half, const float* inv_scale,
                          float* packed_grad, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2: GroupNorm(1 group, no affine) statistics + apply (layers.py:6-7) ---- */
/* stats: (F,B,2) = {mean, rstd} over (C,h,w) of each image and level */
int lgd_gn_finalize(const lgd_pyramid_t* pyr, const float* tile_stats, float* stats, void* stream);
/* in_stats (optional, (F,B,256,2) like lgd_in_stats): InstanceNorm statistics of the STORED y, from the same pass
 * (the teacher pyramid is normalised again per channel by the distillation loss, base_distillator.py:60);
 * needs lgd_gn_apply_workspace() bytes of workspace. */
size_t lgd_gn_apply_workspace(const lgd_pyramid_t* pyr);
int lgd_gn_apply(const lgd_pyramid_t* pyr, const float* x, const float* stats, float* y, int relu, int round_out,
                 void* y_half /* optional fp16 copy */, float* in_stats, void* workspace, size_t workspace_bytes,
                 void* stream);
/* gx = rstd*(g - mean(g) - xhat*mean(g*xhat)) with g = relu ? gy*(y>0) : gy ; two-pass (sums, then apply).
 * Optional by-products from the same pass (either may be NULL), computed from the un-rounded gx: chan_sums (F,B,256) =
 * per-(level,image) channel sums, chan_total (256) = their sum = bias gradient of the convolution in front.
 * gx_half + scale3 (optional, both or neither): scaled fp16 copy of the un-rounded gx for lgd_conv3x3_dgrad_f16 and
 * the device triple {s, 1/s, U} it was written with: U = sqrt(sum_seg rstd^2 * sum g^2) >= ||gx||_2, s = largest
 * power of two with U*s <= 2^14 (decided between the two passes, so no value can overflow). */
int lgd_gn_bwd(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats, int relu, float* gx,
               int round_out, void* gx_half, float* scale3, float* chan_sums, float* chan_total, void* workspace,
               size_t workspace_bytes, void* stream);
size_t lgd_gn_bwd_workspace(const lgd_pyramid_t* pyr);
/* The same when gy was produced by lgd_conv3x3_dgrad_f16_gnsums: its epilogue already emitted the per-tile sums of
 * (g, g*xhat, g^2), so the first pass (2 F1 of reads) is skipped: 4.5 F1 -> 2.5 F1 of traffic. */
int lgd_gn_bwd_tile_sums(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats, int relu,
                         const float* tile_gn, float* gx, int round_out, void* gx_half, float* scale3, float* chan_sums,
                         float* chan_total, void* workspace, size_t workspace_bytes, void* stream);
/* fp16 dgrad (plain: no mask, fp32 output) whose epilogue also reads gn_x -- the input of the GroupNorm(1)(+ReLU) that
 * follows this convolution in the forward -- and writes tile_gn[tile][4] = per-tile sums of (g, g*xhat, g^2), g = out
 * masked by xhat > 0 when gn_relu. lgd_conv3x3_num_tiles() tiles. */
int lgd_conv3x3_dgrad_f16_gnsums(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                 const float* acc_scale, float* out, const float* gn_x, const float* gn_stats,
                                 int gn_relu, float* tile_gn, void* stream);

/* The same sums for a GroupNorm(1) + ReLU site, taken from the fp16 copy of the GroupNorm's OUTPUT y = relu(xhat) that the
 * forward keeps as the operand of the next convolution: y != 0 <=> xhat > 0 and g * xhat = g * y on the passing elements,
 * so the epilogue reads half the bytes and needs no statistics. */
int lgd_conv3x3_dgrad_f16_gnsums_y(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                   const float* acc_scale, float* out, const void* gn_y_half, float* tile_gn, void* stream);

/* ---- K3+K4: label-guided box-mask average pooling (dynamic_teacher.py:81-103) ---- */
/* x: raw student_proj conv output; if gn_stats != NULL the pooled value is relu((x-mean)*rstd).
 * pooled: (F,T,256). workspace: lgd_maskpool_workspace() bytes. */
size_t lgd_maskpool_workspace(const lgd_pyramid_t* pyr, int T);
int lgd_maskpool_fwd(const lgd_pyramid_t* pyr, const float* x, const float* gn_stats, const int32_t* ranges,
                     const int32_t* img_of, int T, float* pooled, void* workspace, size_t workspace_bytes,
                     void* stream);
/* gy (pyramid) = sum over boxes covering the pixel of gpooled[l,t,:]/max(cnt,1): gradient w.r.t. the pooled
 * map (the GN+ReLU output); feed to lgd_gn_bwd(relu=1) afterwards. */
int lgd_maskpool_bwd(const lgd_pyramid_t* pyr, const float* gpooled, const int32_t* ranges, const int32_t* img_start,
                     int T, float* gy, void* stream);

/* ---- K7: rendering (dynamic_teacher.py:106-206): out[pixel] = sum of emb rows of covering boxes ---- */
/* emb: (F,T,256); rows [img_start[b], img_start[b]+n_render[b]) of image b are rendered. */
int lgd_render_fwd(const lgd_pyramid_t* pyr, const float* emb, const int32_t* ranges, const int32_t* img_start,
                   const int32_t* n_render, int T, float* out, int round_out, void* out_half /* optional fp16 copy */,
                   void* stream);
/* gemb[l,t,:] = sum over the box's pixels of gout (zero for rows that were not rendered) */
int lgd_render_bwd(const lgd_pyramid_t* pyr, const float* gout, const int32_t* ranges, const int32_t* img_of,
                   const int32_t* img_start, const int32_t* n_render, int T, float* gemb, void* workspace,
                   size_t workspace_bytes, void* stream);
/* bias table for local_inst_proj_2D: out[l,b,c] = conv_bias[c] + (ctx_row[b] >= 0 ? ctx[l, ctx_row[b], c] : 0) */
int lgd_ctx_bias_table(const float* ctx, const int32_t* ctx_row, const float* conv_bias, int F, int B, int T,
                       float* out, void* stream);
/* transpose of the gather above: gctx[l,t,:] = (t == ctx_row[img_of[t]]) ? gtable[l, img_of[t], :] : 0   (F,T,256) */
int lgd_ctx_bias_table_bwd(const float* gtable, const int32_t* ctx_row, const int32_t* img_of, int F, int B, int T,
                           float* gctx, void* stream);
/* per-(level,image) channel sums of a pyramid buffer: out[l,b,c] = sum_pixels g[l,b,p,c] and, if total != NULL,
 * total[c] = sum_{l,b} out[l,b,c]  (conv-bias / context-vector gradients) */
int lgd_pyramid_channel_sums(const lgd_pyramid_t* pyr, const float* g, float* out, float* total, void* workspace,
                             size_t workspace_bytes, void* stream);
size_t lgd_channel_sums_workspace(const lgd_pyramid_t* pyr);

/* ---- K8: InstanceNorm2d(256) x2 + MSE (base_distillator.py:59-64) ---- */
/* stats: (F,B,256,2) = {mean, rstd} over (h,w) */
int lgd_in_stats(const lgd_pyramid_t* pyr, const float* x, float* stats, void* workspace, size_t workspace_bytes,
                 void* stream);
/* loss[0] = coef/(B*256*P) * sum (IN(t) - IN(s))^2 ; deterministic two-stage reduction */
int lgd_in_mse_fwd(const lgd_pyramid_t* pyr, const float* s, const float* t, const float* stats_s,
                   const float* stats_t, float coef, float* loss, void* workspace, size_t workspace_bytes,
                   void* stream);
/* The same loss from ONE pass over (s, t): five shifted per-channel moments give the InstanceNorm statistics of both
 * sides (written to stats_s / stats_t, (F,B,256,2)), the loss, and bwd_sums (F,B,2,256) = per-channel totals of
 * (d, d*IN(s)), d = IN(s)-IN(t), which lgd_in_mse_bwd otherwise has to reduce in a pass of its own. */
int lgd_in_mse_moments_fwd(const lgd_pyramid_t* pyr, const float* s, const float* t, float coef, float* stats_s,
                           float* stats_t, float* bwd_sums, float* gs_terms /* optional (F*B): sum_c rs^2 sum d^2 */,
                           float* loss, void* workspace, size_t workspace_bytes, void* stream);
/* gs = d loss / d s (through the student-side InstanceNorm), scaled by gloss[0]. bwd_sums: optional, from
 * lgd_in_mse_moments_fwd (skips the reduction pass). chan_sums (F,B,256) / chan_total (256): optional channel sums
 * of the un-rounded gs (bias gradient of the last adapter convolution).
 * gs_terms (from lgd_in_mse_moments_fwd) + gs_half + scale3 (optional, all or none): scaled fp16 copy of gs and its
 * {s, 1/s, U} triple, U = |2 coef/N * gloss| * sqrt(sum gs_terms) >= ||gs||_2. */
int lgd_in_mse_bwd(const lgd_pyramid_t* pyr, const float* s, const float* t, const float* stats_s,
                   const float* stats_t, const float* bwd_sums, float coef, const float* gloss, float* gs,
                   int round_out, const float* gs_terms, void* gs_half, float* scale3, float* chan_sums,
                   float* chan_total, void* workspace, size_t workspace_bytes, void* stream);
size_t lgd_in_workspace(const lgd_pyramid_t* pyr);

/* elementwise helpers on flat fp32 arrays */
int lgd_relu_bwd(const float* gy, const float* y, float* gx, int64_t n, int round_out, void* stream);
int lgd_round_tf32(const float* x, float* y, int64_t n, void* stream);
/* out3 = {s, 1/s, U}: U = m3 * |m1[0]| * |m2[0]| * sqrt(sum_i terms[i*stride]) (NULL factors = 1), s = largest power
 * of two with U*s <= 2^14. Scale of the fp16 copy of a gradient tensor whose l2 norm is bounded by U: e.g. the output
 * of a dgrad (m1 = the weight's gain from lgd_pack_conv_weight_f16, m2 = U of its input), or a measured norm
 * (terms = the sum-of-squares column of a convolution's tile_stats, stride 2). */
int lgd_grad_scale(const float* terms, int n, int stride, const float* m1, const float* m2, float m3, float* out3,
                   void* stream);
/* x <- hi = tf32(x) in place, lo <- tf32(x - hi): operands of the split-operand forward convolution */
int lgd_tf32_split(float* x, float* lo, int64_t n, void* stream);

/* y[i] += x[i] (gradient accumulation of small tensors) */
int lgd_axpy(const float* x, float* y, int64_t n, void* stream);
/* Loss read-back for the training loop's logging (train.py:196, `v.item()` per loss): n <= 1024 device floats written
 * into PINNED host memory (cudaHostAlloc / torch pin_memory) by a kernel, not by a copy engine; the host reads them
 * after synchronising an event recorded behind this call. */
int lgd_store_to_host(const float* src, float* pinned_host_dst, int n, void* stream);
/* Small upload (the box table of box_descriptor_encode, label_encoder.py:40-85: the reference's per-image `.to(device)`
 * copies) pulled from PINNED host memory by a kernel instead of a copy engine, so that it never queues behind a bulk
 * transfer of another stream. nbytes: multiple of 4. The host buffer must stay alive until the kernel has run. */
int lgd_upload_from_host(void* dst, const void* pinned_host_src, int64_t nbytes, void* stream);

/* ==== step runtime: one call per chain (lgd_b200/csrc/chain.cu) ===========================================
 * The per-kernel entry points above are what the chains are made of; these four calls enqueue a whole chain from
 * native code -- the host language only allocates buffers. Replaces, per call:
 *   lgd_teacher_forward   DynamicTeacher.forward              dynamic_teacher/dynamic_teacher.py:285-301
 *   lgd_teacher_backward  autograd of it                      (reference: implicit, train.py:203)
 *   lgd_distill_forward   BaseDistillator.distill             models/base_distillator.py:34-64 (+ sequential_convs.py:7-15)
 *   lgd_distill_backward  autograd of it
 * Supported configuration: INTERACT_PATTERN = stuGuided (every shipped config), context box on or off, fp16
 * tensor-core operands. Other patterns / precision modes are orchestrated per kernel by the host (engine.py).
 *
 * Buffers: `tape` (lgd_*_tape_bytes) is written by the forward and read by the backward of the same step; `scratch`
 * (lgd_*_scratch_bytes) is dead when the call returns (stream-ordered: every side stream of the context is joined
 * into `stream` before the call returns). The context owns the side streams (weight-gradient GEMMs, label-side
 * backward) and the optional per-call event timing; it holds no tensors. One context per device; calls on one
 * context must not overlap on the host (PyTorch's autograd engine runs one device's nodes on one thread).
 * wgrad_workspace: lgd_conv3x3_wgrad_workspace() bytes that stay valid until the wgrad stream has drained (a
 * per-context persistent buffer).
 * params_host / grads_host: host arrays of device pointers in the order of lgd_teacher_param_name(i) /
 * lgd_adapter_param_name(i) (names relative to "teacher." / "adapter.distill.", the reference's state_dict names).
 * Every gradient is written in full (no accumulation); global_ctx_proj_1D.* may be NULL without a context box.
 * box_blob (device int32): boxes (T,4) fp32 clamped XYXY | labels T | img_of T | img_start B+1 | n_render B | ctx_row B.
 * gtea_levels_host / gstu_levels_host: host arrays of num_levels device pointers to contiguous (B,256,h,w) fp32 maps
 * (cotangents of the teacher pyramid in; gradient w.r.t. the student maps out, NULL = not needed). */
typedef struct lgd_ctx lgd_ctx_t;
lgd_ctx_t* lgd_ctx_create(void);
void lgd_ctx_destroy(lgd_ctx_t* ctx);
int lgd_ctx_set_side_streams(lgd_ctx_t* ctx, int enable);
/* Makes `stream` wait until the last lgd_teacher_backward enqueued on this context has produced every parameter
 * gradient EXCEPT student_proj_2D's (which comes out of the last kernels of the chain): the data-parallel exchange of
 * that part (train.py:277-281: DDP's bucketed all-reduce) can then run underneath the rest of the backward. */
int lgd_ctx_wait_early_grads(lgd_ctx_t* ctx, void* stream);
/* 1 (default; env LGD_B200_TOKENPROG=0 turns it off): the label encoder forward and the label-side backward run as ONE
 * persistent cooperative kernel each (tokenprog.cu) instead of ~50 / ~100 dependent launches; same arithmetic, same bits */
int lgd_ctx_set_token_programs(lgd_ctx_t* ctx, int enable);
/* per-call device timing (CUDA events around every kernel-level call, everything on the caller's stream) */
int lgd_ctx_profile(lgd_ctx_t* ctx, int enable);
int lgd_ctx_profile_count(lgd_ctx_t* ctx);
int lgd_ctx_profile_get(lgd_ctx_t* ctx, int i, const char** name, float* ms); /* after a device synchronize */
void lgd_ctx_profile_reset(lgd_ctx_t* ctx);

typedef struct lgd_step_desc {
  lgd_pyramid_t pyr;
  int32_t T;               /* rows of the box table (GT boxes + context / dummy rows) */
  int32_t img_h, img_w;    /* padded batch image size (label_encoder.py:167) */
  int32_t heads;           /* NR_TRANSFORMER_HEADS */
  int32_t max_n;           /* largest number of rows of one image */
  int32_t add_context_box; /* ADD_CONTEXT_BOX */
} lgd_step_desc_t;

int lgd_teacher_param_count(void);
const char* lgd_teacher_param_name(int i);
int lgd_adapter_param_count(void);
const char* lgd_adapter_param_name(int i);
size_t lgd_teacher_tape_bytes(const lgd_step_desc_t* desc);
size_t lgd_teacher_scratch_bytes(const lgd_step_desc_t* desc, int backward);
size_t lgd_distill_tape_bytes(const lgd_step_desc_t* desc);
size_t lgd_distill_scratch_bytes(const lgd_step_desc_t* desc, int backward);
/* diagnostics (parity tests): byte offset / size of a named tensor inside a tape. Teacher: ranges, desc, label_embed,
 * canoni, sp_raw, sp_stats, pooled, a, rend_h, y0_h, r0, st0, y1_h, r1, st1, y2_h, r2, st2. Distillation: a1_h, a2_h, s. */
int lgd_teacher_tape_field(const lgd_step_desc_t* desc, const char* name, size_t* offset, size_t* bytes);
int lgd_distill_tape_field(const lgd_step_desc_t* desc, const char* name, size_t* offset, size_t* bytes);
/* stu_half: fp16 pyramid copy of the student FPN maps (lgd_nchw_to_pyramid); tea: fp32 pyramid out;
 * masks (optional): the reference's float masks, layout of lgd_masks_from_ranges */
int lgd_teacher_forward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const int32_t* box_blob, const void* stu_half,
                        const float* const* params_host, float* tea, float* masks, void* tape, size_t tape_bytes,
                        void* scratch, size_t scratch_bytes, void* stream);
int lgd_teacher_backward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const int32_t* box_blob, const void* stu_half,
                         const float* const* params_host, const float* const* gtea_levels_host,
                         const float* gtea_pyramid /* alternative: cotangents already as one NHWC pyramid buffer */,
                         const void* tape, size_t tape_bytes, float* const* grads_host,
                         float* const* gstu_levels_host, float* gstu_pyramid /* alternative: NHWC pyramid out */,
                         int gstu_accumulate, void* wgrad_workspace, void* scratch, size_t scratch_bytes, void* stream);
/* tea_ready_event (optional cudaEvent_t): waited for right before the loss kernel, so that the adapter convolutions
 * run next to the teacher chain when this chain is on a stream of its own. loss: device scalar. */
int lgd_distill_forward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const void* stu_half, const float* tea,
                        const float* const* params_host, float coef, void* tea_ready_event, float* loss, void* tape,
                        size_t tape_bytes, void* scratch, size_t scratch_bytes, void* stream);
int lgd_distill_backward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const void* stu_half, const float* tea,
                         const float* const* params_host, float coef, const float* gloss, const void* tape,
                         size_t tape_bytes, float* const* grads_host, float* const* gstu_levels_host,
                         float* gstu_pyramid, int gstu_accumulate, void* wgrad_workspace, void* scratch,
                         size_t scratch_bytes, void* stream);

/* ==== rasterised polygon masks: the Mask R-CNN recipe, LOAD_LABELMAP = True ==========
 * Replaces get_segmask_inside_gt's consumers (dynamic_teacher/utils.py:92-132, dynamic_teacher.py:238-253,137-139):
 * the masks are arbitrary bitmaps, so pooling / rendering read the reference's own float 0/1 mask tensors (layout of
 * lgd_masks_from_ranges: level l at T*sum_{j<l} h_j*w_j, row t = h_l*w_l floats) instead of box intervals. */
/* (T,133) descriptors: boxes/W,H | one-hot | mask49 (T,49) 7x7 box-relative bitmasks, all scaled to [-1,1] */
int lgd_encode_descriptors_masks(const float* boxes, const int32_t* labels, const float* mask49, int T, int img_h,
                                 int img_w, float* desc, void* stream);
/* ==== local_inst_proj_2D over the rendered map WITHOUT a convolution (dynamic_teacher.py:137-146) =====================
 * The rendered map sum_t mask_t (x) emb_t is piecewise constant over box rectangles, so
 *   conv3x3(rendered)[y,x] = sum_t sum_tap [(y+dy, x+dx) in box_t] * (W_tap emb_t)
 * (a token-sized GEMM + a paint pass with interior / ring coverage masks), and the backward needs ONE box-sum pass with
 * nine accumulators per channel plus two token-sized GEMMs (taprender.cu). Exact fp32 arithmetic. Box masks only (the
 * LOAD_LABELMAP recipe keeps the convolution); at most LGD_TAP_MAX_ROWS rendered rows per image. */
#define LGD_TAP_MAX_ROWS 256
size_t lgd_tap_render_workspace(const lgd_pyramid_t* pyr, int T, int backward);
/* out = relu(bias + conv3x3(render(emb), weight)): emb (F*T, 256) rows level-major as for lgd_render_fwd, weight =
 * nn.Conv2d weight (256, 256, 3, 3) fp32, bias[l*bias_stride_level + b*bias_stride_img + c] (strides 0: one vector);
 * out_half (fp16 pyramid; positive values never round to zero: it doubles as the ReLU mask) and / or out32. max_rows =
 * largest number of rows of one image (<= LGD_TAP_MAX_ROWS, else LGD_EINVAL: use lgd_render_fwd + lgd_conv3x3_fwd_f16). */
int lgd_tap_render_fwd(const lgd_pyramid_t* pyr, const float* emb, const float* weight, const int32_t* ranges,
                       const int32_t* img_start, const int32_t* n_render, int T, int max_rows, const float* bias,
                       int bias_stride_level, int bias_stride_img, void* out_half, float* out32, void* workspace,
                       size_t workspace_bytes, void* stream);
/* gout: gradient w.r.t. the pre-ReLU output (i.e. already masked by out > 0), fp32 pyramid. Writes gemb (F*T, 256) (rows
 * outside the rendered subset: 0) and gweight (256, 256, 3, 3). The bias gradient is the channel sum of gout. */
int lgd_tap_render_bwd(const lgd_pyramid_t* pyr, const float* gout, const float* emb, const float* weight,
                       const int32_t* ranges, const int32_t* img_of, const int32_t* img_start, const int32_t* n_render,
                       int T, float* gemb, float* gweight, void* workspace, size_t workspace_bytes, void* stream);

/* CATEGORY_FORMAT = norm_classes (label_encoder.py:24-25,91-93): (T,5) descriptors boxes/W,H | class index /
 * num_classes, or (T,54) with mask49 != NULL; labels[t] < 0 (dummy row of an image without GT) encodes as class 0 */
int lgd_encode_descriptors_norm(const float* boxes, const int32_t* labels, const float* mask49, int T, int img_h,
                                int img_w, int num_classes, float* desc, void* stream);
/* 0/1 bytes (host-rasterised, nearest-sampled level masks) -> float masks */
int lgd_masks_from_bytes(const uint8_t* bytes, int64_t n, float* masks, void* stream);
size_t lgd_dense_mask_workspace(const lgd_pyramid_t* pyr, int T);
/* out[l,t,:] = sum_pixels mask[l,t,pixel] * f(x[l,img_of[t],pixel,:]) [/ max(count,1) if divide]; f = identity, or
 * relu((x-mean)*rstd) with gn_stats. n_rows (optional): only the first n_rows[b] rows of image b, others give zero.
 * count (optional, (F,T)): number of mask pixels. Mask average pooling (dynamic_teacher.py:93-101) and the transpose
 * of the rendering. */
int lgd_mask_gather(const lgd_pyramid_t* pyr, const float* x, const float* gn_stats, const float* masks,
                    const int32_t* img_of, const int32_t* img_start, const int32_t* n_rows, int T, int divide, float* out,
                    float* count, void* workspace, size_t workspace_bytes, void* stream);
/* out[l,b,pixel,:] = sum over the rows t of image b (first n_rows[b] if given) of mask[l,t,pixel] * src[l,t,:]
 * [/ max(count[l,t],1)]: the rendering (dynamic_teacher.py:137-139) and the transpose of the pooling. */
int lgd_mask_paint(const lgd_pyramid_t* pyr, const float* src, const float* masks, const int32_t* img_start,
                   const int32_t* n_rows, const float* count, int T, float* out, void* out_half, void* stream);

/* ==== detection head on the teacher pyramid ======================================
 * The student's RetinaNet head (detectron2 RetinaNetHead as used by customized_detectors/retinanet.py:36-45 and
 * distillator.py:107-112: two towers of four conv3x3(256,256)+ReLU, then conv3x3(256, A*K) and conv3x3(256, A*4)) runs
 * on the same tcgen05 convolution, fed from the NHWC teacher pyramid. Its A*K = 720 / A*4 = 36 output channels are
 * written by 256-column launches straight into pixel-major (B*P, A*K) matrices, which per level ARE the
 * (N, H*W*A, K) tensors the losses consume (permute_to_N_HWA_K, retinanet.py:13-22) -- no NCHW round trip. */
/* forward convolution writing out_cols (multiple of 4, <= 256) fp32 columns at column out_col0 of rows of out_ld
 * elements; row r = pixel r of the pyramid order (level-major, image, y*w+x). bias256: 256 floats (zero padded). */
int lgd_conv3x3_fwd_f16_cols(const lgd_pyramid_t* pyr, const void* in_half, const void* packed_w_half,
                             const float* bias256, float* out, int out_ld, int out_col0, int out_cols, int relu,
                             void* stream);
/* lgd_conv3x3_dgrad_f16 with a previously computed partial result added to the accumulator first (the input gradient
 * of a convolution with more than 256 output channels is the sum of its 256-column pieces; also: sum of the two
 * towers' input gradients). addend has the layout of out and may alias it. */
int lgd_conv3x3_dgrad_f16_addend(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                 const float* acc_scale, const float* addend, float* out, const void* relu_mask_half,
                                 void* out_half, const float* half_scale, float* tile_stats, float* chan_sums,
                                 float* chan_total, void* workspace, size_t workspace_bytes, void* stream);
/* fp16 packing of the 256 output channels [co0, co0+256) of a (co_total,256,3,3) weight (rows beyond co_total are zero):
 * fwd_half [tap][co-co0][ci], dgrad_half [8-tap][ci][co-co0] (either may be NULL), bias256 = zero-padded bias slice
 * (optional), gain (optional, with dgrad_half; 9*256 floats of workspace) as lgd_pack_conv_weight_f16. */
int lgd_pack_conv_weight_f16_rows(const float* w, const float* bias, int co_total, int co0, void* fwd_half,
                                  void* dgrad_half, float* bias256, float* gain, void* workspace, size_t workspace_bytes,
                                  void* stream);
/* gw[co0 + co][ci][ky][kx] = packed_grad[tap][co][ci] for co < co_count */
int lgd_unpack_conv_wgrad_rows(const float* packed_grad, float* gw, int co0, int co_count, void* stream);
/* d(head output) -> operands of the backward: grad_levels_host[l] = level l of the gradient, (B, h*w*A, K) fp32 with
 * contiguous (h*w*A*K) rows per image and batch stride batch_strides_host[l] elements; ncols = A*K. Writes
 * ceil(ncols/256) scaled fp16 pyramids back to back into out_half (zero padded columns), the {s, 1/s, U} triple
 * (U = the tensor's l2 norm) and gbias[ncols] = column sums (the bias gradient). */
size_t lgd_head_grad_workspace(const lgd_pyramid_t* pyr, int ncols);
int lgd_head_grad_prepare(const lgd_pyramid_t* pyr, const float* const* grad_levels_host,
                          const int64_t* batch_strides_host, int ncols, void* out_half, float* scale3, float* gbias,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- GroupNorm(32, 256) with affine parameters (+ReLU): the towers of the FCOS-family detection heads
 * (thirdparty_heads/fcos.py:455-476, poto.py:545-566; applied by FCOSHead.forward fcos.py:528-529) ----
 * stats32 (F, B, 32, 2) = {mean, rstd} per (level, image, group of 8 channels), biased variance, eps 1e-5;
 * chsum (F, B, 256) = per-channel sums of x (kept for the backward). y = relu?(xhat * gamma + beta), written as fp16
 * (operand of the next convolution; positive values never round to zero, so the copy shows the activation pattern)
 * and / or fp32. Backward: gx = d(loss)/dx as a power-of-two scaled fp16 copy (+ its {s, 1/s, U} triple) and / or fp32,
 * dgamma / dbeta (256 each), dbias (256) = per-channel sums of gx = bias gradient of the convolution that produced x. */
size_t lgd_gn32_workspace(const lgd_pyramid_t* pyr);
int lgd_gn32_stats(const lgd_pyramid_t* pyr, const float* x, float* stats32, float* chsum, void* workspace,
                   size_t workspace_bytes, void* stream);
int lgd_gn32_apply(const lgd_pyramid_t* pyr, const float* x, const float* stats32, const float* gamma, const float* beta,
                   int relu, void* y_half, float* y, void* stream);
int lgd_gn32_bwd(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats32, const float* chsum,
                 const float* gamma, const float* beta, int relu, void* gx_half, float* scale3, float* gx, float* dgamma,
                 float* dbeta, float* dbias, void* workspace, size_t workspace_bytes, void* stream);

/* ==== multi-tensor optimizer steps (replaces the per-parameter groups of
 * utils/build.py:497-508 + torch.optim.SGD / AdamW, train.py:209-210) =====================================
 * tensors_dev: device array of descriptors; chunks_dev: device array of int32 pairs {tensor index, chunk index},
 * one block per pair, a chunk = lgd_mt_chunk_elems() consecutive elements. state0 = momentum buffer (SGD) or exp_avg
 * (AdamW), state1 = exp_avg_sq (AdamW). Semantics of torch.optim.SGD(momentum, dampening=0, nesterov=False,
 * weight_decay) with first_step selecting buf = d_p, and of torch.optim.AdamW(betas, eps, weight_decay), step >= 1. */
typedef struct lgd_mt_tensor {
  float* param;
  const float* grad;
  float* state0;
  float* state1;
  int64_t numel;
} lgd_mt_tensor_t;
int lgd_mt_chunk_elems(void);
int lgd_mt_sgd(const lgd_mt_tensor_t* tensors_dev, const int32_t* chunks_dev, int num_chunks, float lr,
               float weight_decay, float momentum, int first_step, void* stream);
int lgd_mt_adamw(const lgd_mt_tensor_t* tensors_dev, const int32_t* chunks_dev, int num_chunks, float lr,
                 float weight_decay, float beta1, float beta2, float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LGD_B200_H_ */
