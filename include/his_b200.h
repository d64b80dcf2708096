/*
 * his_b200.h -- C ABI of libhis_b200.so, the B200 (sm_100a) kernels behind the reference's
 * `src/human_edge_detection` model / ROI API (PINTO0309/human-instance-segmentation).
 *
 * The reference has no FFI of its own (it is pure PyTorch); each entry point below replaces
 * the ATen op(s) the reference dispatches at the cited file:line.  Conventions:
 *   - every function returns 0 (HIS_OK) or a negative error code; his_last_error() gives text;
 *   - plain pointers + explicit shapes/strides, no torch types; nothing allocates device memory
 *     (the caller owns every buffer); `stream` is a cudaStream_t passed as void*;
 *   - "NHWC half slice": fp16, channel-last, per-pixel stride `cs` elements (cs % 8 == 0), the
 *     pointer already offset to the first channel of the slice (offset % 8 == 0);
 *   - small (<= 3 channel) tensors that the reference returns to callers are NCHW fp32.
 * Citations are relative to /root/reference/ ("hed/" = src/human_edge_detection/).
 */
#ifndef HIS_B200_H_
#define HIS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define HIS_OK 0
#define HIS_ERR_INVALID_ARG (-1)
#define HIS_ERR_UNSUPPORTED (-2)
#define HIS_ERR_LAUNCH (-3)
#define HIS_ERR_DRIVER (-4)
#define HIS_ERR_NO_DEVICE (-5)

/* activation codes: hed/advanced/activation_utils.py:71-103, hierarchical_segmentation_rgb.py:21-40 */
#define HIS_ACT_NONE 0
#define HIS_ACT_RELU 1
#define HIS_ACT_SILU 2
#define HIS_ACT_SIGMOID 3
#define HIS_ACT_SWISH 4 /* x*sigmoid(beta*x) */
#define HIS_ACT_GELU 5
/* epilogue second operand: ADD = residual before the activation (ResidualBlock, ..._refinement.py:46-55),
 * MUL = gate after the activation (fg_gate ..._refinement.py:569-570, bottleneck attention ..._unet.py:394-395) */
#define HIS_RES_NONE 0
#define HIS_RES_ADD 1
#define HIS_RES_MUL 2

const char* his_last_error(void);
int his_version(void);

/* ---- DynamicRoIAlign.forward, hed/dynamic_roi_align.py:56-171 (index_select + grid_sample).
 * feat: [B,C,H,W] with element strides sN,sC,sH,sW (fp32, or fp16 if feat_is_half).
 * rois: device [n_rois,5] fp32 {batch_idx,x1,y1,x2,y2}, normalised coordinates.
 * Writes an NHWC half slice [n_rois,oh,ow,C] and/or NCHW fp32 [n_rois,C,oh,ow] (either may be NULL). */
int his_roi_align(const void* feat, int feat_is_half, long long sN, long long sC, long long sH, long long sW,
                  int B, int C, int H, int W, const float* rois, int n_rois, int oh, int ow,
                  float scale_h, float scale_w, int aligned, void* out_half, int out_cs, float* out_f32, int split, void* stream);

/* Both aligners of the model forward (rgb.py:751-755) in one launch: source 0 (e.g. the 2-channel UNet logits) and the optional
 * source 1 (the RGB image), fp32 NCHW contiguous [B,C,H,W], C0 + C1 <= 5, same rois / output size, each with its own
 * spatial scale, `aligned` flag and outputs (NHWC half slice and / or NCHW fp32, either may be NULL).  One warp per (ROI, output
 * row) stages the two source rows of that output row in shared memory (coalesced reads) and interpolates from there; results are
 * identical to his_roi_align. */
int his_roi_align_fused(const float* feat0, int C0, float scale_h0, float scale_w0, int aligned0, void* out_half0, int out_cs0, float* out_f0,
                        const float* feat1, int C1, float scale_h1, float scale_w1, int aligned1, void* out_half1, int out_cs1, float* out_f1,
                        int B, int H, int W, const float* rois, int n_rois, int oh, int ow, int split, void* stream);

/* ---- dense conv2d 3x3(pad 1)/1x1, stride 1, and conv_transpose2d k2 s2, as tcgen05 implicit GEMM
 * (hed/advanced/hierarchical_segmentation_rgb.py:657-673,695; ..._refinement.py:37-39,479-523,537-545;
 *  ..._unet.py:44-47,313-372; smp decoder convs / timm 1x1 convs, see oracle/effunet.py).
 * y = act(conv_w(x) + shift[c] (+res)) (*res), fp16 in/out, fp32 accumulate.  An eval-mode BatchNorm that follows the
 * conv is folded by the caller: w = weight * gamma/sqrt(var+eps) per output channel (before the fp16 rounding),
 * shift = beta - mean*scale + bias*scale.
 * w_packed: fp16 [groups][taps][cout_slab][cin_pad] (groups = 4 for transposed: (dy,dx) = (g>>1,g&1)),
 * cout_slab = n_tiles*block_n from his_conv_gemm_tile_n; shift: device fp32 [cout_slab].
 * The plan captures the pointers (TMA descriptors); run it any number of times. */
int his_conv_gemm_tile_n(int cout, int* n_tiles, int* block_n);
int his_conv_gemm_create(void** plan, const void* in, int n_img, int H, int W, int cin, int in_cs,
                         const void* w_packed, int cin_pad, void* out, int cout, int out_cs,
                         const void* res, int res_cs, const float* shift,
                         int ksize, int transposed, int act, float act_beta, int res_mode, int split);
/* `split` = 1 selects the split-fp16 ("strict" precision) form, which matches the reference's fp32 evaluation
 * (hed/train_utils.py:162-180, no autocast) to ~1e-5 instead of ~1.5e-3: every NHWC half slice of the call is then a PAIR of
 * fp16 planes x = hi + lo, hi = fp16(x), lo = fp16(x - hi), stored [hi channels | lo channels] inside one pixel (the lo plane
 * cs/2 elements after the hi plane; cs % 16 == 0); w_packed holds [W_hi | W_lo] along K (cin_pad = 2 x the padded Cin) and the
 * kernel runs three MMA passes per K block (A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, exact products, fp32 accumulation).  The same
 * `split` flag on the other entry points below means the same layout for their NHWC half operands (and fp32 instead of fp16
 * weights for his_conv_direct / his_depthwise_conv).
 *
 * Tuning knobs read from the environment at plan creation (experiments only; defaults are the measured optimum, profiles/):
 *   HIS_GEMM_HALO=0 (disable the halo-window A path), HIS_GEMM_HALO_MAXN=n (widest N tile in halo mode, default 96),
 *   HIS_GEMM_PAIR=n (minimum N for cta_group::2 CTA pairs, default 128; <=0 disables), HIS_GEMM_PAIR_HALO=n (pairs fed by the halo
 *   window), HIS_GEMM_TAPS=1|3|9 (taps per weight stage), HIS_GEMM_ASTAGES / HIS_GEMM_STAGES (ring depths), HIS_GEMM_BRES=0
 *   (no resident weights), HIS_GEMM_NACC (TMEM accumulator ring), HIS_GEMM_DIRECT=0..3 (register->global epilogue policy),
 *   HIS_GEMM_FUSE_UP=0 (no fused nearest upsample), HIS_GEMM_AUX_BUFS=1|2 (fp32 export tiles per epilogue group),
 *   HIS_GEMM_K1_SPLITN=0 (256-wide N tiles for 1x1 layers / merged ConvTranspose phases, two epilogue groups),
 *   HIS_GEMM_DEBUG (bit mask: 1 / 2 / 4 / 8 skip stores / A loads / B loads / MMAs, 16 epilogue clock stamps printed to stderr when
 *   the plan is destroyed, 32 general chunk body everywhere, 64 wait + barrier at the top of every chunk and late residual loads),
 *   HIS_DW_TILED=0 (register-resident depthwise kernel). */
/* Optional fused 1x1 tail to 1-2 channels computed in the epilogue from the fp32 activations (the Cout<=2 convs at
 * ..._refinement.py:293,335,523 and ..._unet.py:371): tail_out[n,o,y,x] = (sigmoid)(sum_c y[c]*tail_w[o][c] + tail_b[o]),
 * NCHW fp32; tail_w: device fp32 [tail_c][cout_slab]; store_main=0 skips writing the wide activation altogether. */
int his_conv_gemm_set_tail(void* plan, const float* tail_w, float tail_b0, float tail_b1, int tail_c, int tail_sigmoid,
                           float* tail_out, int store_main);
/* fp32 NCHW copy [n_img][cout][H][W] of the layer output written from the epilogue (the 256-channel tensors the reference
 * returns in its aux dict: shared_features ..._refinement.py:557, fg_attention :570); with res_mode MUL the copy is the
 * activated value before the product (the gate).  Layers with Cin >= 64 only. */
int his_conv_gemm_set_aux(void* plan, float* aux_out);
/* LayerNorm2d statistics from the epilogue: every work item of the layer writes (sum, sum of squares) of the values it stores to
 * partials[work item][2] (double, device); *parts_per_image receives the work items per sample.  his_layernorm2d_act(...,
 * partials, nparts_given = *parts_per_image, ...) then skips its own pass over the tensor. */
int his_conv_gemm_set_ln_partials(void* plan, double* partials, int* parts_per_image);
int his_conv_gemm_work_items(void* plan);
/* Per-pixel (GEMM row) extras of the SpatialAttentionModule fusion (hed/advanced/attention_modules.py:67-113):
 * row_scale [n_img*H*W] fp32 (may be NULL): y = act(row_scale[pix]*conv(x) + shift ...), i.e. the conv of the gated input;
 * stats_out [n_img*H*W][2] fp32 (may be NULL): channel mean and max of this layer's output per pixel. */
int his_conv_gemm_set_row_ops(void* plan, const float* row_scale, float* stats_out);
/* res_scale [n_img][cout] fp32: the residual / multiplicand operand is multiplied by res_scale[image][channel] when it is read
 * (ChannelAttentionModule folded into the following residual block: y = act(conv(.) + shift + g*x)). */
int his_conv_gemm_set_res_scale(void* plan, const float* res_scale);
/* Fused nearest 2x upsample + concat of the smp UnetDecoderBlock (F.interpolate(x, mode="nearest") then torch.cat with the
 * skip): channels [0, low_c) of this 3x3 halo-mode layer's input are gathered from `low` [n_img, H/2, W/2, low_cs] at
 * (y>>1, x>>1); the remaining channels come from the `in` buffer given at creation (same channel indices).
 * his_conv_gemm_can_fuse_upsample tells whether a layer qualifies (halo mode, even H and W, low_c % K-block == 0). */
int his_conv_gemm_set_upsampled_input(void* plan, const void* low, int low_c, int low_cs);
int his_conv_gemm_can_fuse_upsample(int H, int W, int cin, int cout, int low_c);
/* Per-image weights: image n reads slab n of w_packed_per_image ([n_img] x the his_conv_gemm_create layout).  Used to fold
 * the squeeze-excite gate of timm's MBConv into the projection conv (his_scale_weights). */
int his_conv_gemm_set_image_weights(void* plan, const void* w_packed_per_image);
int his_conv_gemm_run(void* plan, void* stream);
int his_conv_gemm_destroy(void* plan);
long long his_conv_gemm_issued_macs(void* plan);

/* ---- direct convolution for the shapes that are not GEMM-worthy (Cin = 2/3, Cout <= 2 tails, strided stem,
 * segmentation head): rgb.py:657 (3->64), ..._refinement.py:506,523,537 (tails / 2->64), ..._unet.py:371,
 * timm conv_stem, smp segmentation_head.  in_fmt 0: NHWC half slice; 1: NCHW fp32 with optional per-channel
 * input affine x*a[c]+b[c] (device [2*cin], applied to in-bounds samples only = normalise-then-zero-pad); 2: phase-packed NHWC half
 * [N, H/2, W/2, 4*cin] (pixel (y, x) = channel block (y&1)*2 + (x&1) of low pixel (y>>1, x>>1); 3x3 16->1 head only).
 * w: fp16 [kh][kw][cin][cout]. Writes NHWC half slice and/or NCHW fp32. */
int his_conv_direct(const void* in, int in_fmt, const float* in_affine, int N, int H, int W, int cin, int in_cs,
                    const void* w, const float* scale, const float* shift, int cout, int kh, int kw, int stride, int pad,
                    int act, float act_beta, int res_mode, const void* res, int res_cs,
                    void* out_half, int out_cs, float* out_f32, int split, void* stream);

/* ---- timm DepthwiseSeparableConv / InvertedResidual depthwise conv + BN + act (oracle/effunet.py _DS/_IR);
 * symmetric padding ((s-1)+(k-1))/2; w: fp16 [k*k][C]; optionally writes per-block partial sums of the output,
 * fp32 [N][parts][C] with parts = his_depthwise_pool_parts(...), for the squeeze-excite pooling that follows
 * (fixed-order reduction: bit-reproducible run to run, no atomics). */
int his_depthwise_pool_parts(int N, int H, int W, int C, int k, int stride);
int his_depthwise_conv(const void* in, int N, int H, int W, int C, int in_cs, const void* w, const float* scale,
                       const float* shift, int k, int stride, int act, void* out, int out_cs, float* pool_sums, int split, void* stream);

/* ---- squeeze-excite (timm SqueezeExcite) and ChannelAttentionModule (hed/advanced/attention_modules.py:10-64):
 * pool_sum: per-block partial sums fp32 [N][parts][C], parts = his_pool_sum_parts(...); se_gate: mean = sum of the
 * `nparts` partials / HW (the partials are first summed IN PLACE into part 0, so pool_sums is clobbered),
 * gate = sigmoid(W2*act(W1*mean+b1)+b2), fp32 weights w1 [R,C], w2 [C,R], biases may be NULL;
 * scale_channels: out = in * gate[n,c]. */
int his_pool_sum_parts(int N, int HW, int C);
int his_pool_sum(const void* in, int N, int HW, int C, int cs, float* pool_sums, int split, void* stream);
int his_se_gate(float* pool_sums, int nparts, int N, int HW, int C, int R, const float* w1, const float* b1,
                const float* w2, const float* b2, int act, float act_beta, float* hidden_ws /* [N][R] */, float* gate,
                void* stream);
/* out[n][row][k] = w_packed[row][k] * gate[n][k] (k < C, else 0): folds an input-channel gate into packed GEMM weights. */
int his_scale_weights(const void* w_packed, const float* gate, int N, long long rows, int K, int C, void* out, int split, void* stream);
int his_scale_channels(const void* in, int in_cs, const float* gate, int N, int HW, int C, void* out, int out_cs, int split, void* stream);

/* ---- LayerNorm2d, hed/model.py:18-38 (statistics over C,H,W per sample, biased variance) + optional residual +
 * activation, on an NHWC half slice.  partials_ws: device double [N][his_layernorm2d_parts(...)][2].
 * his_convT2x2_small: ConvTranspose2d(k2,s2) for tiny Cin (upsample_bg_fg.0 when it is followed by LayerNorm2d):
 * NCHW fp32 in, PyTorch weight layout [Cin][Cout][2][2] fp32, NHWC half out. */
int his_layernorm2d_parts(int N, int HW, int C);
int his_layernorm2d_act(const void* in, int N, int HW, int C, int in_cs, const float* gamma, const float* beta, float eps,
                        int act, float act_beta, int res_mode, const void* res, int res_cs, double* partials_ws, int nparts_given,
                        void* out, int out_cs, int split, void* stream);
/* nn.GroupNorm / nn.InstanceNorm2d(affine=True) / AdaptiveInstanceNorm2d / SpatialGroupNorm of get_normalization_layer
 * (hed/advanced/normalization_comparison.py:12-74,159-206) + residual + activation on an NHWC half slice: statistics per
 * (sample, group of C/groups channels) over (C/groups, H, W), biased variance.  ws: float [N][his_groupnorm_parts()+1][C][2]. */
int his_groupnorm_parts(int N, int HW, int C);
int his_groupnorm_act(const void* in, int N, int HW, int C, int in_cs, int groups, const float* gamma, const float* beta, float eps, int act,
                      float act_beta, int res_mode, const void* res, int res_cs, float* ws, void* out, int out_cs, int split, void* stream);
/* ForegroundAwareNorm (normalization_comparison.py:84-132): instance statistics of `in` (ws as for his_groupnorm_act), then
 * y = IN(x) * (p*fg_scale + (1-p)*bg_scale) + (p*fg_bias + (1-p)*bg_bias) (+ residual) + activation, with p = prob[N*HW] (fp32, the
 * output of the module's detector convs on the un-normalised input, run by the caller through the conv entry points). */
int his_fgaware_norm_act(const void* in, int N, int HW, int C, int in_cs, const float* fg_scale, const float* fg_bias,
                         const float* bg_scale, const float* bg_bias, const float* prob, float eps, int act, float act_beta,
                         int res_mode, const void* res, int res_cs, float* ws, void* out, int out_cs, int split, void* stream);
int his_convT2x2_small(const float* in, int N, int cin, int h, int w, const float* wt, const float* bias, int cout,
                       void* out, int out_cs, int split, void* stream);

/* ---- SpatialAttentionModule, hed/advanced/attention_modules.py:67-113: out = x*sigmoid(conv_kxk([mean_c,max_c])).
 * w: fp32 [2][k][k]; stats_ws: fp32 workspace [N*H*W*2]. */
int his_spatial_attention(const void* in, int N, int H, int W, int C, int in_cs, const float* w, int k, float* stats_ws,
                          void* out, int out_cs, int split, void* stream);
/* gate = sigmoid(conv_kxk([mean,max])) from per-pixel statistics [N,H,W,2] (no pass over the feature tensor). */
int his_spatial_gate(const float* stats, int N, int H, int W, const float* w, int k, float* gate, void* stream);

/* ---- nn.MaxPool2d(2) (..._unet.py:332), nearest resize (smp UnetDecoderBlock), bilinear align_corners=False
 * (F.interpolate at ..._refinement.py:561-566,581-586,772-802) */
int his_maxpool2(const void* in, int N, int H, int W, int C, int in_cs, void* out, int out_cs, int split, void* stream);
int his_resize_nearest(const void* in, int N, int H, int W, int C, int in_cs, int Ho, int Wo, void* out, int out_cs, int split, void* stream);
int his_resize_bilinear_f32(const float* in, int NC, int H, int W, int Ho, int Wo, float* out, void* stream);
/* F.interpolate(bilinear, align_corners=False) of an NHWC half slice (MultiScaleRGBSegmentationModel, rgb.py:887-893). */
int his_resize_bilinear_half(const void* in, int N, int H, int W, int C, int in_cs, int Ho, int Wo, void* out, int out_cs, int split, void* stream);

/* ---- head tails: upsample_bg_fg (..._refinement.py:501-506) fused ConvT(2->32,k2,s2)+BN+act+1x1(32->2), NCHW fp32;
 * hierarchical combine (:588-596); elementwise sigmoid / sigmoid((x-*param)*10) (:294,:341). */
int his_upsample_bgfg(const float* low, int N, int h, int w, const float* wt, const float* scale, const float* shift,
                      const float* w1, const float* b1, int act, float act_beta, float* out, void* stream);
int his_head_combine(const float* bgfg, const float* tn, int N, int H, int W, float* logits, void* stream);
int his_map_f32(const float* in, long long total, int op, const float* param, float* out, void* stream);
int his_nhwc_half_to_nchw_float(const void* in, int N, int HW, int C, int cs, float* out, int split, void* stream);

/* ---- refinement flags of RefinedHierarchicalSegmentationHead (hed/advanced/hierarchical_segmentation_refinement.py):
 * his_pixel_shuffle2_f32: nn.PixelShuffle(2) of SubPixelDecoder (:218-252), NCHW fp32 [N,in_channels>=4C,h,w] -> [N,C,2h,2w];
 * his_boundary_edges: BoundaryRefinementModule.detect_edges (:94-129) up to the normalisation -- raw edge map [N,H,W] and the
 *   min / max over the whole tensor in minmax_ws[2] (bit patterns); his_boundary_blend: logits + blend * correction * normalised
 *   edges (:131-149), blend_weight is the device scalar parameter. */
int his_pixel_shuffle2_f32(const float* in, int N, int C, int in_channels, int h, int w, float* out, void* stream);
/* NHWC half depth-to-space: out[n][2y+py][2x+px][c] = in[n][y][x][(py*2+px)*C + c] -- interleaves the four phase convolutions that
 * make up ConvTranspose2d(k4, s2, p1) of ProgressiveUpsamplingDecoder (:152-215), which run as one 3x3 conv with 4*C channels. */
int his_depth_to_space2_half(const void* in, int N, int h, int w, int C, int in_cs, void* out, int out_cs, int split, void* stream);
int his_boundary_edges(const float* logits, int N, int H, int W, float* edges, unsigned int* minmax_ws, void* stream);
int his_boundary_blend(const float* logits, const float* correction, const float* edges, const unsigned int* minmax_ws,
                       const float* blend_weight, int N, int H, int W, float* out, void* stream);

/* ---- PretrainedUNetGuidedSegmentationHead glue, hed/advanced/hierarchical_segmentation_rgb.py:125-218 (the head the
 * factory builds when no refinement flag is set, :715-727).
 * his_sigmoid_channel: fg_prob = sigmoid(in[:,c]) of NCHW fp32 [N,C,HW] (:146) -> channel 0 of an NHWC half slice
 *   (the concat slot of input_adjust, :161) and/or fp32 [N,HW]; either output may be NULL.
 * his_scale_pixels: out = x * (attention * (0.5 + 0.5*fg_prob)) per pixel (:169-173); attention, fg_prob fp32 [pixels].
 * his_guided_aux: m = bilinear(in[:,c] -> (Ho,Wo), align_corners=False) (identity when sizes match), fg = sigmoid(m),
 *   bgfg = [log(1-fg+1e-7), log(fg+1e-7)] (:187-205); outputs NCHW fp32 [N,1,Ho,Wo] x2 and [N,2,Ho,Wo]. */
int his_sigmoid_channel(const float* in, int N, int C, int HW, int c, void* out_half, int out_cs, float* out_f32, int split, void* stream);
int his_scale_pixels(const void* in, int in_cs, const float* attention, const float* fg_prob, long long pixels, int C, void* out,
                     int out_cs, int split, void* stream);
int his_guided_aux(const float* in, int N, int C, int c, int H, int W, int Ho, int Wo, float* mask_out, float* fg_out, float* bgfg_out,
                   void* stream);

/* ---- PreTrainedPeopleSegmentationUNet.normalize_input (..._unet.py:1885-1890) without the host sync:
 * affine6 = {a0,a1,a2,b0,b1,b2}, a = 1/(d*std), b = -mean/std, d = 255 iff max(images) > 1.  mean3/std3 are HOST arrays.
 * his_unet_outputs: output_conv 1x1 1->2 (:1963-1971) and the export wrapper's binary mask
 * softmax(two)[:,0] (hed/export_onnx_advanced.py:374-387). */
int his_unet_input_affine(const float* images, long long count, const float* mean3, const float* std3,
                          unsigned int* flag_ws, float* affine6, void* stream);
/* Normalised, space-to-depth copy of the NCHW fp32 images for the stem conv (timm conv_stem 3x3 s2 p1): NHWC half
 * [N, H/2, W/2, 16], channel (sy*2+sx)*3 + c; the stride-2 conv becomes a 2x2 stride-1 conv the tensor-core GEMM can run. */
int his_s2d_input(const float* images, int N, int H, int W, const float* affine6, void* out_half, int split, void* stream);
int his_unet_outputs(const float* one, int B, int H, int W, float w0, float w1, float b0, float b1, float* two,
                     float* binary, void* stream);

int his_memset_async(void* ptr, int value, long long bytes, void* stream);

/* ---- post-processing; masks/logits are NCHW fp32 planes like the reference modules take.
 * instance_mask: where(argmax(masks,1)==1,1,0) (hed/export_onnx_advanced.py:360-364; first-max tie rule), optionally
 *   AND (max softmax prob > score_threshold) (test_hierarchical_instance_peopleseg_onnx.py:250-262; <= 0 disables);
 *   writes fp32 [N,1,H,W] and/or u8 [N,H,W].
 * dilate_logits: MaskDilationModule (export_hierarchical_instance_peopleseg_onnx.py:85-141).
 * dilate_instance_mask: instance_mask(dilate_logits(logits)) in one pass -- the dilated logits are never written (the exported
 *   graph ends in the dilation module and the caller takes the argmax: test_hierarchical_instance_peopleseg_onnx.py:250-262).
 * edge_smooth: BinaryMaskEdgeSmoothing (hed/edge_smoothing.py:10-90), N = B*C planes.
 * binary_bilateral: BinaryMaskBilateralFilter (hed/bilateral_filter.py:299-406); gauss = device fp32 [k*k] (normalised).
 * morph_bilateral: MorphologicalBilateralFilter (hed/bilateral_filter.py:409-501); kernel2d = device fp32 [k*k].
 * paste: NEAREST paste-back (test_hierarchical_instance_peopleseg_onnx.py:144-161,264-278,369-374) of u8 ROI masks
 *   [N,mh,mw] into an int32 label canvas [B,H,W] (caller zeroes): pixel = 1 + index of the last ROI covering it. */
int his_post_instance_mask(const float* logits, int N, int H, int W, float score_threshold, float* out_f32,
                           unsigned char* out_u8, void* stream);
int his_post_dilate_logits(const float* logits, int N, int H, int W, int dilation_pixels, float* out, void* stream);
int his_post_dilate_instance_mask(const float* logits, int N, int H, int W, int dilation_pixels, float score_threshold,
                                  float* out_f32, unsigned char* out_u8, void* stream);
int his_post_edge_smooth(const float* mask, int N, int H, int W, float threshold, float blur_strength, float* out, void* stream);
int his_post_binary_bilateral(const float* mask, int N, int H, int W, const float* gauss, int k, int iterations,
                              float threshold, float* ws0, float* ws1, float* out, void* stream);
int his_post_morph_bilateral(const float* mask, int N, int H, int W, const float* kernel2d, int k, int morph,
                             float* ws0, float* ws1, float* out, void* stream);
int his_post_paste(const unsigned char* masks, int N, int mh, int mw, const float* rois, int* canvas, int B, int H, int W,
                   void* stream);

/* ---- prepare_image after the file decode, test_hierarchical_instance_peopleseg_onnx.py:170-196 (cv2.cvtColor BGR2RGB,
 * cv2.resize uint8 INTER_LINEAR, /255, HWC->CHW) fused, bit-exact with OpenCV's 8-bit bilinear arithmetic.
 * src: device uint8 [N,Hs,Ws,3]; out: fp32 [N,3,Hd,Wd]; xtab / ytab: device int32 [4][Wd] / [4][Hd] = {index0, index1,
 * weight0, weight1} (11-bit fixed point), built like cv2 builds them (human_instance_segmentation_b200/preprocess.py). */
int his_preprocess_u8(const unsigned char* src, int N, int Hs, int Ws, int Hd, int Wd, const int* xtab, const int* ytab, int swap_rb,
                      float* out, void* stream);

/* ---- evaluate_model's metric core, hed/train_utils.py:262-292 (argmax, calculate_confusion_matrix :25-47, the per-sample
 * Python loops of the 2x2 matrices and calculate_iou :14-22): per-ROI 3x3 confusion counts counts[n][gt*3+pred] (int32,
 * zeroed by the call) of 3-class logits [N,3,H,W] against labels [N,H,W] (uint8, or int64 when gt_is_int64). */
int his_eval_confusion(const float* logits, const void* gt, int gt_is_int64, int N, int H, int W, int* counts, void* stream);

/* ---- shared-memory tiled stencils (csrc/post_stencil.cu): one 32x32 tile + halo per CTA, the plane is read once.
 * All take N planes [N,H,W] fp32 (a [B,C,H,W] tensor is B*C planes), N <= 65535 per call.
 * his_post_edge_smooth_tiled: BinaryMaskEdgeSmoothing, hed/edge_smoothing.py:10-90 (same result as his_post_edge_smooth).
 * his_post_edge_directional / _adaptive / _optimized: DirectionalEdgeSmoothing, AdaptiveEdgeSmoothing (per-plane runtime
 *   parameters, device arrays [N]) and OptimizedEdgeSmoothing (fp16 != 0: every operator output rounded to half, the
 *   exported FP16 graph), export_edge_smoothing_onnx.py:63-318.
 * his_post_class_masks: per-class binary masks of [B,C,H,W] predictions -- argmax==c (C == 3) or (softmax?)p > 0.5 --
 *   the front end of MultiClassEdgeSmoothing.smooth_predictions, hed/edge_smoothing.py:93-170.
 * his_post_bilateral_exact: BilateralFilter (reflect padding, spatial x range weights), hed/bilateral_filter.py:9-113;
 *   spatial_kernel: device [k*k] (un-normalised exp(-d^2/2s^2), as the reference builds it).
 * his_post_bilateral_fast: FastBilateralFilter, hed/bilateral_filter.py:116-216; kernel1d: device [k] normalised; ws: [N,H,W].
 * his_post_guided_filter: EdgePreservingFilter, hed/bilateral_filter.py:219-296 (guide == x when the caller has no guide).
 * his_post_binary_bilateral_tiled: BinaryMaskBilateralFilter, all iterations in one pass (needs iterations*(k/2) <= 8).
 * his_post_mask_cleanup_fused: BinaryMaskEdgeSmoothing -> BinaryMaskBilateralFilter in one kernel (BASELINE config 5 chain). */
int his_post_edge_smooth_tiled(const float* mask, int N, int H, int W, float threshold, float blur_strength, float* out, void* stream);
int his_post_edge_directional(const float* mask, int N, int H, int W, float* out, void* stream);
int his_post_edge_adaptive(const float* mask, int N, int H, int W, const float* blur_strength, const float* edge_sensitivity,
                           const float* final_threshold, float* out, void* stream);
int his_post_edge_optimized(const float* mask, int N, int H, int W, int fp16, float* out, void* stream);
int his_post_class_masks(const float* pred, int B, int C, int H, int W, int use_argmax, int softmax_first, float* out, void* stream);
int his_post_bilateral_exact(const float* in, int N, int H, int W, const float* spatial_kernel, int k, float sigma_range, float* out,
                             void* stream);
int his_post_bilateral_fast(const float* in, int N, int H, int W, const float* kernel1d, int k, float sigma_range, int iterations,
                            float* ws, float* out, void* stream);
int his_post_guided_filter(const float* x, const float* guide, int N, int H, int W, int radius, float eps, float* ws_a, float* ws_b,
                           float* out, void* stream);
int his_post_binary_bilateral_tiled(const float* mask, int N, int H, int W, const float* gauss, int k, int iterations, float threshold,
                                    float* out, void* stream);
int his_post_mask_cleanup_fused(const float* mask, int N, int H, int W, float es_threshold, float es_strength, const float* gauss, int k,
                                int iterations, float threshold, float* out, void* stream);
/* The same chain on unsigned char masks (0 / 1 in, 0 / 1 out; a byte is converted to float while loading, so the result is that of
 * the fp32 call on float(mask)): a quarter of the bytes over PCIe and HBM.  kernel size 3 / 5 / 7, 1 or 2 iterations. */
int his_post_mask_cleanup_fused_u8(const unsigned char* mask, int N, int H, int W, float es_threshold, float es_strength,
                                   const float* gauss, int k, int iterations, float threshold, unsigned char* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIS_B200_H_ */
