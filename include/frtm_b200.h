/*
 * frtm_b200.h — C ABI of libfrtm_b200.so: hand-written sm_100a kernels for the FRTM-VOS per-frame inference
 * hot path (backbone feature pass, target-model correlation, refinement network, mask merge, frame memory and
 * the online Gauss-Newton / conjugate-gradient filter update).
 *
 * The reference (andr345/frtm-vos) has no FFI for this path: every device kernel it runs is an ATen / cuDNN
 * call made from Python.  Its only native component is lib/_npp/nppig.cpp, whose conventions
 * this ABI follows (nppig.cpp:48-104): the caller owns and allocates every buffer, inputs are contiguous device
 * tensors on the current device, work is enqueued on the caller's stream, nothing synchronises, nothing is
 * retained.  Each entry point below cites the reference call site it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative FRTM_E* code otherwise; frtm_last_error() gives the text
 *     (thread-local).  No host synchronisation, no allocation, no global state besides the error string and
 *     one-time cudaFuncSetAttribute calls.
 *   - all pointers are device pointers unless the name ends in _host; float = IEEE binary32.
 *   - activations on the conv path are NHWC ("channels last"): element (b,y,x,c) at ((b*H+y)*W+x)*ld + c with a
 *     channel stride ld >= C (lets a producer write straight into a wider concat buffer).
 *   - target-model tensors keep the reference's NCHW layout because they are user-visible attributes
 *     (Discriminator.memory.samples, .current_sample, project/filter weights).
 *   - stream is a cudaStream_t passed as void*.
 */
#ifndef FRTM_B200_H
#define FRTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRTM_OK 0
#define FRTM_EINVAL (-1)  /* bad argument (shape / alignment / null) */
#define FRTM_ELAUNCH (-2) /* CUDA launch or runtime failure           */
#define FRTM_EARCH (-3)   /* device is not sm_100                     */

const char *frtm_last_error(void);
int frtm_version(void);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
int64_t frtm_launch_count(void);
/* Adds n to the launch counter: for callers that replay captured launches of this library as a CUDA graph (a replay runs
 * the kernels without passing through the entry points that count them).  Returns the new count. */
int64_t frtm_count_launches(int64_t n);

/* Upload up to 16 floats and 16 ints from HOST arrays to device memory as kernel arguments (no memcpy, hence no
 * synchronisation with work already queued on the stream).  Used for memory.weights / memory state initialisation
 * (model/memory.py:38-46). */
int frtm_fill_small(float *fdst, const float *fvals_host, int nf, int *idst, const int *ivals_host, int ni, void *stream);
/* The same for up to 16 bytes (the label look-up table of a sequence, kept at a stable address). */
int frtm_fill_u8(uint8_t *dst, const int *vals_host, int n, void *stream);
/* Same for n 64-bit integers (device pointer tables of the batched GN update). */
int frtm_fill_i64(void *dst, const int64_t *vals_host, int n, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backbone + refinement-network building blocks  (model/feature_extractor.py:40-68 -> torchvision ResNet,
 * model/seg_network.py:7-189, lib/utils.py:25-41)
 * ---------------------------------------------------------------------------------------------------------- */

/* uint8 NCHW image(s) (B,3,H,W) -> normalised fp32 NHWC (B,H,W,4), 4th channel 0.
 * Replaces  x = norm_weight * input.float() + norm_bias  (model/feature_extractor.py:27-32,42). */
int frtm_normalize_u8(const uint8_t *img, int B, int H, int W, float *out_nhwc4, void *stream);

/* Generic 2-D convolution, fp32 NHWC in / fp32 NHWC out, implicit GEMM.
 *   x      (B,H,W,ldx) using channels [0,Cin), Cin % 4 == 0
 *   w      packed by frtm_pack_conv_weight layout: [kh][kw*Cin][CoutPad], CoutPad = round_up(Cout,4)
 *   bias   [Cout] or NULL;  res (B,Ho,Wo,ldr) residual added before the activation, or NULL
 *   y      (B,Ho,Wo,ldy) written at channel offset y_coff (NULL to skip); y_nchw (B,Cout,Ho,Wo) optional copy
 *   relu   1 -> max(.,0)
 * Replaces every nn.Conv2d / folded BatchNorm / residual add / ReLU on the path: torchvision resnet.py:92-105,
 * 146-163; model/seg_network.py:13-14,29,48-56,138-146; model/discriminator.py:81 (project, batched over
 * objects by stacking the 1x1 kernels along Cout). */
int frtm_conv2d_nhwc(const float *x, int B, int H, int W, int Cin, int ldx, const float *w, const float *bias,
                     const float *res, int ldr, float *y, int ldy, int y_coff, float *y_nchw, int Cout, int kh,
                     int kw, int stride, int pad, int relu, void *stream);

/* Tensor-core (tcgen05) convolution: 1x1 (pad 0) / 3x3 (pad 1), stride 1 or 2, split-fp16 operands, fp32 accumulation in TMEM.
 *   x_hi/x_lo  fp16 NHWC planes (B,H,W,ldx) of 16*x = hi + lo (frtm_split_f16 or a previous conv's y_hi/y_lo),
 *              Cin % 64 == 0, ldx % 8 == 0;  fetched by TMA, zero OOB fill supplies the padding
 *   wt         weights pre-tiled by frtm_vos_b200.ops.pack_conv_tc: [ntile][tap][Cin/64][hi|lo][bn_tile x 64] fp16 in
 *              the 128-byte-swizzled K-major shared-memory image, scaled per output channel by a power of two;
 *              oscale[n] = 1 / (16 * that scale)
 *   outputs    any of: y fp32 NHWC (ldy, y_coff), y_nchw fp32 (B,Cout,H,W), y_hi/y_lo split planes (ldyh, yh_coff),
 *              y_tap (B,Ho,Wo,12) = the 9 tap maps  sum_c tapw[tap][c] * out[c]  of a following 3x3 -> 1 conv contracted
 *              per pixel in the epilogue (tapw [9][Cout]; needs Cout <= bn_tile; see frtm_upsample_tapsum)
 *              (only the first yh_cout channels go to the split planes; 0 = all), y_extra (B,Ho,Wo) = output channel extra_ch
 *   r1_score   optional 65th input channel (B,Ho,Wo) fp32 handled in the epilogue as an exact fp32 rank-1 term with weights
 *              r1_w [9][Cout] and bias r1_bias (TSE.transform on cat(h, score), seg_network.py:15,19-20; see frtm_rank1_finish)
 *   res / res_hi,res_lo   optional residual (fp32 NHWC or split planes) added before the ReLU
 * Same reference call sites as frtm_conv2d_nhwc; the products hi*hi + hi*lo + lo*hi (issued as A_hi x [B_hi | B_lo] and
 * A_lo x B_hi) keep the result within ~1e-6 relative of an fp32 convolution.
 *   kernel_select  0 = the library picks the kernel for the shape (slab kernel with resident weights for 3x3 / 64 input
 *              channels, persistent streaming kernel for narrow 1x1, CTA-pair kernels (tcgen05.mma.cta_group::2) for Cout a
 *              multiple of 128 — persistent when there are several tiles per SM —, general tile kernel otherwise);
 *              1 = general tile kernel, 2 / 3 = CTA-pair kernel with pair tile N = 128 / 256, 4 = persistent CTA-pair kernel
 *              (A-B measurements and tests; a per-call argument, the library keeps no mode) */
int frtm_conv2d_tc(const void *x_hi, const void *x_lo, int B, int H, int W, int Cin, int ldx, const void *wt,
                   const float *oscale, int bn_tile, const float *bias, const float *res, int ldr, const void *res_hi,
                   const void *res_lo, int ldrh, float *y, int ldy, int y_coff, float *y_nchw, void *y_hi, void *y_lo,
                   int ldyh, int yh_coff, int yh_cout, const float *tapw, float *y_tap, const float *r1_score,
                   const float *r1_w, const float *r1_bias, float *y_extra, int extra_ch, int Cout, int kh, int kw,
                   int stride, int relu, int kernel_select, void *stream);
/* Device-side weight packer for 1x1 convs whose weights change at run time (project.weight, model/discriminator.py:81):
 * W (Cout,Cin) fp32 -> wt / oscale in the layout frtm_conv2d_tc expects (wt: cout_pad*Cin*2 halves, oscale: cout_pad). */
int frtm_pack_tc_1x1(const float *W, int Cout, int Cin, int bn_tile, void *wt, float *oscale, void *stream);
/* fp32 NHWC (npix, ldx)[0,C) -> fp16 planes hi, lo with hi + lo = 16 * x  (channel stride ldh, ldh % 8 == 0). */
int frtm_split_f16(const float *x, int64_t npix, int C, int ldx, void *hi, void *lo, int ldh, void *stream);

/* Completes a 3x3 conv over cat(64 channels, score) (TSE.transform, seg_network.py:15,19-20) after frtm_conv2d_tc has
 * produced the 64-channel part y_in (B/n_obj,H,W,ldin) — shared by the n_obj objects of a frame when it depends on the
 * backbone features only: adds the score channel's contribution (wx [9][Cout], score (B,H,W)), bias and ReLU; writes
 * fp32 y_out (B,H,W,ldout) and/or split planes for output channels [0,64), and output channel 64 (Cout == 65) to `extra`. */
int frtm_rank1_finish(const float *y_in, int ldin, int n_obj, float *y_out, int ldout, const float *score, const float *wx,
                      const float *bias, int B, int H, int W, int Cout, int relu, void *y_hi, void *y_lo, int ldh,
                      float *extra, void *stream);

/* 3x3 / stride 2 / pad 1 max pooling, NHWC (torchvision resnet.py maxpool; feature_extractor.py:53). */
/* im2col of the 7x7 / stride-2 / pad-3 stem (torchvision resnet.py conv1) on the normalised image, as split fp16 planes
 * (B,Ho,Wo,192) of 16*x: k = ky*24 + c*8 + kx (kx = 7 and k >= 168 are zero); Ho = (H-1)/2+1.  The stem is then a 1x1
 * frtm_conv2d_tc with Cin = 192 and conv1.weight permuted the same way (frtm_vos_b200.ops.stem_weight_as_1x1). */
int frtm_stem_patches_u8(const uint8_t *img, int B, int H, int W, void *hi, void *lo, void *stream);
/* The stem in ONE kernel: the patches above are built tile by tile in shared memory (never written to HBM) and consumed by
 * the tensor core; y (B,Ho,Wo,ldy)[0,64) = relu?(conv7x7/s2(normalise(img)) * folded BatchNorm), fp32 NHWC.  wt / oscale / bias:
 * conv1.weight as frtm_vos_b200.ops.stem_weight_as_1x1 packs it for frtm_conv2d_tc with the N tile 64.  Bit-identical to
 * frtm_stem_patches_u8 + frtm_conv2d_tc. */
int frtm_stem_conv_u8(const uint8_t *img, int B, int H, int W, const void *wt, const float *oscale, const float *bias, float *y,
                      int ldy, int relu, void *stream);
int frtm_maxpool3x3s2_nhwc(const float *x, int B, int H, int W, int C, float *y, float *y_nchw, void *stream);
/* The same pooling writing the split fp16 planes (B,Ho,Wo,C) of 16*y that the next tensor-core conv reads (frtm_split_f16's
 * conversion); y (fp32) optional. */
int frtm_maxpool3x3s2_split_nhwc(const float *x, int B, int H, int W, int C, float *y, void *y_hi, void *y_lo, void *stream);

/* Bilinear resize (align_corners=False) of an NHWC tensor; writes C channels at offset y_coff of a tensor with
 * channel stride ldy.  If accumulate != 0 the result is added to y.  (lib/utils.py:33-35, seg_network.py:39,144) */
int frtm_resize_bilinear_nhwc(const float *x, int B, int H, int W, int C, int ldx, float *y, int Ho, int Wo, int ldy,
                              int y_coff, int accumulate, void *stream);
/* Bicubic resize (ATen upsample_bicubic2d semantics: align_corners=False, A = -0.75, clamped 4 x 4 taps) of an
 * NHWC tensor, C % 4 == 0.  The `Upsampler` of the YouTubeVOS all-frames variant (ytvos_validation/seg_network.py:62-74). */
int frtm_resize_bicubic_nhwc(const float *x, int B, int H, int W, int C, int ldx, float *y, int Ho, int Wo, int ldy,
                             void *stream);

/* Fixed x2 bicubic pyramid upsample (replicate pad 2, 4 depthwise 4x4 phases, crop 1) NHWC (B,H,W,C)->(B,2H,2W,C)
 * (model/seg_network.py:75-126). */
int frtm_pyrup_bicubic_nhwc(const float *x, int B, int H, int W, int C, float *y, void *y_hi, void *y_lo, void *stream);
/* (y may be NULL when only the split fp16 planes y_hi / y_lo of 16*result are wanted — the input format of frtm_conv2d_tc.) */

/* Global average pool NHWC (B,H,W,ld)[0,C) -> (B,C)  (seg_network.py:18,34-35). Deterministic two-stage sum. */
int frtm_global_avgpool_nhwc(const float *x, int B, int HW, int C, int ldx, float *out, float *workspace,
                             int64_t workspace_bytes, void *stream);
int64_t frtm_global_avgpool_workspace(int B, int HW, int C);

/* Channel-attention block (seg_network.py:24-41):  gate = sigmoid(W2 relu(W1 [gap(shallow), deep_pool] + b1) + b2)
 * then out = shallow * gate + deeper (deeper already resized to the shallow size, or a (B,C) vector when
 * deeper_is_vector).  w1 (C,2C) row-major, w2 (C,C).  shallow_pool/deep_pool are (B,C). */
int frtm_cab_gate(const float *shallow_pool, const float *deep_pool, const float *w1, const float *b1, const float *w2,
                  const float *b2, int B, int C, float *gate, void *stream);
int frtm_cab_apply_nhwc(const float *shallow, const float *gate, const float *deeper, int deeper_is_vector, int B,
                        int HW, int C, float *out, void *stream);
/* The gate straight from the maps (C = 64): both global average pools — of shallow (B,HWs,lds)[0,C) and of deeper
 * (B,HWd,ldd)[0,C), or with deeper == NULL the pooled vector deep_pool (B,C) in its place — and the gate in two launches
 * (stage 1 of both pools; one block per image that finishes them and runs the two small matrix products) instead of the
 * five of frtm_global_avgpool_nhwc x2 + frtm_cab_gate, bit-identical to them.  pool_out (B,2C) = [sp | dp] or NULL. */
int frtm_cab_gate_from_maps(const float *shallow, int HWs, int lds, const float *deeper, int HWd, int ldd,
                            const float *deep_pool, int B, int C, const float *w1, const float *b1, const float *w2,
                            const float *b2, float *gate, float *pool_out, float *workspace, int64_t workspace_bytes,
                            void *stream);
int64_t frtm_cab_gate_from_maps_workspace(int B, int HWs, int HWd, int C);
/* The same with the deeper level still at its own resolution (B,Hd,Wd,C): the bilinear resize (align_corners=False) to
 * (H,W) is evaluated inside the kernel, with frtm_resize_bilinear_nhwc's arithmetic.  out (fp32 NHWC) and / or the split
 * planes y_hi, y_lo (B,H,W,C halves, frtm_split_f16's conversion) for the tensor-core conv that follows. */
int frtm_cab_apply_resized_nhwc(const float *shallow, const float *gate, const float *deeper, int B, int H, int W, int C, int Hd,
                                int Wd, float *out, void *y_hi, void *y_lo, void *stream);

/* Copy a 1-channel map (B,H,W) into channel `coff` of an NHWC tensor with channel stride ld and zero channels
 * (coff, coff+nzero] (builds cat(h, score), lib/utils.py:38-41 / seg_network.py:19). */
int frtm_scatter_channel_nhwc(const float *src, int B, int HW, float *dst, int ld, int coff, int nzero, void *stream);
/* Replicate an NHWC tensor (F,HW,C) along a new object axis into (F*N,HW,ld)[0,C): out[(f*N+n)] = in[f]. */
int frtm_broadcast_objects_nhwc(const float *src, int F, int N, int HW, int C, int lds, float *dst, int ldd, void *stream);

/* NHWC (B,HW,ld)[0,C) -> NCHW (B,C,HW). */
int frtm_nhwc_to_nchw(const float *x, int B, int HW, int C, int ldx, float *y, void *stream);

/* Final 3x3 conv to a single channel (project.conv2, seg_network.py:145): NHWC (B,H,W,C) -> (B,H,W). w [9][C]. */
int frtm_conv3x3_to1_nhwc(const float *x, int B, int H, int W, int C, const float *w, const float *bias, float *y,
                          void *stream);

/* The same final conv evaluated BEFORE the (linear) upsampling chain: t (npix,12) = 9 tap maps  sum_c w[tap][c] x[c]
 * of the low-resolution input; after upsampling the 12-channel tap maps, frtm_shift_sum9 adds the 9 shifted maps
 * (zero outside the image) and the bias.  Exactly conv3x3(upsample(x)) by linearity, with 9 instead of C full-res maps. */
int frtm_tapmaps_nhwc(const float *x, int64_t npix, int C, const float *w9c, float *y12, void *stream);
int frtm_shift_sum9(const float *v12, int B, int H, int W, const float *bias, float *out, void *stream);
/* The whole tail in one kernel: t12 (B,h,w,12) tap maps -> bicubic x2 (PyrUpBicubic2d) -> bilinear to (H,W) -> sum of the 9
 * tap-shifted maps + bias -> logits (B,H,W).  Reads only t12, writes only the logits.  frtm_upsample_tapsum_supported
 * returns 1 if the sizes fit the kernel's shared-memory windows (2h >= H, 2w >= W, scale close to 1). */
int frtm_upsample_tapsum(const float *t12, int B, int h, int w, int H, int W, const float *bias, float *out, void *stream);
int64_t frtm_upsample_tapsum_supported(int h, int w, int H, int W);

/* ------------------------------------------------------------------------------------------------------------
 * Mask merge  (model/tracker.py:143-150, 203-221)
 * ---------------------------------------------------------------------------------------------------------- */

/* One frame, N objects -> merged masks (N+1,H,W) [slot 0 = background] and uint8 labels (H,W).
 *   src (N,H,W): logits for objects whose bit is set in logit_mask (p = sigmoid(logit) * (1 - suppress)), raw
 *   probabilities (start masks of objects initialised on this frame) for the others;  suppress (H,W) uint8 or NULL;
 *   clamp ; p_0 = min(1-p_i) ; softmax(p/(1-p)) ; first-occurrence argmax ; masks[i] = softmax_i * (argmax == i);
 *   labels = lut[argmax of the same rule applied to the merged masks]  (or masks[1] > 0.5 when single_object).
 * Also adds, per object, the number of pixels with merged mask > 0.5 to counts[N] (the `< 10 px` gate of
 * model/discriminator.py:214) — counts must be zeroed by the caller.  N <= 64. */
int frtm_merge_masks(const float *src, uint64_t logit_mask, const uint8_t *suppress, int N, int HW, const uint8_t *lut,
                     int single_object, float *masks, uint8_t *labels, int *counts, void *stream);
/* The same merge for F consecutive frames in one launch (masks never feed the next frame's forward pass): src (F,N,HW),
 * masks (F,N+1,HW), labels (F,HW), counts (F,counts_stride) zeroed by the caller.  No `suppress` (no object starts inside
 * a block). */
int frtm_merge_masks_frames(const float *src, int F, uint64_t logit_mask, int N, int HW, const uint8_t *lut, int single_object,
                            float *masks, uint8_t *labels, int *counts, int counts_stride, void *stream);
/* Label maps from RAW object probabilities src (F,N,HW), one stage: clamp to [1e-7, 1 - 1e-7], background = min_i (1 - p_i),
 * softmax(p / (1 - p)), first maximum -> lut (model/tracker.py:143-150; ytvos_validation/tracker.py:53-62,106-107, where the
 * rule runs once over all frames of a sequence after the ground truth has been re-inserted). */
int frtm_labels_from_probs(const float *src, int F, int N, int HW, const uint8_t *lut, uint8_t *labels, void *stream);
/* out[i] = x[i] > thr ? 1 : 0 — the binary training labels of the "thresh" update method of the all-frames variant
 * (ytvos_validation/discriminator.py:364-367). */
int frtm_threshold_f32(const float *x, int64_t n, float thr, float *out, void *stream);
/* out (N,HW) = sigmoid(logits) * (1 - suppress): the per-object probabilities the all-frames variant returns per frame, with
 * the pixels of objects that start on this frame suppressed (ytvos_validation/tracker.py:133-139,176-178). */
int frtm_sigmoid_suppress(const float *logits, const uint8_t *suppress, int N, int HW, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Target model: correlation, memory, GN/CG  (model/discriminator.py, model/memory.py, model/optimizer.py)
 * ---------------------------------------------------------------------------------------------------------- */

/* scores[n] = filter[fidx(n)] (3x3, pad 1, no bias) applied to x[n] : x (NB,c,h,w) NCHW, filt (NF,c,3,3),
 * filter_index[NB] (NULL -> n).  out (NB,h,w).  (discriminator.py:82,205; also J·p in optimizer.py:156) */
int frtm_corr3x3_nchw(const float *x, const float *filt, const int *filter_index, int NB, int c, int h, int w,
                      float *out, void *stream);

/* Hinge pixel weights (discriminator.py:107-152): y (K,H*W) in [0,1] -> w (K,H*W).  If threshold != 0 the map is
 * first binarised with y > 0.5 (discriminator.py:217).  workspace: K floats receiving the per-map pixel counts; when
 * `counts` (int[K], e.g. from frtm_merge_masks) is given the counting pass is skipped and workspace may be NULL. */
int frtm_pixel_weights(const float *y, int K, int HW, float tf, int threshold, float *w, float *workspace,
                       const int *counts, void *stream);

/* Stencil form of  U^T diag(pw^2) U  and  U^T (pw^2 * y)  for bilinear upsampling U : (h,w)->(H,W),
 * align_corners=False (discriminator.py:48).  pw,y (K,H,W);  stencil (K,9,h,w) [tap = (dy+1)*3+(dx+1) couples
 * pixel (i,j) with (i+dy,j+dx)],  uty (K,h,w).  Deterministic gather (no atomics). */
int frtm_build_stencil(const float *pw, const float *y, int K, int H, int W, int h, int w, float *stencil, float *uty,
                       void *stream);

/* Device-side sample-weight update + slot choice (memory.py:65-92), predicated on gate_count[0] >= min_px.
 * state = int[4]: {current_size, previous_replace_ind (-1 = none), last_slot (out, -1 if skipped), num_inserts}. */
int frtm_memory_next_slot(float *weights, int capacity, float lr, int *state, const int *gate_count, int min_px,
                          void *stream);
/* Copy one sample into slot state[2] of the frame memory (skipped when state[2] < 0) (memory.py:48-57).
 * mem_split (optional): the memory's operator images (frtm_split_samples layout), kept in step with the sample. */
int frtm_memory_insert(const float *feat, int feat_elems, const float *label, const float *pw, int HW,
                       const float *stencil, const float *uty, int hw, float *mem_samples, float *mem_labels,
                       float *mem_pw, float *mem_stencil, float *mem_uty, void *mem_split, const int *state,
                       void *stream);

/* All memory inserts of a track block (n_frames consecutive frames x n_obj objects) in two launches: the policy steps of
 * every object (sequential over its frames, frtm_memory_next_slot's arithmetic, gated per (frame, object) on
 * gate_counts[f * n_obj + o] >= min_px) and ONE copy launch for all samples (frtm_memory_insert's).  Row (f, o) of feat
 * (feat_elems floats), labels / pw (HW), stencil (9 hw), uty (hw) is row f * n_obj + o.  table = device int64[8][n_obj]:
 * per object the addresses of samples, labels, pixel_weights, stencil, uty, operator images (0 = none), sample weights,
 * policy state (model/memory.py:59-92 applied n_frames times).  with_fullres = 0: the labels / pixel_weights mirrors are
 * not written.  slots = int[n_frames * n_obj] workspace (out: the slots). */
int frtm_memory_insert_block(const void *table, int n_obj, int n_frames, int capacity, float lr, const int *gate_counts,
                             int min_px, const float *feat, int feat_elems, const float *labels, const float *pw, int HW,
                             const float *stencil, const float *uty, int hw, int with_split, int with_fullres, int *slots,
                             void *stream);

/* Operator images of n memory samples for the tensor-core GN/CG operator kernel.  Per sample:
 *   [ntiles][hi|lo][c][64 pixels] fp16 with 16*x = hi + lo, rows in the 128-byte swizzled shared-memory layout, pixels
 *   beyond hw zero, ntiles = hw/64 rounded up to an even count;  then  [ceil(hw/256)][10][256] fp32: the 9 stencil taps
 *   and U^T w^2 y of 256 consecutive pixels.  Same bytes per element as the fp32 arrays it mirrors.
 * samples (n,c,hw), stencil (n,9,hw), uty (n,hw);  frtm_split_sample_bytes = bytes of one image.  c % 8 == 0. */
int frtm_split_samples(const float *samples, const float *stencil, const float *uty, int n, int c, int hw, void *split,
                       void *stream);
int64_t frtm_split_sample_bytes(int c, int hw);

/* Filter-only Gauss-Newton / Polak-Ribiere CG update in closed (stencil) form — replaces
 * GaussNewtonCG.run on the update problem (optimizer.py:55-157, discriminator.py:38-64,221-227):
 *   residual  r = W (U (X * f) - y),  A p = X^T (U^T W^2 U) X p + reg^2 p,  b = -(X^T U^T W^2 (U X f - y) + reg^2 f)
 * samples (cap,c,h,w), stencil (cap,9,h,w), uty (cap,h,w), weights (cap) [inactive = 0];
 * samples_split (optional, NULL = absent): the operator images of the samples (frtm_split_samples); when given the
 * operator streams only the images and both contractions run on the tensor cores: in ONE pass with every sample held on
 * chip by a thread-block cluster (c == 96, 8 <= w <= 95 — gn_apply_cl.cu) or streamed through a sliding window by one CTA
 * (c == 96, 8 <= w <= 84 — gn_apply_mma.cu), or in two passes (tcgen05, c % 16 == 0, 48 <= c <= 128 — gn_apply_tc.cu);
 * otherwise on CUDA cores.  operator_select: 0 = the library picks by shape, 1 = CUDA cores, 2 = two-pass tcgen05,
 * 3 = single-pass sliding window, 4 = cluster (EINVAL if the shape is not supported) — a per-call argument for tests
 * and A-B measurements, the library keeps no mode;
 * filt (c*9) updated in place;  cg_state = float[2*c*9 + 4]: p | r_prev | rho | has_p — persists across calls
 * (zero-initialised by the caller);  cg_iters_host[n_gn] CG iterations per GN iteration (host array);
 * the update is applied only if gate_count == NULL or gate_count[0] >= min_px (device-side predicate, replaces the
 * host sync of discriminator.py:214).  workspace from frtm_gn_update_workspace.  No host synchronisation. */
int frtm_gn_update(const float *samples, const void *samples_split, const float *stencil, const float *uty,
                   const float *weights, int cap, int c, int h, int w, float *filt, float *cg_state,
                   const int *cg_iters_host, int n_gn, float reg, float precond, float forget, const int *gate_count,
                   int min_px, int operator_select, float *workspace, int64_t workspace_bytes, void *stream);
int64_t frtm_gn_update_workspace(int cap, int c, int h, int w);
/* Which operator kernel operator_select = 0 resolves to for this shape when the operator images are given:
 * 3 = single-pass sliding window, 2 = two-pass tcgen05, 1 = CUDA cores (the cluster kernel, 4, runs only when asked for). */
int64_t frtm_gn_operator_kind(int c, int h, int w);
/* The same update for n_obj objects in ONE set of launches (grid.y = object; the objects of a sequence update on the
 * same frames).  table: device int64[8][n_obj] of device pointers, rows = {samples, stencil, uty, weights, filt,
 * cg_state, gate_count (0 = ungated), samples_split (read only if has_split != 0)}; all objects share cap, c, h, w and
 * the schedule.  workspace >= n_obj times frtm_gn_update_workspace. */
int frtm_gn_update_batched(const void *table, int n_obj, int has_split, int cap, int c, int h, int w,
                           const int *cg_iters_host, int n_gn, float reg, float precond, float forget, int min_px,
                           int operator_select, float *workspace, int64_t workspace_bytes, void *stream);

/* Joint (project, filter) Gauss-Newton / CG of Discriminator.init (discriminator.py:154-175; optimizer.py:55-157):
 *   s = F * (P x),  J[dP,dF] = F * (dP x) + dF * (P x)  on the low-resolution grid, normal equations through the
 *   same stencil form.  x_nhwc (K,h,w,C) raw backbone features, stencil (K,9,h,w), uty (K,h,w), sw (K);
 *   P (c,C) = project.weight, F (c*9) = filter.weight, both updated in place.  Fresh CG state per call. */
int frtm_gn_init(const float *x_nhwc, const float *stencil, const float *uty, const float *sw, int K, int C, int c, int h,
                 int w, float *P, float *F, const int *cg_iters_host, int n_gn, float regP, float regF, float precondP,
                 float precondF, float forget, float *workspace, int64_t workspace_bytes, void *stream);
int64_t frtm_gn_init_workspace(int K, int C, int c, int h, int w);
/* frtm_gn_init replays its fixed launch schedule as a CUDA graph when it is called again with the same buffers.  The
 * library keeps at most 16 such graphs (least recently used first out); a caller that frees the buffers of a signature
 * calls this with that signature's workspace pointer first (NULL forgets every graph). */
int frtm_gn_init_release(const void *workspace);
/* Teacher-forced probe of the joint problem at (P,F): out_b = -(J^T r0 + reg^2 theta), out_A = J^T J d + reg^2 d. */
int frtm_gn_init_probe(const float *x_nhwc, const float *stencil, const float *uty, const float *sw, int K, int C, int c,
                       int h, int w, float *P, float *F, const float *dP, const float *dF, float regP, float regF,
                       float *out_bP, float *out_bF, float *out_AP, float *out_AF, float *workspace,
                       int64_t workspace_bytes, void *stream);
/* v[n] = sw[n] * (S[n] s[n] - use_y * uty[n])   9-tap spatially varying stencil, all (NB,h,w). */
int frtm_stencil_apply(const float *stencil, const float *s, const float *uty, const float *sw, int NB, int h, int w,
                       int use_y, float *v, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * First-frame augmentation rendering  (model/augmenter.py:342-396, lib/image.py:38-59, lib/_npp/nppig.cpp:48-104)
 * ---------------------------------------------------------------------------------------------------------- */

/* Affine warp of a (C,H,W) uint8 or float image into (C,Ho,Wo).  M_host: 6 doubles (row-major 2x3) mapping source ->
 * destination exactly as cv2.warpAffine / nppiWarpAffine take it (the kernel samples at the inverse).  Bicubic
 * (a = -0.75) or nearest; samples outside the source read 0; the bicubic result is clamped to [clamp_lo, clamp_hi]
 * (the reference clamps to [0,255] right after the warp, augmenter.py:357,384).  Replaces nppig.cpp warp_affine. */
int frtm_warp_affine(const void *src, int src_is_u8, int C, int H, int W, float *dst_f32, uint8_t *dst_u8, int Ho, int Wo,
                     const double *M_host, int nearest, float clamp_lo, float clamp_hi, void *stream);
/* Nearest-neighbour warp of a (H,W) uint8 mask, bit-identical to cv2.warpAffine(..., INTER_NEAREST) (OpenCV's 1/1024
 * fixed-point coordinates), and the number of output pixels equal to count_value added to *count (device int, zeroed
 * by the caller) — the visibility test of augmenter.py:453-471.  lib/image.py:53 with mode 'nearest'. */
int frtm_warp_mask_nearest(const uint8_t *src, int H, int W, uint8_t *dst, int Ho, int Wo, const double *M_host,
                           int count_value, int *count, void *stream);
/* The same warp for n transforms of ONE source mask in a single launch (all candidate masks of an augmentation round,
 * model/augmenter.py:453-471,520-533): M_host holds n row-major 2x3 matrices, dst is (n,Ho,Wo), counts[n] (zeroed by the
 * caller) receive the per-mask pixel counts.  The matrices travel as kernel arguments: no upload, no synchronisation. */
int frtm_warp_mask_nearest_batch(const uint8_t *src, int H, int W, uint8_t *dst, int Ho, int Wo, int n, const double *M_host,
                                 int count_value, int *counts, void *stream);
/* Per-channel 2-D cross-correlation with zero padding kh/2, kw/2 (the directional blur, augmenter.py:342-350). */
int frtm_filter2d(const float *src, int C, int H, int W, const float *kernel, int kh, int kw, float *dst, void *stream);
/* out (3,H,W) uint8 = rgba[:3] * a + canvas * (1 - a), a = rgba[3] / 255, truncated (augmenter.py:391-394). */
int frtm_alpha_paste(const float *rgba, const float *canvas, int H, int W, uint8_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FRTM_B200_H */
