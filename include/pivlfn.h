/*
 * pivlfn.h -- C ABI of the B200 (sm_100a) PIV-LiteFlowNet forward-pass kernels.
 *
 * Every entry point takes raw DEVICE pointers, plain ints/floats and a CUDA stream
 * (cudaStream_t passed as void*), allocates nothing, and returns 0 on success,
 * a negative PIVLFN_E* code on a validation error, or a positive cudaError_t.
 * Tensors are fp32.  "NHWC view" = (ptr, C, ld): pixel p's channels start at
 * ptr + p*ld floats, ld >= C, so a view can be a channel slice of a wider buffer
 * (that is how torch.cat of the reference disappears).
 *
 * Each function names the reference code (file:line under abrosua/piv_liteflownet-pytorch)
 * whose arithmetic it replaces.
 */
#ifndef PIVLFN_H
#define PIVLFN_H

#ifdef __cplusplus
extern "C" {
#endif

#define PIVLFN_OK 0
#define PIVLFN_EINVAL (-1)      /* bad shape / stride / alignment argument         */
#define PIVLFN_EUNSUPPORTED (-2) /* valid request this build has no kernel for       */
#define PIVLFN_EDRIVER (-3)     /* CUDA driver entry point (tensor-map encode) unavailable */

/* Library / device introspection. */
int pivlfn_abi_version(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM/TMA paths usable). */
int pivlfn_device_is_sm100(void);
/* Number of kernels launched by this library since load (all entry points). */
long long pivlfn_launch_count(void);

/* ---- drop-in operator: FunctionCorrelation(tensorFirst, tensorSecond, intStride) ----------
 * src/correlation.py:285-344 (_FunctionCorrelation.forward) with its two kernels
 * kernel_Correlation_rearrange (:9-34) and kernel_Correlation_updateOutput (:36-104).
 * first, second: [B,C,H,W] contiguous NCHW.  out: [B,49,ceil(H/s),ceil(W/s)] NCHW.
 * out[b,(dy+3)*7+(dx+3),y,x] = (1/C) sum_c first[b,c,y*s,x*s] * second[b,c,y*s+dy*s,x*s+dx*s]. */
int pivlfn_corr_nchw(const float* first, const float* second, float* out,
                     int B, int C, int H, int W, int stride, void* stream);

/* ---- model-internal operators (NHWC views) ---------------------------------------------- */

/* Gradients of pivlfn_corr_nchw (src/correlation.py:106-234,348-405; training only): grad_out [B,49,ceil(H/s),ceil(W/s)] contiguous
 * -> grad_first / grad_second [B,C,H,W] (either may be NULL: not computed).  At stride 2 only the even positions the forward samples
 * receive a gradient, the others are written as zero (the reference allocates them with new_zeros). */
int pivlfn_corr_backward_nchw(const float* first, const float* second, const float* grad_out,
                              float* grad_first, float* grad_second, int B, int C, int H, int W, int stride, void* stream);

/* src/models.py:321-323 (in-place per-channel mean subtraction of BOTH caller tensors) fused with
 * the NCHW->NHWC pack.  img1,img2: [B,3,H,W] NCHW, modified in place.  out: [2B,H,W,4] (images of
 * img1 first, then img2; 4th channel zero).  out_pad (optional): [2B,H,W+8,4] copy with a 4-pixel border left
 * and right that the caller zeroed once (input of pivlfn_conv_stem_tc).  mean6: HOST pointer to 6 floats. */
int pivlfn_prep_images(float* img1, float* img2, float* out_nhwc4, float* out_pad, int B, int H, int W,
                       const float* mean6, void* stream);

/* src/models.py:336-343: bilinear 1/2 downsample with align_corners=False on even sizes
 * == 2x2 mean.  in: [N,H,W,C] dense, out: [N,H/2,W/2,C] dense.  H, W even. */
int pivlfn_avgpool2(const float* in, float* out, int N, int H, int W, int C, void* stream);

/* torch.cat along channels (src/models.py:216,280) done as a strided slice copy: copies C channels of
 * npix pixels from one NHWC view to another.  Views must be 16-byte aligned with ld % 4 == 0 when
 * C % 4 == 0 (vector path); any C otherwise. */
int pivlfn_copy_nhwc(const float* in, int in_ld, float* out, int out_ld, long long npix, int C, void* stream);

/* torch.nn.Conv2d (+ LeakyReLU(0.1)) as used throughout src/models.py:70-106,123-126,154-163,
 * 197-207,228-272.  Zero padding (KH/2, KW/2), stride 1 or 2, generic CUDA-core kernel.
 * w: [KH*KW*Cin, CoutP] row-major, CoutP = Cout rounded up to a multiple of 4, zero padded.
 * y = act(conv(x) + bias) (+ res).  res (optional) is an NHWC view with Cout channels
 * (the "+ xflow" of src/models.py:186,216). */
int pivlfn_conv_simt(const float* x, int x_ld, int N, int H, int W, int Cin,
                     const float* w, const float* bias, float* y, int y_ld, int Cout,
                     int KH, int KW, int stride, int lrelu,
                     const float* res, int res_ld, void* stream);

/* Same operator for every convolution with odd KH, KW <= 7 and stride 1 or 2 on the tcgen05 tensor cores: implicit GEMM,
 * TMA-fed (zero padding = TMA out-of-bounds fill), TMEM accumulators, fused bias + LeakyReLU (+ residual).
 * w_hi / w_lo: [CoutP, KH*KW, CinP] (CoutP = Cout rounded up to 16, CinP = Cin rounded up to 32, zero padded):
 * TF32 split of the weights (w ~= w_hi + w_lo); passes = 1 (plain TF32), 3 (error-compensated 3xTF32 ~ fp32),
 * 2 (TF32 main product + the two low-order products in bf16: same accuracy class, 2/3 of the tensor-core work) or
 * 4 (all three products in fp16, operands split as x = f16(x) + 2^-11 * f16((x - f16(x)) * 2^11): same accuracy class,
 * half the tensor-core work of passes == 3; activations must stay inside the fp16 range, see pivlfn_f16_range_flag).
 * w_c16 (passes == 2 / 4 only, else NULL): 16-bit pack, two tensors of [CoutP, KH*KW, CinP] each:
 * passes == 2: [bf16(w) | bf16(w - w_hi)];  passes == 4: [f16(w) | f16((w - f16(w)) * 2048)];
 * passes == 5: the passes == 4 arithmetic with ONE accumulator per tile (used for Cout > 64, where two accumulators per
 * tile would leave no TMEM for double buffering), THREE tiles of the pre-scaled weights W = 256 w:
 * [f16(W) | f16(W - f16(W)) | f16(f16(W) / 2048)]; the kernel multiplies the result by 1/256.
 * Storage order of w_c16: passes == 2: the tiles one after the other, [part][CoutP][KH*KW][CinP].  passes == 4 / 5: the
 * shared-memory image of the kernel's weight ring, so that one ring stage is one contiguous bulk copy:
 * [CinP/32 chunks][KH*KW taps][part][CoutP rows][64 bytes = 32 channels], the four 16-byte units of row r stored at
 * unit ^ ((r >> 1) & 3) (the 64B swizzle of the MMA operand layout); pivlfn.model.stage_image builds it.
 * Requirements: x 16-byte aligned, x_ld % 4 == 0, Cout <= 128. */
int pivlfn_conv_tc(const float* x, int x_ld, int N, int H, int W, int Cin,
                   const float* w_hi, const float* w_lo, const void* w_c16, const float* bias,
                   float* y, int y_ld, int Cout, int KH, int KW, int stride, int lrelu,
                   const float* res, int res_ld, int passes, void* stream);

/* The 3x3 stride-2 convolutions of NetC (src/models.py:77-106) in the fp16 modes (passes 4 or 5) of pivlfn_conv_tc, restated as
 * a 2x2-tap stride-1 convolution over the four pixel-parity phases of the input; TMA element strides fetch a phase directly,
 * so no space-to-depth copy exists.  H, W: INPUT size (even), output [N,H/2,W/2,Cout]; Cin % 32 == 0; W/2 >= 8.
 * w16: the passes-4 / passes-5 16-bit pack (ring-stage image, see pivlfn_conv_tc) of the restated weights [CoutP][4][4*Cin], tap = (by+1)*2 + (bx+1),
 * channel = (py*2 + px)*Cin + c, where input row 2y + ky - 1 = 2(y + by) + py (zero weights for (by, py) = (-1, 0)). */
int pivlfn_conv_s2_tc(const float* x, int x_ld, int N, int H, int W, int Cin, const void* w16, const float* bias,
                      float* y, int y_ld, int Cout, int lrelu, int passes, void* stream);

/* passes == 4 convolutions convert their input activations to fp16 pairs; a value outside the fp16 range (|x| > 65504,
 * or non-finite) raises a sticky device flag instead of being silently saturated.  Returns the flag (0 / 1, -1 on a
 * CUDA error) and clears it when reset != 0.  Synchronises the device.  The host re-runs with passes == 2 when set. */
int pivlfn_f16_range_flag(int reset);
/* Clears the same flag of the CURRENT device in stream order (cudaMemsetAsync, no synchronisation): issued before a forward so
 * that a stale hit left by another user of the library cannot be attributed to it. */
int pivlfn_f16_range_flag_clear(void* stream);

/* Flow heads, the last layer of conv_M / conv_S (src/models.py:161,205; LiteFlowNet2 :499,547): KxK convolution
 * (K = 3, 5 or 7) from Cin = 32 channels to the 2 flow components, no activation, + bias + optional residual flow
 * ("+ xflow", src/models.py:186,216), in exact fp32 on the CUDA cores (3136 FMA per pixel at K = 7: the halo tile and all
 * weights stay in shared memory, ~9 FMA per shared-memory load).  x: NHWC view, 16-byte aligned, x_ld % 4 == 0.
 * w: [K*K][32][2] (tap, input channel, flow component), 16-byte aligned.  res / out: NHWC views with >= 2 channels.
 * out2 (optional): a second NHWC view that receives the same result (the flow slice of the Subpixel concat buffer,
 * src/models.py:216, so that the torch.cat copy disappears). */
int pivlfn_flow_head(const float* x, int x_ld, int N, int H, int W, int Cin, const float* w, const float* bias,
                     const float* res, int res_ld, float* out, int out_ld, float* out2, int out2_ld, int K, void* stream);

/* The same layer on the tensor cores (kept as an alternative, PIVLFN_HEAD=pairs), restated in two steps so that
 * the tensor cores see N = 2*K*K useful columns instead of 2:
 *   pivlfn_conv1x1_pairs_tc:  D[pixel, tap*2+co] = sum_c x[pixel,c] * w[co,c,tap]  (1x1 convolution, no bias / activation)
 *                             stored as tap planes  planes[tap][pixel][2]
 *   pivlfn_flow_head_sum:     out[p,co] = bias[co] + res[p,co] + sum_tap planes[tap][p + offset(tap)][co], zero outside.
 * w_hi / w_lo / w_c16: packed like pivlfn_conv_tc for a 1x1 convolution with Cout = 2*npair rows (row = tap*2 + co). */
int pivlfn_conv1x1_pairs_tc(const float* x, int x_ld, int N, int H, int W, int Cin,
                            const float* w_hi, const float* w_lo, const void* w_c16,
                            float* planes, int npair, int passes, void* stream);
int pivlfn_flow_head_sum(const float* planes, int K, const float* bias, const float* res, int res_ld,
                         float* out, int out_ld, int N, int H, int W, void* stream);

/* NetC.conv1 (src/models.py:70-73): 7x7, 3 -> 32, stride 1, on the tensor cores.  img_pad: [N,H,W+8,4], the
 * zero-bordered NHWC4 image written by pivlfn_prep_images (pixel x at column x+4).  One GEMM-K row of 32 floats
 * per filter row = the 8 pixels x-3..x+4, fetched through an overlapping-window tensor map.
 * w_hi / w_lo: [32, 7, 32] with column kx*4 + c (kx = 7 and c = 3 are zero). */
int pivlfn_conv_stem_tc(const float* img_pad, int N, int H, int W,
                        const float* w_hi, const float* w_lo, const void* w_c16, const float* bias,
                        float* y, int y_ld, int lrelu, int passes, void* stream);

/* torch.nn.ConvTranspose2d(C, C, 4, stride 2, padding 1, groups=C, bias=False):
 * upConv_M (src/models.py:144-145, C=2) and upCorr_M (:151-152, C=49).
 * in: NHWC view [N,H,W,C], w: [C,4,4], out: NHWC view [N,2H,2W,C]. */
int pivlfn_deconv4x4s2_dw(const float* in, int in_ld, const float* w, float* out, int out_ld,
                          int N, int H, int W, int C, void* stream);

/* backwarp (src/models.py:20-35) as a standalone operator: out[p,c] = bilinear(in[.,c], p + scale*flow[p]),
 * zeros outside.  flow: [N,H,W,2] dense (u,v). */
int pivlfn_warp_nhwc(const float* in, int in_ld, const float* flow, float scale,
                     float* out, int out_ld, int N, int H, int W, int C, void* stream);

/* Matching cost volume (src/models.py:169-184): optional backwarp of f2 by scale*flow fused into the
 * tile load (the warped features never reach HBM), 49-displacement correlation at stride 1 or 2
 * (src/correlation.py:36-104), /C, then LeakyReLU(0.1).  flow may be NULL (level 6).
 * out: NHWC view [N,ceil(H/s),ceil(W/s),49]; when out_ld == 52 (a dedicated buffer padded to 4 floats) the three pad channels
 * are written as zeros (whole float4 rows). */
int pivlfn_corr_nhwc(const float* f1, int f1_ld, const float* f2, int f2_ld,
                     const float* flow, float flow_scale, float* out, int out_ld,
                     int N, int H, int W, int C, int stride, int lrelu, void* stream);

/* Per-sample, per-channel spatial sums of a dense [N,H,W,2] flow in G deterministic partials
 * (src/models.py:275, the .mean(2)).  partial: [N,G,2].  G = pivlfn_flow_mean_parts(). */
int pivlfn_flow_mean_parts(void);
int pivlfn_flow_mean(const float* flow, float* partial, int N, int H, int W, void* stream);

/* Regularization inputs (src/models.py:275-277): rm = flow - mean(flow); brightness error
 * sqrt(sum_c (img1 - backwarp(img2, scale*flow))^2).  Writes 3 channels (err, rm_u, rm_v) at out.
 * img1,img2: [N,H,W,4] dense. */
int pivlfn_reg_input(const float* img1, const float* img2, const float* flow, float scale,
                     const float* partial, float* out, int out_ld, int N, int H, int W, void* stream);

/* Regularization tail (src/models.py:281-302): d = exp(-x^2 - max(-x^2)); out_u = (sum_k wx_k d_k u_N(k) + bx)
 * / sum_k d_k, same for v; N(k) = KxK zero-padded neighbourhood (F.unfold order).  dist: NHWC view with
 * K*K channels.  flow_in/flow_out: [N,H,W,2] dense.  If out_nchw != NULL also writes
 * final_scale * flow as [N,2,H,W] (src/models.py:370). */
int pivlfn_reg_tail(const float* dist, int dist_ld, const float* flow_in,
                    const float* wx, const float* bx, const float* wy, const float* by,
                    float* flow_out, float* out_nchw, float final_scale,
                    int K, int N, int H, int W, void* stream);

/* F.interpolate(mode='bilinear', align_corners=False) of inference.py:46-49 (images up to a multiple of 32)
 * and :57-61 (flow back to the input size, with u *= W/W' and v *= H/H' folded in as mul_even / mul_odd,
 * applied to channels of even / odd index).  in: [NC,H,W], out: [NC,Ho,Wo], both dense NCHW planes. */
int pivlfn_resize_bilinear_nchw(const float* in, float* out, int NC, int H, int W, int Ho, int Wo,
                                float mul_even, float mul_odd, void* stream);

/* ---- the P16 pipeline (default for precision f16c) -----------------------------------------------------------------------
 * P16 is the activation format of the split-operand arithmetic: an fp32 value x is stored as the operands the tensor cores
 * consume, hi = f16(x), lo8 = e5m2((x - hi) * 2^11), hi8 = e5m2(x) (x = hi + lo8 / 2048 up to 2^-14 |x|, |x| < 65504), 4 bytes
 * per element (csrc/p16.cuh).  Channels are grouped by 16; one group of one pixel is 64 contiguous bytes
 * [16 x hi (f16) | 16 x lo8 | 16 x hi8], so that a view
 * (pointer, pixel pitch in 4-byte words, channel count rounded up to 16) addresses a P16 tensor like an fp32 NHWC one,
 * channel slices start at multiples of 16 and a 32-channel chunk of a pixel is one 128-byte MMA-ready row.  Pad channels
 * of the last group are zero.  P16 pointers are 64-byte aligned, pitches multiples of 16 words.
 * range_flag (optional device int): set to 1 by any producer whose result is not finite in fp16 (|x| >= 65520 or NaN);
 * the host then repeats the forward in the fp32-activation pipeline (precision tf32c).  Never cleared by the library. */

/* fp32 NHWC view (C channels, pitch x_ld) <-> P16 view (pitch y_ld words). */
int pivlfn_p16_encode(const float* x, int x_ld, int C, void* y, int y_ld, long long npix, int* range_flag, void* stream);
int pivlfn_p16_decode(const void* x, int x_ld, int C, float* y, int y_ld, long long npix, void* stream);

/* Convolution + bias + LeakyReLU on a P16 input (src/models.py:77-106,124,154-163,197-207,229-272): implicit GEMM on
 * tcgen05 with NO operand split inside the kernel: TMA delivers MMA-ready halo tiles.  Per 16 input channels the kernel issues
 * a_hi * W_hi as one kind::f16 MMA (K = 16) and both correction products, [lo8 | hi8] * [W 2^-11 ; W - W_hi], as ONE
 * kind::f8f6f4 MMA (K = 32; activations e5m2, weights e4m3) into the same fp32 accumulator.  Odd KH, KW <= 7 at stride 1; 3x3 at stride 2 (H, W = INPUT size, even; Cin % 32 == 0;
 * w_img = pack of the parity-restated weights, see pivlfn_conv_s2_tc).  Cin: logical input channels (the buffer holds
 * ceil16(Cin) words per pixel).  mode: 6 (the only product scheme).  w_img: the weight image built by pivlfn.model._pack_f8 --
 * per 32-channel chunk and tap [f16(S w) tile | e4m3 correction tile], CoutP rows of 64 bytes each with the 64B swizzle applied
 * (one ring stage = one contiguous bulk copy), then a 16-byte trailer {1 / S, S, 0, 0} (fp32; S = the layer's power-of-two
 * weight scale, applied by the epilogue).  Cout <= 128.
 * out_fmt 0: P16 view (pitch y_ld words, >= ceil16(Cout); pad channels written as zeros);
 *         1: fp32 NHWC view (pitch y_ld floats);
 *         2: fp32 channel-PAIR planes, plane p = channels (2p, 2p+1) as [pixel][2], plane_stride floats apart. */
int pivlfn_conv_p16(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, int mode,
                    const float* bias, void* y, int y_ld, int Cout, int KH, int KW, int stride, int lrelu,
                    int out_fmt, long long plane_stride, int* range_flag, void* stream);

/* pivlfn_conv_p16 (stride 1, P16 output) with a BACKWARP FUSED INTO ITS INPUT (src/models.py:209-217, the Subpixel consumer):
 * of the Cin logical input channels, [wc0, wc0 + wn) are backwarp(wsrc, wscale * wflow) (src/models.py:20-35) and do not exist in
 * memory: the kernel's gather warps sample wsrc (fp32 NHWC with pitch wsrc_ld floats, or P16 with wsrc_p16 = 1), blend, split
 * into the P16 operands and write the MMA operand tile directly.  x holds the other Cin - wn channels contiguously ([0, wc0) then
 * [wc0 + wn, Cin)).  wc0 and wn are multiples of 32; wflow: dense [N,H,W,2]. */
int pivlfn_conv_p16_warp(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, int mode,
                         const float* bias, void* y, int y_ld, int Cout, int KH, int KW, int lrelu,
                         const void* wsrc, int wsrc_ld, int wsrc_p16, const float* wflow, float wscale,
                         int wc0, int wn, int* range_flag, void* stream);

/* pivlfn_conv_p16 (stride 1, no activation) whose K*K output channels are the Regularization distances, with the rest of the
 * Regularization block FUSED INTO ITS EPILOGUE (src/models.py:279-300 (PIV) / :620-641 (Hui); replaces conv_dist + pivlfn_reg_tail): per pixel
 * softmax_k(-d_k^2) over the K*K channels read straight from the accumulators, the weighted K x K unfold of flow_in, the 1x1
 * moduleScaleX / moduleScaleY convolutions (wx, bx, wy, by: device pointers to K*K weights / 1 bias each) and the division.
 * flow_in, flow_out: dense [N,H,W,2] (distinct buffers); out_nchw: optional [N,2,H,W] copy times final_scale (the network
 * output), or NULL.  K = 3, 5, 7; Cout = K*K is implied.  Same arithmetic in the same order as pivlfn_conv_p16(out_fmt 1) + pivlfn_reg_tail (agrees to rounding). */
int pivlfn_conv_p16_tail(const void* x, int x_ld, int N, int H, int W, int Cin, const void* w_img, const float* bias,
                         int KH, int KW, int K, const float* flow_in, const float* wx, const float* bx, const float* wy,
                         const float* by, float* flow_out, float* out_nchw, float final_scale, void* stream);

/* NetC.conv1 (see pivlfn_conv_stem_tc) with the fp16-split arithmetic and a P16 output; w_img: stage image of the
 * passes-4 pack of the [32, 7, 32] stem weights.  W % 4 == 0, W >= 8. */
int pivlfn_conv_stem_p16(const float* img_pad, int N, int H, int W, const void* w_img, const float* bias,
                         void* y, int y_ld, int lrelu, int* range_flag, void* stream);

/* pivlfn_corr_nhwc with each of f1 / f2 / out either fp32 NHWC (flag 0) or P16 (flag 1; out: 64 channels, 49 + zero pad). */
int pivlfn_corr_p16(const void* f1, int f1_ld, int f1_p16, const void* f2, int f2_ld, int f2_p16,
                    const float* flow, float flow_scale, void* out, int out_ld, int out_p16,
                    int N, int H, int W, int C, int stride, int lrelu, int* range_flag, void* stream);

/* pivlfn_warp_nhwc writing a P16 view (C % 16 == 0); the input is fp32 NHWC (in_p16 = 0) or P16 (in_p16 = 1). */
int pivlfn_warp_p16(const void* in, int in_ld, int in_p16, const float* flow, float scale, void* out, int out_ld,
                    int N, int H, int W, int C, int* range_flag, void* stream);

/* pivlfn_deconv4x4s2_dw (upCorr_M) from an fp32 NHWC view (16-byte aligned, in_ld % 4 == 0) to a P16 view; C <= 64. */
int pivlfn_deconv4x4s2_dw_p16(const float* in, int in_ld, const float* w, void* out, int out_ld,
                              int N, int H, int W, int C, int* range_flag, void* stream);

/* pivlfn_reg_input writing (err, rm_u, rm_v) as the first three channels of one P16 group at out (the other 13 are
 * left untouched: the concat buffer is zero-initialised once). */
int pivlfn_reg_input_p16(const float* img1, const float* img2, const float* flow, float scale,
                         const float* partial, void* out, int out_ld, int N, int H, int W, int* range_flag, void* stream);

/* Second half of the tensor-core flow head (src/models.py:161,205): the KxK 32 -> 2 head runs as a 1xK convolution to 2K
 * channels (row ky*2 + co; pivlfn_conv_p16 with out_fmt 2 -> K planes [pixel][2]); this adds the K planes at their vertical
 * offsets (zero outside the frame) + bias + residual flow (res, dense [N,H,W,2], may be NULL) -> out (dense [N,H,W,2]) and
 * optionally the flow's P16 group (2 channels + untouched pad) at out_p16 (the torch.cat of src/models.py:216). */
int pivlfn_head_rows_sum(const float* planes, int K, const float* bias, const float* res, float* out,
                         void* out_p16, int p16_ld, int N, int H, int W, int* range_flag, void* stream);

/* The transposed split of the same head: a Kx1 convolution to 2K channels (column kx*2 + co), whose K planes are added at their
 * HORIZONTAL offsets (the Kx1 halo tile is 8 pixels wide instead of 8 + K - 1: the convolution runs 1.6x faster). */
int pivlfn_head_cols_sum(const float* planes, int K, const float* bias, const float* res, float* out,
                         void* out_p16, int p16_ld, int N, int H, int W, int* range_flag, void* stream);

/* ---- stereo-PIV post-processing (stereo_run.py:104-163) ------------------------------------------------------------------
 * stereo/dewarp.py:255-270 nl_trans: new_x = P(A[0:6]) / P(A[6:12]), new_y = P(A[12:18]) / P(A[18:24]) with
 * P(a) = a0 x + a1 y + a2 + a3 x^2 + a4 y^2 + a5 x y, evaluated in float32 in the reference's operation order.
 * x, y, new_x, new_y: n device floats; A24: 24 HOST floats. */
int pivlfn_nl_trans(const float* x, const float* y, const float* A24, float* new_x, float* new_y, long long n, void* stream);

/* stereo_run._stereo_cal (:153-163) on both camera flows + stereo/vel3d.py:4-24 willert, fused.  flow_left / flow_right:
 * [B,2,H,W] device (what estimate(..., tensor=True) returns); A_left / A_right: 24 HOST floats each, or both NULL to skip the
 * mapping (willert only); use_calib: multiply the mapped flows by calib and then by fps (two float32 products);
 * tan_*: np.tan of the signed camera angles (theta = off-axis half angle, beta = off-axis half angle in the y-z plane) as
 * float64 scalars; like the reference under its pinned numpy 1.17 they (and their float64 differences) are rounded to float32
 * where they meet the float32 flows.  out: [B,H,W,3] device float32 (U, V, W), the 3-band .flo layout. */
int pivlfn_stereo_2d3c(const float* flow_left, const float* flow_right, const float* A_left, const float* A_right,
                       int use_calib, float calib, float fps, double tan_theta0, double tan_theta1,
                       double tan_beta0, double tan_beta1, float* out, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIVLFN_H */
