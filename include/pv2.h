/*
 * pv2.h -- C ABI of libpranetv2_b200.so: the sm_100a kernels behind the DSRA decoder head and the
 * dual-supervision losses of PraNet-V2.
 *
 * The reference (ai4colonoscopy/PraNet-V2) has no native layer: every op on this path is a stock ATen
 * call made from Python.  Each entry point below therefore cites the reference *Python call site* it
 * replaces (paths relative to the reference root).  A maintainer binds these with ctypes (see
 * INTEGRATION.md); `pranet_v2_b200/_lib.py` is that binding.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers; the small arrays OF pointers (`const void* const* pred`)
 *     are HOST arrays read at call time; tensors are dense NCHW unless stated;
 *   - `stream` is a cudaStream_t passed as void*; every launch goes to that stream, nothing
 *     synchronises, allocates or frees: the caller owns all memory including workspaces;
 *   - return value 0 = ok, anything else = error; pv2_last_error() gives the message (thread local);
 *   - dtype codes: PV2_F32 = 0, PV2_BF16 = 1, PV2_TF32 = 2 (operand kind only).
 */
#ifndef PV2_H_
#define PV2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV2_F32 0
#define PV2_BF16 1
#define PV2_TF32 2   /* fp32 storage read by the tensor cores as tf32; optionally split into hi + lo planes */

#define PV2_MAX_SCALES 4

int pv2_version(void);
const char* pv2_last_error(void);
/* number of kernels this library launched since load (all streams); bench.py reports it */
unsigned long long pv2_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * structure_loss -- binary_seg/MyTrain_med.py:19-38 (called 4x per step, :78-81)
 *
 *   weit = 1 + 5*|avg_pool31x31(mask_fg) - mask_fg|          (zero padding counted: /961)
 *   loss_k = mean_{n,c}[ wbce(pred_k, mask_fg) + wiou(pred_k, mask_fg) + 0.8*wbce(pred_bg_k, mask_bg) ]
 *
 * One launch evaluates `nscales` (1..4) (pred, pred_bg) pairs against the SAME mask, sharing the
 * boundary weight.  mask_bg may be NULL, meaning 1 - mask_fg (what the reference's caller passes,
 * MyTrain_med.py:74).  planes = N*C.  logit_dtype applies to pred/pred_bg and to the gradients.
 *
 *   fwd : writes loss[k] (float, k < nscales) and plane_sums (workspace the backward reads).
 *   bwd : dpred_k = grad_loss[k] * dloss_k/dpred_k, same for dpred_bg_k.  grad_loss is a DEVICE
 *         pointer to nscales floats (the upstream gradient of each scalar loss).
 * ------------------------------------------------------------------------------------------- */
size_t pv2_structure_loss_workspace_bytes(int planes, int H, int W, int nscales);
int pv2_structure_loss_fwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                           const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                           float* loss, void* workspace, size_t workspace_bytes, void* stream);
/* The 31x31 boundary weight depends on the mask only (MyTrain_med.py:21): pv2_structure_loss_prepare computes the 16-bit weight map
 * into `workspace` ahead of time (a training step runs it on a side branch under the backbone), and pv2_structure_loss_fwd_prepared
 * -- same arguments as pv2_structure_loss_fwd, same workspace -- is then a pure stream over logits, mask and weight map.
 * pv2_structure_loss_bwd works after either forward. */
int pv2_structure_loss_prepare(const float* mask_fg, int planes, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
int pv2_structure_loss_fwd_prepared(const void* const* pred, const void* const* pred_bg, const float* mask_fg, const float* mask_bg,
                                    int nscales, int planes, int H, int W, int logit_dtype, float* loss, void* workspace,
                                    size_t workspace_bytes, void* stream);
int pv2_structure_loss_bwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                           const float* mask_bg, const float* grad_loss, void* const* dpred,
                           void* const* dpred_bg, int nscales, int planes, int H, int W, int logit_dtype,
                           const void* workspace, size_t workspace_bytes, void* stream);

/* structure_loss computed FROM THE LOW-RESOLUTION head maps: the final
 * F.interpolate(scale_factor=8|16|32, mode='bilinear', align_corners=False) of binary_seg/lib/pranet.py:349-350,
 * 370-371,392-393,414-415 and the four structure_loss calls of MyTrain_med.py:78-82 as ONE forward launch and one
 * backward launch (+ a fold): the eight full-resolution logit maps and their gradients never exist in HBM.
 *   low_fg[k], low_bg[k] : fp32 (planes, ih[k], iw[k]) maps BEFORE their final upsample; rh[k], rw[k] = ATen's source-index
 *                          ratios of that upsample (1/scale_factor); all HOST arrays of nscales entries.
 *   mask_fg / mask_bg    : fp32 (planes, H, W) as for pv2_structure_loss_fwd (mask_bg NULL = 1 - mask_fg).
 *   fwd writes loss[k]; bwd writes dlow_fg[k], dlow_bg[k] (same shapes as the maps) = grad_loss[k] * dloss_k/dmap.
 * Results equal pv2_bilinear_fwd -> pv2_structure_loss_fwd/bwd -> pv2_bilinear_bwd up to fp32 summation order; the
 * backward is deterministic (no atomics).  Covered: W % 4 == 0, 16-byte aligned masks, ratios <= 1/4 (up-scaling by >= 4);
 * anything else returns an error and the caller takes the unfused pv2 kernels.  The workspace must be the one the forward used. */
size_t pv2_structure_loss_lowres_workspace_bytes(int planes, int H, int W, int nscales);
int pv2_structure_loss_lowres_fwd(const float* const* low_fg, const float* const* low_bg, const int* ih, const int* iw,
                                  const float* rh, const float* rw, const float* mask_fg, const float* mask_bg,
                                  int nscales, int planes, int H, int W, float* loss, void* workspace, size_t workspace_bytes, void* stream);
int pv2_structure_loss_lowres_bwd(const float* const* low_fg, const float* const* low_bg, const int* ih, const int* iw,
                                  const float* rh, const float* rw, const float* mask_fg, const float* mask_bg,
                                  const float* grad_loss, float* const* dlow_fg, float* const* dlow_bg,
                                  int nscales, int planes, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * bilinear resize of NCHW planes -- F.interpolate(mode='bilinear') at binary_seg/lib/pranet.py:349-415
 * (x8/x16/x32 final maps, x0.25 / x2 crops, align_corners=False), nn.Upsample(scale_factor=2,
 * align_corners=True) at pranet.py:93, and the size= form of EMCAD/lib/decoders.py:460-461.
 * rh/rw are the source-index ratios exactly as ATen derives them (1/scale_factor when a scale factor
 * was given, in/out otherwise; (in-1)/(out-1) for align_corners).  dtype covers in and out.
 *   fwd: out[p,oy,ox] = sum of the 4 taps;  bwd: din = transpose(fwd) applied to dout (gather form,
 *   no atomics, deterministic).
 * ------------------------------------------------------------------------------------------- */
int pv2_bilinear_fwd(const void* in, void* out, int planes, int ih, int iw, int oh, int ow,
                     float rh, float rw, int align_corners, int dtype, void* stream);
int pv2_bilinear_bwd(const void* dout, void* din, int planes, int ih, int iw, int oh, int ow,
                     float rh, float rw, int align_corners, int dtype, void* stream);

/* Up to PV2_MAX_MAPS maps resized to the SAME output size by one launch (the 8 final logit maps of PraNet_V2.forward,
 * pranet.py:349-350,370-371,392-393,414-415; EMCAD/lib/networks.py:116-123): in/out (dout/din), ih, iw, rh, rw are HOST arrays. */
#define PV2_MAX_MAPS 8
int pv2_bilinear_multi_fwd(const void* const* in, void* const* out, const int* ih, const int* iw, const float* rh, const float* rw,
                           int nmaps, int planes, int oh, int ow, int align_corners, int dtype, void* stream);
int pv2_bilinear_multi_bwd(const void* const* dout, void* const* din, const int* ih, const int* iw, const float* rh, const float* rw,
                           int nmaps, int planes, int oh, int ow, int align_corners, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DSRA attention fusion -- binary_seg/lib/pranet.py:365-368,385-389,407-411;
 * EMCAD/lib/decoders.py:477,500,523; MERIT/lib/decoders.py:369-372; MIST/lib/MIST.py:431,440,449
 *
 *   out = fg + fg * softmax_c( up(deep_fg) - up(deep_bg) )        (use_softmax = 1)
 *   out = fg + fg * ( up(deep_fg) - up(deep_bg) )                 (use_softmax = 0)
 * with the bilinear resize of the deeper maps (dh x dw -> h x w, align_corners=False, ratios rh/rw)
 * fused, so crop_* is never materialised.  fp32.  C <= 32.
 *   bwd: dfg (B,C,h,w) and dd (B,C,h,w) = gradient w.r.t. (up(deep_fg) - up(deep_bg)); the caller
 *   pushes dd through pv2_bilinear_bwd to get d(deep_fg) = -d(deep_bg).
 * ------------------------------------------------------------------------------------------- */
int pv2_dsra_fuse_fwd(const float* fg, const float* deep_fg, const float* deep_bg, float* out,
                      int B, int C, int h, int w, int dh, int dw, float rh, float rw,
                      int use_softmax, void* stream);
int pv2_dsra_fuse_bwd(const float* dout, const float* fg, const float* deep_fg, const float* deep_bg,
                      float* dfg, float* dd, int B, int C, int h, int w, int dh, int dw,
                      float rh, float rw, int use_softmax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * V1 reverse attention -- binary_seg/lib/PraNet_Res2Net.py:153-154,166-167,177-178
 *   y[b,c,:,:] = (1 - sigmoid(crop[b,0,:,:])) * x[b,c,:,:]
 *   bwd: dx = (1-s)*dy ; dcrop[b,0] = -s(1-s) * sum_c dy*x
 * x/y/dx dtype = dtype; crop/dcrop fp32.
 * ------------------------------------------------------------------------------------------- */
int pv2_ra_v1_scale_fwd(const void* x, const float* crop, void* y, int B, int C, int hw, int dtype, void* stream);
int pv2_ra_v1_scale_bwd(const void* dy, const void* x, const float* crop, void* dx, float* dcrop,
                        int B, int C, int hw, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multiclass dual-supervision loss -- EMCAD/trainer.py:123-140 (== MERIT/train_ACDC.py:259-284 == MIST/trainer.py:112-129)
 * with its pieces powerset (EMCAD/utils/utils.py:20-30), DiceLoss (utils.py:102-138) and the inverted one-hot mask
 * (trainer.py:22-29, built on the fly from the int64 labels, never materialised).
 *   loss = sum_s lc_ce*CE(sum_{i in s} P_fg[i], y) + lc_dice*Dice(softmax(.), y) + lc_bce*mean BCEWithLogits(sum_{i in s} P_bg[i], 1-onehot(y))
 * mode 0: s runs over all non-empty subsets of the nscales (<= 4) maps ('mutation'); mode 1: singletons ('deep_supervision').
 * P_fg / P_bg: HOST arrays of nscales device pointers to fp32 [B][C][H][W]; labels int64 [B][H][W]; 2 <= C <= 12.
 * fwd writes the scalar loss and leaves the Dice sums in the workspace for bwd; bwd writes all 2*nscales gradients.
 * ------------------------------------------------------------------------------------------- */
size_t pv2_mc_dual_loss_workspace_bytes(int B, int C, int H, int W);
int pv2_mc_dual_loss_fwd(const float* const* P_fg, const float* const* P_bg, const long long* labels, int nscales, int mode,
                         int B, int C, int H, int W, float lc_ce, float lc_dice, float lc_bce, float* loss,
                         void* workspace, size_t workspace_bytes, void* stream);
int pv2_mc_dual_loss_bwd(const float* const* P_fg, const float* const* P_bg, const long long* labels, const float* grad_loss,
                         float* const* dP_fg, float* const* dP_bg, int nscales, int mode, int B, int C, int H, int W,
                         float lc_ce, float lc_dice, float lc_bce, const void* workspace, size_t workspace_bytes, void* stream);

/* =============================================================================================
 * Conv engine (tcgen05 / TMEM / TMA) -- nn.Conv2d + nn.BatchNorm2d + F.relu of BasicConv2d
 * (binary_seg/lib/pranet.py:31-43 and its callers :75-82, :109-123, :357-363, :378-383, :400-405;
 * multiclass heads EMCAD/lib/decoders.py:434-444, MERIT/lib/decoders.py:298-322, MIST/lib/MIST.py:403-412).
 *
 * "Operand format": NHWC, channels padded to a multiple of 8 (bf16) / 4 (tf32); `kind` = PV2_BF16 (one plane) or
 * PV2_TF32 (fp32 storage; with nterms = 3 two planes hi, lo `*_plane_stride` ELEMENTS apart and the GEMM runs
 * hi*hi + lo*hi + hi*lo).  Weights in operand format are [Cout][tap][Cin_p] (pv2_weight_pack).
 * "Raw": fp32 rows [pixel][ld].  All convolutions are stride 1 with same padding (pad = dil*(k-1)/2).
 * ============================================================================================= */

/* Training-mode BatchNorm statistics fused into the conv epilogue (nn.BatchNorm2d of BasicConv2d, pranet.py:37,41-42).
 * A conv group (several convs fused along Cout) carries up to PV2_MAX_BN_SEGS BatchNorm modules, each owning the channel
 * range [c_begin, c_end) of the group's Cout.  With splits == 1 the epilogue reduces every 128-pixel tile of the fp32
 * accumulator to per-channel (mean, M2) before the tile leaves the SM, and the last CTA to finish (two-level ticket,
 * fixed combination order: bit-reproducible) folds the tiles with Chan's formula, writes mean / invstd / scale = gamma*invstd
 * / shift = beta - mean*scale for all Cout channels and updates running_mean / running_var / num_batches_tracked exactly
 * like nn.BatchNorm2d (momentum, unbiased running variance).  With split-K the same descriptor is given to
 * pv2_bn_stats_group after the conv.  `counters` must be zero-initialised once; every launch leaves it zeroed. */
#define PV2_BN_ACC_STRIDE 16   /* doubles between the accumulator pairs of consecutive channels (128 bytes) */
#define PV2_SUM_STRIDE 32      /* floats between consecutive entries of pv2_bn_act_bwd's `sums_zeroed` (128 bytes) */
#define PV2_MAX_BN_SEGS 8
#define PV2_BN_COUNTERS 4096
typedef struct pv2_bn_seg {
    const float* gamma; const float* beta; float* running_mean; float* running_var; long long* num_batches_tracked;
    float eps, momentum;
    int c_begin, c_end;
} pv2_bn_seg;
typedef struct pv2_bn_fuse {
    pv2_bn_seg seg[PV2_MAX_BN_SEGS];
    float* mean; float* invstd; float* scale; float* shift;   /* [Cout] each */
    float* part;                                              /* pv2_bn_fuse_workspace_floats(M, Cout) floats */
    unsigned int* counters;                                   /* PV2_BN_COUNTERS uints */
    int nsegs, pad_;
} pv2_bn_fuse;
size_t pv2_bn_fuse_workspace_floats(long long M, int Cout);
/* How pv2_conv_fwd with these arguments produces the statistics:
 *   0  not at all: call pv2_bn_stats_group afterwards (split-K);
 *   2  (the persistent kernel) every CTA reduces the 128-pixel tiles it computed to per-channel (count, mean, M2) in shared memory
 *      (Chan) and, when it is done, ADDS sum x = n*mean and sum x^2 = M2 + n*mean^2 -- in DOUBLE precision -- to the two accumulators
 *      per channel at bn->part (channel c at double index PV2_BN_ACC_STRIDE*c: one 128-byte line per channel, so that the CTAs'
 *      atomics spread over the L2 slices; PV2_BN_ACC_STRIDE*Cout doubles, ZERO on entry): no partial rows, no ticket, no fence, no
 *      serial tail in the GEMM.  The kernel that consumes a channel slice (pv2_act_apply) is given a pv2_bn_defer descriptor, turns the two doubles of
 *      each of its channels into mean / invstd / scale / shift in its prologue and updates the running statistics;
 *   1  (PV2_CONV_V1=1, the one-tile-per-CTA kernel kept for A/B) final statistics written by the launch (two-level ticket). */
int pv2_conv_fuses_bn_stats(int splits, int out_mode);
/* one BatchNorm module's channel slice [c_off, c_off + C) of a conv group whose statistics are still per-CTA partial rows */
typedef struct pv2_bn_defer {
    const float* part;                  /* the conv group's accumulators (sum x, sum x^2 as DOUBLES at PV2_BN_ACC_STRIDE*channel); NULL = not deferred */
    float count; int ldc, c_off, pad_;  /* pixels summed (N*H*W), channels of the whole group (Cout), first channel of the slice */
    const float* gamma; const float* beta; float* running_mean; float* running_var; long long* num_batches_tracked;
    float eps, momentum;
    float* mean; float* invstd;         /* [C] outputs, saved for the backward pass (scale / shift go to the s / b arrays) */
} pv2_bn_defer;
/* statistics of raw[slab][M][ld] (slabs summed into slab 0 first) for every segment of `bn`: two launches per GROUP */
int pv2_bn_stats_group(float* y, long long slab_stride, int nslabs, long long M, int Cout, int ld, const pv2_bn_fuse* bn, void* stream);

/* out_mode 0: raw fp32 out[split][N*H*W][ldo] (split-K partial slabs, summed by pv2_bn_stats / the consumers);
 * out_mode 1: fp32 NCHW out[N][Cout][H][W] + bias (bias may be NULL), splits must be 1;
 * out_mode 2: bf16 rows out[N*H*W][ldo] (ldo % 8 == 0; `out` is really a bf16 pointer) = a torch channels_last bf16 tensor, splits
 *             must be 1 -- the dgrad of the GEMM that reads a backbone feature writes the feature gradient in its final form.
 * Also computes dgrad when given the mode-1 packed weights (Cin_p := padded Cout, Cout := Cin). */
int pv2_conv_splits_hint(int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int kind, int nterms);
int pv2_conv_fwd(const void* x, long long x_plane_stride, const void* w_op, long long w_plane_stride, int kind, int nterms,
                 int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int dil_h, int dil_w,
                 int out_mode, float* out, int ldo, int splits, const float* bias, const pv2_bn_fuse* bn /* or NULL */,
                 unsigned int* tile_counters /* >= (pixel tiles x N tiles) zero-initialised uints when splits > 1, left zeroed; else may be NULL */,
                 void* stream);
/* 1 when pv2_conv_fwd itself sums the split-K partials: with splits > 1 every split CTA ADDS its partial tile to out[0][.][.] with
 * 16-byte fp32 reductions in L2 (red.global.add.v4.f32), so `out` must be ZERO-INITIALISED by the caller, only one slab is needed and
 * consumers read one slab (fp32 summation order of the <= 8 partials is not fixed: results agree to rounding, not bit for bit);
 * 0 when [split] slabs are written and left for the consumers to sum (PV2_CONV_V1=1). */
int pv2_conv_sums_splits(void);
/* Upper bound on the CTAs of the following pv2_conv_fwd launches of this process (0 = none: one CTA per tile up to two per SM).
 * A caller that runs several convolutions CONCURRENTLY (the head's twelve chains on side streams) gives each a share of the
 * machine: a persistent CTA then walks several tiles through one ring -- barrier setup, TMEM allocation and the first TMA round
 * trip are paid once per CTA instead of once per tile -- and CTAs of different launches are resident side by side instead of
 * queueing for the same slots.  Returns the previous value. */
int pv2_conv_set_cta_budget(int max_ctas);
/* BatchNorm backward (pv2_bn_act_bwd, single-source training case) as ONE launch -- reduce, grid-wide barrier, dx -- instead of two.
 * A barrier kernel is only safe when all of its CTAs are resident at once, so the caller, who knows how many such launches can run
 * side by side (parallel chains on side streams), grants each at most `max_ctas` CTAs (capped at one per SM): the sum over
 * concurrent launches must stay below what the device can hold (4 CTAs of 256 threads x 64 registers per SM).  0 (the default) = two launches.
 * Returns the previous value. */
int pv2_bn_set_fused_grid(int max_ctas);
/* dW partials out[split][Cout][KH*KW][Cin_p] = sum over the split's pixels of dY[p][co] * X[p + tap shift][ci] */
int pv2_conv_wgrad_splits_hint(int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int kind);
int pv2_conv_wgrad(const void* dy, long long dy_plane_stride, const void* x, long long x_plane_stride, int kind, int nterms,
                   int N, int H, int W, int Cin_p, int Cout_p, int Cout, int KH, int KW, int dil_h, int dil_w,
                   float* out, int splits, void* stream);
/* OIHW fp32 -> operand weights.  mode 0 (fprop): out[o_off+co][tap][i_off+ci]; mode 1 (dgrad): out[o_off+ci][flipped tap][i_off+co];
 * i_ld = padded inner channel count of the (possibly horizontally fused) destination. */
int pv2_weight_pack(const float* w, void* out, long long plane_stride, int nplanes, int kind, int Cout, int Cin, int KH, int KW,
                    int mode, int i_ld, int i_off, int o_off, void* stream);
int pv2_wgrad_unpack(const float* part, long long split_stride, int splits, float* dw, int Cout, int Cin, int KH, int KW,
                     int Cin_p, int co_off, void* stream);
/* Multi-tensor forms: a handful of launches for every conv of the head.  `descs` is a HOST array of n descriptors in
 * increasing `start` order (start = running element offset of each OIHW tensor in the concatenated index space); the
 * descriptors are passed to the kernels by value, 40 (56) per launch, so the calls are CUDA-graph capturable.
 * pack: writes the fprop layout to out_f and, when out_d != NULL, the dgrad layout to out_d (fields as in pv2_weight_pack). */
typedef struct pv2_pack_desc {
    const float* w; void* out_f; void* out_d;
    long long f_plane, d_plane, start;
    int Cout, Cin, KH, KW, f_ild, f_ioff, f_ooff, d_ild, d_ioff, d_ooff;
} pv2_pack_desc;
typedef struct pv2_unpack_desc {
    const float* part; float* dw;
    long long split_stride, start;
    int splits, Cout, Cin, KH, KW, Cin_p, co_off, pad_;
} pv2_unpack_desc;
int pv2_weight_pack_multi(const pv2_pack_desc* descs, int n, int nplanes, int kind, void* stream);
int pv2_wgrad_unpack_multi(const pv2_unpack_desc* descs, int n, void* stream);
/* NCHW (x_dtype PV2_F32 | PV2_BF16) -> operand NHWC slice [N*HW][ld] at channel c_off; and summed raw slabs -> NCHW */
int pv2_pack_nchw(const void* x, int x_dtype, void* out, long long plane_stride, int nplanes, int kind, int N, int C, int HW,
                  int ld, int c_off, void* stream);
/* channels_last = 1: dx is stored NHWC (a torch channels_last tensor), no transpose */
int pv2_unpack_to_nchw(const float* const* slabs, const int* lds, const int* offs, int nslabs, void* dx, int dx_dtype,
                       int N, int C, int HW, int channels_last, void* stream);
/* BatchNorm batch statistics of raw conv output (sums the split-K slabs into slab 0 first); saves mean / invstd,
 * emits scale = gamma*invstd and shift = beta - mean*scale, updates running stats like nn.BatchNorm2d. */
size_t pv2_bn_workspace_floats(long long M, int C);
int pv2_bn_stats(float* y, long long slab_stride, int nslabs, long long M, int C, int ld, const float* gamma, const float* beta,
                 float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                 float* mean_out, float* invstd_out, float* scale, float* shift, float* workspace, void* stream);
int pv2_bn_eval_affine(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                       float* scale, float* shift, void* stream);
/* a1 = y1*s1+b1; [a2 = y2*s2+b2; v = a1 (+|*) a2 (combine 1|2)]; [v *= mult (operand format)]; [relu]
 * y_i may still be ns_i split-K slabs ss_i elements apart (summed on load).
 * -> operand NHWC slice (out_nchw = 0) or fp32 NCHW (out_nchw = 1).  Covers BN(+ReLU) (pranet.py:41-42,358),
 * relu(x_cat + conv_res) (pranet.py:82), the partial decoder's products (pranet.py:111-113) and biased heads.
 * With a pv2_bn_defer descriptor for source i the launch first folds that source's per-CTA statistics (see
 * pv2_conv_fuses_bn_stats) and WRITES s_i / b_i (and mean / invstd / running statistics) instead of reading them. */
int pv2_act_apply(const float* y1, int ld1, int off1, int ns1, long long ss1, const float* s1, const float* b1,
                  const float* y2, int ld2, int off2, int ns2, long long ss2,
                  const float* s2, const float* b2, int combine, const void* mult, long long mult_plane, int mult_planes,
                  int mult_ld, int mult_off, int relu, long long M, int C, int HW, void* out, long long out_plane,
                  int out_planes, int out_ld, int out_off, int out_nchw,
                  const pv2_bn_defer* defer1 /* or NULL: s1 / b1 are inputs */, const pv2_bn_defer* defer2 /* or NULL */,
                  int kind, void* stream);
/* backward of pv2_act_apply (+ training-mode BN when bn_train = 1): dz comes as <= 8 summed raw slabs or one NCHW
 * tensor; writes dy1 (dy2) in operand format for the dgrad/wgrad GEMMs, d(mult) raw, dgamma/dbeta. */
int pv2_bn_act_bwd(const float* y1, int ld1, int off1, int ns1, long long ss1, const float* s1, const float* b1,
                   const float* y2, int ld2, int off2, int ns2, long long ss2,
                   const float* s2, const float* b2, int combine, const void* mult, long long mult_plane, int mult_planes,
                   int mult_ld, int mult_off, int relu, long long M, int C, int HW,
                   const float* const* dz_slabs, const int* dz_lds, const int* dz_offs, int dz_n, const float* dz_nchw,
                   const float* mean1, const float* inv1, const float* mean2, const float* inv2, int bn_train,
                   float* dmult, int dmult_ld, void* dy1, long long dy1_plane, int dy1_planes, int dy1_ld,
                   void* dy2, long long dy2_plane, int dy2_planes, int dy2_ld,
                   float* dgamma1, float* dbeta1, float* dgamma2, float* dbeta2, float* workspace,
                   float* sums_zeroed /* 4*C*PV2_SUM_STRIDE ZERO-INITIALISED floats, 128-byte aligned: entry i = which_sum*C + channel lives at
                                         float index PV2_SUM_STRIDE*i (its own 128-byte line, so the blocks' L2 reductions do not queue in
                                         one slice); the reduce pass adds its block sums there (no ticket / fold) and the dx pass reads
                                         them; NULL: three-launch scalar path */,
                   int kind, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (pranet.py:93) on operand tensors; backward raw -> raw */
int pv2_up2_nhwc_fwd(const void* in, long long in_plane, int in_planes, int in_ld, int in_off, void* out, long long out_plane,
                     int out_planes, int out_ld, int out_off, int N, int H, int W, int C, int kind, void* stream);
int pv2_up2_nhwc_bwd(const float* const* slabs, const int* lds, const int* offs, int nslabs, float* din, int din_ld,
                     int N, int H, int W, int C, void* stream);

/* =============================================================================================
 * Optimizer tail -- binary_seg/utils/utils.py:7-17 clip_gradient (param.grad.data.clamp_(-clip, clip)) followed by
 * optimizer.step() of torch.optim.Adam(params, lr) (binary_seg/MyTrain_med.py:85-86,148-149) or optim.AdamW(..., lr,
 * weight_decay=1e-4) (EMCAD/trainer.py:86,155-157; MERIT/train_ACDC.py; MIST/trainer.py), as ONE flat stream.
 * params / grads / exp_avg / exp_avg_sq are four fp32 buffers of n elements sharing one layout (n % 4 == 0, 16-byte aligned;
 * padding elements carry zero gradients and stay zero).  Per element:
 *   g = clamp(g * grad_scale, -clip, clip)          grad_scale = 1/world after the all-reduce (sum)
 *   decoupled = 0 (Adam):  g += weight_decay * p     decoupled = 1 (AdamW): p *= 1 - lr*weight_decay
 *   m = m + (g - m)(1 - beta1);  v = beta2*v + (1 - beta2) g*g
 *   p -= lr/(1 - beta1^t) * m / (sqrt(v)/sqrt(1 - beta2^t) + eps),   t = *step + 1
 * `step` (device int64, starts at 0) is advanced by the kernel, so a captured CUDA graph replays correctly; `ticket` is one
 * zero-initialised device uint the launch leaves zeroed.  Algorithmic HBM bytes: 28 per element.
 * ============================================================================================= */
int pv2_adam_clamp_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                        long long* step, unsigned int* ticket, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int decoupled, float clip, float grad_scale, void* stream);

/* =============================================================================================
 * Inference tails from the LOW-RESOLUTION head maps (the 8 full-resolution fp32 maps are never written).
 *
 * Binary -- binary_seg/MyTest_med.py:35-42 and :104-111 (V2: res2+res3+res4+res5; V1 :98-102: a single map):
 *     z   = F.interpolate( sum_k F.interpolate(map_k, scale_factor=s_k, bilinear), size=(GH, GW), bilinear )   align_corners=False
 *     out = uint8( (sigmoid(z) - min) / (max - min + 1e-8) * 255 )      min / max per image (the reference runs batch 1)
 * maps: HOST array of nmaps (<= 4) device pointers to fp32 [B][1][mh_k][mw_k]; rh_k / rw_k = 1/s_k; (SH, SW) = the size the
 * model's own final upsample produces (pranet.py:349-415); rgh = SH/GH, rgw = SW/GW (the size= form of F.interpolate).
 * out: uint8 [B][GH][GW].  workspace: pv2_infer_tail_workspace_bytes(B) bytes.
 *
 * Multiclass -- EMCAD/utils/utils.py:261-273,286-296 (val_single_volume, use_dual):
 *     label = argmax_c softmax( sum_k (P[k] - P_bg[k]) ),  P[k] = F.interpolate(map_k, scale_factor=s_k, bilinear)  (EMCAD/lib/networks.py:116-123)
 * P_fg / P_bg: HOST arrays of nmaps device pointers to fp32 [B][C][mh_k][mw_k], 1 <= C <= 16.  out: uint8 [B][H][W].
 * Algorithmic HBM bytes: 1 per output pixel (+ the KB-sized maps).
 * ============================================================================================= */
size_t pv2_infer_tail_workspace_bytes(int B);
int pv2_infer_tail_binary(const float* const* maps, const int* mh, const int* mw, const float* rh, const float* rw, int nmaps,
                          int B, int SH, int SW, int GH, int GW, float rgh, float rgw, uint8_t* out,
                          void* workspace, size_t workspace_bytes, void* stream);
int pv2_infer_tail_argmax(const float* const* P_fg, const float* const* P_bg, const int* mh, const int* mw, const float* rh,
                          const float* rw, int nmaps, int B, int C, int H, int W, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PV2_H_ */
