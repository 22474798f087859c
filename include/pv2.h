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
 *   - dtype codes: PV2_F32 = 0, PV2_BF16 = 1.
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
int pv2_structure_loss_bwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                           const float* mask_bg, const float* grad_loss, void* const* dpred,
                           void* const* dpred_bg, int nscales, int planes, int H, int W, int logit_dtype,
                           const void* workspace, size_t workspace_bytes, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* PV2_H_ */
