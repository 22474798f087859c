// Inference tails computed straight from the LOW-RESOLUTION head maps: the full-resolution fp32 logit maps
// the reference materialises (8 x B x C x H x W) are never written.
//
//  binary  (binary_seg/MyTest_med.py:35-42 and :104-111; V1 :98-102 uses a single map):
//      out   = p2 + p3 + p4 + p5                      p_k = F.interpolate(map_k, scale_factor=s_k, bilinear, align_corners=False)
//      out   = F.interpolate(out, size=gt.shape, mode='bilinear', align_corners=False)
//      out   = sigmoid(out);  out = (out - out.min()) / (out.max() - out.min() + 1e-8);  uint8(out * 255)
//  multiclass (EMCAD/utils/utils.py:261-273, 286-296):
//      outputs = sum_k (P[k] - P_bg[k]);  label = argmax_c softmax(outputs)        (softmax is monotone: argmax of the sum)
//
// Both are two chained bilinear resamplings of maps a few KB large followed by point-wise work, so every output pixel is
// computed from the low-res maps through L1 (each map is read by thousands of threads); HBM sees the uint8 output only:
// 1 byte per output pixel instead of >= 40 (8 fp32 maps written, 4 re-read, sum written, re-read ...).
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int TAIL_THREADS = 256;
constexpr int TAIL_PX = 4;          // consecutive output pixels per thread: one 4-byte store

struct TailMaps {
    const float* fg[PV2_MAX_SCALES];
    const float* bg[PV2_MAX_SCALES];
    int h[PV2_MAX_SCALES], w[PV2_MAX_SCALES];
    float rh[PV2_MAX_SCALES], rw[PV2_MAX_SCALES];   // ratios low-res -> model output size (1/scale_factor)
    int n;
};

// ATen upsample_bilinear2d: h0*(w0*a + w1*b) + h1*(w0*c + w1*d)
__device__ __forceinline__ float bilerp(const float* __restrict__ p, int w, const Tap& ty, const Tap& tx) {
    const float* r0 = p + (size_t)ty.i0 * w;
    const float* r1 = p + (size_t)ty.i1 * w;
    return ty.w0 * (tx.w0 * __ldg(r0 + tx.i0) + tx.w1 * __ldg(r0 + tx.i1)) + ty.w1 * (tx.w0 * __ldg(r1 + tx.i0) + tx.w1 * __ldg(r1 + tx.i1));
}

// sum over the maps of their upsampled value at pixel (sy, sx) of the model-output grid: ((p2 + p3) + p4) + p5
__device__ __forceinline__ float sum_at(const TailMaps& m, int b, int sy, int sx) {
    float z = 0.0f;
#pragma unroll
    for (int k = 0; k < PV2_MAX_SCALES; ++k) {
        if (k < m.n) {
            const Tap ty = bilinear_tap(sy, m.h[k], m.rh[k], false), tx = bilinear_tap(sx, m.w[k], m.rw[k], false);
            const float v = bilerp(m.fg[k] + (size_t)b * m.h[k] * m.w[k], m.w[k], ty, tx);
            z = (k == 0) ? v : z + v;
        }
    }
    return z;
}

// logit of output pixel (gy, gx): second resize (SH x SW -> GH x GW) of the summed map; identity when the sizes agree
__device__ __forceinline__ float logit_at(const TailMaps& m, int b, int gy, int gx, int SH, int SW, int GH, int GW, float rgh, float rgw) {
    if (SH == GH && SW == GW) return sum_at(m, b, gy, gx);
    const Tap ty = bilinear_tap(gy, SH, rgh, false), tx = bilinear_tap(gx, SW, rgw, false);
    const float a = sum_at(m, b, ty.i0, tx.i0), bb = sum_at(m, b, ty.i0, tx.i1);
    const float c = sum_at(m, b, ty.i1, tx.i0), d = sum_at(m, b, ty.i1, tx.i1);
    return ty.w0 * (tx.w0 * a + tx.w1 * bb) + ty.w1 * (tx.w0 * c + tx.w1 * d);
}

// order-preserving float <-> uint encoding so that integer atomicMin/Max order floats (deterministic: min/max commute)
__device__ __forceinline__ unsigned int enc(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float dec(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// pass 1: per-image min / max of the logits (sigmoid is monotone, so min/max commute with it)
__global__ void __launch_bounds__(TAIL_THREADS)
tail_minmax_kernel(const TailMaps m, unsigned int* __restrict__ mm, int SH, int SW, int GH, int GW, float rgh, float rgw) {
    pdl_prologue();
    const int b = blockIdx.y;
    const long long npx = (long long)GH * GW;
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * TAIL_THREADS + threadIdx.x; i < npx; i += (long long)gridDim.x * TAIL_THREADS) {
        const int gy = (int)(i / GW), gx = (int)(i - (long long)gy * GW);
        const float z = logit_at(m, b, gy, gx, SH, SW, GH, GW, rgh, rgw);
        lo = fminf(lo, z); hi = fmaxf(hi, z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[TAIL_THREADS / 32], s_hi[TAIL_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < TAIL_THREADS / 32; ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
        atomicMin(mm + 2 * b, enc(lo));
        atomicMax(mm + 2 * b + 1, enc(hi));
    }
}

__device__ __forceinline__ float sigmoid_exact(float z) { return 1.0f / (1.0f + expf(-z)); }   // ATen: 1 / (1 + exp(-x)), full-precision expf

// pass 2: recompute the logit, sigmoid, per-image min-max normalisation, x255, truncate to uint8 (numpy astype)
__global__ void __launch_bounds__(TAIL_THREADS)
tail_binary_write_kernel(const TailMaps m, const unsigned int* __restrict__ mm, uint8_t* __restrict__ out, int SH, int SW, int GH, int GW,
                         float rgh, float rgw) {
    pdl_prologue();
    const int b = blockIdx.y;
    const float smin = sigmoid_exact(dec(__ldg(mm + 2 * b))), smax = sigmoid_exact(dec(__ldg(mm + 2 * b + 1)));
    const float den = (smax - smin) + 1e-8f;
    const long long npx = (long long)GH * GW;
    uint8_t* ob = out + (size_t)b * npx;
    const bool vec = (GW % TAIL_PX) == 0;     // rows are 4-byte aligned and a group of 4 never straddles a row
    for (long long g = (long long)blockIdx.x * TAIL_THREADS + threadIdx.x; g * TAIL_PX < npx; g += (long long)gridDim.x * TAIL_THREADS) {
        const long long i0 = g * TAIL_PX;
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < TAIL_PX; ++j) {
            const long long i = i0 + j;
            if (i < npx) {
                const int gy = (int)(i / GW), gx = (int)(i - (long long)gy * GW);
                const float s = sigmoid_exact(logit_at(m, b, gy, gx, SH, SW, GH, GW, rgh, rgw));
                const float o = (s - smin) / den;
                const uint32_t q = (uint32_t)min(max((int)(o * 255.0f), 0), 255);
                if (vec) packed |= q << (8 * j);
                else ob[i] = (uint8_t)q;
            }
        }
        if (vec) *reinterpret_cast<uint32_t*>(ob + i0) = packed;
    }
}

// multiclass: label = argmax_c sum_k (up(P_fg_k)[c] - up(P_bg_k)[c]); the first maximum wins, like torch.argmax
template <int MAXC>
__global__ void __launch_bounds__(TAIL_THREADS)
tail_argmax_kernel(const TailMaps m, uint8_t* __restrict__ out, int C, int H, int W) {
    pdl_prologue();
    const int b = blockIdx.y;
    const long long npx = (long long)H * W;
    for (long long i = (long long)blockIdx.x * TAIL_THREADS + threadIdx.x; i < npx; i += (long long)gridDim.x * TAIL_THREADS) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        float acc[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) acc[c] = 0.0f;
#pragma unroll
        for (int k = 0; k < PV2_MAX_SCALES; ++k) {
            if (k < m.n) {
                const Tap ty = bilinear_tap(y, m.h[k], m.rh[k], false), tx = bilinear_tap(x, m.w[k], m.rw[k], false);
                const size_t plane = (size_t)m.h[k] * m.w[k];
                const float* pf = m.fg[k] + (size_t)b * C * plane;
                const float* pb = m.bg[k] + (size_t)b * C * plane;
#pragma unroll
                for (int c = 0; c < MAXC; ++c) {
                    if (c < C) acc[c] += bilerp(pf + c * plane, m.w[k], ty, tx) - bilerp(pb + c * plane, m.w[k], ty, tx);   // outputs += P[k] - P_bg[k]
                }
            }
        }
        int best = 0;
        float bv = acc[0];
#pragma unroll
        for (int c = 1; c < MAXC; ++c) {
            if (c < C && acc[c] > bv) { bv = acc[c]; best = c; }
        }
        out[(size_t)b * npx + i] = (uint8_t)best;
    }
}

int fill_maps(const char* who, TailMaps* m, const float* const* fg, const float* const* bg, const int* h, const int* w, const float* rh,
              const float* rw, int n) {
    PV2_CHECK(n >= 1 && n <= PV2_MAX_SCALES, "%s: 1..%d maps (got %d)", who, PV2_MAX_SCALES, n);
    PV2_CHECK(fg && h && w && rh && rw, "%s: null array", who);
    m->n = n;
    for (int k = 0; k < PV2_MAX_SCALES; ++k) {
        const int j = k < n ? k : 0;
        PV2_CHECK(fg[j] != nullptr && h[j] > 0 && w[j] > 0 && rh[j] > 0.0f && rw[j] > 0.0f, "%s: bad map %d", who, j);
        m->fg[k] = fg[j]; m->bg[k] = bg ? bg[j] : nullptr;
        m->h[k] = h[j]; m->w[k] = w[j]; m->rh[k] = rh[j]; m->rw[k] = rw[j];
    }
    return 0;
}

inline int tail_grid(long long work_items) {
    const long long want = (work_items + TAIL_THREADS - 1) / TAIL_THREADS;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" size_t pv2_infer_tail_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 1) * 2 * sizeof(unsigned int); }

extern "C" int pv2_infer_tail_binary(const float* const* maps, const int* mh, const int* mw, const float* rh, const float* rw, int nmaps,
                                     int B, int SH, int SW, int GH, int GW, float rgh, float rgw, uint8_t* out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    TailMaps m = {};
    if (int e = fill_maps("infer_tail_binary", &m, maps, nullptr, mh, mw, rh, rw, nmaps)) return e;
    PV2_CHECK(B > 0 && SH > 0 && SW > 0 && GH > 0 && GW > 0 && out && workspace, "infer_tail_binary: bad shape or null pointer");
    PV2_CHECK(workspace_bytes >= pv2_infer_tail_workspace_bytes(B), "infer_tail_binary: workspace too small");
    PV2_CHECK(B <= 65535, "infer_tail_binary: batch %d exceeds grid.y", B);
    cudaStream_t st = (cudaStream_t)stream;
    // min slots <- 0xFFFFFFFF, max slots <- 0: done by pass 0 of the byte pattern below (min, max interleaved)
    unsigned int* mm = (unsigned int*)workspace;
    // interleaved (min, max) pairs cannot be initialised by one memset: two strided 2-D memsets (graph-capturable, no kernel)
    cudaError_t ce = cudaMemset2DAsync(mm, 2 * sizeof(unsigned int), 0xFF, sizeof(unsigned int), (size_t)B, st);
    PV2_CHECK(ce == cudaSuccess, "infer_tail_binary: memset: %s", cudaGetErrorString(ce));
    ce = cudaMemset2DAsync(mm + 1, 2 * sizeof(unsigned int), 0x00, sizeof(unsigned int), (size_t)B, st);
    PV2_CHECK(ce == cudaSuccess, "infer_tail_binary: memset: %s", cudaGetErrorString(ce));
    const long long npx = (long long)GH * GW;
    pv2::launch(tail_minmax_kernel, dim3(tail_grid(npx), B), dim3(TAIL_THREADS), 0, st, m, mm, SH, SW, GH, GW, rgh, rgw);
    PV2_LAUNCH_CHECK("infer_tail_binary(minmax)");
    pv2::launch(tail_binary_write_kernel, dim3(tail_grid((npx + TAIL_PX - 1) / TAIL_PX), B), dim3(TAIL_THREADS), 0, st, m, (const unsigned int*)mm, out,
                SH, SW, GH, GW, rgh, rgw);
    PV2_LAUNCH_CHECK("infer_tail_binary(write)");
    return 0;
}

extern "C" int pv2_infer_tail_argmax(const float* const* P_fg, const float* const* P_bg, const int* mh, const int* mw, const float* rh,
                                     const float* rw, int nmaps, int B, int C, int H, int W, uint8_t* out, void* stream) {
    TailMaps m = {};
    PV2_CHECK(P_bg != nullptr, "infer_tail_argmax: null array");
    if (int e = fill_maps("infer_tail_argmax", &m, P_fg, P_bg, mh, mw, rh, rw, nmaps)) return e;
    for (int k = 0; k < nmaps; ++k) PV2_CHECK(P_bg[k] != nullptr, "infer_tail_argmax: null background map %d", k);
    PV2_CHECK(B > 0 && B <= 65535 && C >= 1 && C <= 16 && H > 0 && W > 0 && out, "infer_tail_argmax: bad shape (1 <= C <= 16) or null pointer");
    const dim3 grid(tail_grid((long long)H * W), B);
    if (C <= 4) pv2::launch(tail_argmax_kernel<4>, grid, dim3(TAIL_THREADS), 0, (cudaStream_t)stream, m, out, C, H, W);
    else if (C <= 9) pv2::launch(tail_argmax_kernel<9>, grid, dim3(TAIL_THREADS), 0, (cudaStream_t)stream, m, out, C, H, W);
    else pv2::launch(tail_argmax_kernel<16>, grid, dim3(TAIL_THREADS), 0, (cudaStream_t)stream, m, out, C, H, W);
    PV2_LAUNCH_CHECK("infer_tail_argmax");
    return 0;
}
