// DSRA attention fusion and the V1 reverse-attention scale.
//
//   V2 (binary_seg/lib/pranet.py:365-368 and the multiclass copies):
//        out = fg + fg * softmax_c( up(deep_fg) - up(deep_bg) )      | use_softmax = 0: no softmax
//   V1 (binary_seg/lib/PraNet_Res2Net.py:153-154):
//        y = (1 - sigmoid(crop)).expand(C) * x
//
// The V2 op works on KB-sized maps: one thread per output pixel walks the C channels, sampling the
// deeper fg/bg maps bilinearly on the fly (the reference materialises two interpolated tensors and
// runs 4 more element-wise launches).  softmax over a single channel is identically 1, so for the
// binary models (num_class = 1) this reproduces the reference's out = 2*fg and a zero gradient to the
// deeper maps without a special case.
// The V1 op is the one large element-wise op of the family (reads and writes a backbone feature map):
// 16-byte vectorised, one read + one write per element; its backward fuses the channel reduction for
// dcrop with the dx pass.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

struct Sampler {
    int i00, i01, i10, i11;
    float w00, w01, w10, w11;
    __device__ __forceinline__ float operator()(const float* __restrict__ p) const {
        return w00 * __ldg(p + i00) + w01 * __ldg(p + i01) + w10 * __ldg(p + i10) + w11 * __ldg(p + i11);
    }
};

__device__ __forceinline__ Sampler make_sampler(int y, int x, int dh, int dw, float rh, float rw) {
    const Tap ty = bilinear_tap(y, dh, rh, false), tx = bilinear_tap(x, dw, rw, false);
    Sampler s;
    s.i00 = ty.i0 * dw + tx.i0; s.i01 = ty.i0 * dw + tx.i1;
    s.i10 = ty.i1 * dw + tx.i0; s.i11 = ty.i1 * dw + tx.i1;
    // same association as ATen: h0*(w0*a + w1*b) + h1*(w0*c + w1*d)
    s.w00 = ty.w0 * tx.w0; s.w01 = ty.w0 * tx.w1; s.w10 = ty.w1 * tx.w0; s.w11 = ty.w1 * tx.w1;
    return s;
}

__global__ void dsra_fuse_fwd_kernel(const float* __restrict__ fg, const float* __restrict__ dfg_map,
                                     const float* __restrict__ dbg_map, float* __restrict__ out, int B, int C,
                                     int h, int w, int dh, int dw, float rh, float rw, int use_softmax) {
    pv2::pdl_prologue();
    const int hw = h * w, idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * hw) return;
    const int b = idx / hw, pix = idx - b * hw, y = pix / w, x = pix - y * w;
    const Sampler s = make_sampler(y, x, dh, dw, rh, rw);
    const float* pf = dfg_map + (size_t)b * C * dh * dw;
    const float* pb = dbg_map + (size_t)b * C * dh * dw;
    const float* f = fg + (size_t)b * C * hw + pix;
    float* o = out + (size_t)b * C * hw + pix;
    if (!use_softmax) {
        for (int c = 0; c < C; ++c) {
            float d = s(pf + c * dh * dw) - s(pb + c * dh * dw);
            float v = f[(size_t)c * hw];
            o[(size_t)c * hw] = v + v * d;
        }
        return;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, s(pf + c * dh * dw) - s(pb + c * dh * dw));
    float den = 0.0f;
    for (int c = 0; c < C; ++c) den += expf(s(pf + c * dh * dw) - s(pb + c * dh * dw) - mx);
    const float inv = 1.0f / den;
    for (int c = 0; c < C; ++c) {
        float p = expf(s(pf + c * dh * dw) - s(pb + c * dh * dw) - mx) * inv;
        float v = f[(size_t)c * hw];
        o[(size_t)c * hw] = v + v * p;
    }
}

__global__ void dsra_fuse_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ fg,
                                     const float* __restrict__ dfg_map, const float* __restrict__ dbg_map,
                                     float* __restrict__ dfg, float* __restrict__ dd, int B, int C, int h, int w,
                                     int dh, int dw, float rh, float rw, int use_softmax) {
    pv2::pdl_prologue();
    const int hw = h * w, idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * hw) return;
    const int b = idx / hw, pix = idx - b * hw, y = pix / w, x = pix - y * w;
    const Sampler s = make_sampler(y, x, dh, dw, rh, rw);
    const float* pf = dfg_map + (size_t)b * C * dh * dw;
    const float* pb = dbg_map + (size_t)b * C * dh * dw;
    const size_t base = (size_t)b * C * hw + pix;
    if (!use_softmax) {
        for (int c = 0; c < C; ++c) {
            float d = s(pf + c * dh * dw) - s(pb + c * dh * dw);
            float g = dout[base + (size_t)c * hw], v = fg[base + (size_t)c * hw];
            dfg[base + (size_t)c * hw] = g + g * d;
            dd[base + (size_t)c * hw] = g * v;
        }
        return;
    }
    if (C == 1) {   // softmax over a single channel is the constant 1: out = 2*fg, no gradient reaches the deeper maps
        const float g = dout[base];
        dfg[base] = g + g;
        dd[base] = 0.0f;
        return;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, s(pf + c * dh * dw) - s(pb + c * dh * dw));
    float den = 0.0f, dot = 0.0f;
    for (int c = 0; c < C; ++c) {
        float e = expf(s(pf + c * dh * dw) - s(pb + c * dh * dw) - mx);
        den += e;
        dot += e * __fmul_rn(dout[base + (size_t)c * hw], fg[base + (size_t)c * hw]);
    }
    const float inv = 1.0f / den;
    dot *= inv;  // sum_j p_j t_j
    for (int c = 0; c < C; ++c) {
        float p = expf(s(pf + c * dh * dw) - s(pb + c * dh * dw) - mx) * inv;
        float g = dout[base + (size_t)c * hw], v = fg[base + (size_t)c * hw];
        dfg[base + (size_t)c * hw] = g + g * p;
        dd[base + (size_t)c * hw] = p * (__fmul_rn(g, v) - dot);
    }
}

// ---- V1 reverse attention -----------------------------------------------------------------------
template <typename T>
__global__ void ra_v1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ crop, T* __restrict__ y,
                                 int C, int hw, size_t total_vec) {
    pv2::pdl_prologue();
    // hw % 4 == 0: one thread = 4 consecutive pixels of one (b,c) plane
    const int vec_per_plane = hw >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (size_t)gridDim.x * blockDim.x) {
        const size_t plane = i / vec_per_plane;
        const int v = (int)(i - plane * vec_per_plane);
        const size_t b = plane / C;
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(crop + b * hw) + v);
        float4 xv = load4<T>(x + plane * hw + (size_t)v * 4);
        xv.x *= 1.0f - 1.0f / (1.0f + __expf(-c4.x));
        xv.y *= 1.0f - 1.0f / (1.0f + __expf(-c4.y));
        xv.z *= 1.0f - 1.0f / (1.0f + __expf(-c4.z));
        xv.w *= 1.0f - 1.0f / (1.0f + __expf(-c4.w));
        store4<T>(y + plane * hw + (size_t)v * 4, xv);
    }
}

template <typename T>
__global__ void ra_v1_fwd_scalar_kernel(const T* __restrict__ x, const float* __restrict__ crop, T* __restrict__ y,
                                        int C, int hw, size_t total) {
    pv2::pdl_prologue();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t plane = i / hw;
        const int p = (int)(i - plane * hw);
        const float a = 1.0f - 1.0f / (1.0f + __expf(-crop[(plane / C) * hw + p]));
        y[i] = from_f<T>(a * to_f(x[i]));
    }
}

// block = 32 pixels x 8 channel groups; grid = (ceil(hw/32), B)
template <typename T>
__global__ void __launch_bounds__(256)
ra_v1_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ crop,
                 T* __restrict__ dx, float* __restrict__ dcrop, int C, int hw) {
    pv2::pdl_prologue();
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, b = blockIdx.y;
    const int p = blockIdx.x * 32 + lane;
    float acc = 0.0f, a = 0.0f, s = 0.0f;
    if (p < hw) {
        s = 1.0f / (1.0f + __expf(-crop[(size_t)b * hw + p]));
        a = 1.0f - s;
        for (int c = grp; c < C; c += 8) {
            const size_t o = ((size_t)b * C + c) * hw + p;
            const float g = to_f(dy[o]);
            acc += g * to_f(x[o]);
            dx[o] = from_f<T>(a * g);
        }
    }
    red[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && p < hw) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        dcrop[(size_t)b * hw + p] = -s * a * t;
    }
}

}  // namespace
}  // namespace pv2

using namespace pv2;

static int fuse_check(int B, int C, int h, int w, int dh, int dw, const char* who) {
    PV2_CHECK(B > 0 && C > 0 && h > 0 && w > 0 && dh > 0 && dw > 0, "%s: empty shape", who);
    PV2_CHECK(C <= 64, "%s: C=%d > 64 unsupported", who, C);
    return 0;
}

extern "C" int pv2_dsra_fuse_fwd(const float* fg, const float* deep_fg, const float* deep_bg, float* out,
                                 int B, int C, int h, int w, int dh, int dw, float rh, float rw,
                                 int use_softmax, void* stream) {
    if (int e = fuse_check(B, C, h, w, dh, dw, "dsra_fuse_fwd")) return e;
    PV2_CHECK(fg && deep_fg && deep_bg && out, "dsra_fuse_fwd: null pointer");
    const int n = B * h * w, threads = 128;
    pv2::launch(dsra_fuse_fwd_kernel, (n + threads - 1) / threads, threads, 0, (cudaStream_t)stream, 
        fg, deep_fg, deep_bg, out, B, C, h, w, dh, dw, rh, rw, use_softmax);
    PV2_LAUNCH_CHECK("dsra_fuse_fwd");
    return 0;
}

extern "C" int pv2_dsra_fuse_bwd(const float* dout, const float* fg, const float* deep_fg, const float* deep_bg,
                                 float* dfg, float* dd, int B, int C, int h, int w, int dh, int dw,
                                 float rh, float rw, int use_softmax, void* stream) {
    if (int e = fuse_check(B, C, h, w, dh, dw, "dsra_fuse_bwd")) return e;
    PV2_CHECK(dout && fg && deep_fg && deep_bg && dfg && dd, "dsra_fuse_bwd: null pointer");
    const int n = B * h * w, threads = 128;
    pv2::launch(dsra_fuse_bwd_kernel, (n + threads - 1) / threads, threads, 0, (cudaStream_t)stream, 
        dout, fg, deep_fg, deep_bg, dfg, dd, B, C, h, w, dh, dw, rh, rw, use_softmax);
    PV2_LAUNCH_CHECK("dsra_fuse_bwd");
    return 0;
}

extern "C" int pv2_ra_v1_scale_fwd(const void* x, const float* crop, void* y, int B, int C, int hw, int dtype, void* stream) {
    PV2_CHECK(x && crop && y, "ra_v1_scale_fwd: null pointer");
    PV2_CHECK(B > 0 && C > 0 && hw > 0, "ra_v1_scale_fwd: empty shape");
    PV2_CHECK(dtype == PV2_F32 || dtype == PV2_BF16, "ra_v1_scale_fwd: bad dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)B * C * hw;
    const int threads = 256;
    if ((hw & 3) == 0) {
        const size_t nv = total / 4;
        int blocks = (int)((nv + threads - 1) / threads);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        if (dtype == PV2_F32) pv2::launch(ra_v1_fwd_kernel<float>, blocks, threads, 0, st, (const float*)x, crop, (float*)y, C, hw, nv);
        else pv2::launch(ra_v1_fwd_kernel<__nv_bfloat16>, blocks, threads, 0, st, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)y, C, hw, nv);
    } else {
        int blocks = (int)((total + threads - 1) / threads);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        if (dtype == PV2_F32) pv2::launch(ra_v1_fwd_scalar_kernel<float>, blocks, threads, 0, st, (const float*)x, crop, (float*)y, C, hw, total);
        else pv2::launch(ra_v1_fwd_scalar_kernel<__nv_bfloat16>, blocks, threads, 0, st, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)y, C, hw, total);
    }
    PV2_LAUNCH_CHECK("ra_v1_scale_fwd");
    return 0;
}

extern "C" int pv2_ra_v1_scale_bwd(const void* dy, const void* x, const float* crop, void* dx, float* dcrop,
                                   int B, int C, int hw, int dtype, void* stream) {
    PV2_CHECK(dy && x && crop && dx && dcrop, "ra_v1_scale_bwd: null pointer");
    PV2_CHECK(B > 0 && C > 0 && hw > 0 && B <= 65535, "ra_v1_scale_bwd: bad shape");
    PV2_CHECK(dtype == PV2_F32 || dtype == PV2_BF16, "ra_v1_scale_bwd: bad dtype %d", dtype);
    dim3 grid((hw + 31) / 32, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PV2_F32) pv2::launch(ra_v1_bwd_kernel<float>, grid, 256, 0, st, (const float*)dy, (const float*)x, crop, (float*)dx, dcrop, C, hw);
    else pv2::launch(ra_v1_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)dx, dcrop, C, hw);
    PV2_LAUNCH_CHECK("ra_v1_scale_bwd");
    return 0;
}
