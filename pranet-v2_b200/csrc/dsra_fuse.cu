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
// block = 64 pixel quads x 4 channel groups; a thread computes the four (1 - sigmoid) factors of its pixels ONCE and applies
// them to RA_CPT channels whose loads are all in flight together.  (One thread per (b, c, quad) recomputed the sigmoid for
// every channel and spent two 64-bit divisions per quad on finding its plane: 41 instructions per element, issue bound.)
constexpr int RA_CPT = 8;
template <typename T>
__global__ void __launch_bounds__(256)
ra_v1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ crop, T* __restrict__ y, int C, int hw, int vec_per_plane) {
    pv2::pdl_prologue();
    const int v = blockIdx.x * 64 + (threadIdx.x & 63), b = blockIdx.y;
    const int c0 = (blockIdx.z * 4 + (threadIdx.x >> 6)) * RA_CPT;
    if (v >= vec_per_plane || c0 >= C) return;
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(crop + (size_t)b * hw) + v);
    const float a0 = 1.0f - 1.0f / (1.0f + __expf(-c4.x)), a1 = 1.0f - 1.0f / (1.0f + __expf(-c4.y));
    const float a2 = 1.0f - 1.0f / (1.0f + __expf(-c4.z)), a3 = 1.0f - 1.0f / (1.0f + __expf(-c4.w));
    const size_t base = ((size_t)b * C + c0) * hw + (size_t)v * 4;
    float4 xv[RA_CPT];
#pragma unroll
    for (int k = 0; k < RA_CPT; ++k)
        if (c0 + k < C) xv[k] = load4<T>(x + base + (size_t)k * hw);
#pragma unroll
    for (int k = 0; k < RA_CPT; ++k)
        if (c0 + k < C) store4<T>(y + base + (size_t)k * hw, make_float4(xv[k].x * a0, xv[k].y * a1, xv[k].z * a2, xv[k].w * a3));
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

// Planes whose pixel count is not a multiple of 4 (11 x 11 = 121 at the deepest level): an image is ONE flat run of C * hw elements,
// read and written as 16-byte vectors regardless of where the planes start.  The image's hw factors (1 - sigmoid(crop)) are computed
// once per CTA into shared memory, extended by one vector's worth of wrap-around copies, so the factors of a vector that starts at
// flat index i are a[(i mod hw) + k] -- one 32-bit remainder per VECTOR; the scalar form above pays a 64-bit division and an exp per
// ELEMENT.  Needs C * hw to be a multiple of the vector length (16-byte aligned images) and hw + VEC floats of shared memory.
constexpr int RA_FLAT_MAX_HW = 4096;
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
ra_v1_fwd_flat_kernel(const T* __restrict__ x, const float* __restrict__ crop, T* __restrict__ y, int C, int hw, int vec_per_image, int vec_per_cta) {
    pv2::pdl_prologue();
    extern __shared__ float fac[];          // [hw + VEC]
    const int b = blockIdx.y;
    for (int p = threadIdx.x; p < hw + VEC; p += 256) {
        const int q = p < hw ? p : p - hw;
        fac[p] = 1.0f - 1.0f / (1.0f + __expf(-crop[(size_t)b * hw + q]));
    }
    __syncthreads();
    const size_t img = (size_t)b * C * hw;
    const int v0 = blockIdx.x * vec_per_cta, v1 = min(vec_per_image, v0 + vec_per_cta);
    for (int v = v0 + threadIdx.x; v < v1; v += 2 * 256) {
        const int vb = v + 256;
        const bool two = vb < v1;
        const size_t e0 = img + (size_t)v * VEC, e1 = img + (size_t)(two ? vb : v) * VEC;
        const int p0 = (int)(((unsigned)v * (unsigned)VEC) % (unsigned)hw), p1 = (int)(((unsigned)(two ? vb : v) * (unsigned)VEC) % (unsigned)hw);
        if (VEC == 4) {
            const float4 a = load4<T>(x + e0), c = load4<T>(x + e1);
            store4<T>(y + e0, make_float4(a.x * fac[p0], a.y * fac[p0 + 1], a.z * fac[p0 + 2], a.w * fac[p0 + 3]));
            if (two) store4<T>(y + e1, make_float4(c.x * fac[p1], c.y * fac[p1 + 1], c.z * fac[p1 + 2], c.w * fac[p1 + 3]));
        } else {        // 8 bf16 per 16 bytes
            const uint4 ua = __ldg(reinterpret_cast<const uint4*>(x + e0)), uc = __ldg(reinterpret_cast<const uint4*>(x + e1));
            auto scale8 = [&](const uint4& u, int p) {
                const unsigned w[4] = {u.x, u.y, u.z, u.w};
                unsigned o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float lo = __uint_as_float(w[j] << 16) * fac[p + 2 * j], hi = __uint_as_float(w[j] & 0xffff0000u) * fac[p + 2 * j + 1];
                    const __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
                    o[j] = *reinterpret_cast<const unsigned*>(&r);
                }
                return make_uint4(o[0], o[1], o[2], o[3]);
            };
            *reinterpret_cast<uint4*>(y + e0) = scale8(ua, p0);
            if (two) *reinterpret_cast<uint4*>(y + e1) = scale8(uc, p1);
        }
    }
}

// block = 32 pixels x 8 channel groups; grid = (ceil(hw/32), B)   (scalar form: hw % 4 != 0)
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

// vector form (hw % 4 == 0): block = 32 pixel quads x 8 channel groups, a thread owns 4 consecutive pixels and walks its
// channels 8 at a time (16 independent 8/16-byte loads in flight); the channel groups are folded in group order through
// shared memory (deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
ra_v1_bwd4_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ crop,
                  T* __restrict__ dx, float* __restrict__ dcrop, int C, int hw, int vec_per_plane) {
    pv2::pdl_prologue();
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, b = blockIdx.y;
    const int v = blockIdx.x * 32 + lane;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f), a = acc, s = acc;
    if (v < vec_per_plane) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(crop + (size_t)b * hw) + v);
        s = make_float4(1.0f / (1.0f + __expf(-c4.x)), 1.0f / (1.0f + __expf(-c4.y)), 1.0f / (1.0f + __expf(-c4.z)), 1.0f / (1.0f + __expf(-c4.w)));
        a = make_float4(1.0f - s.x, 1.0f - s.y, 1.0f - s.z, 1.0f - s.w);
        const size_t base = (size_t)b * C * hw + (size_t)v * 4;
        for (int c = grp; c < C; c += 64) {
            float4 g[8], xv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int cc = c + 8 * k;
                if (cc < C) { g[k] = load4<T>(dy + base + (size_t)cc * hw); xv[k] = load4<T>(x + base + (size_t)cc * hw); }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int cc = c + 8 * k;
                if (cc < C) {
                    acc.x = fmaf(g[k].x, xv[k].x, acc.x); acc.y = fmaf(g[k].y, xv[k].y, acc.y);
                    acc.z = fmaf(g[k].z, xv[k].z, acc.z); acc.w = fmaf(g[k].w, xv[k].w, acc.w);
                    store4<T>(dx + base + (size_t)cc * hw, make_float4(a.x * g[k].x, a.y * g[k].y, a.z * g[k].z, a.w * g[k].w));
                }
            }
        }
    }
    red[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && v < vec_per_plane) {
        float4 t = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) { t.x += red[k][lane].x; t.y += red[k][lane].y; t.z += red[k][lane].z; t.w += red[k][lane].w; }
        reinterpret_cast<float4*>(dcrop + (size_t)b * hw)[v] = make_float4(-s.x * a.x * t.x, -s.y * a.y * t.y, -s.z * a.z * t.z, -s.w * a.w * t.w);
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
    // the flat kernel serves the geometries the plane-wise vector kernel cannot (hw % 4 != 0): 15.2 -> 4.4 us at 16 x 2048 x 11^2 bf16
    // (16 % -> 55 % of the HBM peak).  For planes that ARE 16-byte multiples it was measured slower than the plane-wise kernel
    // (PV2_RA_FLAT=2: 23.2 us against 17.9 at 512 x 44^2, 164 against 126 us at 64 x 1024 x 44^2), so it is not used there
    const bool flat_ok = hw <= RA_FLAT_MAX_HW && B <= 65535 && ((size_t)C * hw) % (dtype == PV2_F32 ? 4 : 8) == 0 && (size_t)C * hw < (1u << 31) &&
                         (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0;
    const int flat_mode = pv2::tune_int("PV2_RA_FLAT", 1);      // 0 never, 1 only when hw % 4 != 0, 2 whenever possible
    if ((hw & 3) == 0 && B <= 65535 && (C + 4 * RA_CPT - 1) / (4 * RA_CPT) <= 65535 && !(flat_ok && flat_mode >= 2)) {
        const int vpp = hw >> 2;
        dim3 grid((vpp + 63) / 64, B, (C + 4 * RA_CPT - 1) / (4 * RA_CPT));
        if (dtype == PV2_F32) pv2::launch(ra_v1_fwd_kernel<float>, grid, threads, 0, st, (const float*)x, crop, (float*)y, C, hw, vpp);
        else pv2::launch(ra_v1_fwd_kernel<__nv_bfloat16>, grid, threads, 0, st, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)y, C, hw, vpp);
    } else if (flat_ok && flat_mode >= 1) {
        const int VEC = dtype == PV2_F32 ? 4 : 8;
        const int vpi = (int)((size_t)C * hw / VEC);
        // ~4 CTAs per SM over the batch, at least two vectors per thread
        int per_img = (4 * kNumSMs + B - 1) / B;
        int vpc = (vpi + per_img - 1) / per_img;
        if (vpc < 512) vpc = 512;
        vpc = (vpc + 511) / 512 * 512;
        dim3 grid((vpi + vpc - 1) / vpc, B);
        const size_t smem = (size_t)(hw + VEC) * sizeof(float);
        if (dtype == PV2_F32) pv2::launch(ra_v1_fwd_flat_kernel<float, 4>, grid, threads, smem, st, (const float*)x, crop, (float*)y, C, hw, vpi, vpc);
        else pv2::launch(ra_v1_fwd_flat_kernel<__nv_bfloat16, 8>, grid, threads, smem, st, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)y, C, hw, vpi, vpc);
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
    cudaStream_t st = (cudaStream_t)stream;
    if ((hw & 3) == 0) {
        const int vpp = hw >> 2;
        dim3 grid4((vpp + 31) / 32, B);
        if (dtype == PV2_F32) pv2::launch(ra_v1_bwd4_kernel<float>, grid4, 256, 0, st, (const float*)dy, (const float*)x, crop, (float*)dx, dcrop, C, hw, vpp);
        else pv2::launch(ra_v1_bwd4_kernel<__nv_bfloat16>, grid4, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)dx, dcrop, C, hw, vpp);
        PV2_LAUNCH_CHECK("ra_v1_scale_bwd");
        return 0;
    }
    dim3 grid((hw + 31) / 32, B);
    if (dtype == PV2_F32) pv2::launch(ra_v1_bwd_kernel<float>, grid, 256, 0, st, (const float*)dy, (const float*)x, crop, (float*)dx, dcrop, C, hw);
    else pv2::launch(ra_v1_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, crop, (__nv_bfloat16*)dx, dcrop, C, hw);
    PV2_LAUNCH_CHECK("ra_v1_scale_bwd");
    return 0;
}
