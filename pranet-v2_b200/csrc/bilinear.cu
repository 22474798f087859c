// Bilinear resize of NCHW planes, forward and (gather-form) backward.
//
// Replaces F.interpolate(mode='bilinear') at binary_seg/lib/pranet.py:349-415 (x8/x16/x32 final maps,
// x0.25 and x2 crops), nn.Upsample(scale_factor=2, align_corners=True) at pranet.py:93 and the size=
// form at EMCAD/lib/decoders.py:460-461.  Index math is ATen's (area_pixel_compute_source_index +
// guard_index_and_lambda), see pv2::bilinear_tap.
//
// fwd: CTA = (plane, band of output rows).  The few source rows the band touches are staged in shared
//      memory once; every thread then produces 4 consecutive outputs and writes them with one 16-byte
//      streaming store.  HBM bytes = (P_in + P_out) * elt: write-bound.
// bwd: CTA = (plane, input row).  Pass 1 folds the <= 2s+2 output rows that touch this input row into
//      one row of column sums in shared memory (coalesced reads of dout), pass 2 folds the columns.
//      Every dout element is read by at most 2 input rows; no atomics, deterministic.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int FWD_THREADS = 256;
constexpr int BAND = 16;  // output rows per CTA (fwd)

template <typename T>
__global__ void __launch_bounds__(FWD_THREADS)
bilinear_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int ih, int iw, int oh, int ow,
                    float rh, float rw, int ac, int max_src_rows) {
    pv2::pdl_prologue();
    extern __shared__ float srows[];  // [nrows][iw]
    const int plane = blockIdx.y;
    const int oy0 = blockIdx.x * BAND, oy1 = min(oy0 + BAND, oh);
    const int r0 = bilinear_tap(oy0, ih, rh, ac).i0;
    const int r1 = bilinear_tap(oy1 - 1, ih, rh, ac).i1;
    const int nrows = r1 - r0 + 1;
    const T* src = in + (size_t)plane * ih * iw;
    const bool staged = nrows <= max_src_rows;
    if (staged) {
        for (int i = threadIdx.x; i < nrows * iw; i += FWD_THREADS) srows[i] = to_f(src[(size_t)r0 * iw + i]);
        __syncthreads();
    }
    T* dst = out + (size_t)plane * oh * ow;
    const int vec_per_row = (ow + 3) >> 2;
    const bool vec_ok = (ow & 3) == 0;
    for (int it = threadIdx.x; it < (oy1 - oy0) * vec_per_row; it += FWD_THREADS) {
        const int oy = oy0 + it / vec_per_row, ox = (it % vec_per_row) * 4;
        const Tap ty = bilinear_tap(oy, ih, rh, ac);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int oxj = min(ox + j, ow - 1);
            const Tap tx = bilinear_tap(oxj, iw, rw, ac);
            float a, b, c, d;
            if (staged) {
                const float* ra = srows + (ty.i0 - r0) * iw;
                const float* rb = srows + (ty.i1 - r0) * iw;
                a = ra[tx.i0]; b = ra[tx.i1]; c = rb[tx.i0]; d = rb[tx.i1];
            } else {
                const T* ra = src + (size_t)ty.i0 * iw;
                const T* rb = src + (size_t)ty.i1 * iw;
                a = to_f(ra[tx.i0]); b = to_f(ra[tx.i1]); c = to_f(rb[tx.i0]); d = to_f(rb[tx.i1]);
            }
            v[j] = ty.w0 * (tx.w0 * a + tx.w1 * b) + ty.w1 * (tx.w0 * c + tx.w1 * d);
        }
        T* o = dst + (size_t)oy * ow + ox;
        if (vec_ok) {
            store4<T>(o, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ox + j < ow) o[j] = from_f<T>(v[j]);
        }
    }
}

// weight with which output index o contributes to input index i
__device__ __forceinline__ float tap_weight(int o, int i, int in_size, float ratio, bool ac) {
    const Tap t = bilinear_tap(o, in_size, ratio, ac);
    return (t.i0 == i ? t.w0 : 0.0f) + (t.i1 == i ? t.w1 : 0.0f);
}

// conservative [lo, hi] range of output indices whose taps can touch input index i
__device__ __forceinline__ void touch_window(int i, int out_size, float ratio, bool ac, int& lo, int& hi) {
    if (!(ratio > 0.0f)) { lo = 0; hi = out_size - 1; return; }
    float a, b;
    if (ac) { a = ((float)i - 1.0f) / ratio; b = ((float)i + 1.0f) / ratio; }
    else { a = ((float)i - 0.5f) / ratio - 0.5f; b = ((float)i + 1.5f) / ratio - 0.5f; }
    lo = max(0, (int)floorf(a) - 1);
    hi = min(out_size - 1, (int)ceilf(b) + 1);
    if (i == 0) lo = 0;  // clamped sources (src < 0) all land on index 0
}

constexpr int BWD_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(BWD_THREADS)
bilinear_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, int ih, int iw, int oh, int ow,
                    float rh, float rw, int ac) {
    pv2::pdl_prologue();
    extern __shared__ float colsum[];  // [ow]
    const int plane = blockIdx.y, iy = blockIdx.x;
    const T* g = dout + (size_t)plane * oh * ow;
    int lo, hi;
    touch_window(iy, oh, rh, ac, lo, hi);
    for (int ox = threadIdx.x; ox < ow; ox += BWD_THREADS) {
        float acc = 0.0f;
        for (int oy = lo; oy <= hi; ++oy) {
            const float wy = tap_weight(oy, iy, ih, rh, ac);
            if (wy != 0.0f) acc += wy * to_f(g[(size_t)oy * ow + ox]);
        }
        colsum[ox] = acc;
    }
    __syncthreads();
    T* d = din + ((size_t)plane * ih + iy) * iw;
    for (int ix = threadIdx.x; ix < iw; ix += BWD_THREADS) {
        int xl, xh;
        touch_window(ix, ow, rw, ac, xl, xh);
        float acc = 0.0f;
        for (int ox = xl; ox <= xh; ++ox) acc += tap_weight(ox, ix, iw, rw, ac) * colsum[ox];
        d[ix] = from_f<T>(acc);
    }
}

int check(const void* a, const void* b, int planes, int ih, int iw, int oh, int ow, int dtype, const char* who) {
    PV2_CHECK(a && b, "%s: null pointer", who);
    PV2_CHECK(planes > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "%s: empty shape", who);
    PV2_CHECK(planes <= 65535, "%s: planes=%d exceeds grid.y limit; split the call", who, planes);
    PV2_CHECK(dtype == PV2_F32 || dtype == PV2_BF16, "%s: bad dtype %d", who, dtype);
    return 0;
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" int pv2_bilinear_fwd(const void* in, void* out, int planes, int ih, int iw, int oh, int ow,
                                float rh, float rw, int align_corners, int dtype, void* stream) {
    if (int e = check(in, out, planes, ih, iw, oh, ow, dtype, "bilinear_fwd")) return e;
    cudaStream_t st = (cudaStream_t)stream;
    // source rows a band can touch: BAND*ratio + 3, capped by what fits in 48 KB of shared memory
    int want = (int)(BAND * (double)rh) + 4;
    if (want > ih) want = ih;
    int cap = (48 * 1024) / (iw * 4);
    int max_rows = want <= cap ? want : 0;
    size_t smem = (size_t)max_rows * iw * 4;
    dim3 grid((oh + BAND - 1) / BAND, planes);
    if (dtype == PV2_F32)
        pv2::launch(bilinear_fwd_kernel<float>, grid, FWD_THREADS, smem, st, (const float*)in, (float*)out, ih, iw, oh, ow, rh, rw, align_corners, max_rows);
    else
        pv2::launch(bilinear_fwd_kernel<__nv_bfloat16>, grid, FWD_THREADS, smem, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ih, iw, oh, ow, rh, rw, align_corners, max_rows);
    PV2_LAUNCH_CHECK("bilinear_fwd");
    return 0;
}

extern "C" int pv2_bilinear_bwd(const void* dout, void* din, int planes, int ih, int iw, int oh, int ow,
                                float rh, float rw, int align_corners, int dtype, void* stream) {
    if (int e = check(dout, din, planes, ih, iw, oh, ow, dtype, "bilinear_bwd")) return e;
    PV2_CHECK((size_t)ow * 4 <= 48 * 1024, "bilinear_bwd: output width %d too large", ow);
    PV2_CHECK(ih <= 65535 * 32, "bilinear_bwd: input height %d too large", ih);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ih, planes);
    size_t smem = (size_t)ow * 4;
    if (dtype == PV2_F32)
        pv2::launch(bilinear_bwd_kernel<float>, grid, BWD_THREADS, smem, st, (const float*)dout, (float*)din, ih, iw, oh, ow, rh, rw, align_corners);
    else
        pv2::launch(bilinear_bwd_kernel<__nv_bfloat16>, grid, BWD_THREADS, smem, st, (const __nv_bfloat16*)dout, (__nv_bfloat16*)din, ih, iw, oh, ow, rh, rw, align_corners);
    PV2_LAUNCH_CHECK("bilinear_bwd");
    return 0;
}
