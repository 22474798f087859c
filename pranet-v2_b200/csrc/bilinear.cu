// Bilinear resize of NCHW planes, forward and (gather-form) backward.
//
// Replaces F.interpolate(mode='bilinear') at binary_seg/lib/pranet.py:349-415 (x8/x16/x32 final maps,
// x0.25 and x2 crops), nn.Upsample(scale_factor=2, align_corners=True) at pranet.py:93 and the size=
// form at EMCAD/lib/decoders.py:460-461.  Index math is ATen's (area_pixel_compute_source_index +
// guard_index_and_lambda), see pv2::bilinear_tap.
//
// fwd: CTA = (plane, band of output rows); the band is sized on the host so that a launch has ~1000 CTAs.  The few source rows the
//      band touches and the band's y taps are staged in shared memory once; every thread then produces 4 consecutive outputs per row
//      (x taps computed once per thread, quads that share their two source columns are interpolated vertically first) and writes
//      them with one 16-byte streaming store.  HBM bytes = (P_in + P_out) * elt: write-bound; 78.6 % of the measured HBM peak on the
//      8 final maps of a B = 16 x 352^2 step (12.4 us).
// bwd: two kernels.  Exact x8 / x16 / x32 / x64 half-pixel up-scalings (the final maps of a step) take bilinear_bwd2_kernel: separable
//      in registers -- a thread folds the 4 x 4 elements of a chunk into 4 partial sums with 13 FMA per 16-byte load, 8 loads in
//      flight, CTA = (band of input rows, plane, map), 288 CTAs in one wave -- 16.2 us on the 8 maps (60 % of the measured HBM peak;
//      the gather kernel: 28.5 us, 34 %).  Everything else (x2, x4, fractional ratios, align_corners) takes the gather kernel: CTA =
//      (R input rows, plane, map), pass 1 folds the output rows that touch these input rows into R rows of column sums in shared
//      memory, pass 2 folds the columns (x taps from a table, lane groups + shuffle tree).  No atomics in either, deterministic.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int FWD_THREADS = 256;
constexpr int BAND = 16;      // output rows per CTA (fwd), default; PV2_BIL_BAND overrides (tuning knob)
constexpr int MAX_BAND = 64;

// up to PV2_MAX_MAPS maps of the same output size in ONE launch (grid.z = map): the 8 final logit maps of a step are
// 8 x 8 MB at B=16 -- far too little per launch to fill HBM, so they share a launch.
struct MultiMaps {
    const void* in[PV2_MAX_MAPS];
    void* out[PV2_MAX_MAPS];
    int ih[PV2_MAX_MAPS], iw[PV2_MAX_MAPS], max_rows[PV2_MAX_MAPS];
    float rh[PV2_MAX_MAPS], rw[PV2_MAX_MAPS];
    float irh[PV2_MAX_MAPS], irw[PV2_MAX_MAPS];     // 1 / ratio (backward only; 0 when the ratio is 0)
};

template <typename T>
__global__ void __launch_bounds__(FWD_THREADS)
bilinear_fwd_kernel(const __grid_constant__ MultiMaps mm, int oh, int ow, int ac, int band) {
    pv2::pdl_prologue();
    extern __shared__ float srows[];  // [nrows][iw]
    const int map = blockIdx.z;
    const T* __restrict__ in = reinterpret_cast<const T*>(mm.in[map]);
    T* __restrict__ out = reinterpret_cast<T*>(mm.out[map]);
    const int ih = mm.ih[map], iw = mm.iw[map], max_src_rows = mm.max_rows[map];
    const float rh = mm.rh[map], rw = mm.rw[map];
    const int plane = blockIdx.y;
    const int oy0 = blockIdx.x * band, oy1 = min(oy0 + band, oh);
    const int r0 = bilinear_tap(oy0, ih, rh, ac).i0;
    const int r1 = bilinear_tap(oy1 - 1, ih, rh, ac).i1;
    const int nrows = r1 - r0 + 1;
    const T* src = in + (size_t)plane * ih * iw;
    const bool staged = nrows <= max_src_rows;
    // y taps of the band's rows, computed once per CTA (every thread used to re-derive its row's tap: 14 of ~100 instructions per
    // quad in a kernel that ran at 74 % issue-slot utilisation)
    __shared__ float2 ytab[MAX_BAND];
    for (int i = threadIdx.x; i < oy1 - oy0; i += FWD_THREADS) {
        const Tap t = bilinear_tap(oy0 + i, ih, rh, ac);
        ytab[i] = make_float2(__int_as_float(t.i0), t.w1);
    }
    if (staged)
        for (int i = threadIdx.x; i < nrows * iw; i += FWD_THREADS) srows[i] = to_f(src[(size_t)r0 * iw + i]);
    __syncthreads();
    T* dst = out + (size_t)plane * oh * ow;
    const int vec_per_row = (ow + 3) >> 2;
    const bool vec_ok = (ow & 3) == 0;
    // thread = one quad of 4 consecutive output columns (x taps computed once) x every rgroups-th row of the band
    const bool wide = vec_per_row >= FWD_THREADS;              // more quads than threads: threads stride over the quads
    const int rgroups = wide ? 1 : FWD_THREADS / vec_per_row;
    const int rg = wide ? 0 : threadIdx.x / vec_per_row;
    const int qstep = wide ? FWD_THREADS : vec_per_row;
    for (int q = wide ? threadIdx.x : threadIdx.x % vec_per_row; rg < rgroups && q < vec_per_row; q += qstep) {
        const int ox = q * 4;
        int xi0[4], xi1[4];
        float xw0[4], xw1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const Tap tx = bilinear_tap(min(ox + j, ow - 1), iw, rw, ac);
            xi0[j] = tx.i0; xi1[j] = tx.i1; xw0[j] = tx.w0; xw1[j] = tx.w1;
        }
        // integer up-scaling by a multiple of 4: the four outputs of a quad read the SAME two source columns, so a row costs 4
        // shared-memory loads instead of 16 (the kernel was issue bound: 160 instructions per quad, 79 % issue slots busy)
        const bool same_cols = xi0[0] == xi0[3] && xi1[0] == xi1[3] && xi0[0] == xi0[1] && xi0[0] == xi0[2] && xi1[0] == xi1[1] && xi1[0] == xi1[2];
        for (int oy = oy0 + rg; oy < oy1; oy += rgroups) {
            const float2 yq = ytab[oy - oy0];
            Tap ty;
            ty.i0 = __float_as_int(yq.x); ty.i1 = min(ty.i0 + 1, ih - 1); ty.w1 = yq.y; ty.w0 = 1.0f - yq.y;
            float v[4];
            if (staged && same_cols) {
                const float* ra = srows + (ty.i0 - r0) * iw;
                const float* rb = srows + (ty.i1 - r0) * iw;
                // the quad shares its two source columns: interpolate them vertically once, then 2 instructions per pixel (12 per
                // quad instead of 24; fp32 rounding order differs from the general path below at the 1e-7 level)
                const float cl = fmaf(ty.w1, rb[xi0[0]], ty.w0 * ra[xi0[0]]), cr = fmaf(ty.w1, rb[xi1[0]], ty.w0 * ra[xi1[0]]);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaf(xw1[j], cr, xw0[j] * cl);
            } else if (staged) {
                const float* ra = srows + (ty.i0 - r0) * iw;
                const float* rb = srows + (ty.i1 - r0) * iw;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = ty.w0 * (xw0[j] * ra[xi0[j]] + xw1[j] * ra[xi1[j]]) + ty.w1 * (xw0[j] * rb[xi0[j]] + xw1[j] * rb[xi1[j]]);
            } else {
                const T* ra = src + (size_t)ty.i0 * iw;
                const T* rb = src + (size_t)ty.i1 * iw;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = ty.w0 * (xw0[j] * to_f(ra[xi0[j]]) + xw1[j] * to_f(ra[xi1[j]])) + ty.w1 * (xw0[j] * to_f(rb[xi0[j]]) + xw1[j] * to_f(rb[xi1[j]]));
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

// input rows per CTA by scale: x8 and below -> 4, x16 -> 2, x32 and beyond -> 1 (same rule on the host for the grid)
__host__ __device__ inline int bwd_rows_per_cta(int ih, int oh) {
    const int scale = (oh + ih - 1) / ih;
    return scale <= 8 ? 4 : (scale <= 16 ? 2 : 1);
}

constexpr int BWD_THREADS = 384;   // upper bound; the launch uses ow4 * rgroups threads so that every thread owns a column quad
constexpr int BWD_R = 4;             // input rows per CTA: neighbouring input rows share output rows, so (R+1)*s rows are read for R rows
// output rows a thread fetches per round trip (BWD_BATCH) and CTAs per SM are template parameters of the kernel; the launch uses (4, 3)
constexpr int BWD_MAX_WIN = 192;     // output rows a block of R input rows can touch: (R+1)*scale + a few (scale <= 32)

// touch_window with the reciprocal ratio precomputed on the host (a float division costs ~10 instructions; the +-1 margins absorb
// the last-bit difference)
__device__ __forceinline__ void touch_window_r(int i, int out_size, float ratio, float inv, bool ac, int& lo, int& hi) {
    if (!(ratio > 0.0f)) { lo = 0; hi = out_size - 1; return; }
    float a, b;
    if (ac) { a = ((float)i - 1.0f) * inv; b = ((float)i + 1.0f) * inv; }
    else { a = ((float)i - 0.5f) * inv - 0.5f; b = ((float)i + 1.5f) * inv - 0.5f; }
    lo = max(0, (int)floorf(a) - 1);
    hi = min(out_size - 1, (int)ceilf(b) + 1);
    if (i == 0) lo = 0;
}

// CTA = (block of R input rows, plane, map); R = 4 / 2 / 1 at x8 / x16 / x32 so that every CTA reads a similar number of rows.
// Pass 1: every output row the block touches is read ONCE (16-byte loads, a thread owns 4 consecutive columns, the row window is
// split over `rgroups` thread groups) and folded into the R rows of column sums with its y tap weights (one 16-byte shared-memory
// load per row); the row groups are then folded; pass 2 folds the columns with x taps from a table.  No atomics, deterministic.
// The kernel was issue bound (22 M warp instructions for 63 MB, 67 % issue slots busy): R is a template parameter (no multiplies by
// zero weights), the row loops carry no per-element predicates, taps and windows are computed once.
template <typename T, int R, int BATCH>
__device__ __forceinline__ void bilinear_bwd_body(const MultiMaps& mm, int map, int oh, int ow, bool ac, int rgroups, float* colsum, float4* wts) {
    const T* __restrict__ dout = reinterpret_cast<const T*>(mm.out[map]);      // for the backward, `out` is the upstream gradient
    T* __restrict__ din = reinterpret_cast<T*>(const_cast<void*>(mm.in[map]));  // and `in` the low-resolution gradient being produced
    const int ih = mm.ih[map], iw = mm.iw[map];
    const float rh = mm.rh[map], rw = mm.rw[map];
    const int plane = blockIdx.y, iy0 = blockIdx.x * R;
    if (iy0 >= ih) return;
    const int nr = min(R, ih - iy0);
    const T* g = dout + (size_t)plane * oh * ow;
    int lo, hi, lo2, hi2;
    touch_window_r(iy0, oh, rh, mm.irh[map], ac, lo, hi2);
    touch_window_r(iy0 + nr - 1, oh, rh, mm.irh[map], ac, lo2, hi);
    const int nwin = hi - lo + 1;
    const int ow4 = (ow + 3) >> 2, pitch = ow4 * 4;
    const bool vec_ok = (ow & 3) == 0;
    float2* xtab = reinterpret_cast<float2*>(colsum + BWD_R * rgroups * pitch);    // x taps of every output column: (i0 as bits, w1)
    for (int ox = threadIdx.x; ox < pitch; ox += blockDim.x) {
        const Tap t = bilinear_tap(min(ox, ow - 1), iw, rw, ac);
        xtab[ox] = make_float2(__int_as_float(t.i0), t.w1);
    }
    if (nwin <= BWD_MAX_WIN) {
        for (int j = threadIdx.x; j < nwin; j += blockDim.x) {                    // y tap weights of window row j towards the R input rows
            const Tap t = bilinear_tap(lo + j, ih, rh, ac);
            float w[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = iy0 + r;
                w[r] = r < nr ? (t.i0 == i ? t.w0 : 0.0f) + (t.i1 == i ? t.w1 : 0.0f) : 0.0f;
            }
            wts[j] = make_float4(w[0], w[1], w[2], w[3]);
        }
        __syncthreads();
        const int q = threadIdx.x % ow4, rg = threadIdx.x / ow4;
        if (rg < rgroups) {
            float acc[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) { acc[r][0] = 0.0f; acc[r][1] = 0.0f; acc[r][2] = 0.0f; acc[r][3] = 0.0f; }
            const int ox = q * 4;
            auto ldrow = [&](const T* p) -> float4 {
                if (vec_ok) return load4<T>(p);
                float4 v;
                v.x = ox < ow ? to_f(p[0]) : 0.0f;     v.y = ox + 1 < ow ? to_f(p[1]) : 0.0f;
                v.z = ox + 2 < ow ? to_f(p[2]) : 0.0f; v.w = ox + 3 < ow ? to_f(p[3]) : 0.0f;
                return v;
            };
            auto fold = [&](const float4& v, const float4& w4) {
                const float wy[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r][0] = fmaf(wy[r], v.x, acc[r][0]); acc[r][1] = fmaf(wy[r], v.y, acc[r][1]);
                    acc[r][2] = fmaf(wy[r], v.z, acc[r][2]); acc[r][3] = fmaf(wy[r], v.w, acc[r][3]);
                }
            };
            const T* gp = g + (size_t)(lo + rg) * ow + ox;       // this thread's rows: window rows rg, rg + rgroups, ...
            const size_t gstride = (size_t)rgroups * ow;
            int j = rg;
            for (; j + (BATCH - 1) * rgroups < nwin; j += BATCH * rgroups) {     // full batches: every load issued before the first FMA
                float4 v[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) v[u] = ldrow(gp + u * gstride);
                gp += BATCH * gstride;
#pragma unroll
                for (int u = 0; u < BATCH; ++u) fold(v[u], wts[j + u * rgroups]);
            }
            for (; j < nwin; j += rgroups) {
                const float4 v = ldrow(gp);
                gp += gstride;
                fold(v, wts[j]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                *reinterpret_cast<float4*>(colsum + (r * rgroups + rg) * pitch + ox) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    } else {   // very large scale factors: one column per thread, weights on the fly
        for (int r = 0; r < nr; ++r) {
            int rl, rh2;
            touch_window(iy0 + r, oh, rh, ac, rl, rh2);
            for (int ox = threadIdx.x; ox < pitch; ox += blockDim.x) {
                float acc = 0.0f;
                if (ox < ow)
                    for (int oy = rl; oy <= rh2; ++oy) {
                        const float wy = tap_weight(oy, iy0 + r, ih, rh, ac);
                        if (wy != 0.0f) acc += wy * to_f(g[(size_t)oy * ow + ox]);
                    }
                colsum[r * pitch + ox] = acc;
            }
        }
        rgroups = 1;
    }
    __syncthreads();
    if (rgroups > 1) {      // the row groups' partial column sums, folded once (group 0, 1, ... in order)
        for (int r = 0; r < nr; ++r)
            for (int q4 = threadIdx.x; q4 < ow4; q4 += blockDim.x) {
                float4* b4 = reinterpret_cast<float4*>(colsum + r * rgroups * pitch) + q4;
                float4 cs = b4[0];
                for (int k = 1; k < rgroups; ++k) {
                    const float4 t = b4[k * ow4];
                    cs.x += t.x; cs.y += t.y; cs.z += t.z; cs.w += t.w;
                }
                b4[0] = cs;
            }
        __syncthreads();
    }
    // pass 2: an input pixel is folded by a group of L lanes (L = the power of two that spreads the nr*iw pixels over the CTA:
    // 32 / 8 / 2 lanes at x32 / x16 / x8), each lane taking every L-th column of the window, then a fixed shuffle tree.
    const int items = nr * iw;
    int L = 32;
    while (L > 1 && items * L > (int)blockDim.x) L >>= 1;
    const int sub = threadIdx.x & (L - 1), per_pass = blockDim.x / L;
    const float irw = mm.irw[map];
    for (int it0 = 0; it0 < items; it0 += per_pass) {       // uniform trip count: the shuffles below need every lane
        const int it = it0 + threadIdx.x / L;
        float acc = 0.0f;
        int r = 0, ix = 0;
        if (it < items) {
            r = it / iw; ix = it - r * iw;
            int xl, xh;
            touch_window_r(ix, ow, rw, irw, ac, xl, xh);
            const float* base = colsum + r * rgroups * pitch;
            for (int ox = xl + sub; ox <= xh; ox += L) {
                const float2 t = xtab[ox];
                const int i0 = __float_as_int(t.x), i1 = min(i0 + 1, iw - 1);
                const float w = (i0 == ix ? 1.0f - t.y : 0.0f) + (i1 == ix ? t.y : 0.0f);      // tap_weight(ox, ix) from the table
                acc = fmaf(w, base[ox], acc);
            }
        }
        for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (it < items && sub == 0) din[((size_t)plane * ih + iy0 + r) * iw + ix] = from_f<T>(acc);
    }
}

template <typename T, int BWD_BATCH, int MIN_CTAS>
__global__ void __launch_bounds__(BWD_THREADS, MIN_CTAS)
bilinear_bwd_kernel(const __grid_constant__ MultiMaps mm, int oh, int ow, int ac, int rgroups) {
    pv2::pdl_prologue();
    extern __shared__ __align__(16) float colsum[];  // [BWD_R][rgroups][pitch] column sums, then [pitch] float2 x taps
    __shared__ float4 wts[BWD_MAX_WIN];
    const int map = blockIdx.z;
    const int R = bwd_rows_per_cta(mm.ih[map], oh);
    if (R == 4) bilinear_bwd_body<T, 4, BWD_BATCH>(mm, map, oh, ow, ac != 0, rgroups, colsum, wts);
    else if (R == 2) bilinear_bwd_body<T, 2, BWD_BATCH>(mm, map, oh, ow, ac != 0, rgroups, colsum, wts);
    else bilinear_bwd_body<T, 1, BWD_BATCH>(mm, map, oh, ow, ac != 0, rgroups, colsum, wts);
}

// ------------------------------------------------------------------------------------------------------------------------------
// Backward, integer up-scaling by s = 8, 16, 32, 64 with half-pixel centres (the final x8 / x16 / x32 maps of a step): separable
// in registers.  With s % 8 == 0 every aligned run of 4 output rows and every aligned quad of 4 output columns has ONE pair of
// source rows / columns (the cell boundaries sit at s/2 + k*s, multiples of 4).  A thread owns a column quad and a subset of the
// CTA's 4-row chunks; per chunk it issues 4 independent 16-byte loads, folds each row's quad into (a, b) = (sum w0x*v, sum w1x*v)
// and the four rows into (a, b) x (lo, hi) with the rows' y taps: 13 FMA per 16-byte load, 4.4 instructions per element (the
// gather kernel above spends ~30).  Pass 2 adds, for every input pixel, the chunk x quad partials of its own cell (lo / a) and of
// the cell before it (hi / b) in a fixed order.  CTA = (band of R input rows, plane, map); the band also reads the cell above it
// (its hi part): (R + 1) / R of the bytes.  No atomics, deterministic.  Taps come from pv2::bilinear_tap (ATen's formula).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int BWD2_THREADS = 384;
constexpr int BWD2_MAX_CHUNKS = 52;      // 4-row chunks a CTA may own: (R + 1) * s / 4 + s / 8 (cell 0 also holds the s/2 clamped rows)
constexpr int BWD2_MAX_IW = 96;

constexpr int BWD2_SMEM = 64 * 1024;     // partial sums of a CTA: 4 components x cells x ow / 4 floats = 4 * (R + 1) * ow bytes

// `bands` = bands per plane the host asks for: 2 when the launch already has ~2 CTAs per SM (the 8 final maps x 16 planes of a step
// are 288 CTAs, one wave at 2 CTAs per SM), more for launches with few planes
__host__ __device__ inline int bwd2_rows_per_cta(int ih, int s, int ow, int bands) {
    // (a cell is s/8 trips and cells are dealt whole to the 4 row groups: at x32 a band of 5 rows is 6 cells = 8 trips for two
    // groups and 4 for the others.  Twice the bands for x32 maps balances that, but the 8-map launch then has 320 CTAs, more than
    // the 296 resident slots: measured 24.5 us against 19.1.)
    int r = (ih + bands - 1) / bands;
    const int cap = (BWD2_MAX_CHUNKS - s / 8) * 4 / s - 1;         // window rows: (R + 1) * s + s / 2 <= 4 * BWD2_MAX_CHUNKS
    const int cap2 = BWD2_SMEM / (4 * ow) - 1;
    if (r > cap) r = cap;
    if (r > cap2) r = cap2;
    return r < 1 ? 1 : r;
}

template <typename T>
__global__ void __launch_bounds__(BWD2_THREADS, 2)
bilinear_bwd2_kernel(const __grid_constant__ MultiMaps mm, int oh, int ow, int G, int bands_hint) {
    pv2::pdl_prologue();
    extern __shared__ __align__(16) float part[];     // [4 components: a_lo, a_hi, b_lo, b_hi][cell][ow4]
    __shared__ float2 ytab[BWD2_MAX_CHUNKS * 4];      // (w0, w1) of every window row
    __shared__ int chunk_i0[BWD2_MAX_CHUNKS];         // source row i0 of a chunk
    __shared__ int cell_k0[BWD2_MAX_CHUNKS + 2];      // first chunk of cell (c0 + i); sentinel at the end
    __shared__ int quad_i0[BWD2_THREADS];
    __shared__ int run_q0[BWD2_MAX_IW + 2];           // first quad whose source column is >= ix; sentinel at iw
    const int map = blockIdx.z, plane = blockIdx.y;
    const int ih = mm.ih[map], iw = mm.iw[map];
    const int s = oh / ih;
    const int R = bwd2_rows_per_cta(ih, s, ow, bands_hint);
    const int ya = blockIdx.x * R;
    if (ya >= ih) return;
    const int yb = min(ya + R, ih), nr = yb - ya;
    const float rh = mm.rh[map], rw = mm.rw[map];
    const T* __restrict__ g = reinterpret_cast<const T*>(mm.out[map]) + (size_t)plane * oh * ow;
    T* __restrict__ din = reinterpret_cast<T*>(const_cast<void*>(mm.in[map]));
    // cells (= source row of the upper tap) c0 .. yb - 1; cell c covers output rows [c*s + s/2, (c+1)*s + s/2), cell 0 also the
    // clamped rows above it, cell ih - 1 ends at oh
    const int c0 = max(ya - 1, 0), ncell = yb - c0;
    const int row0 = c0 == 0 ? 0 : c0 * s + (s >> 1);
    const int row1 = yb == ih ? oh : yb * s + (s >> 1);
    const int nchunk = (row1 - row0) >> 2;
    const int ow4 = ow >> 2;
    const int tid = threadIdx.x;
    for (int j = tid; j < nchunk * 4; j += blockDim.x) {
        const Tap t = bilinear_tap(row0 + j, ih, rh, false);
        ytab[j] = make_float2(t.w0, t.w1);
        if ((j & 3) == 0) chunk_i0[j >> 2] = t.i0;
    }
    const int q = tid % ow4, grp = tid / ow4;
    float wx0[4], wx1[4];
    {
        int i0 = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const Tap t = bilinear_tap(q * 4 + e, iw, rw, false);
            wx0[e] = t.w0; wx1[e] = t.w1;
            if (e == 0) i0 = t.i0;
        }
        if (grp == 0) quad_i0[q] = i0;
    }
    __syncthreads();
    // first chunk of every cell / first quad of every source column (chunks and quads are sorted by their source index)
    for (int k = tid; k <= nchunk; k += blockDim.x) {
        const int prev = k == 0 ? c0 - 1 : chunk_i0[k - 1];
        const int cur = k == nchunk ? c0 + ncell : chunk_i0[k];
        for (int c = prev + 1; c <= cur; ++c) cell_k0[c - c0] = k;
    }
    for (int k = tid; k <= ow4; k += blockDim.x) {
        const int prev = k == 0 ? -1 : quad_i0[k - 1];
        const int cur = k == ow4 ? iw : quad_i0[k];
        for (int c = prev + 1; c <= cur; ++c) run_q0[c] = k;
    }
    __syncthreads();
    const int cstride = ncell * ow4;                   // component stride of part[]
    if (grp < G) {
        const T* gp = g + (size_t)row0 * ow + q * 4;
        // a thread folds whole cells (cell grp, grp + G, ...): the chunks of a cell accumulate in registers, two chunks per trip in ONE
        // interleaved loop so that all 8 independent 16-byte loads are issued before the first FMA (with two separate folds ptxas
        // sinks the second chunk's loads below the first fold: 4 in flight again).  (Dealing two-chunk work items instead of cells to
        // the groups balances a x32 band better on paper and measured slower, 22.8 us against 19.1: one more table, longer pass 2.)
        for (int cell = grp; cell < ncell; cell += G) {
            const int k0 = cell_k0[cell], k1 = cell_k0[cell + 1];
            float alo = 0.0f, ahi = 0.0f, blo = 0.0f, bhi = 0.0f;
            for (int k = k0; k < k1; k += 2) {
                const bool two = k + 1 < k1;
                const int k2 = two ? k + 1 : k;
                const float m2 = two ? 1.0f : 0.0f;
                const T* p = gp + (size_t)(k * 4) * ow;
                const T* p2 = gp + (size_t)(k2 * 4) * ow;
                float4 v[4], v2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = load4<T>(p + (size_t)u * ow);
#pragma unroll
                for (int u = 0; u < 4; ++u) v2[u] = load4<T>(p2 + (size_t)u * ow);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float a = fmaf(wx0[3], v[u].w, fmaf(wx0[2], v[u].z, fmaf(wx0[1], v[u].y, wx0[0] * v[u].x)));
                    const float a2 = fmaf(wx0[3], v2[u].w, fmaf(wx0[2], v2[u].z, fmaf(wx0[1], v2[u].y, wx0[0] * v2[u].x)));
                    const float b = fmaf(wx1[3], v[u].w, fmaf(wx1[2], v[u].z, fmaf(wx1[1], v[u].y, wx1[0] * v[u].x)));
                    const float b2 = fmaf(wx1[3], v2[u].w, fmaf(wx1[2], v2[u].z, fmaf(wx1[1], v2[u].y, wx1[0] * v2[u].x)));
                    const float2 wy = ytab[k * 4 + u];
                    float2 wy2 = ytab[k2 * 4 + u];
                    wy2.x *= m2; wy2.y *= m2;
                    alo = fmaf(wy.x, a, alo); ahi = fmaf(wy.y, a, ahi);
                    blo = fmaf(wy.x, b, blo); bhi = fmaf(wy.y, b, bhi);
                    alo = fmaf(wy2.x, a2, alo); ahi = fmaf(wy2.y, a2, ahi);
                    blo = fmaf(wy2.x, b2, blo); bhi = fmaf(wy2.y, b2, bhi);
                }
            }
            float* o = part + cell * ow4 + q;
            o[0] = alo; o[cstride] = ahi; o[2 * cstride] = blo; o[3 * cstride] = bhi;
        }
    }
    pv2::pdl_done();
    __syncthreads();
    // pass 2b: input pixel (y, ix) = lo parts of cell y + hi parts of cell y - 1 (and of cell y itself on the clamped last row),
    // 'a' parts of the quads of column ix + 'b' parts of the quads of column ix - 1 (and of ix itself on the clamped last column)
    for (int o = tid; o < nr * iw; o += blockDim.x) {
        const int r = o / iw, ix = o - r * iw, y = ya + r;
        float acc = 0.0f;
        auto fold = [&](int cell, int ysel) {          // ysel 0: lo, 1: hi
            const float* pa = part + ysel * cstride + (cell - c0) * ow4;
            const float* pb = part + (2 + ysel) * cstride + (cell - c0) * ow4;
            for (int qq = run_q0[ix]; qq < run_q0[ix + 1]; ++qq) acc += pa[qq];
            if (ix > 0)
                for (int qq = run_q0[ix - 1]; qq < run_q0[ix]; ++qq) acc += pb[qq];
            if (ix == iw - 1)
                for (int qq = run_q0[ix]; qq < run_q0[ix + 1]; ++qq) acc += pb[qq];
        };
        fold(y, 0);
        if (y > 0) fold(y - 1, 1);
        if (y == ih - 1) fold(y, 1);
        din[((size_t)plane * ih + y) * iw + ix] = from_f<T>(acc);
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

// Output rows per CTA: as tall as possible (a CTA's fixed cost -- staging, taps -- is amortised over its rows) while the launch still
// has ~1000 CTAs (7 per SM).  Measured on 8 maps x 16 planes x 352^2: 22.7 / 17.9 / 16.1 / 15.1 / 18.2 us for bands of 8 / 16 / 32 / 44 / 64.
static int fwd_band(int planes, int nmaps, int oh) {
    const int env = pv2::tune_int("PV2_BIL_BAND", 0);
    if (env > 0) return env > MAX_BAND ? MAX_BAND : env;
    const long long pm = (long long)planes * nmaps;
    const int bands = (int)((1024 + pm - 1) / pm);
    int b = (oh + bands - 1) / bands;
    return b < BAND ? BAND : (b > MAX_BAND ? MAX_BAND : b);
}

static int launch_fwd(const MultiMaps& mm, int nmaps, int planes, int oh, int ow, int align_corners, int dtype, size_t smem, int band, cudaStream_t st) {
    dim3 grid((oh + band - 1) / band, planes, nmaps);
    if (dtype == PV2_F32) pv2::launch_streaming(bilinear_fwd_kernel<float>, grid, FWD_THREADS, smem, st, mm, oh, ow, align_corners, band);
    else pv2::launch_streaming(bilinear_fwd_kernel<__nv_bfloat16>, grid, FWD_THREADS, smem, st, mm, oh, ow, align_corners, band);
    PV2_LAUNCH_CHECK("bilinear_fwd");
    return 0;
}

// grid.x of the backward for one map: input rows / rows per CTA (same rule as in the kernel)
static int bwd_row_blocks(int ih, int oh) {
    const int R = bwd_rows_per_cta(ih, oh);
    return (ih + R - 1) / R;
}

// the separable integer-scale backward applies when every map is an exact x8 / x16 / x32 / x64 half-pixel up-scaling
static bool bwd2_ok(const MultiMaps& mm, int nmaps, int oh, int ow, int align_corners) {
    static const bool off = [] { const char* e = getenv("PV2_BIL_BWD2"); return e && e[0] == '0'; }();
    if (off || align_corners || (ow & 3) || ow / 4 > BWD2_THREADS) return false;
    for (int i = 0; i < nmaps; ++i) {
        const int ih = mm.ih[i], iw = mm.iw[i];
        if (oh % ih || ow % iw) return false;
        const int s = oh / ih;
        if (s != ow / iw || (s & 7) || s > 64 || iw > BWD2_MAX_IW) return false;
        if (mm.rh[i] != 1.0f / (float)s || mm.rw[i] != 1.0f / (float)s) return false;
        const int chunks = (bwd2_rows_per_cta(ih, s, ow, 1) + 1) * s / 4 + s / 8;      // bands = 1: the tallest band the caps allow
        if (chunks > BWD2_MAX_CHUNKS || 4 * (bwd2_rows_per_cta(ih, s, ow, 1) + 1) * ow > BWD2_SMEM) return false;
    }
    return true;
}

static int launch_bwd2(const MultiMaps& mm, int nmaps, int planes, int oh, int ow, int dtype, cudaStream_t st) {
    const int ow4 = ow / 4;
    int G = BWD2_THREADS / ow4;
    if (G > 8) G = 8;
    int threads = (ow4 * G + 31) / 32 * 32;
    if (threads < 128) threads = 128;          // pass 2 and the tables still want a few warps
    int hint = 2 * kNumSMs / (planes * nmaps);       // rounded down: one wave
    if (hint < 2) hint = 2;
    int bands = 1, maxcell = 0;
    for (int i = 0; i < nmaps; ++i) {
        const int s = oh / mm.ih[i], R = bwd2_rows_per_cta(mm.ih[i], s, ow, hint);
        const int nb = (mm.ih[i] + R - 1) / R;
        if (nb > bands) bands = nb;
        if (R + 1 > maxcell) maxcell = R + 1;
    }
    const size_t smem = (size_t)4 * maxcell * ow4 * sizeof(float);
    dim3 grid(bands, planes, nmaps);
    if (dtype == PV2_F32) {
        static bool once = [] { return cudaFuncSetAttribute(bilinear_bwd2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD2_SMEM) == cudaSuccess; }();
        (void)once;
        pv2::launch_streaming(bilinear_bwd2_kernel<float>, grid, threads, smem, st, mm, oh, ow, G, hint);
    } else {
        static bool once = [] { return cudaFuncSetAttribute(bilinear_bwd2_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD2_SMEM) == cudaSuccess; }();
        (void)once;
        pv2::launch_streaming(bilinear_bwd2_kernel<__nv_bfloat16>, grid, threads, smem, st, mm, oh, ow, G, hint);
    }
    PV2_LAUNCH_CHECK("bilinear_bwd2");
    return 0;
}

static int launch_bwd(const MultiMaps& mm, int nmaps, int planes, int row_blocks, int oh, int ow, int align_corners, int dtype, cudaStream_t st) {
    if (bwd2_ok(mm, nmaps, oh, ow, align_corners)) return launch_bwd2(mm, nmaps, planes, oh, ow, dtype, st);
    const int ow4 = (ow + 3) / 4;
    PV2_CHECK(ow4 <= BWD_THREADS, "bilinear_bwd: output width %d too large", ow);
    int rgroups = BWD_THREADS / ow4;
    if (rgroups > 8) rgroups = 8;
    int threads = (ow4 * rgroups + 31) / 32 * 32;       // every thread owns a (column quad, row group): no idle lanes beyond the last warp
    if (threads < 64) threads = 64;
    const size_t smem = (size_t)BWD_R * rgroups * ow4 * 4 * sizeof(float) + (size_t)ow4 * 4 * sizeof(float2);
    PV2_CHECK(smem <= 48 * 1024, "bilinear_bwd: output width %d too large", ow);
    dim3 grid(row_blocks, planes, nmaps);
    // 4 rows in flight per thread at 3 CTAs per SM: measured 30.3 / 27.1 / 28.5 us for (8 rows, 2 CTAs) / (4, 3) / (2, 4) on 8 maps x 16 x 352^2
    if (dtype == PV2_F32) pv2::launch(bilinear_bwd_kernel<float, 4, 3>, grid, threads, smem, st, mm, oh, ow, align_corners, rgroups);
    else pv2::launch(bilinear_bwd_kernel<__nv_bfloat16, 4, 3>, grid, threads, smem, st, mm, oh, ow, align_corners, rgroups);
    PV2_LAUNCH_CHECK("bilinear_bwd");
    return 0;
}

// rows of the source a BAND of output rows can touch, if they fit in 48 KB of shared memory (else 0: read through L1/L2)
static int staged_rows(int ih, int iw, float rh, int band) {
    int want = (int)(band * (double)rh) + 4;
    if (want > ih) want = ih;
    const int cap = (48 * 1024) / (iw * 4);
    return want <= cap ? want : 0;
}

extern "C" int pv2_bilinear_fwd(const void* in, void* out, int planes, int ih, int iw, int oh, int ow,
                                float rh, float rw, int align_corners, int dtype, void* stream) {
    if (int e = check(in, out, planes, ih, iw, oh, ow, dtype, "bilinear_fwd")) return e;
    MultiMaps mm = {};
    mm.in[0] = in; mm.out[0] = out; mm.ih[0] = ih; mm.iw[0] = iw; mm.rh[0] = rh; mm.rw[0] = rw;
    const int band = fwd_band(planes, 1, oh);
    mm.max_rows[0] = staged_rows(ih, iw, rh, band);
    return launch_fwd(mm, 1, planes, oh, ow, align_corners, dtype, (size_t)mm.max_rows[0] * iw * 4, band, (cudaStream_t)stream);
}

extern "C" int pv2_bilinear_bwd(const void* dout, void* din, int planes, int ih, int iw, int oh, int ow,
                                float rh, float rw, int align_corners, int dtype, void* stream) {
    if (int e = check(dout, din, planes, ih, iw, oh, ow, dtype, "bilinear_bwd")) return e;
    PV2_CHECK(ih <= 65535 * 32, "bilinear_bwd: input height %d too large", ih);
    MultiMaps mm = {};
    mm.in[0] = din; mm.out[0] = const_cast<void*>(dout); mm.ih[0] = ih; mm.iw[0] = iw; mm.rh[0] = rh; mm.rw[0] = rw;
    mm.irh[0] = rh > 0.0f ? 1.0f / rh : 0.0f; mm.irw[0] = rw > 0.0f ? 1.0f / rw : 0.0f;
    return launch_bwd(mm, 1, planes, bwd_row_blocks(ih, oh), oh, ow, align_corners, dtype, (cudaStream_t)stream);
}

extern "C" int pv2_bilinear_multi_fwd(const void* const* in, void* const* out, const int* ih, const int* iw, const float* rh, const float* rw,
                                      int nmaps, int planes, int oh, int ow, int align_corners, int dtype, void* stream) {
    PV2_CHECK(in && out && ih && iw && rh && rw && nmaps >= 1 && nmaps <= PV2_MAX_MAPS, "bilinear_multi_fwd: 1..%d maps expected", PV2_MAX_MAPS);
    MultiMaps mm = {};
    size_t smem = 0;
    const int band = fwd_band(planes, nmaps, oh);
    for (int i = 0; i < nmaps; ++i) {
        if (int e = check(in[i], out[i], planes, ih[i], iw[i], oh, ow, dtype, "bilinear_multi_fwd")) return e;
        mm.in[i] = in[i]; mm.out[i] = out[i]; mm.ih[i] = ih[i]; mm.iw[i] = iw[i]; mm.rh[i] = rh[i]; mm.rw[i] = rw[i];
        mm.max_rows[i] = staged_rows(ih[i], iw[i], rh[i], band);
        const size_t b = (size_t)mm.max_rows[i] * iw[i] * 4;
        if (b > smem) smem = b;
    }
    return launch_fwd(mm, nmaps, planes, oh, ow, align_corners, dtype, smem, band, (cudaStream_t)stream);
}

extern "C" int pv2_bilinear_multi_bwd(const void* const* dout, void* const* din, const int* ih, const int* iw, const float* rh, const float* rw,
                                      int nmaps, int planes, int oh, int ow, int align_corners, int dtype, void* stream) {
    PV2_CHECK(dout && din && ih && iw && rh && rw && nmaps >= 1 && nmaps <= PV2_MAX_MAPS, "bilinear_multi_bwd: 1..%d maps expected", PV2_MAX_MAPS);
    MultiMaps mm = {};
    int max_blocks = 0;
    for (int i = 0; i < nmaps; ++i) {
        if (int e = check(dout[i], din[i], planes, ih[i], iw[i], oh, ow, dtype, "bilinear_multi_bwd")) return e;
        mm.in[i] = din[i]; mm.out[i] = const_cast<void*>(dout[i]); mm.ih[i] = ih[i]; mm.iw[i] = iw[i]; mm.rh[i] = rh[i]; mm.rw[i] = rw[i];
        mm.irh[i] = rh[i] > 0.0f ? 1.0f / rh[i] : 0.0f; mm.irw[i] = rw[i] > 0.0f ? 1.0f / rw[i] : 0.0f;
        const int nb = bwd_row_blocks(ih[i], oh);
        if (nb > max_blocks) max_blocks = nb;
    }
    return launch_bwd(mm, nmaps, planes, max_blocks, oh, ow, align_corners, dtype, (cudaStream_t)stream);
}
