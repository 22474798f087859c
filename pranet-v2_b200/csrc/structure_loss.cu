// structure_loss forward / backward (binary_seg/MyTrain_med.py:19-38).
//
//   weit = 1 + 5*|avg_pool31x31(mask) - mask|      (zero padding counted in the /961 divisor)
//   loss = mean_{n,c}[ wbce(pred, mask) + wiou(pred, mask) + 0.8*wbce(pred_bg, mask_bg) ]
//
// Kernels, all on the caller's stream (the default forward is 2b; 1 + 2 run when rows are not 16-byte vectorisable or with
// PV2_LOSS_TWO_PASS=1):
//   1. boundary_weight_kernel  -- per 32x64 tile: mask + 15-px halo staged in shared memory, separable running-sum
//      box filter out of shared memory, |avg - m| written ONCE as a 16-bit fixed-point map (2 B/px; weit is in
//      [1,6], quantisation 7.6e-5) plus the per-tile sum of weit.  The reference recomputes the 961-tap pool in each
//      of its 4 loss calls and again in autograd; here it is computed once per step and shared by all scales,
//      forward and backward.
//   2. structure_loss_fwd_kernel<T, NS, VEC> -- pure streaming: each CTA owns a 2048-px chunk of one (n,c) plane,
//      16-byte loads of logits / mask / weight map, MUFU-only transcendental math, the 4 weighted sums per scale
//      reduced warp-shuffle -> shared -> one partial per CTA (no atomics on data).  The last CTA to finish (ticket
//      counter) folds the partials in a fixed order into per-plane sums and the scalar losses: bit-reproducible.
//   2b. structure_loss_fwd_fused_kernel<T, NS, LOWRES> -- boundary weight (summed-area table in shared memory) + loss sums in one
//      pass over 32 x 128 tiles.  LOWRES = true reads the LOW-RESOLUTION head maps and does the model's final bilinear upsamples
//      (pranet.py:349-415) per pixel, so the full-resolution logits never exist.
//   3. structure_loss_bwd_kernel<T, NS, VEC> -- same streaming shape as 2, reads the finished plane sums, writes both
//      gradients with 16-byte stores.
//   3b. structure_loss_lowres_bwd_kernel<NS> + lowres_grad_fold_kernel<NS> -- backward of 2b/LOWRES: gradients scattered straight
//      into the low-resolution maps (register pre-reduction, shared-memory folds, per-tile slots, deterministic fold).
//
// Algorithmic HBM bytes per pixel (fp32 logits, NS scales): fwd 4 + 8*NS, bwd 4 + 16*NS; the weight map adds
// 2 B written once and 2 B read per pass.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

// ---------------------------------------------------------------------------------------------------------
// 1. boundary weight map
// ---------------------------------------------------------------------------------------------------------
constexpr int TH = 32, TW = 64, HALO = 15, KS = 31;
constexpr int SH = TH + 2 * HALO;   // 62 staged rows
constexpr int SW = TW + 2 * HALO;   // 94 staged cols
constexpr int SPITCH = SW + 1;      // odd pitch: lanes walking down rows hit distinct banks
constexpr int WT_THREADS = 256;
constexpr int ROWS_PER_THREAD = TH / (WT_THREADS / TW);  // 8
constexpr float INV_AREA = 1.0f / 961.0f;
constexpr float WQ = 65535.0f, INV_WQ = 1.0f / 65535.0f;

__device__ __forceinline__ float weit_from_q(uint32_t q) { return fmaf((float)q, 5.0f * INV_WQ, 1.0f); }

__global__ void __launch_bounds__(WT_THREADS)
boundary_weight_kernel(const float* __restrict__ mask, uint16_t* __restrict__ wmap, float* __restrict__ wsum_part,
                       unsigned int* __restrict__ ticket, double* __restrict__ plane_acc, int H, int W, int tiles_x, int tiles_per_plane) {
    pv2::pdl_prologue();
    __shared__ float sm[SH * SPITCH];
    __shared__ float hs[SH * TW];
    __shared__ float red[WT_THREADS / 32];
    const int plane = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (plane == 0 && tile == 0 && tid == 0) *ticket = 0u;   // the forward kernel that follows counts on this
    if (tile == 0 && tid < 4 * PV2_MAX_SCALES) plane_acc[(size_t)plane * (4 * PV2_MAX_SCALES) + tid] = 0.0;   // ... and on zeroed plane accumulators
    const int y0 = (tile / tiles_x) * TH, x0 = (tile % tiles_x) * TW;
    const float* mp = mask + (size_t)plane * H * W;
    // stage tile + halo: warp = staged row (stride 8), lane = staged column (stride 32): coalesced, no integer division
    for (int r = tid >> 5; r < SH; r += WT_THREADS / 32) {
        const int gy = y0 + r - HALO;
        const bool row_ok = (gy >= 0) && (gy < H);
        const float* src = mp + (size_t)(row_ok ? gy : 0) * W;
#pragma unroll
        for (int c = tid & 31; c < SW; c += 32) {
            const int gx = x0 + c - HALO;
            sm[r * SPITCH + c] = (row_ok && gx >= 0 && gx < W) ? __ldg(src + gx) : 0.0f;
        }
    }
    __syncthreads();
    // horizontal running sums: item = (row, segment of 8 outputs); rows vary fastest across lanes
    constexpr int SEG = 8, NSEG = TW / SEG;
    for (int it = tid; it < SH * NSEG; it += WT_THREADS) {
        const int r = it % SH, s = it / SH;
        const float* row = sm + r * SPITCH + s * SEG;
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < KS; ++j) acc += row[j];
        float* out = hs + r * TW + s * SEG;
        out[0] = acc;
#pragma unroll
        for (int j = 1; j < SEG; ++j) {
            acc += row[j + KS - 1] - row[j - 1];
            out[j] = acc;
        }
    }
    __syncthreads();
    // vertical running sums: thread = (column, block of 8 rows)
    const int x = tid % TW, rb = (tid / TW) * ROWS_PER_THREAD, gx = x0 + x;
    float acc = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int j = 0; j < KS; ++j) acc += hs[(rb + j) * TW + x];
    uint16_t* wp = wmap + (size_t)plane * H * W;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        if (j > 0) acc += hs[(rb + j + KS - 1) * TW + x] - hs[(rb + j - 1) * TW + x];
        const int gy = y0 + rb + j;
        if (gx < W && gy < H) {
            const float mv = sm[(rb + j + HALO) * SPITCH + x + HALO];
            const float d = fminf(fabsf(acc * INV_AREA - mv), 1.0f);
            const uint32_t q = (uint32_t)__float2int_rn(d * WQ);
            wp[(size_t)gy * W + gx] = (uint16_t)q;
            wsum += weit_from_q(q);       // sum the weights exactly as the loss kernels will see them
        }
    }
    wsum = warp_sum(wsum);
    if ((tid & 31) == 0) red[tid >> 5] = wsum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < WT_THREADS / 32; ++i) t += red[i];
        wsum_part[(size_t)plane * tiles_per_plane + tile] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// 2./3. streaming loss kernels
// ---------------------------------------------------------------------------------------------------------
constexpr int LS_THREADS = 256;
constexpr int CHUNK = 2048;        // pixels per CTA (one plane)
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

struct PtrPack {
    const void* pred[PV2_MAX_SCALES];
    const void* pred_bg[PV2_MAX_SCALES];
    void* dpred[PV2_MAX_SCALES];
    void* dpred_bg[PV2_MAX_SCALES];
};

template <typename T, int VEC> struct Vec;
template <typename T> struct Vec<T, 1> {
    float v[1];
    __device__ __forceinline__ void load(const T* p) { v[0] = to_f(*p); }
    __device__ __forceinline__ void store(T* p) const { *p = from_f<T>(v[0]); }
};
template <typename T> struct Vec<T, 4> {
    float v[4];
    __device__ __forceinline__ void load(const T* p) { const float4 f = load4<T>(p); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
    __device__ __forceinline__ void store(T* p) const { store4<T>(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <int VEC> __device__ __forceinline__ void load_wq(const uint16_t* p, float (&w)[VEC]);
template <> __device__ __forceinline__ void load_wq<1>(const uint16_t* p, float (&w)[1]) { w[0] = weit_from_q(*p); }
template <> __device__ __forceinline__ void load_wq<4>(const uint16_t* p, float (&w)[4]) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = weit_from_q(u.x & 0xffffu); w[1] = weit_from_q(u.x >> 16);
    w[2] = weit_from_q(u.y & 0xffffu); w[3] = weit_from_q(u.y >> 16);
}

// workspace: ticket | plane_sums [planes][1+4*MAX] | plane_loss [planes][MAX] | wsum_part | partials | wmap (256-B aligned pieces)
constexpr int NSUM = 1 + 4 * PV2_MAX_SCALES;

struct Layout {
    unsigned int* ticket;
    float* plane_sums;
    float* plane_loss;
    float* wsum_part;
    float* partials;     // [planes][chunks][4*MAX]
    uint16_t* wmap;
    int wt_tiles_x, wt_tiles, chunks;
    int ft_tiles_x, ft_tiles;       // fused forward: 32 x 128 tiles
    size_t bytes;
};
inline Layout make_layout(void* ws, int planes, int H, int W) {
    Layout L;
    L.wt_tiles_x = (W + TW - 1) / TW;
    L.wt_tiles = L.wt_tiles_x * ((H + TH - 1) / TH);
    L.chunks = (H * W + CHUNK - 1) / CHUNK;
    L.ft_tiles_x = (W + 127) / 128;
    L.ft_tiles = L.ft_tiles_x * ((H + 31) / 32);
    const int np = L.chunks > L.ft_tiles ? L.chunks : L.ft_tiles;      // partial slots: whichever forward runs
    const int nw = L.wt_tiles > L.ft_tiles ? L.wt_tiles : L.ft_tiles;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    uint8_t* b = (uint8_t*)ws;
    L.ticket = (unsigned int*)(b + take(256));
    L.plane_sums = (float*)(b + take(sizeof(float) * (size_t)planes * NSUM));
    L.plane_loss = (float*)(b + take(sizeof(float) * (size_t)planes * PV2_MAX_SCALES));
    L.wsum_part = (float*)(b + take(sizeof(float) * (size_t)planes * nw));
    L.partials = (float*)(b + take(sizeof(float) * (size_t)planes * np * 4 * PV2_MAX_SCALES));
    L.wmap = (uint16_t*)(b + take(sizeof(uint16_t) * (size_t)planes * H * W));
    L.bytes = off;
    return L;
}

template <typename T, int NS, int VEC>
__global__ void __launch_bounds__(LS_THREADS)
structure_loss_fwd_kernel(PtrPack pp, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                          const uint16_t* __restrict__ wmap, int HW, int planes, int chunks, int chunk_px, int wt_tiles,
                          float* __restrict__ partials, const float* __restrict__ wsum_part, float* __restrict__ plane_sums,
                          float* __restrict__ plane_loss, float* __restrict__ loss, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    __shared__ float red[LS_THREADS / 32][4 * NS];
    __shared__ bool is_last;
    const int plane = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const size_t pbase = (size_t)plane * HW;
    const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
    float acc[4 * NS];
#pragma unroll
    for (int i = 0; i < 4 * NS; ++i) acc[i] = 0.0f;
    for (int p = p0 + tid * VEC; p < p1; p += LS_THREADS * VEC) {
        float w[VEC];
        Vec<float, VEC> m, mb;
        Vec<T, VEC> x[NS], xb[NS];
        load_wq<VEC>(wmap + pbase + p, w);
        m.load(mask_fg + pbase + p);
#pragma unroll
        for (int k = 0; k < NS; ++k) {   // all loads of this quad in flight before the math
            x[k].load(reinterpret_cast<const T*>(pp.pred[k]) + pbase + p);
            xb[k].load(reinterpret_cast<const T*>(pp.pred_bg[k]) + pbase + p);
        }
        if (mask_bg != nullptr) mb.load(mask_bg + pbase + p);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) mb.v[j] = 1.0f - m.v[j];
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float xv = x[k].v[j], qv = xb[k].v[j], mv = m.v[j], wv = w[j];
                const float e = fast_ex2(-fabsf(xv) * LOG2E), d = 1.0f + e;
                const float inv = fast_rcp(d);
                const float sig = xv >= 0.0f ? inv : e * inv;
                const float bce = fmaf(-xv, mv, fmaxf(xv, 0.0f)) + fast_lg2(d) * LN2;
                const float e2 = fast_ex2(-fabsf(qv) * LOG2E);
                const float bce2 = fmaf(-qv, mb.v[j], fmaxf(qv, 0.0f)) + fast_lg2(1.0f + e2) * LN2;
                const float sw = sig * wv;
                acc[4 * k + 0] = fmaf(wv, bce, acc[4 * k + 0]);
                acc[4 * k + 1] = fmaf(wv, bce2, acc[4 * k + 1]);
                acc[4 * k + 2] = fmaf(sw, mv, acc[4 * k + 2]);            // inter = sum sig*m*w
                acc[4 * k + 3] = fmaf(mv, wv, acc[4 * k + 3] + sw);       // union = sum (sig+m)*w
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < 4 * NS; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    // The CTA's sums are ADDED, in double precision, to the plane's accumulators (zeroed by boundary_weight_kernel, re-zeroed by the
    // finalizing CTA below): no per-CTA partial rows and no fold over the ~61 chunks of a plane.  In double the order in which the
    // chunks arrive moves the sums by ~1e-16 relative -- below the rounding of the float the loss is reported in.
    double* pacc = reinterpret_cast<double*>(partials);
    if (tid < 4 * NS) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < LS_THREADS / 32; ++wi) v += red[wi][tid];
        atomicAdd(pacc + (size_t)plane * (4 * PV2_MAX_SCALES) + tid, (double)v);
    }
    // ---- the last CTA to finish turns the plane sums into the losses ----
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)(planes * chunks) - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int pl = warp; pl < planes; pl += LS_THREADS / 32) {
        float Wp = 0.0f;
        for (int t = lane; t < wt_tiles; t += 32) Wp += __ldcg(wsum_part + (size_t)pl * wt_tiles + t);
        Wp = warp_sum(Wp);
        double* pa = pacc + (size_t)pl * (4 * PV2_MAX_SCALES);
        float mine = 0.0f;
        if (lane < 4 * NS) { mine = (float)__ldcg(pa + lane); pa[lane] = 0.0; }     // leave the accumulators zeroed for the next forward
        float s[4 * NS];
#pragma unroll
        for (int i = 0; i < 4 * NS; ++i) s[i] = __shfl_sync(0xffffffffu, mine, i);
        if (lane == 0) {
            float* ps = plane_sums + (size_t)pl * NSUM;
            ps[0] = Wp;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                ps[1 + 4 * k + 0] = s[4 * k + 0]; ps[1 + 4 * k + 1] = s[4 * k + 1];
                ps[1 + 4 * k + 2] = s[4 * k + 2]; ps[1 + 4 * k + 3] = s[4 * k + 3];
                const float inter = s[4 * k + 2], uni = s[4 * k + 3];
                plane_loss[(size_t)pl * PV2_MAX_SCALES + k] =
                    s[4 * k + 0] / Wp + 1.0f - (inter + 1.0f) / (uni - inter + 1.0f) + 0.8f * s[4 * k + 1] / Wp;
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (warp == 0) {   // mean over planes, fixed order
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            float v = 0.0f;
            for (int pl = lane; pl < planes; pl += 32) v += __ldcg(plane_loss + (size_t)pl * PV2_MAX_SCALES + k);
            v = warp_sum(v);
            if (lane == 0) loss[k] = v / (float)planes;
        }
        if (lane == 0) *ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------
// 2b. fused forward: boundary weight + loss sums in ONE pass (the path taken whenever rows are 16-byte vectorisable).
// CTA = 32 x 128 pixel tile of one plane.  The mask tile with its 15-px halo is turned into a summed-area table in
// shared memory (row prefix by warp scan while staging, column prefix by one thread per column), so the 31x31 box
// sum of a pixel is 4 shared-memory reads; the weight is quantised exactly as the backward will read it, written to the
// 16-bit map, and used immediately on the logits of that pixel, which stream through once (16-byte loads).  HBM bytes per
// pixel: 4 (mask) + 8*NS (logits) read, 2 written.  The last CTA (ticket) folds the per-tile sums in a fixed order.
// ---------------------------------------------------------------------------------------------------------
constexpr int FT_H = 32, FT_W = 128;
constexpr int FS_H = FT_H + 2 * HALO + 1;      // 63 table rows: row 0 is the zero border of the summed-area table
constexpr int FS_W = FT_W + 2 * HALO + 1;      // 159
constexpr int FS_PITCH = 161;                  // odd: column walks and row scans are both bank-conflict free
constexpr int FS_PER_LANE = 5;                 // 32 lanes x 5 = 160 >= 158 staged columns

// ---- loss from the low-resolution maps: geometry of the final upsamples (align_corners=False), tap tables, interpolation ----
struct LowresGeo {
    int ih[PV2_MAX_SCALES], iw[PV2_MAX_SCALES];
    float rh[PV2_MAX_SCALES], rw[PV2_MAX_SCALES];     // ATen's source-index ratios (1/scale_factor)
};

// (i0, w1) of the bilinear tap of every column / row of a 32 x 128 tile, per scale; i1 = min(i0 + 1, size - 1), w0 = 1 - w1 as in
// pv2::bilinear_tap.  Coordinates beyond the image are clamped to its last row / column (those pixels carry no weight).
template <int NS>
__device__ __forceinline__ void fill_tap_tables(float2 (*xt)[FT_W], float2 (*yt)[FT_H], const LowresGeo& geo, int y0, int x0, int H, int W, int tid) {
    for (int i = tid; i < NS * FT_W; i += LS_THREADS) {
        const int k = i / FT_W, c = i - k * FT_W;
        const Tap t = bilinear_tap(min(x0 + c, W - 1), geo.iw[k], geo.rw[k], false);
        xt[k][c] = make_float2(__int_as_float(t.i0), t.w1);
    }
    for (int i = tid; i < NS * FT_H; i += LS_THREADS) {
        const int k = i / FT_H, r = i - k * FT_H;
        const Tap t = bilinear_tap(min(y0 + r, H - 1), geo.ih[k], geo.rh[k], false);
        yt[k][r] = make_float2(__int_as_float(t.i0), t.w1);
    }
}

// the four upsampled logits of a column quad, foreground and background map at once (same taps); `xq` = the quad's 4 table entries
__device__ __forceinline__ void interp_quad2(const float* __restrict__ fg, const float* __restrict__ bg, int ih, int iw, float2 ytap,
                                             const float2* xq, float (&vf)[4], float (&vb)[4]) {
    const int y0i = __float_as_int(ytap.x), y1i = min(y0i + 1, ih - 1);
    const float wy1 = ytap.y, wy0 = 1.0f - wy1;
    const float4 q01 = *reinterpret_cast<const float4*>(xq), q23 = *reinterpret_cast<const float4*>(xq + 2);
    const int xi[4] = {__float_as_int(q01.x), __float_as_int(q01.z), __float_as_int(q23.x), __float_as_int(q23.z)};
    const float xw1[4] = {q01.y, q01.w, q23.y, q23.w};
    const float* fa = fg + (size_t)y0i * iw;
    const float* fb = fg + (size_t)y1i * iw;
    const float* ba = bg + (size_t)y0i * iw;
    const float* bb = bg + (size_t)y1i * iw;
    if (xi[0] == xi[3]) {            // integer up-scaling by a multiple of 8: the quad shares its two source columns (4 loads per map, not 16)
        const int c0 = xi[0], c1 = min(c0 + 1, iw - 1);
        const float f00 = __ldg(fa + c0), f01 = __ldg(fa + c1), f10 = __ldg(fb + c0), f11 = __ldg(fb + c1);
        const float b00 = __ldg(ba + c0), b01 = __ldg(ba + c1), b10 = __ldg(bb + c0), b11 = __ldg(bb + c1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // same expression, same order as bilinear_fwd_kernel
            const float xw0 = 1.0f - xw1[j];
            vf[j] = wy0 * (xw0 * f00 + xw1[j] * f01) + wy1 * (xw0 * f10 + xw1[j] * f11);
            vb[j] = wy0 * (xw0 * b00 + xw1[j] * b01) + wy1 * (xw0 * b10 + xw1[j] * b11);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c0 = xi[j], c1 = min(c0 + 1, iw - 1);
            const float xw0 = 1.0f - xw1[j];
            vf[j] = wy0 * (xw0 * __ldg(fa + c0) + xw1[j] * __ldg(fa + c1)) + wy1 * (xw0 * __ldg(fb + c0) + xw1[j] * __ldg(fb + c1));
            vb[j] = wy0 * (xw0 * __ldg(ba + c0) + xw1[j] * __ldg(ba + c1)) + wy1 * (xw0 * __ldg(bb + c0) + xw1[j] * __ldg(bb + c1));
        }
    }
}

// LOWRES: pp.pred / pp.pred_bg are the LOW-RESOLUTION head maps (fp32 planes of geo.ih x geo.iw) and the
// final F.interpolate(scale_factor=8|16|32, mode='bilinear') of pranet.py:349-415 happens here, per pixel, from tap tables in
// shared memory: the eight full-resolution logit maps are never written or read (HBM bytes per pixel: 4 read + 2 written).
template <typename T, int NS, bool LOWRES>
__global__ void __launch_bounds__(LS_THREADS, 4)
structure_loss_fwd_fused_kernel(PtrPack pp, const __grid_constant__ LowresGeo geo, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                                uint16_t* __restrict__ wmap, int H, int W, int planes, int tiles_x, int tiles,
                                float* __restrict__ partials, float* __restrict__ wsum_part, float* __restrict__ plane_sums,
                                float* __restrict__ plane_loss, float* __restrict__ loss, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    __shared__ float sat[FS_H * FS_PITCH];
    __shared__ float red[LS_THREADS / 32][4 * NS + 1];
    __shared__ bool is_last;
    __shared__ __align__(16) float2 xt[LOWRES ? NS : 1][LOWRES ? FT_W : 1];   // per scale and tile column / row: (i0 as bits, w1) of the bilinear tap
    __shared__ float2 yt[LOWRES ? NS : 1][LOWRES ? FT_H : 1];
    const int plane = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int y0 = (tile / tiles_x) * FT_H, x0 = (tile % tiles_x) * FT_W;
    const int HW = H * W;
    const size_t pbase = (size_t)plane * HW;
    const float* mp = mask_fg + pbase;
    if constexpr (LOWRES) fill_tap_tables<NS>(xt, yt, geo, y0, x0, H, W, tid);      // visible after the first __syncthreads below
    // ---- stage + row prefix: warp = table row, lane = 5 consecutive columns ----
    if (tid < FS_PITCH) sat[tid] = 0.0f;                       // row 0
    constexpr int ROWS_PW = (FS_H - 1 + LS_THREADS / 32 - 1) / (LS_THREADS / 32);   // 8 table rows per warp
    float v[ROWS_PW][FS_PER_LANE];
    // all global loads of this warp's rows first (independent), then the scans: one memory latency instead of eight
#pragma unroll
    for (int i = 0; i < ROWS_PW; ++i) {
        const int r = 1 + warp + i * (LS_THREADS / 32);
        const int gy = y0 - (HALO + 1) + r;
        const bool row_ok = r < FS_H && gy >= 0 && gy < H;
        const float* src = mp + (size_t)(row_ok ? gy : 0) * W;
#pragma unroll
        for (int j = 0; j < FS_PER_LANE; ++j) {
            const int c = 1 + lane * FS_PER_LANE + j;          // table column
            const int gx = x0 - (HALO + 1) + c;
            v[i][j] = (row_ok && c < FS_W && gx >= 0 && gx < W) ? __ldg(src + gx) : 0.0f;
        }
    }
#pragma unroll
    for (int i = 0; i < ROWS_PW; ++i) {
        const int r = 1 + warp + i * (LS_THREADS / 32);
        if (r >= FS_H) break;
        float run = 0.0f;
#pragma unroll
        for (int j = 0; j < FS_PER_LANE; ++j) { run += v[i][j]; v[i][j] = run; }
        float incl = run;                                      // inclusive scan of the lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const float base = incl - run;
        float* row = sat + r * FS_PITCH;
        if (lane == 0) row[0] = 0.0f;                          // column 0
#pragma unroll
        for (int j = 0; j < FS_PER_LANE; ++j) {
            const int c = 1 + lane * FS_PER_LANE + j;
            if (c < FS_W) row[c] = base + v[i][j];
        }
    }
    __syncthreads();
    // ---- column prefix: one warp per column, lanes over row pairs, shuffle scan (a serial walk would cost 62 dependent
    //      shared-memory round trips) ----
    for (int c = 1 + warp; c < FS_W; c += LS_THREADS / 32) {
        const int r0 = 1 + 2 * lane;                           // rows r0, r0 + 1; 62 rows = 31 lanes
        float a = 0.0f, b2 = 0.0f;
        if (r0 < FS_H) { a = sat[r0 * FS_PITCH + c]; b2 = a + sat[(r0 + 1) * FS_PITCH + c]; }
        float incl = b2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const float base = incl - b2;
        if (r0 < FS_H) { sat[r0 * FS_PITCH + c] = base + a; sat[(r0 + 1) * FS_PITCH + c] = base + b2; }
    }
    __syncthreads();
    // ---- stream the tile: thread = 4 consecutive pixels of a row, warp = row ----
    float acc[4 * NS];
#pragma unroll
    for (int i = 0; i < 4 * NS; ++i) acc[i] = 0.0f;
    float wsum = 0.0f;
    const int tx = lane * 4, gx = x0 + tx;
    for (int ty = warp; ty < FT_H; ty += LS_THREADS / 32) {
        const int gy = y0 + ty;
        if (gy >= H || gx >= W) continue;                      // W % 4 == 0: a quad is inside or outside as a whole
        const size_t p = pbase + (size_t)gy * W + gx;
        Vec<float, 4> m, mb;
        Vec<T, 4> x[NS], xb[NS];
        m.load(mask_fg + p);
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if constexpr (LOWRES) {
                const size_t lp = (size_t)plane * geo.ih[k] * geo.iw[k];
                interp_quad2(reinterpret_cast<const float*>(pp.pred[k]) + lp, reinterpret_cast<const float*>(pp.pred_bg[k]) + lp,
                             geo.ih[k], geo.iw[k], yt[k][ty], &xt[k][tx], x[k].v, xb[k].v);
            } else {
                x[k].load(reinterpret_cast<const T*>(pp.pred[k]) + p);
                xb[k].load(reinterpret_cast<const T*>(pp.pred_bg[k]) + p);
            }
        }
        if (mask_bg != nullptr) mb.load(mask_bg + p);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) mb.v[j] = 1.0f - m.v[j];
        }
        // 31x31 box sums from the table: S[ty+31][tx+31+j] - S[ty][tx+31+j] - S[ty+31][tx+j] + S[ty][tx+j]
        const float* s_lo = sat + ty * FS_PITCH + tx;
        const float* s_hi = sat + (ty + KS) * FS_PITCH + tx;
        float w[4];
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float box = (s_hi[KS + j] - s_lo[KS + j]) - (s_hi[j] - s_lo[j]);
            const float d = fminf(fabsf(box * INV_AREA - m.v[j]), 1.0f);
            q[j] = (uint32_t)__float2int_rn(d * WQ);
            w[j] = weit_from_q(q[j]);
            wsum += w[j];
        }
        *reinterpret_cast<uint2*>(wmap + p) = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
#pragma unroll
        for (int k = 0; k < NS; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xv = x[k].v[j], qv = xb[k].v[j], mv = m.v[j], wv = w[j];
                const float e = fast_ex2(-fabsf(xv) * LOG2E), d = 1.0f + e;
                const float inv = fast_rcp(d);
                const float sig = xv >= 0.0f ? inv : e * inv;
                const float bce = fmaf(-xv, mv, fmaxf(xv, 0.0f)) + fast_lg2(d) * LN2;
                const float e2 = fast_ex2(-fabsf(qv) * LOG2E);
                const float bce2 = fmaf(-qv, mb.v[j], fmaxf(qv, 0.0f)) + fast_lg2(1.0f + e2) * LN2;
                const float sw = sig * wv;
                acc[4 * k + 0] = fmaf(wv, bce, acc[4 * k + 0]);
                acc[4 * k + 1] = fmaf(wv, bce2, acc[4 * k + 1]);
                acc[4 * k + 2] = fmaf(sw, mv, acc[4 * k + 2]);
                acc[4 * k + 3] = fmaf(mv, wv, acc[4 * k + 3] + sw);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * NS; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    wsum = warp_sum(wsum);
    if (lane == 0) red[warp][4 * NS] = wsum;
    __syncthreads();
    if (tid <= 4 * NS) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < LS_THREADS / 32; ++wi) v += red[wi][tid];
        if (tid < 4 * NS) partials[((size_t)plane * tiles + tile) * (4 * PV2_MAX_SCALES) + tid] = v;
        else wsum_part[(size_t)plane * tiles + tile] = v;
    }
    // ---- the last CTA to finish folds everything in a fixed order ----
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)(planes * tiles) - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int pl = warp; pl < planes; pl += LS_THREADS / 32) {
        float Wp = 0.0f;
        for (int t = lane; t < tiles; t += 32) Wp += __ldcg(wsum_part + (size_t)pl * tiles + t);
        Wp = warp_sum(Wp);
        float s[4 * NS];
#pragma unroll
        for (int i = 0; i < 4 * NS; ++i) s[i] = 0.0f;
        for (int c = lane; c < tiles; c += 32) {
            const float* src = partials + ((size_t)pl * tiles + c) * (4 * PV2_MAX_SCALES);
#pragma unroll
            for (int i = 0; i < 4 * NS; ++i) s[i] += __ldcg(src + i);
        }
#pragma unroll
        for (int i = 0; i < 4 * NS; ++i) s[i] = warp_sum(s[i]);
        if (lane == 0) {
            float* ps = plane_sums + (size_t)pl * NSUM;
            ps[0] = Wp;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                ps[1 + 4 * k + 0] = s[4 * k + 0]; ps[1 + 4 * k + 1] = s[4 * k + 1];
                ps[1 + 4 * k + 2] = s[4 * k + 2]; ps[1 + 4 * k + 3] = s[4 * k + 3];
                const float inter = s[4 * k + 2], uni = s[4 * k + 3];
                plane_loss[(size_t)pl * PV2_MAX_SCALES + k] =
                    s[4 * k + 0] / Wp + 1.0f - (inter + 1.0f) / (uni - inter + 1.0f) + 0.8f * s[4 * k + 1] / Wp;
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (warp == 0) {   // mean over planes, fixed order
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            float v = 0.0f;
            for (int pl = lane; pl < planes; pl += 32) v += __ldcg(plane_loss + (size_t)pl * PV2_MAX_SCALES + k);
            v = warp_sum(v);
            if (lane == 0) loss[k] = v / (float)planes;
        }
        if (lane == 0) *ticket = 0u;
    }
}

template <typename T, int NS, int VEC>
__global__ void __launch_bounds__(LS_THREADS)
structure_loss_bwd_kernel(PtrPack pp, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                          const uint16_t* __restrict__ wmap, const float* __restrict__ grad_loss,
                          const float* __restrict__ plane_sums, int HW, int planes) {
    pv2::pdl_prologue();
    const int plane = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const size_t pbase = (size_t)plane * HW;
    const int p0 = chunk * CHUNK, p1 = min(HW, p0 + CHUNK);
    const float* ps = plane_sums + (size_t)plane * NSUM;
    const float invW = 1.0f / ps[0], invn = 1.0f / (float)planes;
    float g[NS], den[NS], ip1[NS], inv_den2[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        g[k] = grad_loss[k] * invn;
        const float inter = ps[1 + 4 * k + 2], uni = ps[1 + 4 * k + 3];
        den[k] = uni - inter + 1.0f;
        ip1[k] = inter + 1.0f;
        inv_den2[k] = 1.0f / (den[k] * den[k]);
    }
    for (int p = p0 + tid * VEC; p < p1; p += LS_THREADS * VEC) {
        float w[VEC];
        Vec<float, VEC> m, mb;
        Vec<T, VEC> x[NS], xb[NS];
        load_wq<VEC>(wmap + pbase + p, w);
        m.load(mask_fg + pbase + p);
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            x[k].load(reinterpret_cast<const T*>(pp.pred[k]) + pbase + p);
            xb[k].load(reinterpret_cast<const T*>(pp.pred_bg[k]) + pbase + p);
        }
        if (mask_bg != nullptr) mb.load(mask_bg + pbase + p);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) mb.v[j] = 1.0f - m.v[j];
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            Vec<T, VEC> gp, gq;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float xv = x[k].v[j], qv = xb[k].v[j], mv = m.v[j], wv = w[j];
                const float e = fast_ex2(-fabsf(xv) * LOG2E), inv = fast_rcp(1.0f + e);
                const float s = xv >= 0.0f ? inv : e * inv;
                const float e2 = fast_ex2(-fabsf(qv) * LOG2E), inv2 = fast_rcp(1.0f + e2);
                const float s2 = qv >= 0.0f ? inv2 : e2 * inv2;
                const float mw = mv * wv;
                // d wiou / d sigma = -[ m w den - (inter+1)(w - m w) ] / den^2
                const float dwiou = -(mw * den[k] - ip1[k] * (wv - mw)) * inv_den2[k];
                gp.v[j] = g[k] * (wv * (s - mv) * invW + dwiou * s * (1.0f - s));
                gq.v[j] = g[k] * 0.8f * wv * (s2 - mb.v[j]) * invW;
            }
            gp.store(reinterpret_cast<T*>(pp.dpred[k]) + pbase + p);
            gq.store(reinterpret_cast<T*>(pp.dpred_bg[k]) + pbase + p);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 3b. backward of the loss from the low-resolution maps: d loss / d (low-res fg_k, bg_k) directly.
//
// CTA = 32 x 128 pixel tile of one plane, warp = tile row, lane = quad of 4 columns.  Per scale: the per-pixel gradients of the
// (recomputed) upsampled logits are multiplied by their x tap weights and pre-reduced in registers over the quad (a quad touches at
// most 3 source columns when the ratio is <= 1/4), folded along x over the lanes, then along y over the 32 rows -- all in shared
// memory, fixed order.  The tile's contribution to the few low-res pixels it touches goes to a private slot; lowres_grad_fold_kernel
// sums the <= 6 slots of every low-res pixel in tile order: no atomics, bit-reproducible.  The full-resolution gradients (8 x 8 MB at
// B = 16 x 352^2) and the bilinear backward that would read them back do not exist on this path.
// ---------------------------------------------------------------------------------------------------------
constexpr int LR_RMAX = 10, LR_CMAX = 34;              // low-res rows / columns a 32 x 128 tile can touch at ratio <= 1/4
constexpr int LR_SLOT = LR_RMAX * LR_CMAX;             // floats per (tile, scale, map) slot
constexpr int LR_QPITCH = 32 * 3 + 1;                  // Q row pitch (odd: the x fold walks rows across lanes)
constexpr int LR_HPITCH = FT_H + 1;

// low-res rows (or columns) a tile edge of `n` pixels starting at o0 touches: [first, first + count)
__device__ __forceinline__ void tile_span(int o0, int n, int out_size, int in_size, float ratio, int& first, int& count) {
    first = bilinear_tap(min(o0, out_size - 1), in_size, ratio, false).i0;
    count = bilinear_tap(min(o0 + n - 1, out_size - 1), in_size, ratio, false).i1 - first + 1;
}

template <int NS>
__global__ void __launch_bounds__(LS_THREADS, 2)      // 128 registers: measured 61.3 / 68.7 / 70.2 us at 2 / 3 / 4 CTAs per SM (the tighter budgets spill)
structure_loss_lowres_bwd_kernel(const __grid_constant__ PtrPack pp, const __grid_constant__ LowresGeo geo, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                                 const uint16_t* __restrict__ wmap, const float* __restrict__ grad_loss, const float* __restrict__ plane_sums,
                                 float* __restrict__ gpart, int H, int W, int planes, int tiles_x, int tiles) {
    pv2::pdl_prologue();
    __shared__ __align__(16) float2 xt[NS][FT_W];
    __shared__ float2 yt[NS][FT_H];
    __shared__ float Q[2][FT_H][LR_QPITCH];            // per map, tile row, lane: the quad's contribution to source columns cb, cb+1, cb+2
    __shared__ float Hs[2][LR_CMAX][LR_HPITCH];        // per map, low-res column, tile row: x-folded
    const int plane = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int y0 = (tile / tiles_x) * FT_H, x0 = (tile % tiles_x) * FT_W;
    const size_t pbase = (size_t)plane * H * W;
    fill_tap_tables<NS>(xt, yt, geo, y0, x0, H, W, tid);
    const float* ps = plane_sums + (size_t)plane * NSUM;
    const float invW = 1.0f / ps[0], invn = 1.0f / (float)planes;
    // mask and boundary weight of this thread's 4 rows x 4 columns (loaded once, used by every scale)
    constexpr int RPT = FT_H / (LS_THREADS / 32);      // 4 rows per thread
    const int tx = lane * 4, gx = x0 + tx;
    float mv[RPT][4], mbv[RPT][4], wv[RPT][4];
    bool ok[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int gy = y0 + warp + i * (LS_THREADS / 32);
        ok[i] = gy < H && gx < W;                      // W % 4 == 0: a quad is inside or outside as a whole
        if (ok[i]) {
            const size_t p = pbase + (size_t)gy * W + gx;
            Vec<float, 4> m;
            m.load(mask_fg + p);
            load_wq<4>(wmap + p, wv[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mv[i][j] = m.v[j];
            if (mask_bg != nullptr) {
                Vec<float, 4> mb;
                mb.load(mask_bg + p);
#pragma unroll
                for (int j = 0; j < 4; ++j) mbv[i][j] = mb.v[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) mbv[i][j] = 1.0f - m.v[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { mv[i][j] = 0.0f; mbv[i][j] = 0.0f; wv[i][j] = 0.0f; }
        }
    }
    __syncthreads();                                   // tap tables
#pragma unroll 1
    for (int k = 0; k < NS; ++k) {
        const int ih = geo.ih[k], iw = geo.iw[k];
        const float g = grad_loss[k] * invn;
        const float inter = ps[1 + 4 * k + 2], uni = ps[1 + 4 * k + 3];
        const float den = uni - inter + 1.0f, ip1 = inter + 1.0f, inv_den2 = 1.0f / (den * den);
        const float* fgp = reinterpret_cast<const float*>(pp.pred[k]) + (size_t)plane * ih * iw;
        const float* bgp = reinterpret_cast<const float*>(pp.pred_bg[k]) + (size_t)plane * ih * iw;
        int r_first, nrows, c_first, ncols;
        tile_span(y0, FT_H, H, ih, geo.rh[k], r_first, nrows);
        tile_span(x0, FT_W, W, iw, geo.rw[k], c_first, ncols);
        nrows = min(nrows, LR_RMAX);                   // cannot bind at ratio <= 1/4 (checked by the host); keeps indices in range regardless
        ncols = min(ncols, LR_CMAX);
        // ---- A. per-pixel gradients, weighted by their x taps and pre-reduced over the quad ----
        const float2* xq = &xt[k][tx];
        const int cb = __float_as_int(xq[0].x);
        int d0[4];
        float xw1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { d0[j] = __float_as_int(xq[j].x) - cb; xw1[j] = xq[j].y; }   // d0 in {0, 1}
        const int last = iw - 1;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int ty = warp + i * (LS_THREADS / 32);
            float af[3] = {0.0f, 0.0f, 0.0f}, ab[3] = {0.0f, 0.0f, 0.0f};
            if (ok[i]) {
                float x[4], xb[4];
                interp_quad2(fgp, bgp, ih, iw, yt[k][ty], xq, x, xb);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xv = x[j], qv = xb[j], m = mv[i][j], w = wv[i][j];
                    const float e = fast_ex2(-fabsf(xv) * LOG2E), inv = fast_rcp(1.0f + e);
                    const float s = xv >= 0.0f ? inv : e * inv;
                    const float e2 = fast_ex2(-fabsf(qv) * LOG2E), inv2 = fast_rcp(1.0f + e2);
                    const float s2 = qv >= 0.0f ? inv2 : e2 * inv2;
                    const float mw = m * w;
                    const float dwiou = -(mw * den - ip1 * (w - mw)) * inv_den2;       // as in structure_loss_bwd_kernel
                    const float gp = g * (w * (s - m) * invW + dwiou * s * (1.0f - s));
                    const float gq = g * 0.8f * w * (s2 - mbv[i][j]) * invW;
                    const float w1 = xw1[j], w0 = 1.0f - w1;
                    const int c0 = cb + d0[j];
                    const int e1 = d0[j] + (c0 < last ? 1 : 0);                        // column offset of the second tap (clamped at the right edge)
                    const float p0 = gp * w0, p1 = gp * w1, q0 = gq * w0, q1 = gq * w1;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        af[d] += (d0[j] == d ? p0 : 0.0f) + (e1 == d ? p1 : 0.0f);
                        ab[d] += (d0[j] == d ? q0 : 0.0f) + (e1 == d ? q1 : 0.0f);
                    }
                }
            }
            float* qf = &Q[0][ty][lane * 3];
            float* qb = &Q[1][ty][lane * 3];
            qf[0] = af[0]; qf[1] = af[1]; qf[2] = af[2];
            qb[0] = ab[0]; qb[1] = ab[1]; qb[2] = ab[2];
        }
        __syncthreads();
        // ---- B. fold along x: task = (map, low-res column, tile row); a warp shares (map, column), its lanes are the 32 rows ----
        for (int t = tid; t < 2 * ncols * FT_H; t += LS_THREADS) {
            const int row = t & (FT_H - 1), mc = t >> 5;
            const int map = mc >= ncols ? 1 : 0, cl = mc - map * ncols, c = c_first + cl;
            // lanes whose first source column is c-2, c-1 or c (monotone in the lane): conservative range from the inverse tap map
            const float inv_r = 1.0f / geo.rw[k];
            int l_lo = (int)floorf((((float)(c - 2) + 0.5f) * inv_r - 0.5f - (float)x0) * 0.25f) - 1;
            int l_hi = (int)ceilf((((float)(c + 1) + 0.5f) * inv_r - 0.5f - (float)x0) * 0.25f) + 1;
            l_lo = max(l_lo, 0);
            l_hi = min(l_hi, 31);
            if (c == 0) l_lo = 0;                      // clamped sources (src < 0) all land on column 0
            const float* qrow = &Q[map][row][0];
            float acc = 0.0f;
            for (int l = l_lo; l <= l_hi; ++l) {
                const int d = c - __float_as_int(xt[k][l * 4].x);
                if (d >= 0 && d <= 2) acc += qrow[l * 3 + d];
            }
            Hs[map][cl][row] = acc;
        }
        __syncthreads();
        // ---- C. fold along y: task = (map, low-res row, low-res column) -> this tile's slot ----
        float* slot = gpart + (((size_t)plane * tiles + tile) * NS + k) * 2 * LR_SLOT;
        for (int t = tid; t < 2 * nrows * ncols; t += LS_THREADS) {
            const int map = t >= nrows * ncols ? 1 : 0, rc = t - map * nrows * ncols;
            const int rl = rc / ncols, cl = rc - rl * ncols, r = r_first + rl;
            const float* col = &Hs[map][cl][0];
            float acc = 0.0f;
#pragma unroll 8
            for (int row = 0; row < FT_H; ++row) {
                const float2 yq = yt[k][row];
                const int i0 = __float_as_int(yq.x), i1 = min(i0 + 1, ih - 1);
                const float wy = (i0 == r ? 1.0f - yq.y : 0.0f) + (i1 == r ? yq.y : 0.0f);
                acc = fmaf(wy, col[row], acc);
            }
            slot[map * LR_SLOT + rl * LR_CMAX + cl] = acc;
        }
        // the next scale's phase A writes Q (last read before the barrier above) and its phase B writes Hs only after its own barrier
    }
}

// dlow[k][map][plane][r][c] = sum over the tiles whose span contains (r, c) of their slot entry, in tile order.
template <int NS>
__global__ void __launch_bounds__(256)
lowres_grad_fold_kernel(const __grid_constant__ PtrPack pp, const __grid_constant__ LowresGeo geo, const float* __restrict__ gpart, int H, int W, int tiles_x, int tiles) {
    pv2::pdl_prologue();
    const int k = blockIdx.z >> 1, map = blockIdx.z & 1, plane = blockIdx.y;
    const int ih = geo.ih[k], iw = geo.iw[k];
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= ih * iw) return;
    const int r = idx / iw, c = idx - r * iw;
    // output rows / columns whose taps can touch (r, c) -> the tile rows / columns to look at (conservative; containment is checked)
    const float irh = 1.0f / geo.rh[k], irw = 1.0f / geo.rw[k];
    int oy_lo = (int)floorf(((float)r - 0.5f) * irh - 0.5f) - 1, oy_hi = (int)ceilf(((float)r + 1.5f) * irh - 0.5f) + 1;
    int ox_lo = (int)floorf(((float)c - 0.5f) * irw - 0.5f) - 1, ox_hi = (int)ceilf(((float)c + 1.5f) * irw - 0.5f) + 1;
    if (r == 0) oy_lo = 0;
    if (c == 0) ox_lo = 0;
    if (r == ih - 1) oy_hi = H - 1;
    if (c == iw - 1) ox_hi = W - 1;
    const int tiles_y = tiles / tiles_x;
    const int ty_lo = max(oy_lo, 0) / FT_H, ty_hi = min(min(oy_hi, H - 1) / FT_H, tiles_y - 1);
    const int tx_lo = max(ox_lo, 0) / FT_W, tx_hi = min(min(ox_hi, W - 1) / FT_W, tiles_x - 1);
    float acc = 0.0f;
    for (int tyi = ty_lo; tyi <= ty_hi; ++tyi) {
        int r_first, nrows;
        tile_span(tyi * FT_H, FT_H, H, ih, geo.rh[k], r_first, nrows);
        nrows = min(nrows, LR_RMAX);
        const int rl = r - r_first;
        if (rl < 0 || rl >= nrows) continue;
        for (int txi = tx_lo; txi <= tx_hi; ++txi) {
            int c_first, ncols;
            tile_span(txi * FT_W, FT_W, W, iw, geo.rw[k], c_first, ncols);
            ncols = min(ncols, LR_CMAX);
            const int cl = c - c_first;
            if (cl < 0 || cl >= ncols) continue;
            const int tile = tyi * tiles_x + txi;
            acc += __ldcg(gpart + ((((size_t)plane * tiles + tile) * NS + k) * 2 + map) * LR_SLOT + rl * LR_CMAX + cl);
        }
    }
    float* dst = reinterpret_cast<float*>(map ? pp.dpred_bg[k] : pp.dpred[k]);
    dst[(size_t)plane * ih * iw + idx] = acc;
}

template <typename T, int VEC>
void launch_fwd(int ns, dim3 grid, int chunk_px, cudaStream_t st, const PtrPack& pp, const float* mf, const float* mb, const Layout& L, int HW, int planes, float* loss) {
#define PV2_FWD(NSV) pv2::launch_streaming(structure_loss_fwd_kernel<T, NSV, VEC>, grid, LS_THREADS, 0, st, pp, mf, mb, L.wmap, HW, planes, (int)grid.x, chunk_px, L.wt_tiles, \
                         L.partials, L.wsum_part, L.plane_sums, L.plane_loss, loss, L.ticket)
    switch (ns) { case 1: PV2_FWD(1); break; case 2: PV2_FWD(2); break; case 3: PV2_FWD(3); break; default: PV2_FWD(4); break; }
#undef PV2_FWD
}
template <typename T, int VEC>
void launch_bwd(int ns, dim3 grid, cudaStream_t st, const PtrPack& pp, const float* mf, const float* mb, const Layout& L, const float* gl, int HW, int planes) {
#define PV2_BWD(NSV) pv2::launch_streaming(structure_loss_bwd_kernel<T, NSV, VEC>, grid, LS_THREADS, 0, st, pp, mf, mb, L.wmap, gl, L.plane_sums, HW, planes)
    switch (ns) { case 1: PV2_BWD(1); break; case 2: PV2_BWD(2); break; case 3: PV2_BWD(3); break; default: PV2_BWD(4); break; }
#undef PV2_BWD
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// the 16-byte vector path needs every plane of every tensor to start 16-byte aligned
bool can_vec(const PtrPack& pp, int ns, const float* mf, const float* mb, int HW, bool grads) {
    if (HW % 8 != 0) return false;
    if (!aligned16(mf) || (mb && !aligned16(mb))) return false;
    for (int k = 0; k < ns; ++k) {
        if (!aligned16(pp.pred[k]) || !aligned16(pp.pred_bg[k])) return false;
        if (grads && (!aligned16(pp.dpred[k]) || !aligned16(pp.dpred_bg[k]))) return false;
    }
    return true;
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" size_t pv2_structure_loss_workspace_bytes(int planes, int H, int W, int nscales) {
    (void)nscales;
    return make_layout(nullptr, planes, H, W).bytes;
}

static int check_common(const void* const* pred, const void* const* pred_bg, const float* mask_fg, int nscales,
                        int planes, int H, int W, int logit_dtype, const void* ws, size_t ws_bytes) {
    PV2_CHECK(nscales >= 1 && nscales <= PV2_MAX_SCALES, "structure_loss: nscales=%d out of range [1,%d]", nscales, PV2_MAX_SCALES);
    PV2_CHECK(planes > 0 && H > 0 && W > 0, "structure_loss: empty input (planes=%d H=%d W=%d)", planes, H, W);
    PV2_CHECK(planes <= 65535, "structure_loss: planes=%d exceeds grid.y limit", planes);
    PV2_CHECK(logit_dtype == PV2_F32 || logit_dtype == PV2_BF16, "structure_loss: bad dtype %d", logit_dtype);
    PV2_CHECK(mask_fg != nullptr && pred != nullptr && pred_bg != nullptr, "structure_loss: null pointer");
    for (int k = 0; k < nscales; ++k) PV2_CHECK(pred[k] && pred_bg[k], "structure_loss: null logits pointer at scale %d", k);
    PV2_CHECK(ws != nullptr && ((uintptr_t)ws & 255u) == 0, "structure_loss: workspace must be non-null and 256-byte aligned");
    PV2_CHECK(ws_bytes >= pv2_structure_loss_workspace_bytes(planes, H, W, nscales),
              "structure_loss: workspace too small (%zu < %zu)", ws_bytes, pv2_structure_loss_workspace_bytes(planes, H, W, nscales));
    return 0;
}

// The boundary weight weit = 1 + 5*|avgpool31(mask) - mask| (MyTrain_med.py:21) depends on the mask only, and the mask is on the device
// before the backbone starts: a training step computes it on a side branch that forks at step start (it hides under the ~18 ms
// backbone) and the loss forward that follows the head is then a pure stream over logits + mask + 2-byte weight map.
extern "C" int pv2_structure_loss_prepare(const float* mask_fg, int planes, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
    PV2_CHECK(mask_fg != nullptr && planes > 0 && planes <= 65535 && H > 0 && W > 0, "structure_loss_prepare: bad arguments");
    PV2_CHECK(workspace != nullptr && ((uintptr_t)workspace & 255u) == 0 && workspace_bytes >= pv2_structure_loss_workspace_bytes(planes, H, W, 1),
              "structure_loss_prepare: workspace must be 256-byte aligned and pv2_structure_loss_workspace_bytes() large");
    const Layout L = make_layout(workspace, planes, H, W);
    pv2::launch(boundary_weight_kernel, dim3(L.wt_tiles, planes), WT_THREADS, 0, (cudaStream_t)stream, mask_fg, L.wmap, L.wsum_part, L.ticket, reinterpret_cast<double*>(L.partials), H, W, L.wt_tiles_x, L.wt_tiles);
    PV2_LAUNCH_CHECK("boundary_weight");
    return 0;
}

static int structure_loss_fwd_impl(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                   const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                   float* loss, void* workspace, size_t workspace_bytes, void* stream, bool prepared);

extern "C" int pv2_structure_loss_fwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                      const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                      float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    return structure_loss_fwd_impl(pred, pred_bg, mask_fg, mask_bg, nscales, planes, H, W, logit_dtype, loss, workspace, workspace_bytes, stream, false);
}

extern "C" int pv2_structure_loss_fwd_prepared(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                               const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                               float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    return structure_loss_fwd_impl(pred, pred_bg, mask_fg, mask_bg, nscales, planes, H, W, logit_dtype, loss, workspace, workspace_bytes, stream, true);
}

static int structure_loss_fwd_impl(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                   const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                   float* loss, void* workspace, size_t workspace_bytes, void* stream, bool prepared) {
    if (int e = check_common(pred, pred_bg, mask_fg, nscales, planes, H, W, logit_dtype, workspace, workspace_bytes)) return e;
    PV2_CHECK(loss != nullptr, "structure_loss_fwd: null loss pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout L = make_layout(workspace, planes, H, W);
    PtrPack pp = {};
    for (int k = 0; k < nscales; ++k) { pp.pred[k] = pred[k]; pp.pred_bg[k] = pred_bg[k]; }
    const int HW = H * W;
    const bool vec = can_vec(pp, nscales, mask_fg, mask_bg, HW, false);
    const bool no_fused = prepared || pv2::tune_int("PV2_LOSS_TWO_PASS", 0) == 1;      // boundary-weight kernel + streaming forward
    if (vec && W % 4 == 0 && !no_fused) {      // one pass: boundary weight + loss sums
        cudaError_t ce = cudaMemsetAsync(L.ticket, 0, sizeof(unsigned int), st);
        PV2_CHECK(ce == cudaSuccess, "structure_loss_fwd: memset: %s", cudaGetErrorString(ce));
        const dim3 fgrid(L.ft_tiles, planes);
        const LowresGeo geo = {};
#define PV2_FUSED(TT, NSV) pv2::launch(structure_loss_fwd_fused_kernel<TT, NSV, false>, fgrid, LS_THREADS, 0, st, pp, geo, mask_fg, mask_bg, L.wmap, H, W, planes, \
                                       L.ft_tiles_x, L.ft_tiles, L.partials, L.wsum_part, L.plane_sums, L.plane_loss, loss, L.ticket)
        if (logit_dtype == PV2_F32) {
            switch (nscales) { case 1: PV2_FUSED(float, 1); break; case 2: PV2_FUSED(float, 2); break; case 3: PV2_FUSED(float, 3); break; default: PV2_FUSED(float, 4); break; }
        } else {
            switch (nscales) { case 1: PV2_FUSED(__nv_bfloat16, 1); break; case 2: PV2_FUSED(__nv_bfloat16, 2); break; case 3: PV2_FUSED(__nv_bfloat16, 3); break; default: PV2_FUSED(__nv_bfloat16, 4); break; }
        }
#undef PV2_FUSED
        PV2_LAUNCH_CHECK("structure_loss_fwd_fused");
        return 0;
    }
    if (!prepared) {
        pv2::launch(boundary_weight_kernel, dim3(L.wt_tiles, planes), WT_THREADS, 0, st, mask_fg, L.wmap, L.wsum_part, L.ticket, reinterpret_cast<double*>(L.partials), H, W, L.wt_tiles_x, L.wt_tiles);
        PV2_LAUNCH_CHECK("boundary_weight");
    }
    // pixels per CTA: 2048 like the backward.  Measured at B = 16 x 352^2 (PV2_LOSS_FWD_CHUNK): 1024 / 2048 / 3072 / 4096 / 6144 / 8192 px
    // -> 24.0 / 23.1 / 24.4 / 24.2 / 26.3 / 26.1 us: neither the reduction tail of a CTA nor the 1.65 waves of 976 CTAs is what bounds it
    int chunk_px = pv2::tune_int("PV2_LOSS_FWD_CHUNK", 0);
    if (chunk_px <= 0) chunk_px = CHUNK;
    if (chunk_px < 1024) chunk_px = 1024;
    chunk_px = (chunk_px + 1023) / 1024 * 1024;
    const dim3 grid((HW + chunk_px - 1) / chunk_px, planes);
    if (logit_dtype == PV2_F32) {
        if (vec) launch_fwd<float, 4>(nscales, grid, chunk_px, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
        else launch_fwd<float, 1>(nscales, grid, chunk_px, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
    } else {
        if (vec) launch_fwd<__nv_bfloat16, 4>(nscales, grid, chunk_px, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
        else launch_fwd<__nv_bfloat16, 1>(nscales, grid, chunk_px, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
    }
    PV2_LAUNCH_CHECK("structure_loss_fwd");
    return 0;
}

extern "C" int pv2_structure_loss_bwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                      const float* mask_bg, const float* grad_loss, void* const* dpred,
                                      void* const* dpred_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                      const void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_common(pred, pred_bg, mask_fg, nscales, planes, H, W, logit_dtype, workspace, workspace_bytes)) return e;
    PV2_CHECK(grad_loss && dpred && dpred_bg, "structure_loss_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout L = make_layout(const_cast<void*>(workspace), planes, H, W);
    PtrPack pp = {};
    for (int k = 0; k < nscales; ++k) {
        PV2_CHECK(dpred[k] && dpred_bg[k], "structure_loss_bwd: null gradient pointer at scale %d", k);
        pp.pred[k] = pred[k]; pp.pred_bg[k] = pred_bg[k]; pp.dpred[k] = dpred[k]; pp.dpred_bg[k] = dpred_bg[k];
    }
    const int HW = H * W;
    const dim3 grid(L.chunks, planes);
    const bool vec = can_vec(pp, nscales, mask_fg, mask_bg, HW, true);
    if (logit_dtype == PV2_F32) {
        if (vec) launch_bwd<float, 4>(nscales, grid, st, pp, mask_fg, mask_bg, L, grad_loss, HW, planes);
        else launch_bwd<float, 1>(nscales, grid, st, pp, mask_fg, mask_bg, L, grad_loss, HW, planes);
    } else {
        if (vec) launch_bwd<__nv_bfloat16, 4>(nscales, grid, st, pp, mask_fg, mask_bg, L, grad_loss, HW, planes);
        else launch_bwd<__nv_bfloat16, 1>(nscales, grid, st, pp, mask_fg, mask_bg, L, grad_loss, HW, planes);
    }
    PV2_LAUNCH_CHECK("structure_loss_bwd");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// loss from the low-resolution maps: pranet.py:349-415 final upsamples + MyTrain_med.py:78-82 in one pass
// ---------------------------------------------------------------------------------------------------------
namespace pv2 {
namespace {

size_t lowres_gpart_floats(int planes, int H, int W, int nscales) {
    const int tiles = ((W + FT_W - 1) / FT_W) * ((H + FT_H - 1) / FT_H);
    return (size_t)planes * tiles * nscales * 2 * LR_SLOT;
}

int check_lowres(const float* const* low_fg, const float* const* low_bg, const int* ih, const int* iw, const float* rh, const float* rw,
                 const float* mask_fg, const float* mask_bg, int nscales, int planes, int H, int W, const void* ws, size_t ws_bytes,
                 PtrPack& pp, LowresGeo& geo) {
    PV2_CHECK(nscales >= 1 && nscales <= PV2_MAX_SCALES, "structure_loss_lowres: nscales=%d out of range [1,%d]", nscales, PV2_MAX_SCALES);
    PV2_CHECK(planes > 0 && planes <= 65535 && H > 0 && W > 0, "structure_loss_lowres: bad shape (planes=%d H=%d W=%d)", planes, H, W);
    PV2_CHECK(low_fg && low_bg && ih && iw && rh && rw && mask_fg, "structure_loss_lowres: null pointer");
    PV2_CHECK(W % 4 == 0 && aligned16(mask_fg) && (!mask_bg || aligned16(mask_bg)),
              "structure_loss_lowres: W must be a multiple of 4 and the masks 16-byte aligned (W=%d); upsample and call pv2_structure_loss_fwd instead", W);
    for (int k = 0; k < nscales; ++k) {
        PV2_CHECK(low_fg[k] && low_bg[k], "structure_loss_lowres: null map pointer at scale %d", k);
        PV2_CHECK(ih[k] > 0 && iw[k] > 0 && rh[k] > 0.0f && rw[k] > 0.0f && rh[k] <= 0.25f && rw[k] <= 0.25f,
                  "structure_loss_lowres: scale %d (%dx%d, ratios %g %g): the fused path covers up-scaling by >= 4; upsample and call pv2_structure_loss_fwd instead",
                  k, ih[k], iw[k], (double)rh[k], (double)rw[k]);
        pp.pred[k] = low_fg[k]; pp.pred_bg[k] = low_bg[k];
        geo.ih[k] = ih[k]; geo.iw[k] = iw[k]; geo.rh[k] = rh[k]; geo.rw[k] = rw[k];
    }
    PV2_CHECK(ws != nullptr && ((uintptr_t)ws & 255u) == 0, "structure_loss_lowres: workspace must be non-null and 256-byte aligned");
    PV2_CHECK(ws_bytes >= pv2_structure_loss_lowres_workspace_bytes(planes, H, W, nscales), "structure_loss_lowres: workspace too small (%zu < %zu)",
              ws_bytes, pv2_structure_loss_lowres_workspace_bytes(planes, H, W, nscales));
    return 0;
}

}  // namespace
}  // namespace pv2

extern "C" size_t pv2_structure_loss_lowres_workspace_bytes(int planes, int H, int W, int nscales) {
    return make_layout(nullptr, planes, H, W).bytes + sizeof(float) * lowres_gpart_floats(planes, H, W, nscales);
}

extern "C" int pv2_structure_loss_lowres_fwd(const float* const* low_fg, const float* const* low_bg, const int* ih, const int* iw,
                                             const float* rh, const float* rw, const float* mask_fg, const float* mask_bg,
                                             int nscales, int planes, int H, int W, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    PtrPack pp = {};
    LowresGeo geo = {};
    if (int e = check_lowres(low_fg, low_bg, ih, iw, rh, rw, mask_fg, mask_bg, nscales, planes, H, W, workspace, workspace_bytes, pp, geo)) return e;
    PV2_CHECK(loss != nullptr, "structure_loss_lowres_fwd: null loss pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout L = make_layout(workspace, planes, H, W);
    cudaError_t ce = cudaMemsetAsync(L.ticket, 0, sizeof(unsigned int), st);
    PV2_CHECK(ce == cudaSuccess, "structure_loss_lowres_fwd: memset: %s", cudaGetErrorString(ce));
    const dim3 fgrid(L.ft_tiles, planes);
#define PV2_LOWRES(NSV) pv2::launch(structure_loss_fwd_fused_kernel<float, NSV, true>, fgrid, LS_THREADS, 0, st, pp, geo, mask_fg, mask_bg, L.wmap, H, W, planes, \
                                    L.ft_tiles_x, L.ft_tiles, L.partials, L.wsum_part, L.plane_sums, L.plane_loss, loss, L.ticket)
    switch (nscales) { case 1: PV2_LOWRES(1); break; case 2: PV2_LOWRES(2); break; case 3: PV2_LOWRES(3); break; default: PV2_LOWRES(4); break; }
#undef PV2_LOWRES
    PV2_LAUNCH_CHECK("structure_loss_lowres_fwd");
    return 0;
}

extern "C" int pv2_structure_loss_lowres_bwd(const float* const* low_fg, const float* const* low_bg, const int* ih, const int* iw,
                                             const float* rh, const float* rw, const float* mask_fg, const float* mask_bg,
                                             const float* grad_loss, float* const* dlow_fg, float* const* dlow_bg,
                                             int nscales, int planes, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
    PtrPack pp = {};
    LowresGeo geo = {};
    if (int e = check_lowres(low_fg, low_bg, ih, iw, rh, rw, mask_fg, mask_bg, nscales, planes, H, W, workspace, workspace_bytes, pp, geo)) return e;
    PV2_CHECK(grad_loss && dlow_fg && dlow_bg, "structure_loss_lowres_bwd: null pointer");
    int max_px = 0;
    for (int k = 0; k < nscales; ++k) {
        PV2_CHECK(dlow_fg[k] && dlow_bg[k], "structure_loss_lowres_bwd: null gradient pointer at scale %d", k);
        pp.dpred[k] = dlow_fg[k]; pp.dpred_bg[k] = dlow_bg[k];
        if (ih[k] * iw[k] > max_px) max_px = ih[k] * iw[k];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const Layout L = make_layout(workspace, planes, H, W);
    float* gpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + L.bytes);
    const dim3 grid(L.ft_tiles, planes), fold_grid((max_px + 255) / 256, planes, 2 * nscales);
#define PV2_LOWRES_BWD(NSV)                                                                                                              \
    pv2::launch(structure_loss_lowres_bwd_kernel<NSV>, grid, LS_THREADS, 0, st, pp, geo, mask_fg, mask_bg, L.wmap, grad_loss, L.plane_sums, gpart, \
                H, W, planes, L.ft_tiles_x, L.ft_tiles);                                                                                \
    PV2_LAUNCH_CHECK("structure_loss_lowres_bwd");                                                                                      \
    pv2::launch(lowres_grad_fold_kernel<NSV>, fold_grid, 256, 0, st, pp, geo, gpart, H, W, L.ft_tiles_x, L.ft_tiles);                   \
    PV2_LAUNCH_CHECK("lowres_grad_fold")
    switch (nscales) { case 1: PV2_LOWRES_BWD(1); break; case 2: PV2_LOWRES_BWD(2); break; case 3: PV2_LOWRES_BWD(3); break; default: PV2_LOWRES_BWD(4); break; }
#undef PV2_LOWRES_BWD
    return 0;
}
