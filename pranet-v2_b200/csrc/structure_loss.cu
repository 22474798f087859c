// structure_loss forward / backward (binary_seg/MyTrain_med.py:19-38).
//
//   weit = 1 + 5*|avg_pool31x31(mask) - mask|      (zero padding counted in the /961 divisor)
//   loss = mean_{n,c}[ wbce(pred, mask) + wiou(pred, mask) + 0.8*wbce(pred_bg, mask_bg) ]
//
// Three kernels, all on the caller's stream:
//   1. boundary_weight_kernel  -- per 32x64 tile: mask + 15-px halo staged in shared memory, separable running-sum
//      box filter out of shared memory, |avg - m| written ONCE as a 16-bit fixed-point map (2 B/px; weit is in
//      [1,6], quantisation 7.6e-5) plus the per-tile sum of weit.  The reference recomputes the 961-tap pool in each
//      of its 4 loss calls and again in autograd; here it is computed once per step and shared by all scales,
//      forward and backward.
//   2. structure_loss_fwd_kernel<T, NS, VEC> -- pure streaming: each CTA owns a 2048-px chunk of one (n,c) plane,
//      16-byte loads of logits / mask / weight map, MUFU-only transcendental math, the 4 weighted sums per scale
//      reduced warp-shuffle -> shared -> one partial per CTA (no atomics on data).  The last CTA to finish (ticket
//      counter) folds the partials in a fixed order into per-plane sums and the scalar losses: bit-reproducible.
//   3. structure_loss_bwd_kernel<T, NS, VEC> -- same streaming shape, reads the finished plane sums, writes both
//      gradients with 16-byte stores.
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
                       unsigned int* __restrict__ ticket, int H, int W, int tiles_x, int tiles_per_plane) {
    pv2::pdl_prologue();
    __shared__ float sm[SH * SPITCH];
    __shared__ float hs[SH * TW];
    __shared__ float red[WT_THREADS / 32];
    const int plane = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (plane == 0 && tile == 0 && tid == 0) *ticket = 0u;   // the forward kernel that follows counts on this
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
                          const uint16_t* __restrict__ wmap, int HW, int planes, int chunks, int wt_tiles,
                          float* __restrict__ partials, const float* __restrict__ wsum_part, float* __restrict__ plane_sums,
                          float* __restrict__ plane_loss, float* __restrict__ loss, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    __shared__ float red[LS_THREADS / 32][4 * NS];
    __shared__ bool is_last;
    const int plane = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const size_t pbase = (size_t)plane * HW;
    const int p0 = chunk * CHUNK, p1 = min(HW, p0 + CHUNK);
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
    if (tid < 4 * NS) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < LS_THREADS / 32; ++wi) v += red[wi][tid];
        partials[((size_t)plane * chunks + chunk) * (4 * PV2_MAX_SCALES) + tid] = v;
    }
    // ---- the last CTA to finish folds everything in a fixed order ----
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
        float s[4 * NS];
#pragma unroll
        for (int i = 0; i < 4 * NS; ++i) s[i] = 0.0f;
        for (int c = lane; c < chunks; c += 32) {
            const float* src = partials + ((size_t)pl * chunks + c) * (4 * PV2_MAX_SCALES);
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

template <typename T, int NS>
__global__ void __launch_bounds__(LS_THREADS, 4)
structure_loss_fwd_fused_kernel(PtrPack pp, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                                uint16_t* __restrict__ wmap, int H, int W, int planes, int tiles_x, int tiles,
                                float* __restrict__ partials, float* __restrict__ wsum_part, float* __restrict__ plane_sums,
                                float* __restrict__ plane_loss, float* __restrict__ loss, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    __shared__ float sat[FS_H * FS_PITCH];
    __shared__ float red[LS_THREADS / 32][4 * NS + 1];
    __shared__ bool is_last;
    const int plane = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int y0 = (tile / tiles_x) * FT_H, x0 = (tile % tiles_x) * FT_W;
    const int HW = H * W;
    const size_t pbase = (size_t)plane * HW;
    const float* mp = mask_fg + pbase;
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
            x[k].load(reinterpret_cast<const T*>(pp.pred[k]) + p);
            xb[k].load(reinterpret_cast<const T*>(pp.pred_bg[k]) + p);
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

template <typename T, int VEC>
void launch_fwd(int ns, dim3 grid, cudaStream_t st, const PtrPack& pp, const float* mf, const float* mb, const Layout& L, int HW, int planes, float* loss) {
#define PV2_FWD(NSV) pv2::launch(structure_loss_fwd_kernel<T, NSV, VEC>, grid, LS_THREADS, 0, st, pp, mf, mb, L.wmap, HW, planes, L.chunks, L.wt_tiles, \
                         L.partials, L.wsum_part, L.plane_sums, L.plane_loss, loss, L.ticket)
    switch (ns) { case 1: PV2_FWD(1); break; case 2: PV2_FWD(2); break; case 3: PV2_FWD(3); break; default: PV2_FWD(4); break; }
#undef PV2_FWD
}
template <typename T, int VEC>
void launch_bwd(int ns, dim3 grid, cudaStream_t st, const PtrPack& pp, const float* mf, const float* mb, const Layout& L, const float* gl, int HW, int planes) {
#define PV2_BWD(NSV) pv2::launch(structure_loss_bwd_kernel<T, NSV, VEC>, grid, LS_THREADS, 0, st, pp, mf, mb, L.wmap, gl, L.plane_sums, HW, planes)
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

extern "C" int pv2_structure_loss_fwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                      const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                      float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_common(pred, pred_bg, mask_fg, nscales, planes, H, W, logit_dtype, workspace, workspace_bytes)) return e;
    PV2_CHECK(loss != nullptr, "structure_loss_fwd: null loss pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout L = make_layout(workspace, planes, H, W);
    PtrPack pp = {};
    for (int k = 0; k < nscales; ++k) { pp.pred[k] = pred[k]; pp.pred_bg[k] = pred_bg[k]; }
    const int HW = H * W;
    const bool vec = can_vec(pp, nscales, mask_fg, mask_bg, HW, false);
    static const bool no_fused = [] { const char* e = getenv("PV2_LOSS_TWO_PASS"); return e && e[0] == '1'; }();
    if (vec && W % 4 == 0 && !no_fused) {      // one pass: boundary weight + loss sums
        cudaError_t ce = cudaMemsetAsync(L.ticket, 0, sizeof(unsigned int), st);
        PV2_CHECK(ce == cudaSuccess, "structure_loss_fwd: memset: %s", cudaGetErrorString(ce));
        const dim3 fgrid(L.ft_tiles, planes);
#define PV2_FUSED(TT, NSV) pv2::launch(structure_loss_fwd_fused_kernel<TT, NSV>, fgrid, LS_THREADS, 0, st, pp, mask_fg, mask_bg, L.wmap, H, W, planes, \
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
    pv2::launch(boundary_weight_kernel, dim3(L.wt_tiles, planes), WT_THREADS, 0, st, mask_fg, L.wmap, L.wsum_part, L.ticket, H, W, L.wt_tiles_x, L.wt_tiles);
    PV2_LAUNCH_CHECK("boundary_weight");
    const dim3 grid(L.chunks, planes);
    if (logit_dtype == PV2_F32) {
        if (vec) launch_fwd<float, 4>(nscales, grid, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
        else launch_fwd<float, 1>(nscales, grid, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
    } else {
        if (vec) launch_fwd<__nv_bfloat16, 4>(nscales, grid, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
        else launch_fwd<__nv_bfloat16, 1>(nscales, grid, st, pp, mask_fg, mask_bg, L, HW, planes, loss);
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
