// structure_loss forward / backward (binary_seg/MyTrain_med.py:19-38), fused.
//
// One launch evaluates up to 4 (pred, pred_bg) pairs against one mask.  Per CTA: a TH x TW tile of
// one (n,c) plane.  The mask tile plus its 15-px halo is staged in shared memory once, the 31x31
// box filter runs as two separable running-sum passes out of shared memory (zero padding counted in
// the /961 divisor, like avg_pool2d's default), and the boundary weight `weit` never leaves
// registers.  The five weighted sums per (plane, scale) are reduced warp-shuffle -> shared -> one
// partial per CTA (no atomics: the finalize kernel adds the per-tile partials in a fixed order, so
// the loss is bit-reproducible run to run).
//
// HBM traffic (algorithmic, fp32): fwd 4 B (mask) + 8 B per scale per pixel; bwd the same reads plus
// 8 B of gradients per scale per pixel.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int TH = 32, TW = 64, HALO = 15, KS = 31;
constexpr int SH = TH + 2 * HALO;   // 62 staged rows
constexpr int SW = TW + 2 * HALO;   // 94 staged cols
constexpr int SPITCH = SW + 1;      // 95: odd pitch -> lanes walking down rows hit distinct banks
constexpr int THREADS = 256;
constexpr int ROWS_PER_THREAD = TH / (THREADS / TW);  // 8
constexpr int NSUM_MAX = 1 + 4 * PV2_MAX_SCALES;      // S_w + (bce, bce2, inter, union) per scale
constexpr float INV_AREA = 1.0f / 961.0f;

struct PtrPack {
    const void* pred[PV2_MAX_SCALES];
    const void* pred_bg[PV2_MAX_SCALES];
    void* dpred[PV2_MAX_SCALES];
    void* dpred_bg[PV2_MAX_SCALES];
};

// Stage mask tile + halo, run the separable box filter; on return each thread holds, for its column
// x = tid % TW and its ROWS_PER_THREAD consecutive rows, the mask value m[] and weit w[].
__device__ __forceinline__ void tile_weit(const float* __restrict__ mask, int H, int W, int y0, int x0,
                                          float* sm, float* hs, float (&m)[ROWS_PER_THREAD],
                                          float (&w)[ROWS_PER_THREAD]) {
    const int tid = threadIdx.x;
    for (int i = tid; i < SH * SW; i += THREADS) {
        int r = i / SW, c = i - r * SW;
        int gy = y0 + r - HALO, gx = x0 + c - HALO;
        float v = 0.0f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(mask + (size_t)gy * W + gx);
        sm[r * SPITCH + c] = v;
    }
    __syncthreads();
    // horizontal running sums: item = (segment of 8 outputs, row); row varies fastest across lanes
    constexpr int SEG = 8, NSEG = TW / SEG;
    for (int it = tid; it < SH * NSEG; it += THREADS) {
        int r = it % SH, s = it / SH;
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
    const int x = tid % TW, rb = (tid / TW) * ROWS_PER_THREAD;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < KS; ++j) acc += hs[(rb + j) * TW + x];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        if (j > 0) acc += hs[(rb + j + KS - 1) * TW + x] - hs[(rb + j - 1) * TW + x];
        float mv = sm[(rb + j + HALO) * SPITCH + x + HALO];
        m[j] = mv;
        w[j] = 1.0f + 5.0f * fabsf(acc * INV_AREA - mv);
    }
}

// MUFU-only transcendental pieces: e = exp(-|x|) via ex2.approx, log1p(e) via lg2.approx(1+e).
// Absolute error of log1p(e) <= ~6e-8 (when e underflows against 1), far inside the 1e-4 loss tolerance.
__device__ __forceinline__ float fast_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// softplus(x) - x*t  and sigmoid(x)
__device__ __forceinline__ void bce_sig(float x, float t, float& bce, float& sig) {
    const float e = fast_ex2(-fabsf(x) * LOG2E);
    const float d = 1.0f + e;
    const float inv = fast_rcp(d);
    sig = x >= 0.0f ? inv : e * inv;
    bce = fmaf(-x, t, fmaxf(x, 0.0f)) + fast_lg2(d) * LN2;
}
__device__ __forceinline__ float bce_only(float x, float t) {
    const float e = fast_ex2(-fabsf(x) * LOG2E);
    return fmaf(-x, t, fmaxf(x, 0.0f)) + fast_lg2(1.0f + e) * LN2;
}
__device__ __forceinline__ float sigmoid_fast(float x) {
    const float e = fast_ex2(-fabsf(x) * LOG2E);
    const float inv = fast_rcp(1.0f + e);
    return x >= 0.0f ? inv : e * inv;
}

template <typename T>
__global__ void __launch_bounds__(THREADS, 3)
structure_loss_fwd_kernel(PtrPack pp, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                          int nscales, int H, int W, int tiles_x, int tiles_per_plane, float* __restrict__ partials) {
    __shared__ float sm[SH * SPITCH];
    __shared__ float hs[SH * TW];
    __shared__ float red[THREADS / 32][NSUM_MAX];

    const int plane = blockIdx.y, tile = blockIdx.x;
    const int y0 = (tile / tiles_x) * TH, x0 = (tile % tiles_x) * TW;
    const size_t poff = (size_t)plane * H * W;
    const int tid = threadIdx.x, x = tid % TW, rb = (tid / TW) * ROWS_PER_THREAD;
    const int gx = x0 + x;

    float m[ROWS_PER_THREAD], w[ROWS_PER_THREAD];
    tile_weit(mask_fg + poff, H, W, y0, x0, sm, hs, m, w);

    float sums[NSUM_MAX];
#pragma unroll
    for (int i = 0; i < NSUM_MAX; ++i) sums[i] = 0.0f;
    bool ok[ROWS_PER_THREAD];
    float mb[ROWS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int gy = y0 + rb + j;
        ok[j] = (gx < W) && (gy < H);
        if (!ok[j]) w[j] = 0.0f;
        sums[0] += w[j];
        mb[j] = 1.0f - m[j];
        if (mask_bg != nullptr && ok[j]) mb[j] = __ldg(mask_bg + poff + (size_t)gy * W + gx);
    }
#pragma unroll
    for (int k = 0; k < PV2_MAX_SCALES; ++k) {
        if (k >= nscales) break;
        const T* p = reinterpret_cast<const T*>(pp.pred[k]) + poff;
        const T* q = reinterpret_cast<const T*>(pp.pred_bg[k]) + poff;
        float pv[ROWS_PER_THREAD], qv[ROWS_PER_THREAD];
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            size_t o = (size_t)(y0 + rb + j) * W + gx;
            pv[j] = ok[j] ? to_f(p[o]) : 0.0f;
            qv[j] = ok[j] ? to_f(q[o]) : 0.0f;
        }
        float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            float bce, sig;
            bce_sig(pv[j], m[j], bce, sig);
            const float bce2 = bce_only(qv[j], mb[j]);
            a += w[j] * bce;
            b += w[j] * bce2;
            c += sig * m[j] * w[j];
            d += (sig + m[j]) * w[j];
        }
        sums[1 + 4 * k + 0] = a;
        sums[1 + 4 * k + 1] = b;
        sums[1 + 4 * k + 2] = c;
        sums[1 + 4 * k + 3] = d;
    }
    const int nsum = 1 + 4 * nscales;
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < NSUM_MAX; ++i) {
        if (i < nsum) {
            float v = warp_sum(sums[i]);
            if (lane == 0) red[warp][i] = v;
        }
    }
    __syncthreads();
    if (tid < nsum) {
        float v = 0.0f;
#pragma unroll
        for (int wi = 0; wi < THREADS / 32; ++wi) v += red[wi][tid];
        partials[((size_t)plane * tiles_per_plane + tile) * NSUM_MAX + tid] = v;
    }
}

// grid = planes; one CTA folds the per-tile partials of its plane in a FIXED order (thread t takes tiles
// t, t+128, ...; then a fixed shuffle/shared tree), writes plane_sums[plane][*] and this plane's loss terms.
constexpr int FIN_THREADS = 128;
__global__ void __launch_bounds__(FIN_THREADS)
structure_loss_plane_kernel(const float* __restrict__ partials, float* __restrict__ plane_sums,
                            float* __restrict__ plane_loss, int tiles_per_plane, int nscales) {
    __shared__ float red[FIN_THREADS / 32][NSUM_MAX];
    const int p = blockIdx.x, nsum = 1 + 4 * nscales;
    float acc[NSUM_MAX];
#pragma unroll
    for (int i = 0; i < NSUM_MAX; ++i) acc[i] = 0.0f;
    for (int t = threadIdx.x; t < tiles_per_plane; t += FIN_THREADS) {
        const float* src = partials + ((size_t)p * tiles_per_plane + t) * NSUM_MAX;
#pragma unroll
        for (int i = 0; i < NSUM_MAX; ++i)
            if (i < nsum) acc[i] += src[i];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NSUM_MAX; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = 0.0f;
        if (lane < nsum) {
#pragma unroll
            for (int wi = 0; wi < FIN_THREADS / 32; ++wi) v += red[wi][lane];
            plane_sums[(size_t)p * NSUM_MAX + lane] = v;
        }
        const float Wsum = __shfl_sync(0xffffffffu, v, 0);
#pragma unroll
        for (int k = 0; k < PV2_MAX_SCALES; ++k) {
            const float sb = __shfl_sync(0xffffffffu, v, 1 + 4 * k + 0);
            const float sb2 = __shfl_sync(0xffffffffu, v, 1 + 4 * k + 1);
            const float inter = __shfl_sync(0xffffffffu, v, 1 + 4 * k + 2);
            const float uni = __shfl_sync(0xffffffffu, v, 1 + 4 * k + 3);
            if (lane == 0 && k < nscales)
                plane_loss[(size_t)p * PV2_MAX_SCALES + k] = sb / Wsum + 1.0f - (inter + 1.0f) / (uni - inter + 1.0f) + 0.8f * sb2 / Wsum;
        }
    }
}

// one small CTA: loss[k] = mean over planes (fixed order)
__global__ void structure_loss_mean_kernel(const float* __restrict__ plane_loss, float* __restrict__ loss, int planes, int nscales) {
    __shared__ float red[8][PV2_MAX_SCALES];
    float acc[PV2_MAX_SCALES] = {0.f, 0.f, 0.f, 0.f};
    for (int p = threadIdx.x; p < planes; p += blockDim.x)
#pragma unroll
        for (int k = 0; k < PV2_MAX_SCALES; ++k)
            if (k < nscales) acc[k] += plane_loss[(size_t)p * PV2_MAX_SCALES + k];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < PV2_MAX_SCALES; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < nscales) {
        float v = 0.0f;
        for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) v += red[wi][threadIdx.x];
        loss[threadIdx.x] = v / (float)planes;
    }
}

template <typename T>
__global__ void __launch_bounds__(THREADS, 3)
structure_loss_bwd_kernel(PtrPack pp, const float* __restrict__ mask_fg, const float* __restrict__ mask_bg,
                          const float* __restrict__ grad_loss, const float* __restrict__ plane_sums,
                          int nscales, int planes, int H, int W, int tiles_x) {
    __shared__ float sm[SH * SPITCH];
    __shared__ float hs[SH * TW];
    const int plane = blockIdx.y, tile = blockIdx.x;
    const int y0 = (tile / tiles_x) * TH, x0 = (tile % tiles_x) * TW;
    const size_t poff = (size_t)plane * H * W;
    const int tid = threadIdx.x, x = tid % TW, rb = (tid / TW) * ROWS_PER_THREAD;
    const int gx = x0 + x;

    float m[ROWS_PER_THREAD], w[ROWS_PER_THREAD];
    tile_weit(mask_fg + poff, H, W, y0, x0, sm, hs, m, w);
    bool ok[ROWS_PER_THREAD];
    float mb[ROWS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        int gy = y0 + rb + j;
        ok[j] = (gx < W) && (gy < H);
        mb[j] = 1.0f - m[j];
        if (mask_bg != nullptr && ok[j]) mb[j] = __ldg(mask_bg + poff + (size_t)gy * W + gx);
    }
    const float* ps = plane_sums + (size_t)plane * NSUM_MAX;
    const float invW = 1.0f / ps[0];
    const float invn = 1.0f / (float)planes;
#pragma unroll
    for (int k = 0; k < PV2_MAX_SCALES; ++k) {
        if (k >= nscales) break;
        const T* p = reinterpret_cast<const T*>(pp.pred[k]) + poff;
        const T* q = reinterpret_cast<const T*>(pp.pred_bg[k]) + poff;
        T* dp = reinterpret_cast<T*>(pp.dpred[k]) + poff;
        T* dq = reinterpret_cast<T*>(pp.dpred_bg[k]) + poff;
        const float g = grad_loss[k] * invn;
        const float inter = ps[1 + 4 * k + 2], uni = ps[1 + 4 * k + 3];
        const float den = uni - inter + 1.0f, inv_den2 = 1.0f / (den * den), ip1 = inter + 1.0f;
        float pv[ROWS_PER_THREAD], qv[ROWS_PER_THREAD];
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            size_t o = (size_t)(y0 + rb + j) * W + gx;
            pv[j] = ok[j] ? to_f(p[o]) : 0.0f;
            qv[j] = ok[j] ? to_f(q[o]) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            if (!ok[j]) continue;
            const float s = sigmoid_fast(pv[j]), s2 = sigmoid_fast(qv[j]);
            float mw = m[j] * w[j];
            // d wiou / d sigma = -[ m w den - (inter+1)(w - m w) ] / den^2
            float dwiou = -(mw * den - ip1 * (w[j] - mw)) * inv_den2;
            float gp = g * (w[j] * (s - m[j]) * invW + dwiou * s * (1.0f - s));
            float gq = g * 0.8f * w[j] * (s2 - mb[j]) * invW;
            size_t o = (size_t)(y0 + rb + j) * W + gx;
            dp[o] = from_f<T>(gp);
            dq[o] = from_f<T>(gq);
        }
    }
}

inline int tiles_of(int H, int W, int* tx) {
    *tx = (W + TW - 1) / TW;
    return *tx * ((H + TH - 1) / TH);
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" size_t pv2_structure_loss_workspace_bytes(int planes, int H, int W, int nscales) {
    (void)nscales;
    int tx;
    int tiles = tiles_of(H, W, &tx);
    return sizeof(float) * ((size_t)NSUM_MAX * ((size_t)planes + (size_t)planes * tiles) + (size_t)PV2_MAX_SCALES * planes);
}

static int check_common(const void* const* pred, const void* const* pred_bg, const float* mask_fg, int nscales,
                        int planes, int H, int W, int logit_dtype, const void* ws, size_t ws_bytes) {
    PV2_CHECK(nscales >= 1 && nscales <= PV2_MAX_SCALES, "structure_loss: nscales=%d out of range [1,%d]", nscales, PV2_MAX_SCALES);
    PV2_CHECK(planes > 0 && H > 0 && W > 0, "structure_loss: empty input (planes=%d H=%d W=%d)", planes, H, W);
    PV2_CHECK(planes <= 65535, "structure_loss: planes=%d exceeds grid.y limit", planes);
    PV2_CHECK(logit_dtype == PV2_F32 || logit_dtype == PV2_BF16, "structure_loss: bad dtype %d", logit_dtype);
    PV2_CHECK(mask_fg != nullptr && pred != nullptr && pred_bg != nullptr, "structure_loss: null pointer");
    for (int k = 0; k < nscales; ++k) PV2_CHECK(pred[k] && pred_bg[k], "structure_loss: null logits pointer at scale %d", k);
    PV2_CHECK(ws != nullptr && ws_bytes >= pv2_structure_loss_workspace_bytes(planes, H, W, nscales),
              "structure_loss: workspace too small (%zu < %zu)", ws_bytes, pv2_structure_loss_workspace_bytes(planes, H, W, nscales));
    return 0;
}

extern "C" int pv2_structure_loss_fwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                      const float* mask_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                      float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_common(pred, pred_bg, mask_fg, nscales, planes, H, W, logit_dtype, workspace, workspace_bytes)) return e;
    PV2_CHECK(loss != nullptr, "structure_loss_fwd: null loss pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int tx;
    int tiles = tiles_of(H, W, &tx);
    PtrPack pp = {};
    for (int k = 0; k < nscales; ++k) { pp.pred[k] = pred[k]; pp.pred_bg[k] = pred_bg[k]; }
    float* plane_sums = (float*)workspace;
    float* partials = plane_sums + (size_t)planes * NSUM_MAX;
    float* plane_loss = partials + (size_t)planes * tiles * NSUM_MAX;
    dim3 grid(tiles, planes);
    if (logit_dtype == PV2_F32)
        structure_loss_fwd_kernel<float><<<grid, THREADS, 0, st>>>(pp, mask_fg, mask_bg, nscales, H, W, tx, tiles, partials);
    else
        structure_loss_fwd_kernel<__nv_bfloat16><<<grid, THREADS, 0, st>>>(pp, mask_fg, mask_bg, nscales, H, W, tx, tiles, partials);
    PV2_LAUNCH_CHECK("structure_loss_fwd");
    structure_loss_plane_kernel<<<planes, FIN_THREADS, 0, st>>>(partials, plane_sums, plane_loss, tiles, nscales);
    PV2_LAUNCH_CHECK("structure_loss_plane");
    structure_loss_mean_kernel<<<1, 256, 0, st>>>(plane_loss, loss, planes, nscales);
    PV2_LAUNCH_CHECK("structure_loss_mean");
    return 0;
}

extern "C" int pv2_structure_loss_bwd(const void* const* pred, const void* const* pred_bg, const float* mask_fg,
                                      const float* mask_bg, const float* grad_loss, void* const* dpred,
                                      void* const* dpred_bg, int nscales, int planes, int H, int W, int logit_dtype,
                                      const void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_common(pred, pred_bg, mask_fg, nscales, planes, H, W, logit_dtype, workspace, workspace_bytes)) return e;
    PV2_CHECK(grad_loss && dpred && dpred_bg, "structure_loss_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int tx;
    int tiles = tiles_of(H, W, &tx);
    PtrPack pp = {};
    for (int k = 0; k < nscales; ++k) {
        PV2_CHECK(dpred[k] && dpred_bg[k], "structure_loss_bwd: null gradient pointer at scale %d", k);
        pp.pred[k] = pred[k]; pp.pred_bg[k] = pred_bg[k]; pp.dpred[k] = dpred[k]; pp.dpred_bg[k] = dpred_bg[k];
    }
    const float* plane_sums = (const float*)workspace;
    dim3 grid(tiles, planes);
    if (logit_dtype == PV2_F32)
        structure_loss_bwd_kernel<float><<<grid, THREADS, 0, st>>>(pp, mask_fg, mask_bg, grad_loss, plane_sums, nscales, planes, H, W, tx);
    else
        structure_loss_bwd_kernel<__nv_bfloat16><<<grid, THREADS, 0, st>>>(pp, mask_fg, mask_bg, grad_loss, plane_sums, nscales, planes, H, W, tx);
    PV2_LAUNCH_CHECK("structure_loss_bwd");
    return 0;
}
