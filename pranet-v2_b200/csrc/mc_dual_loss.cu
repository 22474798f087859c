// Multiclass dual-supervision loss (EMCAD/trainer.py:123-140 == MERIT/train_ACDC.py:259-284 == MIST/trainer.py:112-129):
//
//   loss = sum over the non-empty subsets s of the n (= 4) scales of
//          lc_ce * CE(sum_{i in s} P_fg[i], y) + lc_dice * Dice(softmax(sum P_fg[i]), onehot(y)) + lc_bce * mean BCEWithLogits(sum P_bg[i], 1 - onehot(y))
//   Dice = mean_c [ 1 - (2*sum p_c t_c + 1e-5) / (sum p_c^2 + sum t_c^2 + 1e-5) ]   over the WHOLE batch (EMCAD/utils/utils.py:102-138)
//
// The reference evaluates this as 15 x (CrossEntropyLoss + softmax + a 9-iteration python Dice loop with .item() syncs +
// BCEWithLogitsLoss) on tensors it first sums in separate kernels, and builds the inverted one-hot mask on the CPU
// (trainer.py:22-29, 99-103).  Here: ONE forward launch (+ a tiny fold) and ONE backward launch.
//   * one thread = one pixel; the 8*C logits of the pixel are read once (coalesced per class plane) into registers;
//   * the 2^n - 1 subsets are walked in Gray-code order, so every step adds or removes ONE scale from the running sums;
//   * the inverted one-hot target is derived from the label on the fly (no mask tensor, no H2D copy);
//   * per subset the 2 + 2C partial sums are reduced warp-shuffle -> per-warp shared slots -> one partial row per CTA;
//     the fold kernel adds the CTA rows in a fixed order (bit-reproducible) and leaves the Dice sums for the backward.
// Bound: instruction issue (15 subsets x C classes of softmax / softplus arithmetic per pixel), not HBM.  The per-subset sums -- 2 + 2C
// values per pixel -- are reduced over the warp with ONE multi-value reduction (multi_warp_sum: 21 shuffles for 20 values instead of
// 100); with one 5-step warp sum per value the forward spent half of its ~9 400 instructions per pixel there (0.42 -> 0.29 ms).
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int MAXN = 4;               // scales
constexpr int MAXS = (1 << MAXN) - 1; // subsets
constexpr int MC_THREADS = 128;
constexpr float DICE_EPS = 1e-5f;

struct McPtrs {
    const float* fg[MAXN];
    const float* bg[MAXN];
    float* dfg[MAXN];
    float* dbg[MAXN];
};

__device__ __forceinline__ int subset_of(int k, int mode) {   // k = 1 .. nsub
    return mode == 0 ? (k ^ (k >> 1)) : (1 << (k - 1));         // Gray code over all non-empty subsets | singletons
}

// Warp reduction of NV values per lane at once: at every halving step a lane keeps one half of its values, sends the other half to
// its partner and adds what it receives, so NV values cost NV-ish shuffles in total (10 + 5 + 3 + 2 + 1 = 21 for NV = 20) instead of
// 5 each (100), and the totals end up spread over the lanes: lane L holds the total of value multi_slot(L) (or nothing).  Fixed
// order, deterministic.
template <int N>
__device__ __forceinline__ void multi_step(float (&v)[N], int off, bool up) {
    constexpr int H = (N + 1) / 2;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float lo = v[j], hi = (j + H < N) ? v[j + H] : 0.0f;
        const float recv = __shfl_xor_sync(0xffffffffu, up ? lo : hi, off);
        v[j] = (up ? hi : lo) + recv;
    }
}
template <int NV>
__device__ __forceinline__ float multi_warp_sum(float (&v)[NV], int lane) {
    static_assert(NV <= 32, "at most one value per lane");
    constexpr int N1 = (NV + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
    multi_step<NV>(v, 16, (lane & 16) != 0);
    float a[N1];
#pragma unroll
    for (int j = 0; j < N1; ++j) a[j] = v[j];
    multi_step<N1>(a, 8, (lane & 8) != 0);
    float b[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) b[j] = a[j];
    multi_step<N2>(b, 4, (lane & 4) != 0);
    float c[N3];
#pragma unroll
    for (int j = 0; j < N3; ++j) c[j] = b[j];
    multi_step<N3>(c, 2, (lane & 2) != 0);
    float d[N4];
#pragma unroll
    for (int j = 0; j < N4; ++j) d[j] = c[j];
    multi_step<N4>(d, 1, (lane & 1) != 0);
    return d[0];
}
// which value a lane ends up with (-1: a padding slot).  Every lane halves the same N per step; `real` counts how many of the values a
// lane kept are real ones (the upper half of an odd N is one short and padded with a zero)
template <int NV>
__device__ __forceinline__ int multi_slot(int lane) {
    int N = NV, real = NV, idx = 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const int H = (N + 1) / 2;
        if (lane & off) { idx += H; real = real > H ? real - H : 0; } else { real = real < H ? real : H; }
        N = H;
    }
    return real >= 1 ? idx : -1;
}

// row layout of the partial / total sums for subset index j (0-based): [ce, bce, I_0..I_{C-1}, Z_0..Z_{C-1}]
template <int C> struct Row { static constexpr int N = 2 + 2 * C; };

template <int C>
__global__ void __launch_bounds__(MC_THREADS)
mc_loss_fwd_kernel(McPtrs p, const long long* __restrict__ labels, int n, int mode, long long npix, int HW,
                   float* __restrict__ partials /* [grid][nsub*Row + C] */, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    constexpr int R = Row<C>::N;
    const int nsub = mode == 0 ? (1 << n) - 1 : n;
    extern __shared__ float sacc[];                     // [warps][nsub*R + C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = MC_THREADS / 32;
    const int width = nsub * R + C;
    if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;      // the fold kernel that follows counts on it
    for (int i = threadIdx.x; i < nw * width; i += MC_THREADS) sacc[i] = 0.0f;
    __syncthreads();
    float* my = sacc + warp * width;
    const int slot = multi_slot<R>(lane);
    for (long long pix0 = (long long)blockIdx.x * MC_THREADS; pix0 < npix; pix0 += (long long)gridDim.x * MC_THREADS) {
        const long long pix = pix0 + threadIdx.x;
        const bool ok = pix < npix;
        const long long b = ok ? pix / HW : 0, hw = ok ? pix - b * HW : 0;
        const int y = ok ? (int)labels[pix] : -1;
        float f[MAXN][C], g[MAXN][C];
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const long long off = (b * C + c) * HW + hw;
                f[i][c] = (ok && i < n) ? __ldg(p.fg[i] + off) : 0.0f;
                g[i][c] = (ok && i < n) ? __ldg(p.bg[i] + off) : 0.0f;
            }
        // class counts (sum t_c^2 = sum t_c), once per pixel: one ballot per class
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, ok && y == c);
            if (lane == 0) my[nsub * R + c] += (float)__popc(m);
        }
        float so[C], sb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { so[c] = 0.0f; sb[c] = 0.0f; }
        int prev = 0;
        for (int k = 1; k <= nsub; ++k) {
            const int s = subset_of(k, mode), diff = s ^ prev;
            prev = s;
#pragma unroll
            for (int i = 0; i < MAXN; ++i) {
                if (diff & (1 << i)) {
                    const float sg = (s & (1 << i)) ? 1.0f : -1.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) { so[c] = fmaf(sg, f[i][c], so[c]); sb[c] = fmaf(sg, g[i][c], sb[c]); }
                }
            }
            float mx = so[0];
#pragma unroll
            for (int c = 1; c < C; ++c) mx = fmaxf(mx, so[c]);
            float e[C], den = 0.0f, sy = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) { e[c] = __expf(so[c] - mx); den += e[c]; if (c == y) sy = so[c]; }
            const float inv = 1.0f / den;
            float vals[R];                                   // [ce, bce, I_0.., Z_0..] of this pixel
            float bce = 0.0f;
            float* row = my + (k - 1) * R;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float pc = ok ? e[c] * inv : 0.0f;
                const float t = (c == y) ? 1.0f : 0.0f;
                const float x = sb[c];
                bce += ok ? (fmaxf(x, 0.0f) - x * (1.0f - t) + __logf(1.0f + __expf(-fabsf(x)))) : 0.0f;
                vals[2 + c] = pc * t; vals[2 + C + c] = pc * pc;
            }
            vals[0] = ok ? (__logf(den) + mx - sy) : 0.0f;
            vals[1] = bce;
            const float tot = multi_warp_sum<R>(vals, lane);      // lane `slot` holds the warp total of value `slot`
            if (slot >= 0) row[slot] += tot;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += MC_THREADS) {
        float v = 0.0f;
        for (int w = 0; w < nw; ++w) v += sacc[w * width + i];
        partials[(size_t)blockIdx.x * width + i] = v;
    }
}

// totals[i] = sum over the forward kernel's CTA rows, loss from the totals.  CTA = 8 columns x 32 row lanes: a lane adds rows
// lane, lane+32, ... with four independent accumulators (one CTA walking all 1184 rows one dependent load at a time took 217 us),
// the lanes are folded in lane order: a fixed order, bit-reproducible.  The CTA that draws the last ticket computes the scalar.
template <int C>
__global__ void __launch_bounds__(256)
mc_loss_fold_kernel(const float* __restrict__ partials, int nrows, int nsub, long long npix,
                    float lc_ce, float lc_dice, float lc_bce, float* __restrict__ totals, float* __restrict__ loss, unsigned int* __restrict__ ticket) {
    pv2::pdl_prologue();
    constexpr int R = Row<C>::N;
    const int width = nsub * R + C;
    __shared__ float sh[32][9];
    __shared__ float sterm[32];
    __shared__ bool is_last;
    const int cl = threadIdx.x & 7, lane = threadIdx.x >> 3;
    const int col = blockIdx.x * 8 + cl;
    float v = 0.0f;
    if (col < width) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        const float* src = partials + col;
        int r = lane;
        for (; r + 96 < nrows; r += 128) {
            a0 += src[(size_t)r * width]; a1 += src[(size_t)(r + 32) * width];
            a2 += src[(size_t)(r + 64) * width]; a3 += src[(size_t)(r + 96) * width];
        }
        for (; r < nrows; r += 32) a0 += src[(size_t)r * width];
        v = (a0 + a1) + (a2 + a3);
    }
    sh[lane][cl] = v;
    __syncthreads();
    if (threadIdx.x < 8 && col < width) {
        float t = 0.0f;
#pragma unroll
        for (int l = 0; l < 32; ++l) t += sh[l][threadIdx.x];
        totals[col] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float term = 0.0f;
    if ((int)threadIdx.x < nsub) {       // one thread per subset (<= 15)
        const float* row = totals + threadIdx.x * R;
        float dice = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float Y = __ldcg(totals + nsub * R + c);
            dice += 1.0f - (2.0f * __ldcg(row + 2 + c) + DICE_EPS) / (__ldcg(row + 2 + C + c) + Y + DICE_EPS);
        }
        term = lc_ce * __ldcg(row) / (float)npix + lc_dice * dice / (float)C + lc_bce * __ldcg(row + 1) / ((float)npix * (float)C);
    }
    if (threadIdx.x < 32) sterm[threadIdx.x] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
        float L = 0.0f;
        for (int j = 0; j < nsub; ++j) L += sterm[j];      // subset order, as before
        *loss = L;
        *ticket = 0u;
    }
}

template <int C>
__global__ void __launch_bounds__(MC_THREADS)
mc_loss_bwd_kernel(McPtrs p, const long long* __restrict__ labels, const float* __restrict__ grad_loss,
                   const float* __restrict__ totals, int n, int mode, long long npix, int HW,
                   float lc_ce, float lc_dice, float lc_bce) {
    pv2::pdl_prologue();
    constexpr int R = Row<C>::N;
    const int nsub = mode == 0 ? (1 << n) - 1 : n;
    const float gl = *grad_loss;
    const float w_ce = gl * lc_ce / (float)npix, w_bce = gl * lc_bce / ((float)npix * (float)C), w_dice = gl * lc_dice / (float)C;
    // Dice coefficients of every (subset, class), once per CTA: d/dp_c of -(num/D) = -(2 t D - 2 num p) / D^2 = -(cA t - cB p)
    __shared__ float cA[MAXS * C], cB[MAXS * C];
    for (int i = threadIdx.x; i < nsub * C; i += MC_THREADS) {
        const int k = i / C, c = i - k * C;
        const float* row = totals + k * R;
        const float num = 2.0f * row[2 + c] + DICE_EPS, D = row[2 + C + c] + totals[nsub * R + c] + DICE_EPS;
        cA[i] = w_dice * 2.0f / D;
        cB[i] = w_dice * 2.0f * num / (D * D);
    }
    __syncthreads();
    const long long pix = (long long)blockIdx.x * MC_THREADS + threadIdx.x;
    if (pix >= npix) return;
    const long long b = pix / HW, hw = pix - b * HW;
    const int y = (int)labels[pix];
    // ---- foreground: CE + Dice through the softmax ----
    {
        float f[MAXN][C], d[MAXN][C], so[C];
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) { f[i][c] = i < n ? __ldg(p.fg[i] + (b * C + c) * HW + hw) : 0.0f; d[i][c] = 0.0f; }
#pragma unroll
        for (int c = 0; c < C; ++c) so[c] = 0.0f;
        int prev = 0;
        for (int k = 1; k <= nsub; ++k) {
            const int s = subset_of(k, mode), diff = s ^ prev;
            prev = s;
#pragma unroll
            for (int i = 0; i < MAXN; ++i)
                if (diff & (1 << i)) {
                    const float sg = (s & (1 << i)) ? 1.0f : -1.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) so[c] = fmaf(sg, f[i][c], so[c]);
                }
            float mx = so[0];
#pragma unroll
            for (int c = 1; c < C; ++c) mx = fmaxf(mx, so[c]);
            float pr[C], den = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) { pr[c] = __expf(so[c] - mx); den += pr[c]; }
            const float inv = 1.0f / den;
            const float* ka = cA + (k - 1) * C;
            const float* kb = cB + (k - 1) * C;
            float dp[C], dot = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                pr[c] *= inv;
                dp[c] = kb[c] * pr[c] - ((c == y) ? ka[c] : 0.0f);
                dot += pr[c] * dp[c];
            }
            float gs[C];
#pragma unroll
            for (int c = 0; c < C; ++c) gs[c] = pr[c] * (dp[c] - dot) + w_ce * (pr[c] - ((c == y) ? 1.0f : 0.0f));
#pragma unroll
            for (int i = 0; i < MAXN; ++i)
                if (s & (1 << i)) {
#pragma unroll
                    for (int c = 0; c < C; ++c) d[i][c] += gs[c];
                }
        }
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
            if (i < n) {
#pragma unroll
                for (int c = 0; c < C; ++c) p.dfg[i][(b * C + c) * HW + hw] = d[i][c];
            }
    }
    // ---- background: BCE against the inverted one-hot ----
    {
        float g[MAXN][C], d[MAXN][C], sb[C];
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) { g[i][c] = i < n ? __ldg(p.bg[i] + (b * C + c) * HW + hw) : 0.0f; d[i][c] = 0.0f; }
#pragma unroll
        for (int c = 0; c < C; ++c) sb[c] = 0.0f;
        int prev = 0;
        for (int k = 1; k <= nsub; ++k) {
            const int s = subset_of(k, mode), diff = s ^ prev;
            prev = s;
#pragma unroll
            for (int i = 0; i < MAXN; ++i)
                if (diff & (1 << i)) {
                    const float sg = (s & (1 << i)) ? 1.0f : -1.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) sb[c] = fmaf(sg, g[i][c], sb[c]);
                }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float sig = 1.0f / (1.0f + __expf(-sb[c]));
                const float gq = w_bce * (sig - ((c == y) ? 0.0f : 1.0f));
#pragma unroll
                for (int i = 0; i < MAXN; ++i)
                    if (s & (1 << i)) d[i][c] += gq;
            }
        }
#pragma unroll
        for (int i = 0; i < MAXN; ++i)
            if (i < n) {
#pragma unroll
                for (int c = 0; c < C; ++c) p.dbg[i][(b * C + c) * HW + hw] = d[i][c];
            }
    }
}

inline int fwd_grid(long long npix) {
    long long g = (npix + MC_THREADS - 1) / MC_THREADS;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(g > cap ? cap : g);
}

template <int C>
int launch_all(bool backward, const McPtrs& p, const long long* labels, const float* grad_loss, int n, int mode, int B, int H, int W,
               float lc_ce, float lc_dice, float lc_bce, float* loss, float* ws, cudaStream_t st) {
    const long long npix = (long long)B * H * W;
    const int nsub = mode == 0 ? (1 << n) - 1 : n, width = nsub * Row<C>::N + C, grid = fwd_grid(npix);
    float* totals = ws;
    float* partials = ws + ((width + 63) / 64) * 64;
    if (!backward) {
        const size_t smem = sizeof(float) * (MC_THREADS / 32) * width;
        pv2::launch(mc_loss_fwd_kernel<C>, grid, MC_THREADS, smem, st, p, labels, n, mode, npix, H * W, partials,
                    reinterpret_cast<unsigned int*>(totals + width));
        PV2_LAUNCH_CHECK("mc_loss_fwd");
        unsigned int* ticket = reinterpret_cast<unsigned int*>(totals + width);      // inside the 64-float-aligned totals block; zeroed by the forward kernel
        if (width % 64 == 0) { set_error("mc_dual_loss: no room for the fold ticket (width %d)", width); return 1; }   // 30 + 31*C is never a multiple of 64 for C <= 12
        pv2::launch(mc_loss_fold_kernel<C>, (width + 7) / 8, 256, 0, st, partials, grid, nsub, npix, lc_ce, lc_dice, lc_bce, totals, loss, ticket);
        PV2_LAUNCH_CHECK("mc_loss_fold");
    } else {
        pv2::launch(mc_loss_bwd_kernel<C>, (unsigned)((npix + MC_THREADS - 1) / MC_THREADS), MC_THREADS, 0, st, p, labels, grad_loss, totals, n, mode, npix, H * W,
                                                                                                        lc_ce, lc_dice, lc_bce);
        PV2_LAUNCH_CHECK("mc_loss_bwd");
    }
    return 0;
}

int dispatch(bool backward, int C, const McPtrs& p, const long long* labels, const float* grad_loss, int n, int mode, int B, int H, int W,
             float a, float b2, float c2, float* loss, float* ws, cudaStream_t st) {
#define PV2_MC(CV) case CV: return launch_all<CV>(backward, p, labels, grad_loss, n, mode, B, H, W, a, b2, c2, loss, ws, st)
    switch (C) {
        PV2_MC(2); PV2_MC(3); PV2_MC(4); PV2_MC(5); PV2_MC(6); PV2_MC(7); PV2_MC(8); PV2_MC(9); PV2_MC(10); PV2_MC(11); PV2_MC(12);
        default: break;
    }
#undef PV2_MC
    set_error("mc_dual_loss: num_classes=%d not in the compiled range [2,12]", C);
    return 1;
}

int mc_check(const float* const* fg, const float* const* bg, const long long* labels, int n, int mode, int B, int C, int H, int W,
             const void* ws, size_t ws_bytes) {
    PV2_CHECK(fg && bg && labels && ws, "mc_dual_loss: null pointer");
    PV2_CHECK(n >= 1 && n <= MAXN, "mc_dual_loss: number of scales %d out of range [1,%d]", n, MAXN);
    PV2_CHECK(mode == 0 || mode == 1, "mc_dual_loss: mode must be 0 (all non-empty subsets) or 1 (deep supervision)");
    PV2_CHECK(B > 0 && C >= 2 && H > 0 && W > 0, "mc_dual_loss: bad shape");
    for (int i = 0; i < n; ++i) PV2_CHECK(fg[i] && bg[i], "mc_dual_loss: null logits pointer at scale %d", i);
    PV2_CHECK(ws_bytes >= pv2_mc_dual_loss_workspace_bytes(B, C, H, W), "mc_dual_loss: workspace too small");
    return 0;
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" size_t pv2_mc_dual_loss_workspace_bytes(int B, int C, int H, int W) {
    const int width = MAXS * (2 + 2 * C) + C;
    const long long npix = (long long)B * H * W;
    return sizeof(float) * ((size_t)((width + 63) / 64) * 64 + (size_t)fwd_grid(npix) * width);
}

extern "C" int pv2_mc_dual_loss_fwd(const float* const* P_fg, const float* const* P_bg, const long long* labels, int nscales, int mode,
                                    int B, int C, int H, int W, float lc_ce, float lc_dice, float lc_bce, float* loss,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = mc_check(P_fg, P_bg, labels, nscales, mode, B, C, H, W, workspace, workspace_bytes)) return e;
    PV2_CHECK(loss != nullptr, "mc_dual_loss_fwd: null loss pointer");
    McPtrs p = {};
    for (int i = 0; i < nscales; ++i) { p.fg[i] = P_fg[i]; p.bg[i] = P_bg[i]; }
    return dispatch(false, C, p, labels, nullptr, nscales, mode, B, H, W, lc_ce, lc_dice, lc_bce, loss, (float*)workspace, (cudaStream_t)stream);
}

extern "C" int pv2_mc_dual_loss_bwd(const float* const* P_fg, const float* const* P_bg, const long long* labels, const float* grad_loss,
                                    float* const* dP_fg, float* const* dP_bg, int nscales, int mode, int B, int C, int H, int W,
                                    float lc_ce, float lc_dice, float lc_bce, const void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = mc_check(P_fg, P_bg, labels, nscales, mode, B, C, H, W, workspace, workspace_bytes)) return e;
    PV2_CHECK(grad_loss && dP_fg && dP_bg, "mc_dual_loss_bwd: null pointer");
    McPtrs p = {};
    for (int i = 0; i < nscales; ++i) {
        PV2_CHECK(dP_fg[i] && dP_bg[i], "mc_dual_loss_bwd: null gradient pointer at scale %d", i);
        p.fg[i] = P_fg[i]; p.bg[i] = P_bg[i]; p.dfg[i] = dP_fg[i]; p.dbg[i] = dP_bg[i];
    }
    return dispatch(true, C, p, labels, grad_loss, nscales, mode, B, H, W, lc_ce, lc_dice, lc_bce, nullptr, (float*)const_cast<void*>(workspace),
                    (cudaStream_t)stream);
}
