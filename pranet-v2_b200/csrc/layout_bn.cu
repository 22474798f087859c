// Memory-bound companions of the tcgen05 conv kernels: operand packing, weight packing, BatchNorm statistics,
// the fused BN-affine / activation / product / residual "apply" pass and its backward, and the x2 align-corners
// upsample of the partial decoder in NHWC.
//
// Reference ops replaced: nn.BatchNorm2d inside BasicConv2d (binary_seg/lib/pranet.py:37,41-42), the callers'
// F.relu (pranet.py:358-360), RFB_modified's relu(x_cat + conv_res(x)) (pranet.py:82), aggregation's element-wise
// products and concats (pranet.py:111-119) and its nn.Upsample(scale_factor=2, align_corners=True) (pranet.py:93).
//
// "Operand format" = NHWC, channels padded, element bf16 (1 plane) or tf32 hi/lo (2 fp32 planes `plane_stride`
// elements apart) -- what the conv kernels' TMA maps read.  "Raw" = fp32 NHWC rows [pixel][ld] as the conv epilogue
// writes them, possibly as several slabs (split-K partials / several gradient contributions) that are summed on load.
#include "pv2_common.cuh"
#include "ticket.cuh"

namespace pv2 {
namespace {

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// KIND 0: bf16, 1 plane.  KIND 1: tf32, nplanes = 1 (hi) or 2 (hi, lo)
template <int KIND>
__device__ __forceinline__ void store_op(void* base, long long plane_stride, int nplanes, long long idx, float v) {
    if constexpr (KIND == 0) {
        reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    } else {
        float* p = reinterpret_cast<float*>(base);
        const float hi = tf32_rna(v);
        p[idx] = hi;
        if (nplanes > 1) p[idx + plane_stride] = tf32_rna(v - hi);
    }
}
template <int KIND>
__device__ __forceinline__ float load_op(const void* base, long long plane_stride, int nplanes, long long idx) {
    if constexpr (KIND == 0) {
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
    } else {
        const float* p = reinterpret_cast<const float*>(base);
        float v = p[idx];
        if (nplanes > 1) v += p[idx + plane_stride];
        return v;
    }
}

// ------------------------------------------------------------------------------------------------------
// weights: OIHW fp32 -> operand [o][tap][i_ld] (+ i_off).  mode 0 (fprop): o = co, i = ci, tap = kh*KW+kw.
// mode 1 (dgrad): o = ci, i = co, tap flipped.  Padding channels are zeroed by the caller (memset) once.
// ------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void weight_pack_kernel(const float* __restrict__ w, void* out, long long plane_stride, int nplanes,
                                   int Cout, int Cin, int KH, int KW, int mode, int i_ld, int i_off, int o_off) {
    pv2::pdl_prologue();
    const long long total = (long long)Cout * Cin * KH * KW;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int kw = (int)(e % KW);
        const int kh = (int)((e / KW) % KH);
        const int ci = (int)((e / ((long long)KW * KH)) % Cin);
        const int co = (int)(e / ((long long)KW * KH * Cin));
        const int taps = KH * KW;
        long long idx;
        if (mode == 0) idx = ((long long)(co + o_off) * taps + kh * KW + kw) * i_ld + i_off + ci;
        else idx = ((long long)(ci + o_off) * taps + (KH - 1 - kh) * KW + (KW - 1 - kw)) * i_ld + i_off + co;
        store_op<KIND>(out, plane_stride, nplanes, idx, w[e]);
    }
}

// wgrad partials [split][Cout_total][taps][Cin_p] -> OIHW fp32 gradient of one conv (rows co_off .. co_off+Cout)
__global__ void wgrad_unpack_kernel(const float* __restrict__ part, long long split_stride, int splits, float* __restrict__ dw,
                                    int Cout, int Cin, int KH, int KW, int Cin_p, int co_off) {
    pv2::pdl_prologue();
    const long long total = (long long)Cout * Cin * KH * KW;
    const int taps = KH * KW;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int kw = (int)(e % KW);
        const int kh = (int)((e / KW) % KH);
        const int ci = (int)((e / ((long long)KW * KH)) % Cin);
        const int co = (int)(e / ((long long)KW * KH * Cin));
        const long long idx = ((long long)(co + co_off) * taps + kh * KW + kw) * Cin_p + ci;
        float acc = 0.0f;
        for (int s = 0; s < splits; ++s) acc += part[s * split_stride + idx];
        dw[e] = acc;
    }
}

// multi-tensor variants: a handful of launches pack every conv weight of the head (both layouts) / unpack every weight
// gradient.  The descriptors travel BY VALUE in the kernel parameters (<= 4 KB per launch), so nothing has to be staged
// in device memory and the launches are CUDA-graph capturable as they are.
// Work unit = one TILE of one tensor: (co tile, ci tile, all taps), staged in shared memory so that the OIHW side AND the
// operand side are both accessed in contiguous runs (the three layouts [co][ci][tap], [co][tap][ci] and [ci][tap'][co] are
// permutations of each other: an element-per-thread walk is coalesced on one side only and pays one 32-byte sector per 2-byte
// element on the other).  A CTA finds its tensor once (binary search over the per-tensor tile prefix), not once per element,
// and all in-tensor index arithmetic is 32-bit.
constexpr int PACK_BATCH = 40;     // 40 * 88 B + 41 * 4 B = 3684 B of kernel parameters
constexpr int UNPACK_BATCH = 56;   // 56 * 64 B + 57 * 4 B = 3812 B
constexpr int PK_TCI = 32, UP_TCI = 32;    // input channels per tile
// output channels per tile: ~12800 (pack) / ~2048 (unpack) elements whatever the filter size, so a 1x1 tile is not 25x
// smaller than a 5x5 one.  pack keeps >= 16 so that the dgrad layout's co runs stay >= one 32-byte sector.
__host__ __device__ inline int pk_tco(int taps) { const int t = 12800 / (PK_TCI * taps); return t < 16 ? 16 : (t > 128 ? 128 : t); }
constexpr int UP_EPT = 8, UP_SPU = 8;       // elements per thread of an unpack tile (2048 / 256); splits fetched per round trip
__host__ __device__ inline int up_tco(int taps) { const int t = 2048 / (UP_TCI * taps); return t < 1 ? 1 : (t > 64 ? 64 : t); }
struct PackBatch { pv2_pack_desc d[PACK_BATCH]; int tile_start[PACK_BATCH + 1]; };
struct UnpackBatch { pv2_unpack_desc d[UNPACK_BATCH]; int tile_start[UNPACK_BATCH + 1]; };

template <int NB>
__device__ __forceinline__ int find_tensor(const int* tile_start, int n, int tile) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile_start[mid] <= tile) lo = mid; else hi = mid - 1; }
    return lo;
}

template <int KIND>
__global__ void __launch_bounds__(256)
weight_pack_multi_kernel(const __grid_constant__ PackBatch b, int n, int nplanes) {
    pv2::pdl_prologue();
    extern __shared__ float pk_tile[];           // [pk_tco(taps)][PK_TCI * taps + 1]
    __shared__ int s_t;
    if (threadIdx.x == 0) s_t = find_tensor<PACK_BATCH>(b.tile_start, n, (int)blockIdx.x);
    __syncthreads();
    const pv2_pack_desc& t = b.d[s_t];
    const int taps = t.KH * t.KW;
    const int ci_tiles = (t.Cin + PK_TCI - 1) / PK_TCI;
    const int lt = (int)blockIdx.x - b.tile_start[s_t];
    const int TCO = pk_tco(taps);
    const int co0 = (lt / ci_tiles) * TCO, ci0 = (lt % ci_tiles) * PK_TCI;
    const int nco = min(TCO, t.Cout - co0), nci = min(PK_TCI, t.Cin - ci0);
    const int run = nci * taps, pitch = PK_TCI * taps + 1;
    // Warp-per-row loops, lanes along the contiguous axis of whichever side is being touched: no integer division per element
    // (three passes of i / run, i % nci, r % taps with run-time divisors made this kernel ~60 us for the head's 27 MB of weights).
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // OIHW side: for every co of the tile the (ci, tap) run is contiguous
    for (int col = warp; col < nco; col += 8) {
        const float* src = t.w + ((size_t)(co0 + col) * t.Cin + ci0) * taps;
        for (int j = lane; j < run; j += 32) pk_tile[col * pitch + j] = __ldg(src + j);
    }
    __syncthreads();
    // fprop layout [co][tap][ci]: ci fastest (lane = ci; shared-memory stride `taps` is odd or 1: conflict free)
    for (int col = warp; col < nco; col += 8) {
        const long long row = (long long)(co0 + col + t.f_ooff) * taps;
        for (int tap = 0; tap < taps; ++tap)
            if (lane < nci)
                store_op<KIND>(t.out_f, t.f_plane, nplanes, (row + tap) * t.f_ild + t.f_ioff + ci0 + lane, pk_tile[col * pitch + lane * taps + tap]);
    }
    // dgrad layout [ci][flipped tap][co]: co fastest (lane = co; stride `pitch` is odd)
    if (t.out_d) {
        for (int cil = warp; cil < nci; cil += 8) {
            const long long row = (long long)(ci0 + cil + t.d_ooff) * taps;
            for (int tap = 0; tap < taps; ++tap)
                for (int col = lane; col < nco; col += 32)
                    store_op<KIND>(t.out_d, t.d_plane, nplanes, (row + (taps - 1 - tap)) * t.d_ild + t.d_ioff + co0 + col, pk_tile[col * pitch + cil * taps + tap]);
        }
    }
}

__global__ void __launch_bounds__(256)
wgrad_unpack_multi_kernel(const __grid_constant__ UnpackBatch b, int n) {
    pv2::pdl_prologue();
    extern __shared__ float up_tile[];           // [up_tco(taps)][taps][UP_TCI + 1]
    __shared__ int s_t;
    if (threadIdx.x == 0) s_t = find_tensor<UNPACK_BATCH>(b.tile_start, n, (int)blockIdx.x);
    __syncthreads();
    const pv2_unpack_desc& t = b.d[s_t];
    const int taps = t.KH * t.KW;
    const int ci_tiles = (t.Cin + UP_TCI - 1) / UP_TCI;
    const int lt = (int)blockIdx.x - b.tile_start[s_t];
    const int TCO = up_tco(taps);
    const int co0 = (lt / ci_tiles) * TCO, ci0 = (lt % ci_tiles) * UP_TCI;
    const int nco = min(TCO, t.Cout - co0), nci = min(UP_TCI, t.Cin - ci0);
    // partial side [split][co][tap][ci]: ci fastest.  A thread owns up to UP_EPT elements of the tile and walks the splits in
    // the OUTER loop with one accumulator per element: UP_EPT independent loads in flight per thread (a per-element split
    // loop is a chain of dependent L2 round trips -- 31 splits x 16 elements made a CTA live > 100 us), and every element
    // still sums its splits in split order (deterministic).
    const int nelem = nco * taps * nci;          // <= up_tco(taps) * taps * UP_TCI <= 2048 = 256 threads * UP_EPT
    int soff[UP_EPT], doff[UP_EPT];
    float acc[UP_EPT];
#pragma unroll
    for (int e = 0; e < UP_EPT; ++e) {
        const int i = threadIdx.x + e * 256;
        acc[e] = 0.0f;
        soff[e] = -1; doff[e] = 0;
        if (i < nelem) {
            const int cil = i % nci, r = i / nci, tap = r % taps, col = r / taps;
            soff[e] = ((co0 + col + t.co_off) * taps + tap) * t.Cin_p + ci0 + cil;     // < 2^31: one conv's partial slab
            doff[e] = (col * taps + tap) * (UP_TCI + 1) + cil;
        }
    }
    // UP_SPU splits x UP_EPT elements = 64 independent loads per thread per round trip (a 31-split tensor is 4 round trips)
    for (int sp0 = 0; sp0 < t.splits; sp0 += UP_SPU) {
        float v[UP_SPU][UP_EPT];
#pragma unroll
        for (int u = 0; u < UP_SPU; ++u) {
            const float* src = t.part + (size_t)(sp0 + u) * t.split_stride;
            const bool live = sp0 + u < t.splits;
#pragma unroll
            for (int e = 0; e < UP_EPT; ++e) v[u][e] = (live && soff[e] >= 0) ? __ldg(src + soff[e]) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < UP_SPU; ++u) {
#pragma unroll
            for (int e = 0; e < UP_EPT; ++e) acc[e] += v[u][e];      // split order preserved per element
        }
    }
#pragma unroll
    for (int e = 0; e < UP_EPT; ++e)
        if (soff[e] >= 0) up_tile[doff[e]] = acc[e];
    __syncthreads();
    // OIHW side: for every co the (ci, tap) run is contiguous
    const int run = nci * taps;
    for (int i = threadIdx.x; i < nco * run; i += 256) {
        const int col = i / run, j = i - col * run, cil = j / taps, tap = j - cil * taps;
        t.dw[((size_t)(co0 + col) * t.Cin + ci0) * taps + j] = up_tile[(col * taps + tap) * (UP_TCI + 1) + cil];
    }
}

// ------------------------------------------------------------------------------------------------------
// NCHW (fp32 | bf16) -> operand NHWC, tiled transpose through shared memory; and the reverse for gradients
// ------------------------------------------------------------------------------------------------------
template <typename TIN, int KIND>
__global__ void __launch_bounds__(256)
pack_nchw_kernel(const TIN* __restrict__ x, void* out, long long plane_stride, int nplanes, int C, int HW, int ld, int c_off) {
    pv2::pdl_prologue();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? to_f(x[((long long)n * C + c) * HW + p]) : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + tx;
        if (p < HW && c < C) store_op<KIND>(out, plane_stride, nplanes, ((long long)n * HW + p) * ld + c_off + c, tile[tx][j]);
    }
}

struct Slabs {               // up to 8 fp32 row-major [M][ld] gradient / partial slabs that are summed on load
    const float* p[8];
    int ld[8];
    int off[8];
    int n;
};
__device__ __forceinline__ float slab_sum(const Slabs& s, long long row, int c) {
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < s.n) v += s.p[i][row * s.ld[i] + s.off[i] + c];
    return v;
}

template <typename TOUT>
__global__ void __launch_bounds__(256)
unpack_to_nchw_kernel(const Slabs g, TOUT* __restrict__ dx, int C, int HW) {
    pv2::pdl_prologue();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + tx;
        tile[j][tx] = (p < HW && c < C) ? slab_sum(g, (long long)n * HW + p, c) : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + tx;
        if (c < C && p < HW) dx[((long long)n * C + c) * HW + p] = from_f<TOUT>(tile[tx][j]);
    }
}

// ------------------------------------------------------------------------------------------------------
// BatchNorm batch statistics over raw conv output (sums split-K slabs first and writes the total back to slab 0)
//   partial kernel: grid (row_blocks, ceil(C/32)); block (32 channels x 8 row lanes); shifted sums -> (n, mean, M2)
//   finalize: Chan-combine the row blocks; mean/invstd saved for backward; scale/shift for the apply pass;
//             running stats updated like nn.BatchNorm2d (momentum, unbiased variance), num_batches_tracked += 1
// ------------------------------------------------------------------------------------------------------
constexpr int ST_ROWS_MIN = 32, ST_ROWS_MAX = 256;   // rows per partial block (chosen per call so that ~2 waves of CTAs exist)

inline int pick_rows(long long M, int C) {
    const int cg = (C + 31) / 32;
    long long rb_target = (2LL * kNumSMs + cg - 1) / cg;
    if (rb_target < 1) rb_target = 1;
    long long rows = (M + rb_target - 1) / rb_target;
    rows = (rows + 7) / 8 * 8;
    if (rows < ST_ROWS_MIN) rows = ST_ROWS_MIN;
    if (rows > ST_ROWS_MAX) rows = ST_ROWS_MAX;
    return (int)rows;
}

__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(float* __restrict__ y, long long slab_stride, int nslabs, long long M, int C, int ld, int rows_pb,
                        float* __restrict__ part /* [row_blocks][C][3] */) {
    pv2::pdl_prologue();
    __shared__ float sh[8][32][3];
    const int c = blockIdx.y * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * rows_pb, r1 = min(M, r0 + rows_pb);
    float cnt = 0.0f, mean = 0.0f, m2 = 0.0f;
    if (c < C) {
        float shift = 0.0f, s1 = 0.0f, s2 = 0.0f;
        bool first = true;
        // Two rows x up to 8 slabs = 16 independent loads per round trip; a row-by-row, slab-by-slab walk is a chain of
        // dependent L2 latencies (7 rows x 8 slabs ~ 30 us for the 11x11 maps of the split-K 5x5 convs).  Slabs are summed in
        // slab order, rows enter the running statistics in row order: same arithmetic as the serial walk.
        for (long long r = r0 + ty; r < r1; r += 16) {
            float v[2][8];
            const bool two = r + 8 < r1;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long rr = (u == 0 || two) ? r + 8 * u : r;
#pragma unroll
                for (int sl = 0; sl < 8; ++sl) v[u][sl] = sl < nslabs ? y[sl * slab_stride + rr * ld + c] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                float t = v[u][0];
#pragma unroll
                for (int sl = 1; sl < 8; ++sl) t += v[u][sl];       // slabs beyond nslabs contribute +0.0f
                for (int sl = 8; sl < nslabs; ++sl) t += y[sl * slab_stride + (r + 8 * u) * ld + c];
                if (nslabs > 1) y[(r + 8 * u) * ld + c] = t;
                if (first) { shift = t; first = false; }
                const float d = t - shift;
                s1 += d; s2 += d * d; cnt += 1.0f;
            }
        }
        if (cnt > 0.0f) { mean = shift + s1 / cnt; m2 = fmaxf(s2 - s1 * s1 / cnt, 0.0f); }
    }
    sh[ty][threadIdx.x & 31][0] = cnt; sh[ty][threadIdx.x & 31][1] = mean; sh[ty][threadIdx.x & 31][2] = m2;
    __syncthreads();
    if (ty == 0 && c < C) {
        float n = 0.0f, mu = 0.0f, M2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) chan_combine(n, mu, M2, sh[k][threadIdx.x][0], sh[k][threadIdx.x][1], sh[k][threadIdx.x][2]);
        float* o = part + ((long long)blockIdx.x * C + c) * 3;
        o[0] = n; o[1] = mu; o[2] = M2;
    }
}

// one warp per channel: lanes take row blocks b = lane, lane+32, ... (fixed order), then a fixed shuffle tree
__global__ void __launch_bounds__(128)
bn_stats_finalize_kernel(const float* __restrict__ part, int row_blocks, int C, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                         float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                         float* __restrict__ mean_out, float* __restrict__ invstd_out, float* __restrict__ scale,
                         float* __restrict__ shift) {
    pv2::pdl_prologue();
    const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c == 0 && lane == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    float n = 0.0f, mu = 0.0f, M2 = 0.0f;
    for (int b = lane; b < row_blocks; b += 32) {
        const float* p = part + ((long long)b * C + c) * 3;
        chan_combine(n, mu, M2, p[0], p[1], p[2]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o), mub = __shfl_xor_sync(0xffffffffu, mu, o), M2b = __shfl_xor_sync(0xffffffffu, M2, o);
        // combine in a lane-symmetric way so every lane ends with the same value
        const float nt = n + nb;
        if (nt > 0.0f) {
            const float d = mub - mu;
            const float mu_new = (n * mu + nb * mub) / nt;
            M2 = M2 + M2b + d * d * n * nb / nt;
            mu = mu_new;
        }
        n = nt;
    }
    if (lane != 0) return;
    const float var = M2 / n;
    const float inv = rsqrtf(var + eps);
    mean_out[c] = mu;
    invstd_out[c] = inv;
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * inv;
    shift[c] = b - mu * g * inv;
    if (running_mean) {
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mu;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (n > 1.0f ? M2 / (n - 1.0f) : var);
    }
}

// the same fold for a whole conv group: one warp per channel of the group, BatchNorm module found through the segment table
__global__ void __launch_bounds__(128)
bn_group_finalize_kernel(const float* __restrict__ part, int row_blocks, int C, const __grid_constant__ BnFuseDev bn) {
    pv2::pdl_prologue();
    const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    float n = 0.0f, mu = 0.0f, M2 = 0.0f;
    for (int b = lane; b < row_blocks; b += 32) {
        const float* p = part + ((long long)b * C + c) * 3;
        chan_combine(n, mu, M2, p[0], p[1], p[2]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float nb = __shfl_xor_sync(0xffffffffu, n, o), mub = __shfl_xor_sync(0xffffffffu, mu, o), M2b = __shfl_xor_sync(0xffffffffu, M2, o);
        const float nt = n + nb;
        if (nt > 0.0f) {
            const float d = mub - mu;
            const float mu_new = (n * mu + nb * mub) / nt;
            M2 = M2 + M2b + d * d * n * nb / nt;
            mu = mu_new;
        }
        n = nt;
    }
    if (lane == 0) bn_write_channel(bn.f, c, n, mu, M2);
}

// eval-mode BN / plain bias as an affine: scale = gamma / sqrt(rv + eps), shift = beta - rm * scale
__global__ void bn_eval_affine_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
    pv2::pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = (gamma ? gamma[c] : 1.0f) * rsqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = (beta ? beta[c] : 0.0f) - rm[c] * s;
}

// ------------------------------------------------------------------------------------------------------
// apply:  a1 = y1*s1+b1 ; [a2 = y2*s2+b2 ; v = a1 (+|*) a2] ; [v *= mult] ; [relu] -> operand NHWC slice or NCHW fp32
// ------------------------------------------------------------------------------------------------------
struct ApplyArgs {
    const float* y1; int ld1, off1; const float* s1; const float* b1;
    const float* y2; int ld2, off2; const float* s2; const float* b2;
    int ns1, ns2; long long ss1, ss2;   // split-K slabs of y1 / y2 still to be summed on load (eval path), stride in elements
    int combine;                 // 0 none, 1 add, 2 mul
    const void* mult; long long mult_plane; int mult_planes, mult_ld, mult_off;   // operand-format multiplier (or null)
    int relu;
    long long M; int C, HW;
    void* out; long long out_plane; int out_planes, out_ld, out_off;
    int out_nchw;                // 1: out is fp32 NCHW [N][C][HW]
};

__device__ __forceinline__ float raw_load(const float* y, int ld, int off, int ns, long long ss, long long r, int c) {
    float v = y[r * ld + off + c];
    for (int s = 1; s < ns; ++s) v += y[s * ss + r * ld + off + c];
    return v;
}

template <int KIND>
__device__ __forceinline__ float apply_value(const ApplyArgs& a, long long r, int c, float* a1o, float* a2o, float* mo) {
    const float a1 = raw_load(a.y1, a.ld1, a.off1, a.ns1, a.ss1, r, c) * a.s1[c] + a.b1[c];
    float a2 = 0.0f, v = a1;
    if (a.combine) {
        a2 = raw_load(a.y2, a.ld2, a.off2, a.ns2, a.ss2, r, c) * a.s2[c] + a.b2[c];
        v = a.combine == 1 ? a1 + a2 : a1 * a2;
    }
    float m = 1.0f;
    if (a.mult) { m = load_op<KIND>(a.mult, a.mult_plane, a.mult_planes, r * a.mult_ld + a.mult_off + c); v *= m; }
    *a1o = a1; *a2o = a2; *mo = m;
    return v;
}

// ------------------------------------------------------------------------------------------------------
// Deferred BatchNorm statistics.  The persistent conv kernel leaves, per channel, sum x and sum x^2 over all pixels as two
// doubles (added by its CTAs, see conv_fwd2_kernel); the kernel that consumes a channel slice turns them into scale / shift here,
// in its prologue: two loads per channel.  Every CTA keeps the slice's scale / shift in its own shared memory; CTA 0 also
// publishes mean / invstd / scale / shift (saved for the backward pass) and updates the running statistics exactly like
// nn.BatchNorm2d (momentum, unbiased running variance).
// ------------------------------------------------------------------------------------------------------
constexpr int DEFER_MAX_C = 512;
struct Defer2 { pv2_bn_defer d[2]; };

__device__ __forceinline__ void bn_fold_deferred(const pv2_bn_defer& d, int C, float* s_scale, float* s_shift,
                                                 bool writer, float* g_scale, float* g_shift) {
    const double* sums = reinterpret_cast<const double*>(d.part) + (size_t)PV2_BN_ACC_STRIDE * d.c_off;
    const double N = (double)d.count;
    for (int cl = threadIdx.x; cl < C; cl += blockDim.x) {
        const double S1 = __ldcg(sums + (size_t)PV2_BN_ACC_STRIDE * cl), S2 = __ldcg(sums + (size_t)PV2_BN_ACC_STRIDE * cl + 1);
        const double dmean = S1 / N;
        double dvar = S2 / N - dmean * dmean;
        if (dvar < 0.0) dvar = 0.0;
        const float mu = (float)dmean, var = (float)dvar, n = (float)d.count;
        const float inv = rsqrtf(var + d.eps);
        const float g = d.gamma ? d.gamma[cl] : 1.0f, b = d.beta ? d.beta[cl] : 0.0f;
        const float sc = g * inv, sh = b - mu * g * inv;
        s_scale[cl] = sc; s_shift[cl] = sh;
        if (writer) {
            g_scale[cl] = sc; g_shift[cl] = sh;
            d.mean[cl] = mu; d.invstd[cl] = inv;
            if (d.running_mean) {
                d.running_mean[cl] = (1.0f - d.momentum) * d.running_mean[cl] + d.momentum * mu;
                d.running_var[cl] = (1.0f - d.momentum) * d.running_var[cl] + d.momentum * (n > 1.0f ? var * n / (n - 1.0f) : var);
            }
            if (cl == 0 && d.num_batches_tracked) *d.num_batches_tracked += 1;
        }
    }
    __syncthreads();
}

// shared prologue of the apply kernels: returns the ApplyArgs to use (scale / shift redirected to shared memory when deferred)
#define PV2_APPLY_DEFER_PROLOGUE(a, df, la)                                                                             \
    __shared__ __align__(16) float s_aff[2][2][DEFER_MAX_C];                                                            \
    ApplyArgs la = a;                                                                                                   \
    if (df.d[0].part) {                                                                                                 \
        bn_fold_deferred(df.d[0], a.C, s_aff[0][0], s_aff[0][1], blockIdx.x == 0, const_cast<float*>(a.s1), const_cast<float*>(a.b1)); \
        la.s1 = s_aff[0][0]; la.b1 = s_aff[0][1];                                                                       \
    }                                                                                                                   \
    if (df.d[1].part) {                                                                                                 \
        bn_fold_deferred(df.d[1], a.C, s_aff[1][0], s_aff[1][1], blockIdx.x == 0, const_cast<float*>(a.s2), const_cast<float*>(a.b2)); \
        la.s2 = s_aff[1][0]; la.b2 = s_aff[1][1];                                                                       \
    }

template <int KIND>
__global__ void __launch_bounds__(256)
act_apply_kernel(const ApplyArgs a0, const Defer2 df) {
    pv2::pdl_prologue();
    PV2_APPLY_DEFER_PROLOGUE(a0, df, a)
    const long long total = a.M * a.C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / a.C;
        const int c = (int)(e - r * a.C);
        float a1, a2, m;
        float v = apply_value<KIND>(a, r, c, &a1, &a2, &m);
        if (a.relu) v = fmaxf(v, 0.0f);
        if (a.out_nchw) {
            const long long n = r / a.HW, p = r - n * a.HW;
            reinterpret_cast<float*>(a.out)[(n * a.C + c) * a.HW + p] = v;
        } else {
            store_op<KIND>(a.out, a.out_plane, a.out_planes, r * a.out_ld + a.out_off + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// backward of apply + BN.  dz arrives as summed slabs (raw NHWC) or as one fp32 NCHW tensor.
//   reduce pass : per channel S1_i = sum da_i, S2_i = sum da_i * yhat_i  (i = 1,2) as row-block partials;
//                 also writes d(mult) raw fp32 [M][C] when a multiplier is present
//   finalize    : fold partials -> sums[4][C]; dgamma_i = S2_i, dbeta_i = S1_i
//   dx pass     : dy_i = gamma_i*invstd_i * (da_i - S1_i/n - yhat_i*S2_i/n)  (bn_train) or scale_i*da_i (affine)
//                 written in operand format for the dgrad / wgrad GEMMs
// ------------------------------------------------------------------------------------------------------
struct BwdArgs {
    ApplyArgs f;                  // the forward description (out fields unused)
    Slabs dz; const float* dz_nchw;
    const float* mean1; const float* inv1; const float* mean2; const float* inv2;
    float* dmult; int dmult_ld;   // raw fp32 gradient of the multiplier (or null)
    float* part;                  // [row_blocks][4][C]
    const float* sums;            // [4][C] (dx pass)
    int bn_train;                 // 1: batch-stat BN backward, 0: plain affine
    int rows_pb;                  // rows per reduce block
    void* dy1; long long dy1_plane; int dy1_planes, dy1_ld;
    void* dy2; long long dy2_plane; int dy2_planes, dy2_ld;
};

template <int KIND>
__device__ __forceinline__ void bwd_da(const BwdArgs& b, long long r, int c, float* da1, float* da2, float* yh1, float* yh2, float* dm) {
    const ApplyArgs& a = b.f;
    float a1, a2, m;
    const float v = apply_value<KIND>(a, r, c, &a1, &a2, &m);
    float g;
    if (b.dz_nchw) {
        const long long n = r / a.HW, p = r - n * a.HW;
        g = b.dz_nchw[(n * a.C + c) * a.HW + p];
    } else {
        g = slab_sum(b.dz, r, c);
    }
    if (a.relu && !(v > 0.0f)) g = 0.0f;
    float comb = a1;
    if (a.combine == 1) comb = a1 + a2; else if (a.combine == 2) comb = a1 * a2;
    *dm = g * comb;
    const float dc = a.mult ? g * m : g;
    *da1 = a.combine == 2 ? dc * a2 : dc;
    *da2 = a.combine == 0 ? 0.0f : (a.combine == 2 ? dc * a1 : dc);
    const float y1v = raw_load(a.y1, a.ld1, a.off1, a.ns1, a.ss1, r, c);
    *yh1 = b.mean1 ? (y1v - b.mean1[c]) * b.inv1[c] : y1v;
    const float y2v = a.combine ? raw_load(a.y2, a.ld2, a.off2, a.ns2, a.ss2, r, c) : 0.0f;
    *yh2 = (a.combine && b.mean2) ? (y2v - b.mean2[c]) * b.inv2[c] : y2v;
}

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const BwdArgs b) {
    pv2::pdl_prologue();
    __shared__ float sh[8][32][4];
    const int C = b.f.C;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * b.rows_pb, r1 = min(b.f.M, r0 + b.rows_pb);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        for (long long r = r0 + ty; r < r1; r += 8) {
            float da1, da2, yh1, yh2, dm;
            bwd_da<KIND>(b, r, c, &da1, &da2, &yh1, &yh2, &dm);
            s[0] += da1; s[1] += da1 * yh1; s[2] += da2; s[3] += da2 * yh2;
            if (b.dmult) b.dmult[r * b.dmult_ld + c] = dm;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sh[ty][threadIdx.x & 31][k] = s[k];
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x][k];
            b.part[((long long)blockIdx.x * 4 + k) * C + c] = t;
        }
    }
}

// one warp per channel, lanes over row blocks (fixed order) + shuffle tree
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ part, int row_blocks, int C, float* __restrict__ sums,
                       float* __restrict__ dgamma1, float* __restrict__ dbeta1, float* __restrict__ dgamma2,
                       float* __restrict__ dbeta2) {
    pv2::pdl_prologue();
    const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = lane; b < row_blocks; b += 32)
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] += part[((long long)b * 4 + k) * C + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
    if (lane != 0) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) sums[k * C + c] = s[k];
    if (dbeta1) dbeta1[c] = s[0];
    if (dgamma1) dgamma1[c] = s[1];
    if (dbeta2) dbeta2[c] = s[2];
    if (dgamma2) dgamma2[c] = s[3];
}

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_dx_kernel(const BwdArgs b) {
    pv2::pdl_prologue();
    const ApplyArgs& a = b.f;
    const long long total = a.M * a.C;
    const float invn = 1.0f / (float)a.M;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / a.C;
        const int c = (int)(e - r * a.C);
        float da1, da2, yh1, yh2, dm;
        bwd_da<KIND>(b, r, c, &da1, &da2, &yh1, &yh2, &dm);
        float d1, d2 = 0.0f;
        if (b.bn_train) {
            d1 = a.s1[c] * (da1 - b.sums[c] * invn - yh1 * b.sums[a.C + c] * invn);          // s1 = gamma*invstd
            if (a.combine) d2 = a.s2[c] * (da2 - b.sums[2 * a.C + c] * invn - yh2 * b.sums[3 * a.C + c] * invn);
        } else {
            d1 = a.s1[c] * da1;
            if (a.combine) d2 = a.s2[c] * da2;
        }
        store_op<KIND>(b.dy1, b.dy1_plane, b.dy1_planes, r * b.dy1_ld + c, d1);
        if (a.combine) store_op<KIND>(b.dy2, b.dy2_plane, b.dy2_planes, r * b.dy2_ld + c, d2);
    }
}

// ------------------------------------------------------------------------------------------------------
// 4-channel vector forms of apply / backward (every BasicConv2d of the head qualifies: C % 4 == 0, slices at channel
// offsets that are multiples of 4).  One thread = one pixel x 4 channels: 16-byte loads of the raw fp32 rows, 8-byte
// bf16 stores.  The scalar kernels above remain for C = 1 / 9 head maps and NCHW outputs.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) { return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_set(float v) { return make_float4(v, v, v, v); }

template <int KIND>
__device__ __forceinline__ void store_op4(void* base, long long plane_stride, int nplanes, long long idx, float4 v) {
    if constexpr (KIND == 0) {
        store4<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(base) + idx, v);
    } else {
        float* p = reinterpret_cast<float*>(base);
        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        *reinterpret_cast<float4*>(p + idx) = hi;
        if (nplanes > 1)
            *reinterpret_cast<float4*>(p + idx + plane_stride) =
                make_float4(tf32_rna(v.x - hi.x), tf32_rna(v.y - hi.y), tf32_rna(v.z - hi.z), tf32_rna(v.w - hi.w));
    }
}
template <int KIND>
__device__ __forceinline__ float4 load_op4(const void* base, long long plane_stride, int nplanes, long long idx) {
    if constexpr (KIND == 0) {
        return load4<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    } else {
        const float* p = reinterpret_cast<const float*>(base);
        float4 v = f4_ld(p + idx);
        if (nplanes > 1) v = f4_add(v, f4_ld(p + idx + plane_stride));
        return v;
    }
}
__device__ __forceinline__ float4 raw_load4(const float* y, int ld, int off, int ns, long long ss, long long r, int c) {
    float4 v = f4_ld(y + r * ld + off + c);
    for (int s = 1; s < ns; ++s) v = f4_add(v, f4_ld(y + s * ss + r * ld + off + c));
    return v;
}
__device__ __forceinline__ float4 slab_sum4(const Slabs& s, long long row, int c) {
    float4 v = f4_set(0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < s.n) v = f4_add(v, f4_ld(s.p[i] + row * s.ld[i] + s.off[i] + c));
    return v;
}

// scale / shift vectors may live in shared memory (deferred BatchNorm fold): generic loads, not ld.global.nc
__device__ __forceinline__ float4 f4_ldp(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int KIND>
__device__ __forceinline__ float4 apply_value4(const ApplyArgs& a, long long r, int c, float4* a1o, float4* a2o, float4* mo) {
    const float4 a1 = f4_fma(raw_load4(a.y1, a.ld1, a.off1, a.ns1, a.ss1, r, c), f4_ldp(a.s1 + c), f4_ldp(a.b1 + c));
    float4 a2 = f4_set(0.0f), v = a1;
    if (a.combine) {
        a2 = f4_fma(raw_load4(a.y2, a.ld2, a.off2, a.ns2, a.ss2, r, c), f4_ldp(a.s2 + c), f4_ldp(a.b2 + c));
        v = a.combine == 1 ? f4_add(a1, a2) : f4_mul(a1, a2);
    }
    float4 m = f4_set(1.0f);
    if (a.mult) { m = load_op4<KIND>(a.mult, a.mult_plane, a.mult_planes, r * a.mult_ld + a.mult_off + c); v = f4_mul(v, m); }
    *a1o = a1; *a2o = a2; *mo = m;
    return v;
}

template <int KIND>
__global__ void __launch_bounds__(256)
act_apply4_kernel(const ApplyArgs a0, const Defer2 df) {
    pv2::pdl_prologue();
    PV2_APPLY_DEFER_PROLOGUE(a0, df, a)
    const unsigned C4 = (unsigned)a.C >> 2;
    const unsigned total = (unsigned)a.M * C4;
    for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < total; e += gridDim.x * 256u) {
        const unsigned r = e / C4;
        const int c = (int)(e - r * C4) << 2;
        float4 a1, a2, m;
        float4 v = apply_value4<KIND>(a, r, c, &a1, &a2, &m);
        if (a.relu) v = make_float4(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f), fmaxf(v.z, 0.0f), fmaxf(v.w, 0.0f));
        store_op4<KIND>(a.out, a.out_plane, a.out_planes, (long long)r * a.out_ld + a.out_off + c, v);
    }
    pv2::pdl_done();
}

// Lean form of the common case -- one source, one slab, no multiplier, NHWC out: thread = (channel quad, row lane) like the lean
// BatchNorm-backward kernels, so the quad's scale / shift live in registers and the loop has no division (the general kernel above
// divides by C/4 for every 16 bytes and re-reads the affine from shared memory); 4 rows in flight per thread.
template <int KIND>
__global__ void __launch_bounds__(256)
act_apply4_lean_kernel(const ApplyArgs a0, const Defer2 df, int RP) {
    pv2::pdl_prologue();
    PV2_APPLY_DEFER_PROLOGUE(a0, df, a)
    const int C4 = a.C >> 2;
    const int quad = threadIdx.x % C4, rp = threadIdx.x / C4;
    if (rp < RP) {
        const int c = quad << 2;
        const float4 sc = f4_ldp(a.s1 + c), sh = f4_ldp(a.b1 + c);
        const float* src = a.y1 + a.off1 + c;
        const long long step = (long long)gridDim.x * RP;
        for (long long r = (long long)blockIdx.x * RP + rp; r < a.M; r += 4 * step) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + u * step;
                v[u] = f4_ld(src + (ru < a.M ? ru : r) * a.ld1);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + u * step;
                if (ru >= a.M) break;
                float4 o = f4_fma(v[u], sc, sh);
                if (a.relu) o = make_float4(fmaxf(o.x, 0.0f), fmaxf(o.y, 0.0f), fmaxf(o.z, 0.0f), fmaxf(o.w, 0.0f));
                store_op4<KIND>(a.out, a.out_plane, a.out_planes, ru * a.out_ld + a.out_off + c, o);
            }
        }
    }
    pv2::pdl_done();
}

struct Da4 { float4 da1, da2, yh1, yh2, dm; };

template <int KIND>
__device__ __forceinline__ Da4 bwd_da4(const BwdArgs& b, long long r, int c) {
    const ApplyArgs& a = b.f;
    float4 a1, a2, m;
    const float4 v = apply_value4<KIND>(a, r, c, &a1, &a2, &m);
    float4 g = slab_sum4(b.dz, r, c);
    if (a.relu) {
        if (!(v.x > 0.0f)) g.x = 0.0f;
        if (!(v.y > 0.0f)) g.y = 0.0f;
        if (!(v.z > 0.0f)) g.z = 0.0f;
        if (!(v.w > 0.0f)) g.w = 0.0f;
    }
    float4 comb = a1;
    if (a.combine == 1) comb = f4_add(a1, a2); else if (a.combine == 2) comb = f4_mul(a1, a2);
    Da4 o;
    o.dm = f4_mul(g, comb);
    const float4 dc = a.mult ? f4_mul(g, m) : g;
    o.da1 = a.combine == 2 ? f4_mul(dc, a2) : dc;
    o.da2 = a.combine == 0 ? f4_set(0.0f) : (a.combine == 2 ? f4_mul(dc, a1) : dc);
    const float4 y1v = raw_load4(a.y1, a.ld1, a.off1, a.ns1, a.ss1, r, c);
    o.yh1 = b.mean1 ? f4_mul(f4_sub(y1v, f4_ld(b.mean1 + c)), f4_ld(b.inv1 + c)) : y1v;
    const float4 y2v = a.combine ? raw_load4(a.y2, a.ld2, a.off2, a.ns2, a.ss2, r, c) : f4_set(0.0f);
    o.yh2 = (a.combine && b.mean2) ? f4_mul(f4_sub(y2v, f4_ld(b.mean2 + c)), f4_ld(b.inv2 + c)) : y2v;
    return o;
}

// reduce pass, vector form.  block = RP rows x C4 channel quads; each block folds its rows in shared memory and ADDS its
// 4 x C block sums to sums[4][C] = (sum da1, sum da1*yhat1, sum da2, sum da2*yhat2) -- which are also dbeta / dgamma -- with fp32
// reductions in L2 (sums must be zero on entry).  No ticket, no fence, no serial fold: the dx pass reads the finished sums.
// (The <= 296 block sums per value are added in arrival order: results agree to fp32 rounding, not bit for bit.)
struct Reduce4Plan { int nblk, rows_pb, RP; float* dg1; float* db1; float* dg2; float* db2; float* sums_out; };

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_reduce4_kernel(const BwdArgs b, const Reduce4Plan pl) {
    pv2::pdl_prologue();
    __shared__ float sh[256 * 17];     // pitch 17: conflict-free writes (thread-major) and reads (value-major)
    const int C = b.f.C, C4 = C >> 2;
    const int tid = threadIdx.x;
    const int quad = tid % C4, rp = tid / C4;
    const long long r0 = (long long)blockIdx.x * pl.rows_pb, r1 = min(b.f.M, r0 + pl.rows_pb);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
    auto add = [&](const Da4& d) {
        acc[0] += d.da1.x; acc[1] += d.da1.y; acc[2] += d.da1.z; acc[3] += d.da1.w;
        acc[4] = fmaf(d.da1.x, d.yh1.x, acc[4]); acc[5] = fmaf(d.da1.y, d.yh1.y, acc[5]);
        acc[6] = fmaf(d.da1.z, d.yh1.z, acc[6]); acc[7] = fmaf(d.da1.w, d.yh1.w, acc[7]);
        acc[8] += d.da2.x; acc[9] += d.da2.y; acc[10] += d.da2.z; acc[11] += d.da2.w;
        acc[12] = fmaf(d.da2.x, d.yh2.x, acc[12]); acc[13] = fmaf(d.da2.y, d.yh2.y, acc[13]);
        acc[14] = fmaf(d.da2.z, d.yh2.z, acc[14]); acc[15] = fmaf(d.da2.w, d.yh2.w, acc[15]);
    };
    if (rp < pl.RP) {
        const int c = quad << 2;
        // two rows per trip: their loads are independent, and the d(mult) stores no longer sit between one row's loads and the next's
        for (long long r = r0 + rp; r < r1; r += 2 * pl.RP) {
            const long long rb = r + pl.RP;
            const bool two = rb < r1;
            const Da4 d0 = bwd_da4<KIND>(b, r, c);
            const Da4 d1 = bwd_da4<KIND>(b, two ? rb : r, c);
            add(d0);
            if (b.dmult) *reinterpret_cast<float4*>(b.dmult + r * b.dmult_ld + c) = d0.dm;
            if (two) {
                add(d1);
                if (b.dmult) *reinterpret_cast<float4*>(b.dmult + rb * b.dmult_ld + c) = d1.dm;
            }
        }
    }
    pv2::pdl_done();
#pragma unroll
    for (int i = 0; i < 16; ++i) sh[tid * 17 + i] = acc[i];
    __syncthreads();
    const int nk = b.f.combine ? 16 : 8;          // sums 2, 3 belong to the second source
    // (quad, k) -> sum over the RP row lanes in order; k = 4*which_sum + channel
    for (int idx = tid; idx < C4 * 16; idx += 256) {
        const int qd = idx >> 4, k = idx & 15;
        if (k >= nk) continue;
        float t = 0.0f;
        for (int j = 0; j < pl.RP; ++j) t += sh[(j * C4 + qd) * 17 + k];
        atomicAdd(pl.sums_out + (size_t)PV2_SUM_STRIDE * ((k >> 2) * C + (qd << 2) + (k & 3)), t);     // one 128-byte line per entry
    }
}

// ------------------------------------------------------------------------------------------------------
// Lean forms of the two passes for the common layer of the head -- one BatchNorm'd source, optional ReLU, no product / second
// source, the incoming gradient in ONE slab (the persistent conv sums its split-K partials itself): ~40 registers instead of
// 128, so 6+ CTAs per SM, and four rows of independent 16-byte loads in flight per thread.  Same arithmetic as the general
// kernels above (which stay for the combined / multiplied / multi-slab cases).
// ------------------------------------------------------------------------------------------------------
struct LeanBwd {
    const float* y; int ldy, offy;            // raw conv output rows
    const float* dz; int ldz, offz;           // gradient w.r.t. the activation, raw rows
    const float* scale; const float* shift; const float* mean; const float* inv;
    int relu; long long M; int C;
    float* sums;                              // [2*C] entries PV2_SUM_STRIDE floats apart: sum da, sum da*yhat
    void* dy; long long dy_plane; int dy_planes, dy_ld;
    float* dgamma; float* dbeta;
    int rows_pb, RP;
};

__device__ __forceinline__ void lean_da(const LeanBwd& a, long long r, int c, const float4& sc, const float4& sh, const float4& mu, const float4& iv,
                                        float4* da, float4* yh) {
    const float4 y = f4_ld(a.y + r * a.ldy + a.offy + c);
    float4 g = f4_ld(a.dz + r * a.ldz + a.offz + c);
    if (a.relu) {
        if (!(fmaf(y.x, sc.x, sh.x) > 0.0f)) g.x = 0.0f;
        if (!(fmaf(y.y, sc.y, sh.y) > 0.0f)) g.y = 0.0f;
        if (!(fmaf(y.z, sc.z, sh.z) > 0.0f)) g.z = 0.0f;
        if (!(fmaf(y.w, sc.w, sh.w) > 0.0f)) g.w = 0.0f;
    }
    *da = g;
    *yh = f4_mul(f4_sub(y, mu), iv);
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce4_lean_kernel(const LeanBwd a) {
    pv2::pdl_prologue();
    __shared__ float sh[256 * 9];
    const int C4 = a.C >> 2, tid = threadIdx.x;
    const int quad = tid % C4, rp = tid / C4;
    const long long r0 = (long long)blockIdx.x * a.rows_pb, r1 = min(a.M, r0 + a.rows_pb);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (rp < a.RP) {
        const int c = quad << 2;
        const float4 sc = f4_ld(a.scale + c), shf = f4_ld(a.shift + c), mu = f4_ld(a.mean + c), iv = f4_ld(a.inv + c);
        for (long long r = r0 + rp; r < r1; r += 4LL * a.RP) {
            float4 da[4], yh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + (long long)u * a.RP;
                lean_da(a, ru < r1 ? ru : r, c, sc, shf, mu, iv, &da[u], &yh[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r + (long long)u * a.RP >= r1) break;
                acc[0] += da[u].x; acc[1] += da[u].y; acc[2] += da[u].z; acc[3] += da[u].w;
                acc[4] = fmaf(da[u].x, yh[u].x, acc[4]); acc[5] = fmaf(da[u].y, yh[u].y, acc[5]);
                acc[6] = fmaf(da[u].z, yh[u].z, acc[6]); acc[7] = fmaf(da[u].w, yh[u].w, acc[7]);
            }
        }
    }
    pv2::pdl_done();
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[tid * 9 + i] = acc[i];
    __syncthreads();
    for (int idx = tid; idx < C4 * 8; idx += 256) {
        const int qd = idx >> 3, k = idx & 7;
        float t = 0.0f;
        for (int j = 0; j < a.RP; ++j) t += sh[(j * C4 + qd) * 9 + k];
        atomicAdd(a.sums + (size_t)PV2_SUM_STRIDE * ((k >> 2) * a.C + (qd << 2) + (k & 3)), t);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_dx4_lean_kernel(const LeanBwd a) {
    pv2::pdl_prologue();
    // thread = (channel quad, row lane), like the reduce pass: the quad's affine, statistics and finished sums live in registers and
    // the row loop has no division (the element-indexed form divided by C/4 for every 16 bytes); 4 rows in flight per thread
    const int C4 = a.C >> 2, tid = threadIdx.x;
    const int quad = tid % C4, rp = tid / C4, c = quad << 2;
    if (blockIdx.x == 0) {
        for (int i = tid; i < a.C; i += 256) {
            if (a.dbeta) a.dbeta[i] = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * i);
            if (a.dgamma) a.dgamma[i] = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (a.C + i));
        }
    }
    if (rp < a.RP) {
        const float invn = 1.0f / (float)a.M;
        const float4 sc = f4_ld(a.scale + c), shf = f4_ld(a.shift + c), mu = f4_ld(a.mean + c), iv = f4_ld(a.inv + c);
        float4 k1, k2;
        k1.x = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (c + 0)) * invn; k1.y = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (c + 1)) * invn;
        k1.z = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (c + 2)) * invn; k1.w = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (c + 3)) * invn;
        k2.x = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (a.C + c + 0)) * invn; k2.y = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (a.C + c + 1)) * invn;
        k2.z = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (a.C + c + 2)) * invn; k2.w = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * (a.C + c + 3)) * invn;
        const long long step = (long long)gridDim.x * a.RP;
        for (long long r = (long long)blockIdx.x * a.RP + rp; r < a.M; r += 4 * step) {
            float4 da[4], yh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + u * step;
                lean_da(a, ru < a.M ? ru : r, c, sc, shf, mu, iv, &da[u], &yh[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + u * step;
                if (ru >= a.M) break;
                const float4 d1 = make_float4(sc.x * (da[u].x - k1.x - yh[u].x * k2.x), sc.y * (da[u].y - k1.y - yh[u].y * k2.y),
                                              sc.z * (da[u].z - k1.z - yh[u].z * k2.z), sc.w * (da[u].w - k1.w - yh[u].w * k2.w));
                store_op4<KIND>(a.dy, a.dy_plane, a.dy_planes, ru * a.dy_ld + c, d1);
            }
        }
    }
    pv2::pdl_done();
}

// Reduce + dx of the lean case in ONE launch behind a grid-wide barrier (the sums are global: a CTA needs every other CTA's
// contribution before it can write dx).  Safe only because the grid is bounded by the host to a number of CTAs that are certainly
// co-resident (pv2_bn_set_fused_grid: the caller knows how many such launches run side by side and divides the machine) -- a
// spinning CTA holds its slot, and two barrier kernels that each hold part of the slots and wait for the rest would deadlock.
// `launch_dependents` is issued after the barrier only, so a programmatically launched successor can never take slots this
// kernel still needs.  A CTA walks the same rows in both phases (its second read of y / dz comes from L1 / L2).  The counter lives
// in the padding of the zero-initialised sums (float 16 of line 0), so it is zero at launch and nobody has to reset it.
template <int KIND>
__global__ void __launch_bounds__(256, 4)
bn_bwd_fused_lean_kernel(const LeanBwd a) {
    pv2::pdl_prologue();
    __shared__ float sh[256 * 9];
    __shared__ __align__(16) float s_sums[2 * 256];
    const int C4 = a.C >> 2, tid = threadIdx.x;
    const int quad = tid % C4, rp = tid / C4;
    const int c = quad << 2;
    const long long r0 = (long long)blockIdx.x * a.rows_pb, r1 = min(a.M, r0 + a.rows_pb);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    float4 sc = f4_set(0.0f), shf = sc, mu = sc, iv = sc;
    if (rp < a.RP) {
        sc = f4_ld(a.scale + c); shf = f4_ld(a.shift + c); mu = f4_ld(a.mean + c); iv = f4_ld(a.inv + c);
        for (long long r = r0 + rp; r < r1; r += 4LL * a.RP) {
            float4 da[4], yh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + (long long)u * a.RP;
                lean_da(a, ru < r1 ? ru : r, c, sc, shf, mu, iv, &da[u], &yh[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r + (long long)u * a.RP >= r1) break;
                acc[0] += da[u].x; acc[1] += da[u].y; acc[2] += da[u].z; acc[3] += da[u].w;
                acc[4] = fmaf(da[u].x, yh[u].x, acc[4]); acc[5] = fmaf(da[u].y, yh[u].y, acc[5]);
                acc[6] = fmaf(da[u].z, yh[u].z, acc[6]); acc[7] = fmaf(da[u].w, yh[u].w, acc[7]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[tid * 9 + i] = acc[i];
    __syncthreads();
    for (int idx = tid; idx < C4 * 8; idx += 256) {
        const int qd = idx >> 3, k = idx & 7;
        float t = 0.0f;
        for (int j = 0; j < a.RP; ++j) t += sh[(j * C4 + qd) * 9 + k];
        atomicAdd(a.sums + (size_t)PV2_SUM_STRIDE * ((k >> 2) * a.C + (qd << 2) + (k & 3)), t);
    }
    // ---- grid barrier ----
    __syncthreads();
    if (tid == 0) {
        unsigned int* ctr = reinterpret_cast<unsigned int*>(a.sums) + 16;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            if (seen < gridDim.x) __nanosleep(40);
        } while (seen < gridDim.x);
        __threadfence();
    }
    __syncthreads();
    pv2::pdl_done();
    for (int i = tid; i < 2 * a.C; i += 256) s_sums[i] = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * i);
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = tid; i < a.C; i += 256) {
            if (a.dbeta) a.dbeta[i] = s_sums[i];
            if (a.dgamma) a.dgamma[i] = s_sums[a.C + i];
        }
    }
    if (rp < a.RP) {
        const float invn = 1.0f / (float)a.M;
        const float4 S1 = f4_ldp(s_sums + c), S2 = f4_ldp(s_sums + a.C + c);
        const float4 k1 = make_float4(S1.x * invn, S1.y * invn, S1.z * invn, S1.w * invn);
        const float4 k2 = make_float4(S2.x * invn, S2.y * invn, S2.z * invn, S2.w * invn);
        for (long long r = r0 + rp; r < r1; r += 4LL * a.RP) {
            float4 da[4], yh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + (long long)u * a.RP;
                lean_da(a, ru < r1 ? ru : r, c, sc, shf, mu, iv, &da[u], &yh[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + (long long)u * a.RP;
                if (ru >= r1) break;
                const float4 d1 = make_float4(sc.x * (da[u].x - k1.x - yh[u].x * k2.x), sc.y * (da[u].y - k1.y - yh[u].y * k2.y),
                                              sc.z * (da[u].z - k1.z - yh[u].z * k2.z), sc.w * (da[u].w - k1.w - yh[u].w * k2.w));
                store_op4<KIND>(a.dy, a.dy_plane, a.dy_planes, ru * a.dy_ld + c, d1);
            }
        }
    }
}

int g_bn_fused_grid = 0;       // CTAs a fused (barrier) BN-backward launch may use; 0 = two launches (see pv2_bn_set_fused_grid)

// Small-C forms (the fg / bg logit heads: C = 1 .. 8 maps, gradient arriving as an fp32 NCHW tensor): thread = pixel row, all C
// channels in registers; the block's 2*C sums go to the strided accumulators with one reduction each.  The channel-major
// general kernel gives such a layer 8 of its 256 threads per block something to do.
constexpr int SMALL_C = 8;
struct SmallBwd {
    const float* y; int ldy, offy;
    const float* dz_nchw; int HW;
    const float* scale; const float* shift; const float* mean; const float* inv;
    int relu; long long M; int C;
    float* sums;
    void* dy; long long dy_plane; int dy_planes, dy_ld;
    float* dgamma; float* dbeta;
};

__global__ void __launch_bounds__(256)
bn_bwd_reduce_small_kernel(const SmallBwd a) {
    pv2::pdl_prologue();
    __shared__ float red[8][2 * SMALL_C];
    float acc[2 * SMALL_C];
#pragma unroll
    for (int i = 0; i < 2 * SMALL_C; ++i) acc[i] = 0.0f;
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < a.M; r += (long long)gridDim.x * 256) {
        const long long n = r / a.HW, p = r - n * a.HW;
#pragma unroll
        for (int c = 0; c < SMALL_C; ++c) {
            if (c >= a.C) break;
            const float y = __ldg(a.y + r * a.ldy + a.offy + c);
            float g = __ldg(a.dz_nchw + (n * a.C + c) * a.HW + p);
            if (a.relu && !(fmaf(y, a.scale[c], a.shift[c]) > 0.0f)) g = 0.0f;
            acc[c] += g;
            acc[SMALL_C + c] = fmaf(g, (y - a.mean[c]) * a.inv[c], acc[SMALL_C + c]);
        }
    }
    pv2::pdl_done();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 2 * SMALL_C; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 2 * SMALL_C) {
        const int k = threadIdx.x / SMALL_C, c = threadIdx.x % SMALL_C;
        if (c < a.C) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
            atomicAdd(a.sums + (size_t)PV2_SUM_STRIDE * (k * a.C + c), t);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_dx_small_kernel(const SmallBwd a) {
    pv2::pdl_prologue();
    __shared__ float s_sums[2 * SMALL_C];
    if (threadIdx.x < 2 * a.C) s_sums[threadIdx.x] = __ldcg(a.sums + (size_t)PV2_SUM_STRIDE * threadIdx.x);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < a.C) {
        if (a.dbeta) a.dbeta[threadIdx.x] = s_sums[threadIdx.x];
        if (a.dgamma) a.dgamma[threadIdx.x] = s_sums[a.C + threadIdx.x];
    }
    const float invn = 1.0f / (float)a.M;
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < a.M; r += (long long)gridDim.x * 256) {
        const long long n = r / a.HW, p = r - n * a.HW;
#pragma unroll
        for (int c = 0; c < SMALL_C; ++c) {
            if (c >= a.C) break;
            const float y = __ldg(a.y + r * a.ldy + a.offy + c);
            float g = __ldg(a.dz_nchw + (n * a.C + c) * a.HW + p);
            if (a.relu && !(fmaf(y, a.scale[c], a.shift[c]) > 0.0f)) g = 0.0f;
            const float yh = (y - a.mean[c]) * a.inv[c];
            store_op<KIND>(a.dy, a.dy_plane, a.dy_planes, r * a.dy_ld + c, a.scale[c] * (g - s_sums[c] * invn - yh * s_sums[a.C + c] * invn));
        }
    }
    pv2::pdl_done();
}

template <int KIND>
__global__ void __launch_bounds__(256)
bn_bwd_dx4_kernel(const BwdArgs b, const Reduce4Plan pl) {
    pv2::pdl_prologue();
    const ApplyArgs& a = b.f;
    // the finished sums (one 128-byte line each in global memory) gathered into shared memory once per CTA
    __shared__ __align__(16) float s_sums[4 * 256];
    const int nv = (a.combine ? 4 : 2) * a.C;
    for (int i = threadIdx.x; i < nv; i += 256) s_sums[i] = __ldcg(b.sums + (size_t)PV2_SUM_STRIDE * i);
    __syncthreads();
    if (blockIdx.x == 0) {      // ... they ARE the parameter gradients: dbeta_i = S1_i, dgamma_i = S2_i
        for (int i = threadIdx.x; i < a.C; i += 256) {
            if (pl.db1) pl.db1[i] = s_sums[i];
            if (pl.dg1) pl.dg1[i] = s_sums[a.C + i];
            if (a.combine && pl.db2) pl.db2[i] = s_sums[2 * a.C + i];
            if (a.combine && pl.dg2) pl.dg2[i] = s_sums[3 * a.C + i];
        }
    }
    const unsigned C4 = (unsigned)a.C >> 2;
    const unsigned total = (unsigned)a.M * C4;
    const float invn = 1.0f / (float)a.M;
    for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < total; e += gridDim.x * 256u) {
        const unsigned r = e / C4;
        const int c = (int)(e - r * C4) << 2;
        const Da4 d = bwd_da4<KIND>(b, r, c);
        float4 d1, d2 = f4_set(0.0f);
        const float4 s1 = f4_ld(a.s1 + c);
        if (b.bn_train) {
            const float4 S1 = f4_ldp(s_sums + c), S2 = f4_ldp(s_sums + a.C + c);
            d1 = make_float4(s1.x * (d.da1.x - S1.x * invn - d.yh1.x * S2.x * invn), s1.y * (d.da1.y - S1.y * invn - d.yh1.y * S2.y * invn),
                             s1.z * (d.da1.z - S1.z * invn - d.yh1.z * S2.z * invn), s1.w * (d.da1.w - S1.w * invn - d.yh1.w * S2.w * invn));
            if (a.combine) {
                const float4 s2 = f4_ld(a.s2 + c), T1 = f4_ldp(s_sums + 2 * a.C + c), T2 = f4_ldp(s_sums + 3 * a.C + c);
                d2 = make_float4(s2.x * (d.da2.x - T1.x * invn - d.yh2.x * T2.x * invn), s2.y * (d.da2.y - T1.y * invn - d.yh2.y * T2.y * invn),
                                 s2.z * (d.da2.z - T1.z * invn - d.yh2.z * T2.z * invn), s2.w * (d.da2.w - T1.w * invn - d.yh2.w * T2.w * invn));
            }
        } else {
            d1 = f4_mul(s1, d.da1);
            if (a.combine) d2 = f4_mul(f4_ld(a.s2 + c), d.da2);
        }
        store_op4<KIND>(b.dy1, b.dy1_plane, b.dy1_planes, (long long)r * b.dy1_ld + c, d1);
        if (a.combine) store_op4<KIND>(b.dy2, b.dy2_plane, b.dy2_planes, (long long)r * b.dy2_ld + c, d2);
    }
    pv2::pdl_done();
}

// ------------------------------------------------------------------------------------------------------
// x2 bilinear upsample, align_corners=True, NHWC operand -> NHWC operand slice; backward raw -> raw
// ------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
up2_nhwc_fwd_kernel(const void* in, long long in_plane, int in_planes, int in_ld, int in_off, void* out, long long out_plane,
                    int out_planes, int out_ld, int out_off, int N, int H, int W, int C, float rh, float rw) {
    pv2::pdl_prologue();
    const int OH = 2 * H, OW = 2 * W;
    const long long total = (long long)N * OH * OW * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        long long t = e / C;
        const int ox = (int)(t % OW); t /= OW;
        const int oy = (int)(t % OH);
        const long long n = t / OH;
        const Tap ty = bilinear_tap(oy, H, rh, true), tx = bilinear_tap(ox, W, rw, true);
        const long long b0 = (n * H + ty.i0) * W, b1 = (n * H + ty.i1) * W;
        const float v00 = load_op<KIND>(in, in_plane, in_planes, (b0 + tx.i0) * in_ld + in_off + c);
        const float v01 = load_op<KIND>(in, in_plane, in_planes, (b0 + tx.i1) * in_ld + in_off + c);
        const float v10 = load_op<KIND>(in, in_plane, in_planes, (b1 + tx.i0) * in_ld + in_off + c);
        const float v11 = load_op<KIND>(in, in_plane, in_planes, (b1 + tx.i1) * in_ld + in_off + c);
        const float v = ty.w0 * (tx.w0 * v00 + tx.w1 * v01) + ty.w1 * (tx.w0 * v10 + tx.w1 * v11);
        store_op<KIND>(out, out_plane, out_planes, ((n * OH + oy) * OW + ox) * out_ld + out_off + c, v);
    }
    pv2::pdl_done();
}

__device__ __forceinline__ float ac_weight(int o, int i, int in_size, float ratio) {
    const Tap t = bilinear_tap(o, in_size, ratio, true);
    return (t.i0 == i ? t.w0 : 0.0f) + (t.i1 == i ? t.w1 : 0.0f);
}

__global__ void __launch_bounds__(256)
up2_nhwc_bwd_kernel(const Slabs g, float* __restrict__ din, int din_ld, int N, int H, int W, int C, float rh, float rw) {
    pv2::pdl_prologue();
    const int OH = 2 * H, OW = 2 * W;
    const long long total = (long long)N * H * W * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        long long t = e / C;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const long long n = t / H;
        // outputs whose taps can touch (iy, ix): src = ratio*o in (i-1, i+1)
        const int ylo = max(0, rh > 0.f ? (int)floorf((iy - 1) / rh) : 0), yhi = min(OH - 1, rh > 0.f ? (int)ceilf((iy + 1) / rh) : OH - 1);
        const int xlo = max(0, rw > 0.f ? (int)floorf((ix - 1) / rw) : 0), xhi = min(OW - 1, rw > 0.f ? (int)ceilf((ix + 1) / rw) : OW - 1);
        float acc = 0.0f;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = ac_weight(oy, iy, H, rh);
            if (wy == 0.0f) continue;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = ac_weight(ox, ix, W, rw);
                if (wx != 0.0f) acc += wy * wx * slab_sum(g, (n * OH + oy) * OW + ox, c);
            }
        }
        din[((n * H + iy) * W + ix) * din_ld + c] = acc;
    }
}

// 4-channel vector form: one thread = one INPUT pixel x 4 channels.  The contributing output rows / columns and their tap
// weights are computed once per thread (<= 7 candidates per axis, the same bilinear_tap arithmetic as the forward), then the
// non-zero (wy * wx) taps are gathered with 16-byte loads -- ~16 independent loads per thread instead of a nested scalar
// walk with the weights recomputed per element.  Accumulation order = (oy, ox) ascending, as in the scalar kernel.
__global__ void __launch_bounds__(256)
up2_nhwc_bwd4_kernel(const Slabs g, float* __restrict__ din, int din_ld, int N, int H, int W, int C, float rh, float rw) {
    pv2::pdl_prologue();
    constexpr int NC = 7;
    const int OH = 2 * H, OW = 2 * W, C4 = C >> 2;
    const long long total = (long long)N * H * W * C4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C4) << 2;
        long long t = e / C4;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const long long n = t / H;
        const int ylo = max(0, rh > 0.f ? (int)floorf((iy - 1) / rh) : 0), yhi = min(OH - 1, rh > 0.f ? (int)ceilf((iy + 1) / rh) : OH - 1);
        const int xlo = max(0, rw > 0.f ? (int)floorf((ix - 1) / rw) : 0), xhi = min(OW - 1, rw > 0.f ? (int)ceilf((ix + 1) / rw) : OW - 1);
        float wy[NC], wx[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            wy[k] = (ylo + k <= yhi) ? ac_weight(ylo + k, iy, H, rh) : 0.0f;
            wx[k] = (xlo + k <= xhi) ? ac_weight(xlo + k, ix, W, rw) : 0.0f;
        }
        float4 acc = f4_set(0.0f);
        // candidates beyond NC (ratio < 2/7: never for a x2 upsample of >= 2 pixels) cannot occur.  (Loading all NC column candidates
        // of a contributing row unconditionally -- 7 independent loads per row instead of ~4 conditional ones -- was measured
        // SLOWER, 34 us against 27: the kernel is bound by its L2 -> SM traffic, ~4.5x the gradient's bytes, not by load latency.)
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            if (wy[a] == 0.0f) continue;
#pragma unroll
            for (int b = 0; b < NC; ++b) {
                if (wx[b] == 0.0f) continue;
                const float w = wy[a] * wx[b];
                const float4 v = slab_sum4(g, (n * OH + ylo + a) * OW + xlo + b, c);
                acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
            }
        }
        *reinterpret_cast<float4*>(din + ((n * H + iy) * W + ix) * din_ld + c) = acc;
    }
    pv2::pdl_done();
}

// Tiled form of the x2 backward: CTA = 8 x 8 input pixels x 16 channels of one image.  The window of output pixels the tile touches
// (<= 22 x 22) is summed over the slabs ONCE while it is staged in shared memory -- every gradient element is read ~1.6x per slab
// instead of the ~4.5x of the gather above (which is bound by exactly that L2 -> SM traffic, times the number of slabs) -- with
// all of a thread's loads independent; the <= 5 x 5 taps of an input pixel then come from shared memory.
constexpr int UT = 8, UC = 16, UWIN = 22;
__global__ void __launch_bounds__(256)
up2_nhwc_bwd4_tiled_kernel(const Slabs g, float* __restrict__ din, int din_ld, int H, int W, int C, float rh, float rw, int tiles_x, int tiles_y, int cgroups) {
    pv2::pdl_prologue();
    __shared__ __align__(16) float win[UWIN * UWIN * UC];
    constexpr int NC = 7;
    const int OH = 2 * H, OW = 2 * W;
    int b = blockIdx.x;
    const int cg = b % cgroups; b /= cgroups;
    const int tx = b % tiles_x; b /= tiles_x;
    const int ty = b % tiles_y;
    const long long n = b / tiles_y;
    const int iy0 = ty * UT, ix0 = tx * UT, c0 = cg * UC;
    const int iy1 = min(iy0 + UT, H) - 1, ix1 = min(ix0 + UT, W) - 1;
    auto lo_of = [](int i, float r) { return max(0, r > 0.f ? (int)floorf((i - 1) / r) : 0); };
    auto hi_of = [](int i, float r, int on) { return min(on - 1, r > 0.f ? (int)ceilf((i + 1) / r) : on - 1); };
    const int wy0 = lo_of(iy0, rh), wy1 = hi_of(iy1, rh, OH), wx0 = lo_of(ix0, rw), wx1 = hi_of(ix1, rw, OW);
    const int wh = wy1 - wy0 + 1, ww = wx1 - wx0 + 1;
    const bool staged = wh <= UWIN && ww <= UWIN;          // always, for a x2 up-sampling of >= 2 pixels
    const int lanes = UC / 4;                                // float4 lanes per pixel
    if (staged) {
        for (int i = threadIdx.x; i < wh * ww * lanes; i += 256) {
            const int l = i % lanes, px = i / lanes, wx = px % ww, wy = px / ww;
            const int c = c0 + l * 4;
            float4 v = f4_set(0.0f);
            if (c < C) v = slab_sum4(g, (n * OH + wy0 + wy) * OW + wx0 + wx, c);
            *reinterpret_cast<float4*>(win + (wy * UWIN + wx) * UC + l * 4) = v;
        }
    }
    pv2::pdl_done();
    __syncthreads();
    const int l = threadIdx.x % lanes, pix = threadIdx.x / lanes;      // 64 pixels x 4 lanes
    const int iy = iy0 + pix / UT, ix = ix0 + pix % UT, c = c0 + l * 4;
    if (iy > iy1 || ix > ix1 || c >= C) return;
    const int ylo = lo_of(iy, rh), yhi = hi_of(iy, rh, OH), xlo = lo_of(ix, rw), xhi = hi_of(ix, rw, OW);
    float wyv[NC], wxv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        wyv[k] = (ylo + k <= yhi) ? ac_weight(ylo + k, iy, H, rh) : 0.0f;
        wxv[k] = (xlo + k <= xhi) ? ac_weight(xlo + k, ix, W, rw) : 0.0f;
    }
    float4 acc = f4_set(0.0f);
#pragma unroll
    for (int a = 0; a < NC; ++a) {
        if (wyv[a] == 0.0f) continue;
#pragma unroll
        for (int bb = 0; bb < NC; ++bb) {
            if (wxv[bb] == 0.0f) continue;
            const float w = wyv[a] * wxv[bb];
            const float4 v = staged ? *reinterpret_cast<const float4*>(win + ((ylo + a - wy0) * UWIN + (xlo + bb - wx0)) * UC + l * 4)
                                    : slab_sum4(g, (n * OH + ylo + a) * OW + xlo + bb, c);
            acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
        }
    }
    *reinterpret_cast<float4*>(din + ((n * H + iy) * W + ix) * din_ld + c) = acc;
}

inline int grid_for(long long total, int threads = 256) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int fill_slabs(Slabs* s, const float* const* ptrs, const int* lds, const int* offs, int n, const char* who) {
    PV2_CHECK(n >= 1 && n <= 8, "%s: between 1 and 8 gradient slabs expected, got %d", who, n);
    s->n = n;
    for (int i = 0; i < 8; ++i) { s->p[i] = nullptr; s->ld[i] = 0; s->off[i] = 0; }
    for (int i = 0; i < n; ++i) {
        PV2_CHECK(ptrs[i] != nullptr, "%s: null slab %d", who, i);
        s->p[i] = ptrs[i]; s->ld[i] = lds[i]; s->off[i] = offs[i];
    }
    return 0;
}

}  // namespace
}  // namespace pv2

using namespace pv2;

#define KIND_CHECK(who)                                                                                              \
    PV2_CHECK(kind == PV2_BF16 || kind == PV2_TF32, who ": operand kind must be PV2_BF16 or PV2_TF32 (got %d)", kind); \
    PV2_CHECK(nplanes == 1 || (kind == PV2_TF32 && nplanes == 2), who ": nplanes must be 1 (or 2 with tf32)")

extern "C" int pv2_weight_pack(const float* w, void* out, long long plane_stride, int nplanes, int kind, int Cout, int Cin, int KH, int KW,
                               int mode, int i_ld, int i_off, int o_off, void* stream) {
    KIND_CHECK("weight_pack");
    PV2_CHECK(w && out && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "weight_pack: bad arguments");
    const long long total = (long long)Cout * Cin * KH * KW;
    if (kind == PV2_BF16) pv2::launch(weight_pack_kernel<0>, grid_for(total), 256, 0, (cudaStream_t)stream, w, out, plane_stride, nplanes, Cout, Cin, KH, KW, mode, i_ld, i_off, o_off);
    else pv2::launch(weight_pack_kernel<1>, grid_for(total), 256, 0, (cudaStream_t)stream, w, out, plane_stride, nplanes, Cout, Cin, KH, KW, mode, i_ld, i_off, o_off);
    PV2_LAUNCH_CHECK("weight_pack");
    return 0;
}

extern "C" int pv2_wgrad_unpack(const float* part, long long split_stride, int splits, float* dw, int Cout, int Cin, int KH, int KW,
                                int Cin_p, int co_off, void* stream) {
    PV2_CHECK(part && dw && splits >= 1, "wgrad_unpack: bad arguments");
    const long long total = (long long)Cout * Cin * KH * KW;
    pv2::launch(wgrad_unpack_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, part, split_stride, splits, dw, Cout, Cin, KH, KW, Cin_p, co_off);
    PV2_LAUNCH_CHECK("wgrad_unpack");
    return 0;
}

extern "C" int pv2_weight_pack_multi(const pv2_pack_desc* descs, int n, int nplanes, int kind, void* stream) {
    KIND_CHECK("weight_pack_multi");
    PV2_CHECK(descs && n > 0, "weight_pack_multi: bad arguments");
    for (int i0 = 0; i0 < n; i0 += PACK_BATCH) {
        const int nb = n - i0 < PACK_BATCH ? n - i0 : PACK_BATCH;
        PackBatch b = {};
        int tiles = 0;
        size_t smem = 0;
        for (int i = 0; i < nb; ++i) {
            b.d[i] = descs[i0 + i];
            const pv2_pack_desc& t = b.d[i];
            PV2_CHECK(t.w && t.out_f, "weight_pack_multi: null pointer in descriptor %d", i0 + i);
            PV2_CHECK(t.Cout > 0 && t.Cin > 0 && t.KH > 0 && t.KW > 0 && t.KH * t.KW <= 64, "weight_pack_multi: bad shape in descriptor %d", i0 + i);
            b.tile_start[i] = tiles;
            const int taps = t.KH * t.KW, tco = pk_tco(taps);
            tiles += ((t.Cout + tco - 1) / tco) * ((t.Cin + PK_TCI - 1) / PK_TCI);
            const size_t need = (size_t)tco * (PK_TCI * taps + 1) * sizeof(float);
            if (need > smem) smem = need;
        }
        for (int i = nb; i <= PACK_BATCH; ++i) b.tile_start[i] = tiles;
        PV2_CHECK(smem <= 200 * 1024, "weight_pack_multi: tile needs %zu B of shared memory", smem);
        cudaError_t ce = kind == PV2_BF16 ? cudaFuncSetAttribute(weight_pack_multi_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                          : cudaFuncSetAttribute(weight_pack_multi_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "weight_pack_multi: smem attribute: %s", cudaGetErrorString(ce));
        if (kind == PV2_BF16) pv2::launch(weight_pack_multi_kernel<0>, dim3(tiles), 256, smem, (cudaStream_t)stream, b, nb, nplanes);
        else pv2::launch(weight_pack_multi_kernel<1>, dim3(tiles), 256, smem, (cudaStream_t)stream, b, nb, nplanes);
        PV2_LAUNCH_CHECK("weight_pack_multi");
    }
    return 0;
}

extern "C" int pv2_wgrad_unpack_multi(const pv2_unpack_desc* descs, int n, void* stream) {
    PV2_CHECK(descs && n > 0, "wgrad_unpack_multi: bad arguments");
    for (int i0 = 0; i0 < n; i0 += UNPACK_BATCH) {
        const int nb = n - i0 < UNPACK_BATCH ? n - i0 : UNPACK_BATCH;
        UnpackBatch b = {};
        int tiles = 0;
        size_t smem = 0;
        for (int i = 0; i < nb; ++i) {
            b.d[i] = descs[i0 + i];
            const pv2_unpack_desc& t = b.d[i];
            PV2_CHECK(t.part && t.dw, "wgrad_unpack_multi: null pointer in descriptor %d", i0 + i);
            PV2_CHECK(t.Cout > 0 && t.Cin > 0 && t.KH > 0 && t.KW > 0 && t.KH * t.KW <= 64 && t.splits >= 1, "wgrad_unpack_multi: bad descriptor %d", i0 + i);
            b.tile_start[i] = tiles;
            const int taps = t.KH * t.KW, tco = up_tco(taps);
            tiles += ((t.Cout + tco - 1) / tco) * ((t.Cin + UP_TCI - 1) / UP_TCI);
            const size_t need = (size_t)tco * taps * (UP_TCI + 1) * sizeof(float);
            if (need > smem) smem = need;
        }
        for (int i = nb; i <= UNPACK_BATCH; ++i) b.tile_start[i] = tiles;
        cudaError_t ce = cudaFuncSetAttribute(wgrad_unpack_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "wgrad_unpack_multi: smem attribute: %s", cudaGetErrorString(ce));
        pv2::launch(wgrad_unpack_multi_kernel, dim3(tiles), 256, smem, (cudaStream_t)stream, b, nb);
        PV2_LAUNCH_CHECK("wgrad_unpack_multi");
    }
    return 0;
}

extern "C" int pv2_pack_nchw(const void* x, int x_dtype, void* out, long long plane_stride, int nplanes, int kind, int N, int C, int HW,
                             int ld, int c_off, void* stream) {
    KIND_CHECK("pack_nchw");
    PV2_CHECK(x && out && N > 0 && C > 0 && HW > 0 && N <= 65535, "pack_nchw: bad arguments");
    PV2_CHECK(x_dtype == PV2_F32 || x_dtype == PV2_BF16, "pack_nchw: bad input dtype %d", x_dtype);
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == PV2_F32 && kind == PV2_BF16) pv2::launch(pack_nchw_kernel<float, 0>, grid, 256, 0, st, (const float*)x, out, plane_stride, nplanes, C, HW, ld, c_off);
    else if (x_dtype == PV2_F32) pv2::launch(pack_nchw_kernel<float, 1>, grid, 256, 0, st, (const float*)x, out, plane_stride, nplanes, C, HW, ld, c_off);
    else if (kind == PV2_BF16) pv2::launch(pack_nchw_kernel<__nv_bfloat16, 0>, grid, 256, 0, st, (const __nv_bfloat16*)x, out, plane_stride, nplanes, C, HW, ld, c_off);
    else pv2::launch(pack_nchw_kernel<__nv_bfloat16, 1>, grid, 256, 0, st, (const __nv_bfloat16*)x, out, plane_stride, nplanes, C, HW, ld, c_off);
    PV2_LAUNCH_CHECK("pack_nchw");
    return 0;
}

template <typename TOUT>
__global__ void __launch_bounds__(256)
slabs_to_nhwc_kernel(const pv2::Slabs g, TOUT* __restrict__ dx, long long M, int C) {
    pv2::pdl_prologue();
    const long long total = M * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / C;
        dx[e] = pv2::from_f<TOUT>(pv2::slab_sum(g, r, (int)(e - r * C)));
    }
}

extern "C" int pv2_unpack_to_nchw(const float* const* slabs, const int* lds, const int* offs, int nslabs, void* dx, int dx_dtype,
                                  int N, int C, int HW, int channels_last, void* stream) {
    if (channels_last) {
        Slabs s2;
        if (int e = fill_slabs(&s2, slabs, lds, offs, nslabs, "unpack_to_nchw")) return e;
        PV2_CHECK(dx && N > 0 && C > 0 && HW > 0, "unpack_to_nchw: bad arguments");
        const long long M = (long long)N * HW;
        if (dx_dtype == PV2_F32) pv2::launch(slabs_to_nhwc_kernel<float>, grid_for(M * C), 256, 0, (cudaStream_t)stream, s2, (float*)dx, M, C);
        else if (dx_dtype == PV2_BF16) pv2::launch(slabs_to_nhwc_kernel<__nv_bfloat16>, grid_for(M * C), 256, 0, (cudaStream_t)stream, s2, (__nv_bfloat16*)dx, M, C);
        else PV2_CHECK(false, "unpack_to_nchw: bad dtype %d", dx_dtype);
        PV2_LAUNCH_CHECK("slabs_to_nhwc");
        return 0;
    }
    Slabs s;
    if (int e = fill_slabs(&s, slabs, lds, offs, nslabs, "unpack_to_nchw")) return e;
    PV2_CHECK(dx && N > 0 && C > 0 && HW > 0 && N <= 65535, "unpack_to_nchw: bad arguments");
    PV2_CHECK(dx_dtype == PV2_F32 || dx_dtype == PV2_BF16, "unpack_to_nchw: bad dtype %d", dx_dtype);
    dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
    if (dx_dtype == PV2_F32) pv2::launch(unpack_to_nchw_kernel<float>, grid, 256, 0, (cudaStream_t)stream, s, (float*)dx, C, HW);
    else pv2::launch(unpack_to_nchw_kernel<__nv_bfloat16>, grid, 256, 0, (cudaStream_t)stream, s, (__nv_bfloat16*)dx, C, HW);
    PV2_LAUNCH_CHECK("unpack_to_nchw");
    return 0;
}

extern "C" int pv2_bn_set_fused_grid(int max_ctas) {
    const int old = g_bn_fused_grid;
    g_bn_fused_grid = max_ctas > 0 ? (max_ctas > kNumSMs ? kNumSMs : max_ctas) : 0;
    return old;
}

extern "C" size_t pv2_bn_workspace_floats(long long M, int C) {
    const int rows = pick_rows(M, C);
    const long long rb = (M + rows - 1) / rows;
    const long long scalar = rb * C * 4 + 4 * (long long)C;
    // vector backward: <= 2*148 row blocks + their sqrt-sized groups + the folded sums
    const long long nblk = 2 * kNumSMs;
    const long long vec = (nblk + make_fold_plan((int)nblk).ngroups + 1) * 4 * (long long)C;
    return (size_t)(scalar > vec ? scalar : vec);
}

extern "C" int pv2_bn_stats(float* y, long long slab_stride, int nslabs, long long M, int C, int ld, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                            float* mean_out, float* invstd_out, float* scale, float* shift, float* workspace, void* stream) {
    PV2_CHECK(y && mean_out && invstd_out && scale && shift && workspace, "bn_stats: null pointer");
    PV2_CHECK(M > 0 && C > 0 && ld >= C && nslabs >= 1, "bn_stats: bad shape");
    const int rows = pick_rows(M, C);
    const int rb = (int)((M + rows - 1) / rows);
    PV2_CHECK((C + 31) / 32 <= 65535, "bn_stats: too many channels");
    cudaStream_t st = (cudaStream_t)stream;
    pv2::launch(bn_stats_partial_kernel, dim3(rb, (C + 31) / 32), 256, 0, st, y, slab_stride, nslabs, M, C, ld, rows, workspace);
    PV2_LAUNCH_CHECK("bn_stats_partial");
    pv2::launch(bn_stats_finalize_kernel, (C + 3) / 4, 128, 0, st, workspace, rb, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                               num_batches_tracked, mean_out, invstd_out, scale, shift);
    PV2_LAUNCH_CHECK("bn_stats_finalize");
    return 0;
}

extern "C" size_t pv2_bn_fuse_workspace_floats(long long M, int Cout) {
    const long long m_tiles = (M + 127) / 128;
    const FoldPlan fp = make_fold_plan((int)m_tiles);
    const long long fused = m_tiles * Cout * 2 + (long long)fp.ngroups * Cout * 3;       // conv epilogue: tile + group partials
    const int rows = pick_rows(M, Cout);
    const long long standalone = ((M + rows - 1) / rows) * Cout * 3;                       // pv2_bn_stats_group row-block partials
    return (size_t)(fused > standalone ? fused : standalone);
}

extern "C" int pv2_bn_stats_group(float* y, long long slab_stride, int nslabs, long long M, int Cout, int ld, const pv2_bn_fuse* bn, void* stream) {
    PV2_CHECK(y && bn && bn->nsegs > 0 && bn->nsegs <= PV2_MAX_BN_SEGS, "bn_stats_group: bad arguments");
    PV2_CHECK(bn->mean && bn->invstd && bn->scale && bn->shift && bn->part, "bn_stats_group: incomplete pv2_bn_fuse descriptor");
    PV2_CHECK(M > 0 && Cout > 0 && ld >= Cout && nslabs >= 1, "bn_stats_group: bad shape");
    const int rows = pick_rows(M, Cout);
    const int rb = (int)((M + rows - 1) / rows);
    cudaStream_t st = (cudaStream_t)stream;
    pv2::launch(bn_stats_partial_kernel, dim3(rb, (Cout + 31) / 32), 256, 0, st, y, slab_stride, nslabs, M, Cout, ld, rows, bn->part);
    PV2_LAUNCH_CHECK("bn_stats_partial");
    BnFuseDev d;
    d.f = *bn;
    pv2::launch(bn_group_finalize_kernel, (Cout + 3) / 4, 128, 0, st, (const float*)bn->part, rb, Cout, d);
    PV2_LAUNCH_CHECK("bn_group_finalize");
    return 0;
}

extern "C" int pv2_bn_eval_affine(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                  float* scale, float* shift, void* stream) {
    PV2_CHECK(C > 0 && rm && rv && scale && shift, "bn_eval_affine: bad arguments");
    pv2::launch(bn_eval_affine_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, C, gamma, beta, rm, rv, eps, scale, shift);
    PV2_LAUNCH_CHECK("bn_eval_affine");
    return 0;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// the 4-channel vector kernels need every per-channel array and every row start on a 16-byte (bf16 operand: 8-byte) boundary
static bool apply_vec_ok(const pv2::ApplyArgs& a) {
    if (a.C % 4 != 0 || a.M * (long long)a.C >= (1LL << 31)) return false;
    if (a.ld1 % 4 != 0 || a.off1 % 4 != 0 || a.ss1 % 4 != 0 || !al16(a.y1) || !al16(a.s1) || !al16(a.b1)) return false;
    if (a.combine && (a.ld2 % 4 != 0 || a.off2 % 4 != 0 || a.ss2 % 4 != 0 || !al16(a.y2) || !al16(a.s2) || !al16(a.b2))) return false;
    if (a.mult && (a.mult_ld % 4 != 0 || a.mult_off % 4 != 0 || a.mult_plane % 4 != 0 || !al16(a.mult))) return false;
    return true;
}

// flat argument list -> ApplyArgs (ctypes-friendly)
static int fill_apply(ApplyArgs* a, const float* y1, int ld1, int off1, int ns1, long long ss1, const float* s1, const float* b1,
                      const float* y2, int ld2, int off2, int ns2, long long ss2, const float* s2, const float* b2, int combine, const void* mult, long long mult_plane, int mult_planes, int mult_ld,
                      int mult_off, int relu, long long M, int C, int HW) {
    PV2_CHECK(y1 && s1 && b1 && M > 0 && C > 0 && HW > 0, "apply: bad arguments");
    PV2_CHECK(combine >= 0 && combine <= 2 && (combine == 0 || (y2 && s2 && b2)), "apply: bad combine arguments");
    a->y1 = y1; a->ld1 = ld1; a->off1 = off1; a->s1 = s1; a->b1 = b1;
    a->y2 = y2; a->ld2 = ld2; a->off2 = off2; a->s2 = s2; a->b2 = b2;
    a->ns1 = ns1 < 1 ? 1 : ns1; a->ss1 = ss1; a->ns2 = ns2 < 1 ? 1 : ns2; a->ss2 = ss2;
    a->combine = combine;
    a->mult = mult; a->mult_plane = mult_plane; a->mult_planes = mult_planes; a->mult_ld = mult_ld; a->mult_off = mult_off;
    a->relu = relu; a->M = M; a->C = C; a->HW = HW;
    a->out = nullptr; a->out_plane = 0; a->out_planes = 1; a->out_ld = 0; a->out_off = 0; a->out_nchw = 0;
    return 0;
}

extern "C" int pv2_act_apply(const float* y1, int ld1, int off1, int ns1, long long ss1, const float* s1, const float* b1,
                             const float* y2, int ld2, int off2, int ns2, long long ss2,
                             const float* s2, const float* b2, int combine, const void* mult, long long mult_plane, int mult_planes,
                             int mult_ld, int mult_off, int relu, long long M, int C, int HW, void* out, long long out_plane,
                             int out_planes, int out_ld, int out_off, int out_nchw, const pv2_bn_defer* d1, const pv2_bn_defer* d2,
                             int kind, void* stream) {
    PV2_CHECK(kind == PV2_BF16 || kind == PV2_TF32, "act_apply: bad operand kind %d", kind);
    Defer2 df = {};
    if (d1) df.d[0] = *d1;
    if (d2) df.d[1] = *d2;
    const bool deferred = df.d[0].part != nullptr || df.d[1].part != nullptr;
    for (int i = 0; i < 2; ++i) {
        const pv2_bn_defer& d = df.d[i];
        if (!d.part) continue;
        PV2_CHECK(C <= DEFER_MAX_C, "act_apply: a deferred BatchNorm fold handles at most %d channels per slice (got %d)", DEFER_MAX_C, C);
        PV2_CHECK(d.count >= 1.0f && d.ldc >= d.c_off + C && d.mean && d.invstd && (((uintptr_t)d.part) & 15) == 0,
                  "act_apply: incomplete pv2_bn_defer descriptor");
        PV2_CHECK(i == 0 ? (s1 && b1) : (combine && s2 && b2), "act_apply: deferred fold needs the scale / shift output arrays");
    }
    ApplyArgs a;
    if (int e = fill_apply(&a, y1, ld1, off1, ns1, ss1, s1, b1, y2, ld2, off2, ns2, ss2, s2, b2, combine, mult, mult_plane, mult_planes, mult_ld, mult_off, relu, M, C, HW)) return e;
    PV2_CHECK(out != nullptr, "act_apply: null output");
    a.out = out; a.out_plane = out_plane; a.out_planes = out_planes; a.out_ld = out_ld; a.out_off = out_off; a.out_nchw = out_nchw;
    const long long total = M * C;
    auto grid_of = [&](long long work) { return grid_for(work); };
    (void)deferred;
    if (!out_nchw && apply_vec_ok(a) && out_ld % 4 == 0 && out_off % 4 == 0 && al16(out)) {
        static const bool lean_off = [] { const char* e = getenv("PV2_ACT_LEAN"); return e && e[0] == '0'; }();
        if (!lean_off && a.combine == 0 && a.mult == nullptr && a.ns1 == 1 && C / 4 <= 256) {
            const int RP = 256 / (C / 4);
            long long nb = (M + 4LL * RP - 1) / (4LL * RP);          // >= 4 rows per thread when there are that many
            const long long cap = (long long)kNumSMs * 8;
            if (nb > cap) nb = cap;
            if (nb < 1) nb = 1;
            if (kind == PV2_BF16) pv2::launch(act_apply4_lean_kernel<0>, (int)nb, 256, 0, (cudaStream_t)stream, a, df, RP);
            else pv2::launch(act_apply4_lean_kernel<1>, (int)nb, 256, 0, (cudaStream_t)stream, a, df, RP);
            PV2_LAUNCH_CHECK("act_apply4_lean");
            return 0;
        }
        if (kind == PV2_BF16) pv2::launch(act_apply4_kernel<0>, grid_of(total / 4), 256, 0, (cudaStream_t)stream, a, df);
        else pv2::launch(act_apply4_kernel<1>, grid_of(total / 4), 256, 0, (cudaStream_t)stream, a, df);
        PV2_LAUNCH_CHECK("act_apply4");
        return 0;
    }
    if (kind == PV2_BF16) pv2::launch(act_apply_kernel<0>, grid_of(total), 256, 0, (cudaStream_t)stream, a, df);
    else pv2::launch(act_apply_kernel<1>, grid_of(total), 256, 0, (cudaStream_t)stream, a, df);
    PV2_LAUNCH_CHECK("act_apply");
    return 0;
}

extern "C" int pv2_bn_act_bwd(const float* y1, int ld1, int off1, int ns1, long long ss1, const float* s1, const float* b1,
                              const float* y2, int ld2, int off2, int ns2, long long ss2,
                              const float* s2, const float* b2, int combine, const void* mult, long long mult_plane, int mult_planes,
                              int mult_ld, int mult_off, int relu, long long M, int C, int HW,
                              const float* const* dz_slabs, const int* dz_lds, const int* dz_offs, int dz_n, const float* dz_nchw,
                              const float* mean1, const float* inv1, const float* mean2, const float* inv2, int bn_train,
                              float* dmult, int dmult_ld, void* dy1, long long dy1_plane, int dy1_planes, int dy1_ld,
                              void* dy2, long long dy2_plane, int dy2_planes, int dy2_ld,
                              float* dgamma1, float* dbeta1, float* dgamma2, float* dbeta2, float* workspace, float* sums_zeroed,
                              int kind, void* stream) {
    PV2_CHECK(kind == PV2_BF16 || kind == PV2_TF32, "bn_act_bwd: bad operand kind %d", kind);
    BwdArgs b = {};
    if (int e = fill_apply(&b.f, y1, ld1, off1, ns1, ss1, s1, b1, y2, ld2, off2, ns2, ss2, s2, b2, combine, mult, mult_plane, mult_planes, mult_ld, mult_off, relu, M, C, HW)) return e;
    if (dz_nchw == nullptr) {
        if (int e = fill_slabs(&b.dz, dz_slabs, dz_lds, dz_offs, dz_n, "bn_act_bwd")) return e;
    } else {
        b.dz.n = 0;
    }
    b.dz_nchw = dz_nchw;
    PV2_CHECK(dy1 && workspace, "bn_act_bwd: null pointer");
    PV2_CHECK(combine == 0 || dy2, "bn_act_bwd: dy2 missing for a combined op");
    PV2_CHECK(!bn_train || (mean1 && inv1 && (combine == 0 || (mean2 && inv2))), "bn_act_bwd: saved batch statistics missing");
    b.mean1 = mean1; b.inv1 = inv1; b.mean2 = mean2; b.inv2 = inv2;
    b.dmult = mult ? dmult : nullptr; b.dmult_ld = dmult_ld;
    b.bn_train = bn_train;
    b.dy1 = dy1; b.dy1_plane = dy1_plane; b.dy1_planes = dy1_planes; b.dy1_ld = dy1_ld;
    b.dy2 = dy2; b.dy2_plane = dy2_plane; b.dy2_planes = dy2_planes; b.dy2_ld = dy2_ld;
    {   // vector path: reduce (+ ticket fold) and dx, two launches
        bool ok = dz_nchw == nullptr && apply_vec_ok(b.f) && dy1_ld % 4 == 0 && al16(dy1) && dy1_plane % 4 == 0 && sums_zeroed != nullptr &&
                  (reinterpret_cast<uintptr_t>(sums_zeroed) & 127u) == 0;
        if (ok && combine) ok = dy2_ld % 4 == 0 && al16(dy2) && dy2_plane % 4 == 0;
        if (ok && bn_train) ok = al16(mean1) && al16(inv1) && (!combine || (al16(mean2) && al16(inv2)));
        if (ok && b.dmult) ok = dmult_ld % 4 == 0 && al16(dmult);
        for (int i = 0; ok && i < b.dz.n; ++i) ok = b.dz.ld[i] % 4 == 0 && b.dz.off[i] % 4 == 0 && al16(b.dz.p[i]);
        if (ok) {
            const int C4 = C / 4;
            ok = C4 <= 64;        // the dx pass keeps the 4*C finished sums in 4 KB of shared memory
            if (ok) {
                Reduce4Plan pl;
                pl.RP = 256 / C4;
                long long nblk = (M + pl.RP - 1) / pl.RP;                 // at least one pass of rows per block
                const long long cap = 2 * kNumSMs;
                if (nblk > cap) nblk = cap;
                long long rows = (M + nblk - 1) / nblk;
                rows = (rows + pl.RP - 1) / pl.RP * pl.RP;
                nblk = (M + rows - 1) / rows;
                pl.nblk = (int)nblk; pl.rows_pb = (int)rows;
                pl.sums_out = sums_zeroed; b.sums = sums_zeroed;
                pl.dg1 = dgamma1; pl.db1 = dbeta1; pl.dg2 = dgamma2; pl.db2 = dbeta2;
                b.rows_pb = pl.rows_pb;
                cudaStream_t st4 = (cudaStream_t)stream;
                static const bool lean_off = [] { const char* e = getenv("PV2_BN_BWD_LEAN"); return e && e[0] == '0'; }();
                if (!lean_off && bn_train && combine == 0 && mult == nullptr && b.dz.n == 1 && b.f.ns1 == 1 && M * C4 < (1LL << 31)) {
                    LeanBwd a = {};
                    a.y = y1; a.ldy = ld1; a.offy = off1;
                    a.dz = b.dz.p[0]; a.ldz = b.dz.ld[0]; a.offz = b.dz.off[0];
                    a.scale = s1; a.shift = b1; a.mean = mean1; a.inv = inv1;
                    a.relu = relu; a.M = M; a.C = C; a.sums = sums_zeroed;
                    a.dy = dy1; a.dy_plane = dy1_plane; a.dy_planes = dy1_planes; a.dy_ld = dy1_ld;
                    a.dgamma = dgamma1; a.dbeta = dbeta1;
                    a.RP = pl.RP;
                    // more, smaller row blocks than the general kernel: the lean kernel keeps 6+ CTAs per SM resident
                    long long nb = (M + a.RP - 1) / a.RP;
                    if (nb > 4 * kNumSMs) nb = 4 * kNumSMs;
                    long long rws = (M + nb - 1) / nb;
                    rws = (rws + a.RP - 1) / a.RP * a.RP;
                    nb = (M + rws - 1) / rws;
                    a.rows_pb = (int)rws;
                    // opt-in (PV2_BN_BWD_FUSED=1): measured 23.9 us per launch against 6.0 + 5.2 us for the two lean launches at B = 16 x 352^2 --
                    // <= 148 CTAs walk 13+ rows per thread twice and idle at the barrier; head step 1.76-1.78 ms against 1.74-1.75
                    static const bool fused_on = [] { const char* e = getenv("PV2_BN_BWD_FUSED"); return e && e[0] == '1'; }();
                    if (fused_on && g_bn_fused_grid >= 8) {
                        // one launch: every CTA of the grid is resident (the caller's bound), rows split evenly over them
                        long long nf = (M + a.RP - 1) / a.RP;
                        if (nf > g_bn_fused_grid) nf = g_bn_fused_grid;
                        long long rf = (M + nf - 1) / nf;
                        rf = (rf + a.RP - 1) / a.RP * a.RP;
                        nf = (M + rf - 1) / rf;
                        a.rows_pb = (int)rf;
                        if (kind == PV2_BF16) pv2::launch(bn_bwd_fused_lean_kernel<0>, (int)nf, 256, 0, st4, a); else pv2::launch(bn_bwd_fused_lean_kernel<1>, (int)nf, 256, 0, st4, a);
                        PV2_LAUNCH_CHECK("bn_bwd_fused_lean");
                        return 0;
                    }
                    pv2::launch(bn_bwd_reduce4_lean_kernel, (int)nb, 256, 0, st4, a);
                    PV2_LAUNCH_CHECK("bn_bwd_reduce4_lean");
                    const long long total4l = M * C4;
                    (void)total4l;
                    long long nd = (M + 4LL * a.RP - 1) / (4LL * a.RP);       // >= 4 rows per thread when there are that many
                    if (nd > 8LL * kNumSMs) nd = 8LL * kNumSMs;
                    if (kind == PV2_BF16) pv2::launch(bn_bwd_dx4_lean_kernel<0>, (int)nd, 256, 0, st4, a); else pv2::launch(bn_bwd_dx4_lean_kernel<1>, (int)nd, 256, 0, st4, a);
                    PV2_LAUNCH_CHECK("bn_bwd_dx4_lean");
                    return 0;
                }
                if (kind == PV2_BF16) pv2::launch(bn_bwd_reduce4_kernel<0>, pl.nblk, 256, 0, st4, b, pl); else pv2::launch(bn_bwd_reduce4_kernel<1>, pl.nblk, 256, 0, st4, b, pl);
                PV2_LAUNCH_CHECK("bn_bwd_reduce4");
                const long long total4 = M * C4;
                if (kind == PV2_BF16) pv2::launch(bn_bwd_dx4_kernel<0>, grid_for(total4), 256, 0, st4, b, pl); else pv2::launch(bn_bwd_dx4_kernel<1>, grid_for(total4), 256, 0, st4, b, pl);
                PV2_LAUNCH_CHECK("bn_bwd_dx4");
                return 0;
            }
        }
    }
    if (dz_nchw != nullptr && C <= SMALL_C && bn_train && combine == 0 && mult == nullptr && b.f.ns1 == 1 && sums_zeroed != nullptr) {
        SmallBwd a = {};
        a.y = y1; a.ldy = ld1; a.offy = off1; a.dz_nchw = dz_nchw; a.HW = HW;
        a.scale = s1; a.shift = b1; a.mean = mean1; a.inv = inv1; a.relu = relu; a.M = M; a.C = C;
        a.sums = sums_zeroed; a.dy = dy1; a.dy_plane = dy1_plane; a.dy_planes = dy1_planes; a.dy_ld = dy1_ld;
        a.dgamma = dgamma1; a.dbeta = dbeta1;
        cudaStream_t sts = (cudaStream_t)stream;
        long long nb = (M + 255) / 256;
        if (nb > 2 * kNumSMs) nb = 2 * kNumSMs;
        pv2::launch(bn_bwd_reduce_small_kernel, (int)nb, 256, 0, sts, a);
        PV2_LAUNCH_CHECK("bn_bwd_reduce_small");
        if (kind == PV2_BF16) pv2::launch(bn_bwd_dx_small_kernel<0>, grid_for(M), 256, 0, sts, a); else pv2::launch(bn_bwd_dx_small_kernel<1>, grid_for(M), 256, 0, sts, a);
        PV2_LAUNCH_CHECK("bn_bwd_dx_small");
        return 0;
    }
    const int rows = pick_rows(M, C);
    const int rb = (int)((M + rows - 1) / rows);
    b.rows_pb = rows;
    float* part = workspace;
    float* sums = workspace + (size_t)rb * 4 * C;
    b.part = part; b.sums = sums;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 rgrid(rb, (C + 31) / 32);
    if (kind == PV2_BF16) pv2::launch(bn_bwd_reduce_kernel<0>, rgrid, 256, 0, st, b); else pv2::launch(bn_bwd_reduce_kernel<1>, rgrid, 256, 0, st, b);
    PV2_LAUNCH_CHECK("bn_bwd_reduce");
    pv2::launch(bn_bwd_finalize_kernel, (C + 3) / 4, 128, 0, st, part, rb, C, sums, dgamma1, dbeta1, dgamma2, dbeta2);
    PV2_LAUNCH_CHECK("bn_bwd_finalize");
    const long long total = M * C;
    if (kind == PV2_BF16) pv2::launch(bn_bwd_dx_kernel<0>, grid_for(total), 256, 0, st, b); else pv2::launch(bn_bwd_dx_kernel<1>, grid_for(total), 256, 0, st, b);
    PV2_LAUNCH_CHECK("bn_bwd_dx");
    return 0;
}

extern "C" int pv2_up2_nhwc_fwd(const void* in, long long in_plane, int in_planes, int in_ld, int in_off, void* out, long long out_plane,
                                int out_planes, int out_ld, int out_off, int N, int H, int W, int C, int kind, void* stream) {
    PV2_CHECK(kind == PV2_BF16 || kind == PV2_TF32, "up2_nhwc_fwd: bad operand kind %d", kind);
    PV2_CHECK(in && out && N > 0 && H > 0 && W > 0 && C > 0, "up2_nhwc_fwd: bad arguments");
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    const long long total = (long long)N * 4 * H * W * C;
    if (kind == PV2_BF16) pv2::launch(up2_nhwc_fwd_kernel<0>, grid_for(total), 256, 0, (cudaStream_t)stream, in, in_plane, in_planes, in_ld, in_off, out, out_plane, out_planes, out_ld, out_off, N, H, W, C, rh, rw);
    else pv2::launch(up2_nhwc_fwd_kernel<1>, grid_for(total), 256, 0, (cudaStream_t)stream, in, in_plane, in_planes, in_ld, in_off, out, out_plane, out_planes, out_ld, out_off, N, H, W, C, rh, rw);
    PV2_LAUNCH_CHECK("up2_nhwc_fwd");
    return 0;
}

extern "C" int pv2_up2_nhwc_bwd(const float* const* slabs, const int* lds, const int* offs, int nslabs, float* din, int din_ld,
                                int N, int H, int W, int C, void* stream) {
    Slabs s;
    if (int e = fill_slabs(&s, slabs, lds, offs, nslabs, "up2_nhwc_bwd")) return e;
    PV2_CHECK(din && N > 0 && H > 0 && W > 0 && C > 0, "up2_nhwc_bwd: bad arguments");
    const float rh = (float)(H - 1) / (float)(2 * H - 1), rw = (float)(W - 1) / (float)(2 * W - 1);
    bool vec = C % 4 == 0 && din_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(din) & 15u) == 0 && H >= 2 && W >= 2;
    for (int i = 0; vec && i < s.n; ++i) vec = s.ld[i] % 4 == 0 && s.off[i] % 4 == 0 && (reinterpret_cast<uintptr_t>(s.p[i]) & 15u) == 0;
    static const bool tiled_off = [] { const char* e = getenv("PV2_UP2_TILED"); return e && e[0] == '0'; }();
    if (vec && !tiled_off) {
        const int tiles_x = (W + UT - 1) / UT, tiles_y = (H + UT - 1) / UT, cgroups = (C + UC - 1) / UC;
        const long long blocks = (long long)N * tiles_y * tiles_x * cgroups;
        if (blocks < (1LL << 31)) {
            pv2::launch(up2_nhwc_bwd4_tiled_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, s, din, din_ld, H, W, C, rh, rw, tiles_x, tiles_y, cgroups);
            PV2_LAUNCH_CHECK("up2_nhwc_bwd4_tiled");
            return 0;
        }
    }
    if (vec) {
        pv2::launch(up2_nhwc_bwd4_kernel, grid_for((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream, s, din, din_ld, N, H, W, C, rh, rw);
        PV2_LAUNCH_CHECK("up2_nhwc_bwd4");
        return 0;
    }
    pv2::launch(up2_nhwc_bwd_kernel, grid_for((long long)N * H * W * C), 256, 0, (cudaStream_t)stream, s, din, din_ld, N, H, W, C, rh, rw);
    PV2_LAUNCH_CHECK("up2_nhwc_bwd");
    return 0;
}
