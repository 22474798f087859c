// Implicit-GEMM convolutions of the DSRA head on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces every nn.Conv2d of the head (binary_seg/lib/pranet.py:34-36 via BasicConv2d, and the biased 1x1
// heads pranet.py:103-104): 1x1, 3x3, 5x5, 1xk, kx1 and dilated 3x3, all stride 1 with "same" padding.
//
// Data layout ("operand format"): activations NHWC with the channel count padded to a multiple of 8 (bf16) / 4
// (tf32), weights [Cout][tap][Cin_p].  GEMM view: M = pixels, N = Cout, K = taps x Cin.
//   * an M tile is 128 CONSECUTIVE output pixels of the flattened (n, y, x) axis; for filter tap (kh, kw) the A operand is
//     those pixels displaced by (kh*dil - pad, kw*dil - pad): ONE im2col-mode TMA load per (tap, 64-channel chunk) --
//     the TMA unit walks rows / images inside the bounding box and zero-fills out-of-image elements, so the
//     convolution's zero padding costs nothing, every MMA row is a real pixel (M = 30976 / 7744 / 1936 at B=16 are
//     242 / 60.5 / 15.1 tiles) and no im2col buffer exists anywhere.  (PV2_CONV_PATCH=1 selects the older tiling:
//     TWb x THb patches of one image loaded with 4-D tiled boxes.);
//   * B (weights) is a 3-D TMA box [BN rows][64 ch] of tap `tap`;
//   * both land in 128-byte-swizzled shared memory, which is exactly the K-major UMMA canonical layout;
//   * one elected thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 16 bf16 / 8 tf32), the fp32 accumulator
//     lives in TMEM; a 4-stage mbarrier ring overlaps TMA with MMA; 4 epilogue warps drain TMEM with tcgen05.ld.
//   * fp32 parity mode ("tf32x3"): every fp32 operand is stored as hi + lo tf32 planes and the K loop runs the
//     three products hi*hi + lo*hi + hi*lo into the same accumulator, which restores ~fp32 accuracy on the
//     tensor cores (single-pass TF32 misses the 1e-3 logit tolerance through ~25 stacked conv+BN layers).
//   * split-K over grid.z when M*N tiles cannot fill 148 SMs (the 11x11 5x5 256->256 layers: K = 6400).
//
// The same kernel computes dgrad (weights repacked flipped/transposed by pv2_weight_pack).  wgrad is the second
// kernel below: D[co][ci] = sum_pixels dY[p][co] * X[p + shift][ci] with both operands MN-major straight out of
// the same NHWC boxes.
#include <cudaTypedefs.h>

#include "pv2_common.cuh"
#include "sm100_ptx.cuh"
#include "ticket.cuh"

namespace pv2 {
namespace {

using namespace ptx;

constexpr int BM = 128;                 // pixel rows per tile = UMMA M
constexpr int ROW_BYTES = 128;          // one swizzle row: 64 bf16 or 32 tf32 channels
constexpr int A_BYTES = BM * ROW_BYTES; // 16 KB
constexpr int MAX_STAGES = 6;           // mbarrier ring capacity; the launch picks 2..6 stages so that several CTAs share an SM
constexpr int THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

struct ConvArgs {
    int H, W, Cout;
    int KW, taps, dil_h, dil_w, pad_h, pad_w;
    int TWb, THb, tiles_x, tiles_y;
    int im2col;             // 1: flat 128-pixel tiles loaded in TMA im2col mode; 0: TWb x THb patches (tiled mode)
    long long M;            // N*H*W
    int kc_per_tap, iters_total, iters_per_split;
    int stats;              // 1: the epilogue also produces the BatchNorm batch statistics (im2col tiles, no split-K)
    int m_tiles, G, ngroups;   // two-level fold of the per-tile statistics
    int BN;                 // N tile (multiple of 16, <= 256)
    int stages;             // depth of the TMA -> MMA ring (2..MAX_STAGES)
    uint32_t tmem_cols;     // power of two >= max(32, BN)
    int out_mode;           // 0: raw fp32 [split][pixel][ldo]   1: NCHW fp32 + bias   2: bf16 [pixel][ldo] (channels_last tensor)
    float* out;
    long long split_stride; // elements between split slabs (mode 0)
    int ldo;
    const float* bias;
};

template <int KIND>
__global__ void __launch_bounds__(THREADS, 4)   // up to MAX_CTAS_PER_SM co-resident CTAs (plan_stages)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const ConvArgs a,
                const __grid_constant__ BnFuseDev bn) {
    constexpr int KC = (KIND == 0) ? 64 : 32;   // channels per 128-byte row
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.BN * ROW_BYTES;
    const int STAGES = a.stages;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_flag;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n_img, y0, x0;
    const long long m0 = (long long)blockIdx.x * BM;
    if (a.im2col) {
        const int hw = a.H * a.W;
        n_img = (int)(m0 / hw);
        const int rem = (int)(m0 - (long long)n_img * hw);
        y0 = rem / a.W; x0 = rem - y0 * a.W;
    } else {
        const int tiles_per_img = a.tiles_x * a.tiles_y;
        n_img = blockIdx.x / tiles_per_img;
        const int trem = blockIdx.x % tiles_per_img;
        y0 = (trem / a.tiles_x) * a.THb; x0 = (trem % a.tiles_x) * a.TWb;
    }
    const int n0 = blockIdx.y * a.BN;
    const int it0 = blockIdx.z * a.iters_per_split;
    const int it1 = min(a.iters_total, it0 + a.iters_per_split);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0); prefetch_tmap(&tmB0);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Dependents may be scheduled only now that this CTA owns its TMEM columns: CTAs of a dependent grid co-reside with ours
    // (shallow rings leave shared memory free) and would otherwise be able to take the columns we still need while they
    // sit in griddepcontrol.wait for us -- a circular wait.
    if (PV2_PDL_EARLY) pdl_trigger();
    pdl_wait();   // everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        const int per_term = a.taps * a.kc_per_tap;
        for (int it = it0; it < it1; ++it) {
            const int s = (it - it0) % STAGES;
            const uint32_t ph = ((it - it0) / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int term = it / per_term, rem = it - term * per_term;
            const int tap = rem / a.kc_per_tap, kc = rem - tap * a.kc_per_tap;
            const int kh = tap / a.KW, kw = tap - kh * a.KW;
            mbar_expect_tx(&full_bar[s], (uint32_t)(A_BYTES + b_bytes));
            // tf32x3 terms: (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo)
            if (a.im2col)
                tma_load_im2col_4d(sA + s * A_BYTES, term == 1 ? &tmA1 : &tmA0, &full_bar[s], kc * KC, x0 - a.pad_w, y0 - a.pad_h, n_img,
                                   (uint16_t)(kw * a.dil_w), (uint16_t)(kh * a.dil_h));
            else
                tma_load_4d(sA + s * A_BYTES, term == 1 ? &tmA1 : &tmA0, &full_bar[s], kc * KC,
                            x0 + kw * a.dil_w - a.pad_w, y0 + kh * a.dil_h - a.pad_h, n_img);
            tma_load_3d(sB + (size_t)s * b_bytes, term == 2 ? &tmB1 : &tmB0, &full_bar[s], kc * KC, tap, n0);
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = instr_desc(KIND == 0 ? 1 : 2, 0, 0, BM, a.BN);
        for (int it = it0; it < it1; ++it) {
            const int s = (it - it0) % STAGES;
            const uint32_t ph = ((it - it0) / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + s * A_BYTES), b_addr = smem_u32(sB + (size_t)s * b_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 4 x (UMMA_K = 32 bytes of K) per 128-byte row
                umma<KIND>(tmem_base, smem_desc_sw128(a_addr + k * 32, 16, 1024), smem_desc_sw128(b_addr + k * 32, 16, 1024),
                           idesc, (it > it0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);   // frees the smem slot once these MMAs have read it
        }
        umma_commit(&acc_bar);            // accumulator complete
        pdl_done();
    } else if (warp >= 2) {
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                  // tile row = pixel within the tile
        bool valid;
        long long pix;
        int y, x, n_pix = n_img;
        if (a.im2col) {
            pix = m0 + r;
            valid = pix < a.M;
            const int hw = a.H * a.W;
            n_pix = (int)(pix / hw);
            const int rem = (int)(pix - (long long)n_pix * hw);
            y = rem / a.W; x = rem - y * a.W;
        } else {
            const int ty = r / a.TWb, tx = r - ty * a.TWb;
            y = y0 + ty; x = x0 + tx;
            valid = (y < a.H) && (x < a.W);
            pix = ((long long)n_img * a.H + y) * a.W + x;
        }
        // fused BatchNorm statistics: the pipeline's shared memory is idle now (every MMA has completed) and serves as scratch
        float* scratch = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);   // this warp's 32 rows x 32 columns, transposed read
        float* wstat = reinterpret_cast<float*>(smem) + 4 * 32 * 33;                // [4 row quarters][BN][mean, M2]
        int nvalid_w = 0;                                                           // valid rows of this warp: a prefix (flat tiles)
        if (a.stats) {
            const long long rem = a.M - m0 - q * 32;
            nvalid_w = rem < 0 ? 0 : (rem > 32 ? 32 : (int)rem);
        }
        for (int c0 = 0; c0 < a.BN; c0 += 32) {
            uint32_t v[32];
            if (a.BN - c0 >= 32) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {   // BN is a multiple of 16: a 16-column tail
                uint32_t t[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, t);
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[j] = t[j]; v[16 + j] = 0u; }
            }
            tmem_ld_wait();
            if (a.stats) {
#pragma unroll
                for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                float sum = 0.0f;
#pragma unroll 8
                for (int r2 = 0; r2 < nvalid_w; ++r2) sum += scratch[r2 * 33 + lane];
                const float mean = nvalid_w > 0 ? sum / (float)nvalid_w : 0.0f;
                float m2 = 0.0f;
#pragma unroll 8
                for (int r2 = 0; r2 < nvalid_w; ++r2) { const float d = scratch[r2 * 33 + lane] - mean; m2 = fmaf(d, d, m2); }
                __syncwarp();
                if (c0 + lane < a.BN) { wstat[(q * a.BN + c0 + lane) * 2] = mean; wstat[(q * a.BN + c0 + lane) * 2 + 1] = m2; }
            }
            if (!valid) continue;
            if (a.out_mode == 0) {
                float* dst = a.out + (long long)blockIdx.z * a.split_stride + pix * a.ldo + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n0 + c0 + j + 3 < a.ldo)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            } else if (a.out_mode == 2) {   // bf16 rows [pixel][ldo]: a channels_last bf16 tensor (the gradient of a backbone feature), final
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.ldo + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (n0 + c0 + j + 7 < a.ldo) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        u.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        u.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                        u.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                        *reinterpret_cast<uint4*>(dst + j) = u;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = n0 + c0 + j;
                    if (co < a.Cout)
                        a.out[(((long long)n_pix * a.Cout + co) * a.H + y) * a.W + x] = __uint_as_float(v[j]) + (a.bias ? a.bias[co] : 0.0f);
                }
            }
        }
        if (a.stats) {
            const int et = threadIdx.x - 64;                 // 0..127 over the four epilogue warps
            const int tile_m = blockIdx.x;
            bar_sync(1, 128);
            // tile statistics: the four row quarters combined in row order
            for (int c = et; c < a.BN; c += 128) {
                const int cg = n0 + c;
                if (cg >= a.Cout) continue;
                float n = 0.0f, mu = 0.0f, M2 = 0.0f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const long long rem = a.M - m0 - qq * 32;
                    const float nv = rem < 0 ? 0.0f : (rem > 32 ? 32.0f : (float)rem);
                    chan_combine(n, mu, M2, nv, wstat[(qq * a.BN + c) * 2], wstat[(qq * a.BN + c) * 2 + 1]);
                }
                float* p = bn.f.part + ((size_t)tile_m * a.Cout + cg) * 2;
                p[0] = mu; p[1] = M2;
            }
            unsigned int* cnt = bn.f.counters + (size_t)blockIdx.y * (a.ngroups + 1);
            const int g = tile_m / a.G;
            const int t0 = g * a.G, t1 = min(t0 + a.G, a.m_tiles);
            if (ticket_last(cnt + g, (unsigned)(t1 - t0), et == 0, &s_flag, 1, 128)) {
                // group fold: one thread per channel, the group's tile partials fetched as one batch of independent loads
                float* gpart = bn.f.part + (size_t)a.m_tiles * a.Cout * 2;
                for (int c = et; c < a.BN; c += 128) {
                    const int cg = n0 + c;
                    if (cg >= a.Cout) continue;
                    float n, mu, M2;
                    pooled_stats(t0, t1, [&](int t, float& pn, float& pmu, float& pm2) {
                        const long long rem = a.M - (long long)t * BM;
                        const float* p = bn.f.part + ((size_t)t * a.Cout + cg) * 2;
                        pn = rem > BM ? (float)BM : (float)rem; pmu = __ldcg(p); pm2 = __ldcg(p + 1);
                    }, n, mu, M2);
                    if (a.ngroups == 1) {                      // few tiles: this fold is already the final one
                        bn_write_channel(bn.f, cg, n, mu, M2);
                    } else {
                        float* gp = gpart + ((size_t)g * a.Cout + cg) * 3;
                        gp[0] = n; gp[1] = mu; gp[2] = M2;
                    }
                }
                if (a.ngroups > 1 && ticket_last(cnt + a.ngroups, (unsigned)a.ngroups, et == 0, &s_flag, 1, 128)) {
                    for (int c = et; c < a.BN; c += 128) {
                        const int cg = n0 + c;
                        if (cg >= a.Cout) continue;
                        float n, mu, M2;
                        pooled_stats(0, a.ngroups, [&](int gg, float& pn, float& pmu, float& pm2) {
                            const float* gp = gpart + ((size_t)gg * a.Cout + cg) * 3;
                            pn = __ldcg(gp); pmu = __ldcg(gp + 1); pm2 = __ldcg(gp + 2);
                        }, n, mu, M2);
                        bn_write_channel(bn.f, cg, n, mu, M2);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// v2 forward / dgrad kernel: PERSISTENT CTAs, deep TMA ring, two TMEM accumulators, staged epilogue.
//
//   * grid = min(work items, SMs x CTAs/SM); a CTA walks work items w = blockIdx.x, + gridDim.x, ... where
//     w = (split, n tile, m tile).  The producer streams the K slabs of consecutive work items through ONE ring, so the
//     loads of the next tile are in flight while the current one is still being multiplied / drained;
//   * the ring is as deep as shared memory allows (up to 12 slabs, ~160 KB): at the head's sizes a tile has 8..100 K slabs of
//     20..48 KB and the L2 -> SM latency is ~1 us, so "all slabs of the tile in flight at once" is what turns the K loop from
//     a chain of round trips into one;
//   * the accumulator is double buffered in TMEM (2 x BN columns): warp 1 starts the MMAs of tile i+1 as soon as its slabs
//     land while warps 2-5 drain tile i;
//   * a K slab whose channel chunk is partly padding (Cin = 32 in a 64-channel box) issues only the MMAs that see data;
//   * epilogue: tcgen05.ld (lane = pixel row) -> 128-byte-swizzled staging tile in shared memory -> (a) BatchNorm column
//     statistics read column-wise (conflict free, all 32 rows in registers, two passes: mean, then centred squares) and
//     (b) row-contiguous 16-byte global stores, 4 complete 128-byte rows per warp instruction.
// ------------------------------------------------------------------------------------------------------
constexpr int V2_MAX_STAGES = 12;
constexpr int EPI_BUF_BYTES = 32 * 128;         // one warp's staging tile: 32 rows x 128 bytes
constexpr int EPI_BYTES = 4 * EPI_BUF_BYTES;     // one staging tile per epilogue warp

struct ConvArgs2 {
    ConvArgs c;
    int n_tiles, splits, work_total;
    int ksteps_last;        // MMAs (of 32 K-bytes each) that see data in the LAST channel chunk of a tap (1..4)
    int BNr;                // TMEM column stride between the two accumulators (BN rounded up to 32)
    int row_bytes;          // bytes of K per shared-memory row: 128 (128-byte swizzle), or 64 when the channel count is 32 mod 64 (bf16:
                            // 32-channel TMA boxes, 64-byte swizzle -- a 64-channel box would be half zero fill for the 32-channel chains)
    unsigned int* tile_counters;   // split-K: one zero-initialised ticket per (n tile, m tile); left zeroed by the launch
    int dbg;                       // profiling switches (PV2_CONV_DBG): 1 = skip the final statistics atomics, 2 = skip the per-chunk column pass
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
// 16-byte fp32 reduction into global memory (L2 atomic unit): split-K partial tiles are ADDED to the one output slab
__device__ __forceinline__ void red_add_v4(float* p, uint4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
    return r;
}

template <int KIND>
__global__ void __launch_bounds__(THREADS, 2)
conv_fwd2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const ConvArgs2 a2,
                 const __grid_constant__ BnFuseDev bn) {
    const ConvArgs& a = a2.c;
    const int RB = a2.row_bytes;
    const int KC = RB / (KIND == 0 ? 2 : 4);    // channels per shared-memory row
    const int a_bytes = BM * RB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.BN * RB;
    const int STAGES = a.stages;
    const int stage_bytes = a_bytes + b_bytes;
    uint8_t* epi = smem + (size_t)STAGES * stage_bytes;                          // [4 warps][32 rows][128 B]
    float* wstat = reinterpret_cast<float*>(epi + EPI_BYTES);                    // [4 row quarters][BN][mean, M2]
    float* cacc = wstat + 4 * a.BN * 2;                                          // [Cout][count, mean, M2] of this CTA's tiles
    __shared__ __align__(8) uint64_t full_bar[V2_MAX_STAGES], empty_bar[V2_MAX_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_flag;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = a.H * a.W;
    const int m_tiles = a.m_tiles;
    const int per_term = a.taps * a.kc_per_tap;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0); prefetch_tmap(&tmB0);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, a.tmem_cols);
    if (a.stats)
        for (int i = threadIdx.x; i < 3 * a.Cout; i += THREADS) cacc[i] = 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (PV2_PDL_EARLY) pdl_trigger();   // (early mode: only now that this CTA owns its TMEM columns, see conv_fwd_kernel)
    pdl_wait();

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int s = 0; uint32_t ph = 0;
        for (int w = blockIdx.x; w < a2.work_total; w += gridDim.x) {
            const int m_tile = w % m_tiles, rest = w / m_tiles;
            const int n_tile = rest % a2.n_tiles, split = rest / a2.n_tiles;
            const long long m0 = (long long)m_tile * BM;
            const int n_img = (int)(m0 / hw);
            const int rem = (int)(m0 - (long long)n_img * hw);
            const int y0 = rem / a.W, x0 = rem - y0 * a.W;
            const int n0 = n_tile * a.BN;
            const int it0 = split * a.iters_per_split, it1 = min(a.iters_total, it0 + a.iters_per_split);
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int term = it / per_term, r2 = it - term * per_term;
                const int tap = r2 / a.kc_per_tap, kc = r2 - tap * a.kc_per_tap;
                const int kh = tap / a.KW, kw = tap - kh * a.KW;
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                tma_load_im2col_4d(sa, term == 1 ? &tmA1 : &tmA0, &full_bar[s], kc * KC, x0 - a.pad_w, y0 - a.pad_h, n_img,
                                   (uint16_t)(kw * a.dil_w), (uint16_t)(kh * a.dil_h));
                tma_load_3d(sa + a_bytes, term == 2 ? &tmB1 : &tmB0, &full_bar[s], kc * KC, tap, n0);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = instr_desc(KIND == 0 ? 1 : 2, 0, 0, BM, a.BN);
        int s = 0; uint32_t ph = 0;
        int lt = 0;
        for (int w = blockIdx.x; w < a2.work_total; w += gridDim.x, ++lt) {
            const int split = (w / m_tiles) / a2.n_tiles;
            const int it0 = split * a.iters_per_split, it1 = min(a.iters_total, it0 + a.iters_per_split);
            const int buf = lt & 1;
            mbar_wait(&tempty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * a2.BNr);
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + (uint32_t)a_bytes;
                const int kc = it % a.kc_per_tap;
                const int ksteps = (kc == a.kc_per_tap - 1) ? a2.ksteps_last : (RB >> 5);
                // K-major canonical layouts: 8-row core groups of 8 * row bytes; layout type 2 = 128-byte swizzle, 4 = 64-byte swizzle
                const uint32_t sbo = 8u * (uint32_t)RB, lt = RB == 128 ? 2u : 4u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < ksteps)
                        umma<KIND>(d_tmem, smem_desc(a_addr + k * 32, 16, sbo, lt), smem_desc(b_addr + k * 32, 16, sbo, lt),
                                   idesc, (it > it0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(&tfull_bar[buf]);
        }
        pdl_done();      // every MMA of this CTA is issued: what is left is the last tile's epilogue, the dependent's prologue may overlap it
    } else if (warp >= 2) {
        // ---------------- epilogue ----------------
        const int q = warp & 3;                          // TMEM lane quarter of this warp
        const uint32_t stg = smem_u32(epi + (size_t)(warp - 2) * EPI_BUF_BYTES);
        const uint32_t my_row_off = (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int et = threadIdx.x - 64;
        int lt = 0;
        // BatchNorm statistics of this warp's 32 staged rows x 32 columns (lane = column): all rows into registers, two passes
        auto stats_chunk = [&](uint32_t stg, int nvalid_w, int c0) {
            float x[32];
            const uint32_t cb = stg + (uint32_t)((lane & 3) << 2);
            const uint32_t cc = (uint32_t)(lane >> 2);
#pragma unroll
            for (int r = 0; r < 32; ++r) x[r] = ld_shared_f32(cb + (uint32_t)r * 128u + ((cc ^ (uint32_t)(r & 7)) << 4));
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; r += 4) {
                s0 += (r < nvalid_w) ? x[r] : 0.f; s1 += (r + 1 < nvalid_w) ? x[r + 1] : 0.f;
                s2 += (r + 2 < nvalid_w) ? x[r + 2] : 0.f; s3 += (r + 3 < nvalid_w) ? x[r + 3] : 0.f;
            }
            const float mean = nvalid_w > 0 ? ((s0 + s1) + (s2 + s3)) / (float)nvalid_w : 0.0f;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; r += 4) {
                const float d0 = x[r] - mean, d1 = x[r + 1] - mean, d2 = x[r + 2] - mean, d3 = x[r + 3] - mean;
                q0 = (r < nvalid_w) ? fmaf(d0, d0, q0) : q0; q1 = (r + 1 < nvalid_w) ? fmaf(d1, d1, q1) : q1;
                q2 = (r + 2 < nvalid_w) ? fmaf(d2, d2, q2) : q2; q3 = (r + 3 < nvalid_w) ? fmaf(d3, d3, q3) : q3;
            }
            if (c0 + lane < a.BN) {
                wstat[(q * a.BN + c0 + lane) * 2] = mean;
                wstat[(q * a.BN + c0 + lane) * 2 + 1] = (q0 + q1) + (q2 + q3);
            }
        };
        // tile statistics: the four row quarters combined in row order, then folded (Chan) into this CTA's running statistics
        // of the channel -- always by the same thread, so no further synchronisation is needed
        auto tile_stats = [&](long long m0, int n0) {
            bar_sync(1, 128);
            for (int c = et; c < a.BN; c += 128) {
                const int cg = n0 + c;
                if (cg >= a.Cout) continue;
                float n = 0.0f, mu = 0.0f, M2 = 0.0f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const long long remr = a.M - m0 - qq * 32;
                    const float nv = remr < 0 ? 0.0f : (remr > 32 ? 32.0f : (float)remr);
                    chan_combine(n, mu, M2, nv, wstat[(qq * a.BN + c) * 2], wstat[(qq * a.BN + c) * 2 + 1]);
                }
                float* acc = cacc + 3 * cg;
                float an = acc[0], amu = acc[1], aM2 = acc[2];
                chan_combine(an, amu, aM2, n, mu, M2);
                acc[0] = an; acc[1] = amu; acc[2] = aM2;
            }
            bar_sync(1, 128);     // wstat is rewritten by the next tile
        };
        const bool fix = a2.splits > 1;      // split-K: partial tiles are ADDED to the (zero-initialised) output with 16-byte reductions
        for (int w = blockIdx.x; w < a2.work_total; w += gridDim.x, ++lt) {
            const int m_tile = w % m_tiles, rest = w / m_tiles;
            const int n_tile = rest % a2.n_tiles, split = rest / a2.n_tiles;
            const long long m0 = (long long)m_tile * BM;
            const int n0 = n_tile * a.BN;
            const int buf = lt & 1;
            mbar_wait(&tfull_bar[buf], ((uint32_t)lt >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * a2.BNr);
            int nvalid_w = 0;
            {
                const long long remr = a.M - m0 - q * 32;
                nvalid_w = remr < 0 ? 0 : (remr > 32 ? 32 : (int)remr);
            }
            if (a.out_mode == 1) {
                // biased NCHW head maps (Cout <= a few): direct stores, one pixel row per lane
                const long long pix = m0 + q * 32 + lane;
                const bool valid = pix < a.M;
                const int n_pix = (int)(pix / hw);
                const int rem = (int)(pix - (long long)n_pix * hw);
                const int y = rem / a.W, x = rem - y * a.W;
                for (int c0 = 0; c0 < a.BN; c0 += 16) {
                    uint32_t t[16];
                    tmem_ld_32x16(t_addr + (uint32_t)c0, t);
                    tmem_ld_wait();
                    if (c0 + 16 >= a.BN) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&tempty_bar[buf]); }
                    if (!valid) continue;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = n0 + c0 + j;
                        if (co < a.Cout)
                            a.out[(((long long)n_pix * a.Cout + co) * a.H + y) * a.W + x] = __uint_as_float(t[j]) + (a.bias ? a.bias[co] : 0.0f);
                    }
                }
                continue;
            }
            const int CW = a.out_mode == 2 ? 64 : 32;       // accumulator columns per 128-byte staging row
            for (int c0 = 0; c0 < a.BN; c0 += CW) {
                uint32_t v[32];
                // ---- TMEM -> registers -> swizzled staging rows (this lane's pixel row) ----
                if (a.out_mode == 0) {
                    if (a.BN - c0 >= 32) {
                        tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    } else {
                        uint32_t t[16];
                        tmem_ld_32x16(t_addr + (uint32_t)c0, t);
#pragma unroll
                        for (int j = 0; j < 16; ++j) { v[j] = t[j]; v[16 + j] = 0u; }
                    }
                    tmem_ld_wait();
                } else {
                    uint32_t u[32];
                    const int ncols = min(64, a.BN - c0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int ch = ncols - 32 * h;          // columns of this half that exist
                        if (ch >= 32) {
                            tmem_ld_32x32(t_addr + (uint32_t)(c0 + 32 * h), u);
                        } else if (ch >= 16) {
                            uint32_t t[16];
                            tmem_ld_32x16(t_addr + (uint32_t)(c0 + 32 * h), t);
#pragma unroll
                            for (int j = 0; j < 16; ++j) { u[j] = t[j]; u[16 + j] = 0u; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) u[j] = 0u;
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[16 * h + j] = pack_bf16x2(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1]));
                    }
                }
                if (c0 + CW >= a.BN) {      // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st_shared_v4(stg + my_row_off + ((((uint32_t)j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                if (a.stats && !fix && !(a2.dbg & 2)) stats_chunk(stg, nvalid_w, c0);
                // ---- staging -> global: 4 complete 128-byte rows per warp instruction ----
                const int colu = lane & 7;                                   // 16-byte unit of the row this lane moves
                const int col = n0 + c0 + colu * (a.out_mode == 2 ? 8 : 4);  // first output column of that unit
                const bool col_ok = (col < n0 + a.BN) && (col + (a.out_mode == 2 ? 7 : 3) < a.ldo);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + (lane >> 3);
                    const uint4 val = ld_shared_v4(stg + (uint32_t)rr * 128u + ((((uint32_t)colu) ^ (uint32_t)(rr & 7)) << 4));
                    const long long pix = m0 + q * 32 + rr;
                    if (col_ok && pix < a.M) {
                        if (a.out_mode == 0) {
                            if (fix) red_add_v4(a.out + pix * a.ldo + col, val);
                            else *reinterpret_cast<uint4*>(a.out + pix * a.ldo + col) = val;
                        } else
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + pix * a.ldo + col) = val;
                    }
                }
                __syncwarp();      // the staging tile is rewritten by the next chunk
            }
            if (fix && a.stats) {
                // split-K + BatchNorm: every split CTA has ADDED its partial tile to the output; the one that draws the tile's last
                // ticket reads the finished tile back (L2) and reduces it to the tile statistics exactly like the unsplit path
                unsigned int* cnt = a2.tile_counters + (size_t)n_tile * m_tiles + m_tile;
                if (ticket_last(cnt, (unsigned)a2.splits, et == 0, &s_flag, 1, 128)) {
                    const int colu = lane & 7;
                    // software pipeline over the 32-column chunks: the loads of chunk c+1 are in flight while chunk c is staged and
                    // reduced (a chunk-by-chunk walk is a chain of L2 round trips: 7-8 of them for a 224..256-column tile)
                    auto load_chunk = [&](int c0, float4 (&t)[8]) {
                        const int col = n0 + c0 + colu * 4;
                        const bool col_ok = (c0 < a.BN) && (col < n0 + a.BN) && (col + 3 < a.ldo);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const long long pix = m0 + q * 32 + i * 4 + (lane >> 3);
                            t[i] = (col_ok && pix < a.M) ? __ldcg(reinterpret_cast<const float4*>(a.out + pix * a.ldo + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    };
                    float4 cur[8], nxt[8];
                    load_chunk(0, cur);
                    for (int c0 = 0; c0 < a.BN; c0 += 32) {
                        load_chunk(c0 + 32, nxt);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + (lane >> 3);
                            st_shared_v4(stg + (uint32_t)rr * 128u + ((((uint32_t)colu) ^ (uint32_t)(rr & 7)) << 4),
                                         __float_as_uint(cur[i].x), __float_as_uint(cur[i].y), __float_as_uint(cur[i].z), __float_as_uint(cur[i].w));
                        }
                        __syncwarp();
                        stats_chunk(stg, nvalid_w, c0);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
                    }
                    tile_stats(m0, n0);
                }
            } else if (a.stats) {
                tile_stats(m0, n0);
            }
        }
        if (a.stats) {
            // This CTA's (count, mean, M2) of every channel it saw become raw moments sum x = n*mean and sum x^2 = M2 + n*mean^2 and
            // are ADDED, in double precision, to the layer's two accumulators per channel (zero on entry).  No partial rows, no
            // ticket, no fold: the kernel that consumes a channel slice reads 2 doubles per channel (pv2_bn_defer).  In double the
            // cancellation of E[x^2] - E[x]^2 costs ~1e-16 * (1 + mean^2 / var): nothing at fp32 output precision, and the order in
            // which the <= 296 CTAs arrive changes the result below the rounding of the final float.
            // Each channel's two accumulators sit in their OWN 128-byte line (PV2_BN_ACC_STRIDE doubles apart): same-line atomics
            // serialise in one L2 slice -- with 16 doubles per line 242 CTAs x 16 addresses queued behind each other (~5-9 us).
            double* acc2 = reinterpret_cast<double*>(bn.f.part);
            for (int cg = et; cg < a.Cout; cg += 128) {
                const float n = cacc[3 * cg], mu = cacc[3 * cg + 1], M2 = cacc[3 * cg + 2];
                if (n > 0.0f && !(a2.dbg & 1)) {
                    const double dn = (double)n, dm = (double)mu;
                    atomicAdd(acc2 + (size_t)PV2_BN_ACC_STRIDE * cg, dn * dm);
                    atomicAdd(acc2 + (size_t)PV2_BN_ACC_STRIDE * cg + 1, (double)M2 + dn * dm * dm);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// wgrad: per filter tap, dW[co][ci] = sum over pixels dY[p][co] * X[p + shift(tap)][ci].
// CTA = (tap, (co tile, ci tile), pixel split).  A = dY patch, B = shifted X patch, both [128 pixels][128 B of
// channels] boxes -> MN-major UMMA operands (K = pixels).  M = 128 output channels (rows beyond Cout are TMA
// zero fill), N = BNW input channels.
// ------------------------------------------------------------------------------------------------------
struct WgradArgs {
    int H, W, Cout, Cin_p;
    int KW, taps, dil_h, dil_w, pad_h, pad_w;
    int TWb, THb, tiles_x, tiles_y, tiles_total, tiles_per_split;
    int im2col;             // 1: flat 128-pixel K tiles (dY: 2-D flat boxes, X: im2col-mode loads); 0: patches
    int nterms;
    int BN;                 // ci tile (multiple of KC, <= 256)
    int stages;             // ring depth (2..WgradCfg::STAGES_)
    int ci_tiles;
    uint32_t tmem_cols;
    float* out;             // [split][Cout][taps][Cin_p]
    long long split_stride;
};

// stage = (dY boxes + X boxes) x 16 KB: bf16 (2 + 2) x 16 KB x 3 stages, tf32 (4 + 2) x 16 KB x 2 stages = 192 KB
template <int KIND> struct WgradCfg { static constexpr int STAGES_ = KIND == 0 ? 3 : 2; static constexpr int BN_MAX = KIND == 0 ? 128 : 64; };   // STAGES_: upper bound

template <int KIND>
__global__ void __launch_bounds__(THREADS, 4)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1,
                  const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1, const WgradArgs a) {
    constexpr int KC = (KIND == 0) ? 64 : 32;
    constexpr int KSTEP_ROWS = (KIND == 0) ? 16 : 8;      // pixels per MMA (UMMA_K)
    constexpr int A_BOXES = BM / KC;                       // dY boxes per stage (128 output channels)
    const int WSTAGES = a.stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_boxes = a.BN / KC;
    const int stage_bytes = (A_BOXES + b_boxes) * A_BYTES;
    __shared__ __align__(8) uint64_t full_bar[WgradCfg<KIND>::STAGES_], empty_bar[WgradCfg<KIND>::STAGES_], acc_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap = blockIdx.x, kh = tap / a.KW, kw = tap - kh * a.KW;
    const int co0 = (blockIdx.y / a.ci_tiles) * BM, ci0 = (blockIdx.y % a.ci_tiles) * a.BN;
    const int t0 = blockIdx.z * a.tiles_per_split, t1 = min(a.tiles_total, t0 + a.tiles_per_split);
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int n_iters = (t1 - t0) * a.nterms;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmG0); prefetch_tmap(&tmX0);
        for (int s = 0; s < WSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Dependents may be scheduled only now that this CTA owns its TMEM columns: CTAs of a dependent grid co-reside with ours
    // (shallow rings leave shared memory free) and would otherwise be able to take the columns we still need while they
    // sit in griddepcontrol.wait for us -- a circular wait.
    if (PV2_PDL_EARLY) pdl_trigger();
    pdl_wait();   // everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < n_iters; ++i) {
            const int s = i % WSTAGES;
            const uint32_t ph = (i / WSTAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int t = t0 + i / a.nterms, term = i % a.nterms;
            uint8_t* base = smem + (size_t)s * stage_bytes;
            mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            const CUtensorMap* mg = (term == 1) ? &tmG1 : &tmG0;   // (G_hi,X_hi) (G_lo,X_hi) (G_hi,X_lo)
            const CUtensorMap* mx = (term == 2) ? &tmX1 : &tmX0;
            if (a.im2col) {
                const long long m0 = (long long)t * BM;
                const int hw = a.H * a.W;
                const int n_img = (int)(m0 / hw), rem = (int)(m0 - (long long)n_img * hw);
                const int y0 = rem / a.W, x0 = rem - y0 * a.W;
                for (int j = 0; j < A_BOXES; ++j) tma_load_2d(base + j * A_BYTES, mg, &full_bar[s], co0 + j * KC, (int)m0);
                for (int j = 0; j < b_boxes; ++j)
                    tma_load_im2col_4d(base + (A_BOXES + j) * A_BYTES, mx, &full_bar[s], ci0 + j * KC, x0 - a.pad_w, y0 - a.pad_h, n_img,
                                       (uint16_t)(kw * a.dil_w), (uint16_t)(kh * a.dil_h));
            } else {
                const int n_img = t / tiles_per_img, trem = t % tiles_per_img;
                const int y0 = (trem / a.tiles_x) * a.THb, x0 = (trem % a.tiles_x) * a.TWb;
                for (int j = 0; j < A_BOXES; ++j) tma_load_4d(base + j * A_BYTES, mg, &full_bar[s], co0 + j * KC, x0, y0, n_img);
                for (int j = 0; j < b_boxes; ++j)
                    tma_load_4d(base + (A_BOXES + j) * A_BYTES, mx, &full_bar[s], ci0 + j * KC,
                                x0 + kw * a.dil_w - a.pad_w, y0 + kh * a.dil_h - a.pad_h, n_img);
            }
        }
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = instr_desc(KIND == 0 ? 1 : 2, 1, 1, BM, a.BN);
        for (int i = 0; i < n_iters; ++i) {
            const int s = i % WSTAGES;
            const uint32_t ph = (i / WSTAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + A_BOXES * A_BYTES;
#pragma unroll 4
            for (int k = 0; k < BM / KSTEP_ROWS; ++k) {
                const uint32_t off = (uint32_t)k * KSTEP_ROWS * ROW_BYTES;
                if constexpr (KIND == 0)
                    umma<KIND>(tmem_base, smem_desc_sw128(a_addr + off, A_BYTES, 1024), smem_desc_sw128(b_addr + off, A_BYTES, 1024),
                               idesc, (i > 0 || k > 0) ? 1u : 0u);
                else   // 32-bit MN-major: 32-byte-atom swizzle, 4-row core groups
                    umma<KIND>(tmem_base, smem_desc_sw128_base32(a_addr + off, A_BYTES, 512), smem_desc_sw128_base32(b_addr + off, A_BYTES, 512),
                               idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_bar);
        pdl_done();
    } else if (warp >= 2) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        for (int c0 = 0; c0 < a.BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (co >= a.Cout) continue;
            float* dst = a.out + (long long)blockIdx.z * a.split_stride + ((long long)co * a.taps + tap) * a.Cin_p + ci0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (ci0 + c0 + j + 3 < a.Cin_p)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// NHWC activation map: dims (C, W, H, N), box (KC, TWb, THb, 1), 128B swizzle, zero OOB fill
int make_act_map(CUtensorMap* m, const void* base, int kind, int Cp, int W, int H, int N, int TWb, int THb, bool atom32 = false) {
    auto enc = get_encode();
    PV2_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const size_t es = kind == 0 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cp * es, (cuuint64_t)W * Cp * es, (cuuint64_t)H * W * Cp * es};
    cuuint32_t box[4] = {(cuuint32_t)(kind == 0 ? 64 : 32), (cuuint32_t)TWb, (cuuint32_t)THb, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation C=%d W=%d H=%d N=%d box %dx%d) failed: %d", Cp, W, H, N, TWb, THb, (int)r);
    return 0;
}

PFN_cuTensorMapEncodeIm2col_v12000 get_encode_im2col() {
    static PFN_cuTensorMapEncodeIm2col_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeIm2col_v12000)p;
    }
    return fn;
}

bool use_im2col() {
    static const bool on = [] { const char* e = getenv("PV2_CONV_PATCH"); return !(e && e[0] == '1'); }();
    return on;
}

// NHWC activation map in im2col mode: dims (C, W, H, N); the base pixel walks the W x H bounding box whose lower corner is
// (-pad_w, -pad_h) and whose upper corner is pulled in by (K-1)*dil - pad = pad ("same" convolution), KC channels per pixel,
// 128 pixels per load; the per-tap displacement (kw*dil_w, kh*dil_h) is given to each load instruction.
int make_im2col_map(CUtensorMap* m, const void* base, int kind, int Cp, int W, int H, int N, int pad_w, int pad_h, bool atom32 = false, bool narrow = false) {
    auto enc = get_encode_im2col();
    PV2_CHECK(enc != nullptr, "cuTensorMapEncodeIm2col not available from the driver");
    PV2_CHECK(pad_w <= 127 && pad_h <= 127, "im2col map: padding %dx%d exceeds the 8-bit corner range", pad_h, pad_w);
    const size_t es = kind == 0 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cp * es, (cuuint64_t)W * Cp * es, (cuuint64_t)H * W * Cp * es};
    int lower[2] = {-pad_w, -pad_h}, upper[2] = {-pad_w, -pad_h};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims,
                     strides, lower, upper, (cuuint32_t)(narrow ? 32 : (kind == 0 ? 64 : 32)), (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     narrow ? CU_TENSOR_MAP_SWIZZLE_64B : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col(C=%d W=%d H=%d N=%d pad %dx%d) failed: %d", Cp, W, H, N, pad_h, pad_w, (int)r);
    return 0;
}

// flat [M][Cp] map (wgrad's dY operand with flat pixel tiles): dims (Cp, M), box (KC, 128)
int make_flat_map(CUtensorMap* m, const void* base, int kind, int Cp, long long M, bool atom32 = false) {
    auto enc = get_encode();
    PV2_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const size_t es = kind == 0 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)Cp, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)Cp * es};
    cuuint32_t box[2] = {(cuuint32_t)(kind == 0 ? 64 : 32), (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(flat C=%d M=%lld) failed: %d", Cp, M, (int)r);
    return 0;
}

// weight map: dims (Cin_p, taps, Cout), box (KC, 1, BN)
int make_w_map(CUtensorMap* m, const void* base, int kind, int Cin_p, int taps, int Cout, int BN, bool narrow = false) {
    auto enc = get_encode();
    PV2_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const size_t es = kind == 0 ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)Cin_p, (cuuint64_t)taps, (cuuint64_t)Cout};
    cuuint64_t strides[2] = {(cuuint64_t)Cin_p * es, (cuuint64_t)taps * Cin_p * es};
    cuuint32_t box[3] = {(cuuint32_t)(narrow ? 32 : (kind == 0 ? 64 : 32)), 1, (cuuint32_t)BN};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights Cin=%d taps=%d Cout=%d BN=%d) failed: %d", Cin_p, taps, Cout, BN, (int)r);
    return 0;
}

void pick_patch(int W, int* TWb, int* THb) {
    int tw = 8;
    while (tw < W && tw < 128) tw <<= 1;
    *TWb = tw;
    *THb = BM / tw;
}

// A 1x1 convolution has no spatial structure: its pixels are one flat GEMM M dimension, so the 128-row tiles are taken
// from the flattened [N*H*W] axis (every MMA row is a real pixel) instead of per-image TWb x THb patches.
void flatten_1x1(int KH, int KW, int* N, int* H, int* W) {
    if (KH == 1 && KW == 1) { *W = *N * *H * *W; *H = 1; *N = 1; }
}

int m_tiles_of(int N, int H, int W, int KH, int KW, bool flatten_ok) {
    if (use_im2col()) return (int)(((long long)N * H * W + BM - 1) / BM);
    if (flatten_ok) flatten_1x1(KH, KW, &N, &H, &W);
    int TWb, THb;
    pick_patch(W, &TWb, &THb);
    return N * ((W + TWb - 1) / TWb) * ((H + THb - 1) / THb);
}

uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// How many pipeline stages a CTA gets.  The head's GEMMs are short (8..100 K iterations per CTA), their epilogue (TMEM ->
// registers -> global, plus the fused BatchNorm statistics) is as long as the main loop, and up to twelve independent chains
// of the head run concurrently on side streams.  So instead of one CTA owning an SM with a deep ring, a CTA takes the
// SMALLEST footprint that still pipelines (>= 2 stages) and up to MAX_CTAS_PER_SM CTAs -- of this grid or of a sibling
// chain's kernel -- share the SM: one CTA's epilogue / prologue overlaps another's MMAs, and grids of 150..600 tiles run as
// one wave.  Measured on the B=16 x 352^2 head step (fwd + loss + bwd, CUDA graph): 1 CTA/SM 3.15 ms, 2: 2.84 ms, 3: 2.73 ms.
// Bounded by shared memory (227 KB), TMEM (512 columns) and the scratch the statistics epilogue needs.
// PV2_CONV_CTAS / PV2_CONV_STAGES override (profiling sweeps).
constexpr int MAX_CTAS_PER_SM = 4;      // == the kernels' __launch_bounds__ minimum-blocks argument
int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}
int plan_stages(long long ctas, int iters, size_t stage_bytes, uint32_t tmem_cols, size_t min_bytes, int max_stages) {
    static const int force_stages = env_int("PV2_CONV_STAGES", 0), force_ctas = env_int("PV2_CONV_CTAS", 0);
    static const int small_grid = env_int("PV2_CONV_SMALL", 0);   // grids up to this many CTAs take the deepest ring (latency mode)
    int target = force_ctas > 0 ? force_ctas : 3;
    if (ctas <= small_grid) target = 1;
    if (target > MAX_CTAS_PER_SM) target = MAX_CTAS_PER_SM;
    while (target > 1 && (uint32_t)target * tmem_cols > 512u) --target;
    int stages = max_stages;
    for (;; --target) {
        const size_t budget = (size_t)(227 * 1024) / target - 2048;     // 1 KB reserved per CTA + 1 KB alignment slack
        stages = (int)(budget / stage_bytes);
        if (stages > max_stages) stages = max_stages;
        if ((stages >= 2 && (size_t)stages * stage_bytes >= min_bytes) || target == 1) break;
    }
    if (stages < 2) stages = 2;
    if (force_stages > 0) stages = force_stages < 2 ? 2 : (force_stages > max_stages ? max_stages : force_stages);
    if (stages > iters && iters >= 2) stages = iters;
    while ((size_t)stages * stage_bytes < min_bytes && stages < max_stages) ++stages;
    return stages;
}

// N tiling shared by the kernels, the split-K hint and the statistics workspace: the fewest tiles of at most 256 columns,
// evenly sized (Cout = 416 -> 2 x 208 instead of 256 + 160)
bool use_v1();
bool use_narrow(int kind, int Cin_p);
inline void n_tiling(int Cout, int* BN, int* n_tiles) {
    if (use_v1()) {     // the one-tile-per-CTA kernel stores whole 32-column chunks: its tiles are 256 wide or the last one
        *BN = Cout >= 256 ? 256 : ((Cout + 15) / 16) * 16;
        *n_tiles = (Cout + *BN - 1) / *BN;
        return;
    }
    const int nt = (Cout + 255) / 256;
    const int per = (Cout + nt - 1) / nt;
    *BN = ((per + 15) / 16) * 16;
    *n_tiles = (Cout + *BN - 1) / *BN;
}

// bf16 convs with 32 input channels (the RFB / aggregation chains: 30 of the head's 51 convs, and their dgrads) load 32-channel boxes
// into 64-byte-swizzled rows: a 64-channel box would be half zero fill -- TMA and shared-memory traffic for nothing
bool use_narrow(int kind, int Cin_p) {
    static const bool off = [] { const char* e = getenv("PV2_CONV_NARROW"); return e && e[0] == '0'; }();
    return !off && !use_v1() && kind == PV2_BF16 && Cin_p == 32;      // (at 96 channels three 32-channel slabs per tap cost more barrier round trips than the zero fill saves)
}

bool use_v1() {
    static const bool on = [] { const char* e = getenv("PV2_CONV_V1"); return (e && e[0] == '1') || !use_im2col(); }();
    return on;
}

inline size_t v2_tail_bytes(bool stats, int BN, int Cout) {     // staging tiles + (statistics) per-quarter scratch + CTA accumulators
    return (size_t)EPI_BYTES + (stats ? ((size_t)4 * BN * 2 + (size_t)3 * Cout) * sizeof(float) : 0);
}

// Persistent launch plan of conv_fwd2_kernel: CTAs per SM (1, or 2 when a work item's K slabs are few and small), ring depth.
struct Plan2 { int grid, stages; size_t smem; };
int g_cta_budget = 0;
int cta_budget() {
    const int e = tune_int("PV2_CONV_MAXGRID", 0);
    return e > 0 ? e : g_cta_budget;
}
// Policy.  A work item whose K slabs add up to more than ~256 KB (the level GEMMs, the 5x5 and 96-channel stacks) gets an SM to
// itself and the deepest ring that fits (latency: as many slabs in flight as possible).  Smaller items -- the 32/64-channel
// chains, twelve of which run concurrently in the head -- take at most half an SM, so that CTAs of two launches (or two tiles
// of one) share it and one's epilogue / TMA round trips overlap the other's MMAs; 242-tile grids are then one resident wave.
Plan2 plan_v2(int work_total, int slabs_per_item, size_t stage_bytes, uint32_t tmem_cols, size_t tail_bytes) {
    const int force_stages = tune_int("PV2_CONV_STAGES", 0), force_cps = tune_int("PV2_CONV_CPS", 0);
    int cps = ((size_t)slabs_per_item * stage_bytes <= (size_t)256 * 1024 && 2u * tmem_cols <= 512u) ? 2 : 1;
    if (force_cps == 1 || force_cps == 2) cps = (force_cps == 2 && 2u * tmem_cols > 512u) ? 1 : force_cps;
    if (cps == 2 && (size_t)112 * 1024 < 1024 + tail_bytes + stage_bytes) cps = 1;
    Plan2 p;
    p.grid = work_total < kNumSMs * cps ? work_total : kNumSMs * cps;
    const int cap = cta_budget();      // concurrent chains: fewer CTAs, each walking more tiles through its ring (see pv2_conv_set_cta_budget)
    if (cap > 0 && p.grid > cap) p.grid = cap;
    const int items_per_cta = (work_total + p.grid - 1) / p.grid;
    const size_t budget = (cps == 1 ? (size_t)226 * 1024 : (size_t)112 * 1024) - 1024 - tail_bytes;   // 1 KB alignment slack
    int st = (int)(budget / stage_bytes);
    if (st > V2_MAX_STAGES) st = V2_MAX_STAGES;
    const long long need = (long long)slabs_per_item * items_per_cta;
    if (st > need) st = (int)need;
    if (force_stages > 0 && force_stages < st) st = force_stages;
    if (st < 1) st = 1;
    p.stages = st;
    p.smem = (size_t)st * stage_bytes + tail_bytes + 1024;
    return p;
}

int common_checks(const char* who, int kind, int nterms, int N, int H, int W, int Cin_p, int Cout, int KH, int KW) {
    PV2_CHECK(kind == PV2_BF16 || kind == PV2_TF32, "%s: operand kind must be PV2_BF16 or PV2_TF32 (got %d)", who, kind);
    PV2_CHECK(nterms == 1 || (kind == PV2_TF32 && nterms == 3), "%s: nterms must be 1, or 3 with tf32 operands", who);
    PV2_CHECK(N > 0 && H > 0 && W > 0 && Cin_p > 0 && Cout > 0 && KH > 0 && KW > 0, "%s: empty shape", who);
    PV2_CHECK((KH & 1) && (KW & 1), "%s: only odd kernel sizes with same padding are supported (%dx%d)", who, KH, KW);
    PV2_CHECK(Cin_p % (kind == PV2_BF16 ? 8 : 4) == 0, "%s: padded input channels %d break the 16-byte TMA stride rule", who, Cin_p);
    return 0;
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" int pv2_conv_fuses_bn_stats(int splits, int out_mode) {
    static const bool off = [] { const char* e = getenv("PV2_NO_FUSED_STATS"); return e && e[0] == '1'; }();   // A/B switch for profiling
    if (off || !use_im2col() || out_mode != 0) return 0;
    if (use_v1()) return splits == 1 ? 1 : 0;      // 1: final statistics written by the launch
    return 2;                                      // 2: per-CTA partials (also under split-K: the fix-up CTA of a tile produces them)
}

extern "C" int pv2_conv_sums_splits(void) { return use_v1() ? 0 : 1; }

extern "C" int pv2_conv_set_cta_budget(int max_ctas) {
    const int old = g_cta_budget;
    g_cta_budget = max_ctas > 0 ? max_ctas : 0;
    return old;
}

extern "C" int pv2_conv_splits_hint(int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int kind, int nterms) {
    const int KC = use_narrow(kind, Cin_p) ? 32 : (kind == PV2_BF16 ? 64 : 32);
    const int m_tiles = m_tiles_of(N, H, W, KH, KW, true);
    int BN, n_tiles;
    n_tiling(Cout, &BN, &n_tiles);
    const int iters = nterms * KH * KW * ((Cin_p + KC - 1) / KC);
    int splits = 1;
    if (use_v1()) {
        // fill the 148 SMs when the output tiling alone cannot, keeping >= 8 K iterations per split
        while (splits < 8 && m_tiles * n_tiles * splits * 2 <= kNumSMs && iters / (splits * 2) >= 8) splits *= 2;   // <= 8: consumers sum at most 8 slabs
    } else {
        // Persistent kernel: splitting K costs the partial-tile reductions in L2 plus -- for a BatchNorm'd conv -- a ticket and a re-read
        // of the finished tile (~10 us), so a problem with 48+ output tiles runs unsplit even though it leaves SMs idle (measured at
        // B = 16: the 61-tile level-3 GEMM 26 us split in two, 20 us unsplit); fewer tiles than that (the 16-tile 5x5 stack, the
        // 32-tile level-4 GEMM: 1.3 - 4.8 MB of K slabs per tile) are split until the SMs are filled, keeping >= 8 K slabs per split.
        if (m_tiles * n_tiles < 48)
            while (splits < 8 && m_tiles * n_tiles * splits * 2 <= kNumSMs && iters / (splits * 2) >= 8) splits *= 2;
    }
    const int per = (iters + splits - 1) / splits;
    return (iters + per - 1) / per;   // no empty split
}

extern "C" int pv2_conv_fwd(const void* x, long long x_plane_stride, const void* w_op, long long w_plane_stride, int kind, int nterms,
                            int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int dil_h, int dil_w,
                            int out_mode, float* out, int ldo, int splits, const float* bias, const pv2_bn_fuse* bn,
                            unsigned int* tile_counters, void* stream) {
    if (int e = common_checks("conv_fwd", kind, nterms, N, H, W, Cin_p, Cout, KH, KW)) return e;
    PV2_CHECK(x && w_op && out, "conv_fwd: null pointer");
    PV2_CHECK(out_mode >= 0 && out_mode <= 2, "conv_fwd: bad out_mode %d", out_mode);
    PV2_CHECK(out_mode == 1 || (ldo % 4 == 0 && ldo >= Cout), "conv_fwd: ldo=%d must be a multiple of 4 and >= Cout=%d", ldo, Cout);
    PV2_CHECK(out_mode != 2 || (ldo % 8 == 0 && ((uintptr_t)out & 15) == 0), "conv_fwd: bf16 row output needs ldo %% 8 == 0 and a 16-byte aligned base");
    PV2_CHECK(out_mode == 0 || splits == 1, "conv_fwd: split-K needs the raw output mode");
    const bool im2col = use_im2col();
    if (out_mode == 0 && !im2col) flatten_1x1(KH, KW, &N, &H, &W);
    const bool narrow = use_narrow(kind, Cin_p);
    const int k = kind == PV2_BF16 ? 0 : 1, KC = narrow ? 32 : (k == 0 ? 64 : 32);
    const size_t es = k == 0 ? 2 : 4;
    const int row_bytes = narrow ? 64 : ROW_BYTES;
    ConvArgs a = {};
    a.H = H; a.W = W; a.Cout = Cout;
    a.KW = KW; a.taps = KH * KW; a.dil_h = dil_h; a.dil_w = dil_w;
    a.pad_h = dil_h * (KH - 1) / 2; a.pad_w = dil_w * (KW - 1) / 2;
    pick_patch(W, &a.TWb, &a.THb);
    a.tiles_x = (W + a.TWb - 1) / a.TWb; a.tiles_y = (H + a.THb - 1) / a.THb;
    a.im2col = im2col ? 1 : 0;
    a.M = (long long)N * H * W;
    a.kc_per_tap = (Cin_p + KC - 1) / KC;
    a.iters_total = nterms * a.taps * a.kc_per_tap;
    PV2_CHECK(splits >= 1 && splits <= a.iters_total, "conv_fwd: splits=%d out of range (K iterations %d)", splits, a.iters_total);
    a.iters_per_split = (a.iters_total + splits - 1) / splits;
    PV2_CHECK((long long)a.iters_per_split * (splits - 1) < a.iters_total, "conv_fwd: splits=%d leaves an empty split", splits);
    int n_tiles;
    n_tiling(Cout, &a.BN, &n_tiles);
    a.tmem_cols = pow2_cols(a.BN);
    a.out_mode = out_mode; a.out = out; a.ldo = ldo; a.bias = bias;
    a.split_stride = (long long)N * H * W * ldo;
    BnFuseDev bnd = {};
    a.m_tiles = (int)((a.M + BM - 1) / BM);
    if (bn != nullptr && bn->nsegs > 0 && pv2_conv_fuses_bn_stats(splits, out_mode)) {
        PV2_CHECK(bn->nsegs <= PV2_MAX_BN_SEGS && bn->mean && bn->invstd && bn->scale && bn->shift && bn->part && bn->counters,
                  "conv_fwd: incomplete pv2_bn_fuse descriptor");
        const FoldPlan fp = make_fold_plan(a.m_tiles);
        a.stats = 1; a.G = fp.G; a.ngroups = fp.ngroups;
        PV2_CHECK((long long)((Cout + a.BN - 1) / a.BN) * (fp.ngroups + 1) <= PV2_BN_COUNTERS, "conv_fwd: ticket counters too small for %d tiles", a.m_tiles);
        bnd.f = *bn;
    }
    CUtensorMap mA0, mA1, mB0, mB1;
    if (int e = im2col ? make_im2col_map(&mA0, x, k, Cin_p, W, H, N, a.pad_w, a.pad_h, false, narrow) : make_act_map(&mA0, x, k, Cin_p, W, H, N, a.TWb, a.THb)) return e;
    if (int e = make_w_map(&mB0, w_op, k, Cin_p, a.taps, Cout, a.BN, narrow)) return e;
    mA1 = mA0; mB1 = mB0;
    if (nterms == 3) {
        const void* x1 = (const uint8_t*)x + x_plane_stride * es;
        if (int e = im2col ? make_im2col_map(&mA1, x1, k, Cin_p, W, H, N, a.pad_w, a.pad_h) : make_act_map(&mA1, x1, k, Cin_p, W, H, N, a.TWb, a.THb)) return e;
        if (int e = make_w_map(&mB1, (const uint8_t*)w_op + w_plane_stride * es, k, Cin_p, a.taps, Cout, a.BN)) return e;
    }
    const size_t stage_bytes = (size_t)BM * row_bytes + (size_t)a.BN * row_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce;
    if (!use_v1()) {
        ConvArgs2 a2 = {};
        a2.row_bytes = row_bytes;
        a2.n_tiles = n_tiles; a2.splits = splits;
        a2.work_total = a.m_tiles * n_tiles * splits;
        const int last_ch = Cin_p - (a.kc_per_tap - 1) * KC;              // channels in the last chunk of a tap
        const int step_ch = k == 0 ? 16 : 8;                              // channels one MMA consumes (32 bytes of K)
        a2.ksteps_last = (last_ch + step_ch - 1) / step_ch;
        a2.BNr = ((a.BN + 31) / 32) * 32;
        a2.tile_counters = tile_counters;
        a2.dbg = tune_int("PV2_CONV_DBG", 0);
        PV2_CHECK(splits <= 8, "conv_fwd: at most 8 K splits (got %d)", splits);
        PV2_CHECK(splits == 1 || !a.stats || (tile_counters != nullptr && (long long)a.m_tiles * n_tiles <= PV2_BN_COUNTERS),
                  "conv_fwd: split-K with fused statistics needs %lld zero-initialised tile counters (<= %d)", (long long)a.m_tiles * n_tiles, PV2_BN_COUNTERS);
        a.tmem_cols = pow2_cols(2 * a2.BNr);
        const size_t tail = v2_tail_bytes(a.stats != 0, a.BN, Cout);
        const Plan2 pl = plan_v2(a2.work_total, a.iters_per_split, stage_bytes, a.tmem_cols, tail);
        a.stages = pl.stages;
        a2.c = a;
        PV2_CHECK(pl.smem <= 227 * 1024, "conv_fwd: bad v2 stage plan (%d stages of %zu B)", pl.stages, stage_bytes);
        if (k == 0) {
            ce = cudaFuncSetAttribute(conv_fwd2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
            PV2_CHECK(ce == cudaSuccess, "conv_fwd: smem attribute: %s", cudaGetErrorString(ce));
            pv2::launch(conv_fwd2_kernel<0>, dim3(pl.grid), THREADS, pl.smem, st, mA0, mA1, mB0, mB1, a2, bnd);
        } else {
            ce = cudaFuncSetAttribute(conv_fwd2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
            PV2_CHECK(ce == cudaSuccess, "conv_fwd: smem attribute: %s", cudaGetErrorString(ce));
            pv2::launch(conv_fwd2_kernel<1>, dim3(pl.grid), THREADS, pl.smem, st, mA0, mA1, mB0, mB1, a2, bnd);
        }
        PV2_LAUNCH_CHECK("conv_fwd");
        return 0;
    }
    dim3 grid(im2col ? (unsigned)((a.M + BM - 1) / BM) : (unsigned)(N * a.tiles_x * a.tiles_y), (Cout + a.BN - 1) / a.BN, splits);
    const size_t stats_scratch = a.stats ? (size_t)(4 * 32 * 33 + 4 * a.BN * 2) * sizeof(float) : 0;   // epilogue reuses the ring
    a.stages = plan_stages((long long)grid.x * grid.y * grid.z, a.iters_per_split, stage_bytes, a.tmem_cols, stats_scratch, MAX_STAGES);
    const size_t smem = (size_t)a.stages * stage_bytes + 1024;
    PV2_CHECK(smem <= 227 * 1024 && smem - 1024 >= stats_scratch, "conv_fwd: bad stage plan (%d stages of %zu B)", a.stages, stage_bytes);
    if (k == 0) {
        ce = cudaFuncSetAttribute(conv_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "conv_fwd: smem attribute: %s", cudaGetErrorString(ce));
        pv2::launch(conv_fwd_kernel<0>, grid, THREADS, smem, st, mA0, mA1, mB0, mB1, a, bnd);
    } else {
        ce = cudaFuncSetAttribute(conv_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "conv_fwd: smem attribute: %s", cudaGetErrorString(ce));
        pv2::launch(conv_fwd_kernel<1>, grid, THREADS, smem, st, mA0, mA1, mB0, mB1, a, bnd);
    }
    PV2_LAUNCH_CHECK("conv_fwd");
    return 0;
}

extern "C" int pv2_conv_wgrad_splits_hint(int N, int H, int W, int Cin_p, int Cout, int KH, int KW, int kind) {
    const int KC = kind == PV2_BF16 ? 64 : 32;
    const int tiles = m_tiles_of(N, H, W, KH, KW, true);
    const int cin_r = ((Cin_p + KC - 1) / KC) * KC;
    const int bn_max = kind == PV2_BF16 ? 128 : 64;
    const int BN = cin_r >= bn_max ? bn_max : cin_r;
    const int ctas = KH * KW * ((Cout + BM - 1) / BM) * ((cin_r + BN - 1) / BN);
    int splits = 1;
    while (splits < 64 && ctas * splits * 2 <= 2 * kNumSMs && tiles / (splits * 2) >= 2) splits *= 2;
    const int per = (tiles + splits - 1) / splits;
    return (tiles + per - 1) / per;   // no empty split
}

extern "C" int pv2_conv_wgrad(const void* dy, long long dy_plane_stride, const void* x, long long x_plane_stride, int kind, int nterms,
                              int N, int H, int W, int Cin_p, int Cout_p, int Cout, int KH, int KW, int dil_h, int dil_w,
                              float* out, int splits, void* stream) {
    if (int e = common_checks("conv_wgrad", kind, nterms, N, H, W, Cin_p, Cout, KH, KW)) return e;
    PV2_CHECK(dy && x && out, "conv_wgrad: null pointer");
    PV2_CHECK(Cout_p % (kind == PV2_BF16 ? 8 : 4) == 0 && Cout_p >= Cout, "conv_wgrad: bad padded Cout %d", Cout_p);
    const bool im2col = use_im2col();
    if (!im2col) flatten_1x1(KH, KW, &N, &H, &W);
    const int k = kind == PV2_BF16 ? 0 : 1, KC = k == 0 ? 64 : 32;
    const size_t es = k == 0 ? 2 : 4;
    WgradArgs a = {};
    a.im2col = im2col ? 1 : 0;
    a.H = H; a.W = W; a.Cout = Cout; a.Cin_p = Cin_p;
    a.KW = KW; a.taps = KH * KW; a.dil_h = dil_h; a.dil_w = dil_w;
    a.pad_h = dil_h * (KH - 1) / 2; a.pad_w = dil_w * (KW - 1) / 2;
    pick_patch(W, &a.TWb, &a.THb);
    a.tiles_x = (W + a.TWb - 1) / a.TWb; a.tiles_y = (H + a.THb - 1) / a.THb;
    const long long Mtot = (long long)N * H * W;
    a.tiles_total = im2col ? (int)((Mtot + BM - 1) / BM) : N * a.tiles_x * a.tiles_y;
    PV2_CHECK(splits >= 1 && splits <= a.tiles_total, "conv_wgrad: splits=%d out of range (pixel tiles %d)", splits, a.tiles_total);
    a.tiles_per_split = (a.tiles_total + splits - 1) / splits;
    PV2_CHECK((long long)a.tiles_per_split * (splits - 1) < a.tiles_total, "conv_wgrad: splits=%d leaves an empty split", splits);
    a.nterms = nterms;
    const int cin_r = ((Cin_p + KC - 1) / KC) * KC;
    const int bn_max = k == 0 ? WgradCfg<0>::BN_MAX : WgradCfg<1>::BN_MAX;
    a.BN = cin_r >= bn_max ? bn_max : cin_r;
    a.ci_tiles = (cin_r + a.BN - 1) / a.BN;
    a.tmem_cols = pow2_cols(a.BN);
    a.out = out;
    a.split_stride = (long long)Cout * a.taps * Cin_p;
    CUtensorMap mG0, mG1, mX0, mX1;
    const bool a32 = (k == 1);   // tf32 MN-major operands: 32-byte-atom swizzle
    if (int e = im2col ? make_flat_map(&mG0, dy, k, Cout_p, Mtot, a32) : make_act_map(&mG0, dy, k, Cout_p, W, H, N, a.TWb, a.THb, a32)) return e;
    if (int e = im2col ? make_im2col_map(&mX0, x, k, Cin_p, W, H, N, a.pad_w, a.pad_h, a32) : make_act_map(&mX0, x, k, Cin_p, W, H, N, a.TWb, a.THb, a32)) return e;
    mG1 = mG0; mX1 = mX0;
    if (nterms == 3) {
        const void* dy1 = (const uint8_t*)dy + dy_plane_stride * es;
        const void* x1 = (const uint8_t*)x + x_plane_stride * es;
        if (int e = im2col ? make_flat_map(&mG1, dy1, k, Cout_p, Mtot, a32) : make_act_map(&mG1, dy1, k, Cout_p, W, H, N, a.TWb, a.THb, a32)) return e;
        if (int e = im2col ? make_im2col_map(&mX1, x1, k, Cin_p, W, H, N, a.pad_w, a.pad_h, a32) : make_act_map(&mX1, x1, k, Cin_p, W, H, N, a.TWb, a.THb, a32)) return e;
    }
    dim3 grid(a.taps, ((Cout + BM - 1) / BM) * a.ci_tiles, splits);
    const size_t wstage_bytes = (size_t)((BM / KC) + (a.BN / KC)) * A_BYTES;
    a.stages = plan_stages((long long)grid.x * grid.y * grid.z, a.tiles_per_split * nterms, wstage_bytes, a.tmem_cols, 0,
                           k == 0 ? WgradCfg<0>::STAGES_ : WgradCfg<1>::STAGES_);
    const size_t smem = (size_t)a.stages * wstage_bytes + 1024;
    PV2_CHECK(smem <= 227 * 1024, "conv_wgrad: stage too large (%zu B)", smem);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce;
    if (k == 0) {
        ce = cudaFuncSetAttribute(conv_wgrad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "conv_wgrad: smem attribute: %s", cudaGetErrorString(ce));
        pv2::launch(conv_wgrad_kernel<0>, grid, THREADS, smem, st, mG0, mG1, mX0, mX1, a);
    } else {
        ce = cudaFuncSetAttribute(conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        PV2_CHECK(ce == cudaSuccess, "conv_wgrad: smem attribute: %s", cudaGetErrorString(ce));
        pv2::launch(conv_wgrad_kernel<1>, grid, THREADS, smem, st, mG0, mG1, mX0, mX1, a);
    }
    PV2_LAUNCH_CHECK("conv_wgrad");
    return 0;
}
