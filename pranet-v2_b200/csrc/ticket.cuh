// Deterministic cross-CTA folds without a second launch: every CTA writes its partial, takes a ticket, and the CTA that
// draws the last ticket folds the partials in index order.  Two levels (groups of ~sqrt(n) partials, then the groups) keep
// the serial tail short for any grid size.  Counters live in a caller-provided zero-initialised buffer and are reset by
// the CTA that draws the last ticket, so the same buffer serves every launch on a stream.
#pragma once
#include "pv2_common.cuh"

namespace pv2 {

__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// All `nthreads` threads that synchronise on barrier `bar_id` call this after writing their part of the CTA's partial.
// Returns true (to all of them) in exactly one CTA: the one whose arrival completes `total`.
__device__ __forceinline__ bool ticket_last(unsigned int* counter, unsigned int total, bool leader, volatile int* s_flag, int bar_id, int nthreads) {
    __threadfence();
    bar_sync(bar_id, nthreads);
    if (leader) {
        const unsigned int prev = atomicAdd(counter, 1u);
        const int last = (prev == total - 1u) ? 1 : 0;
        if (last) *counter = 0u;          // nobody else arrives in this launch: leave the buffer zeroed for the next one
        *s_flag = last;
    }
    bar_sync(bar_id, nthreads);
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    bar_sync(bar_id, nthreads);           // s_flag may be reused by the next level
    return last;
}

struct FoldPlan {       // host-computed: n partials in ngroups groups of G
    int n, G, ngroups;
};
inline FoldPlan make_fold_plan(int n) {
    FoldPlan p;
    p.n = n;
    if (n <= 16) { p.G = n < 1 ? 1 : n; p.ngroups = 1; return p; }     // one batch of loads (FOLD_BATCH): a single level is enough
    int g = 1;
    while (g * g < n) ++g;
    p.G = g;
    p.ngroups = (n + g - 1) / g;
    return p;
}

// Chan's parallel combination of (count, mean, M2)
__device__ __forceinline__ void chan_combine(float& n, float& mu, float& M2, float nb, float mub, float M2b) {
    if (nb > 0.0f) {
        const float d = mub - mu, nt = n + nb;
        mu += d * nb / nt;
        M2 += M2b + d * d * n * nb / nt;
        n = nt;
    }
}

// The folds themselves must not be serial chains of L2 loads (each ~0.6 us): a warp folds one value at a time with its LANES
// over the partials (one independent load per lane) and a fixed shuffle tree, so a fold level costs about one L2 latency.
// The folds must not be serial chains of L2 loads (each ~0.6 us).  One THREAD folds one value; the partials it needs are
// fetched in batches of independent (predicated) loads, so a fold level costs one or two L2 latencies whatever the count.
constexpr int FOLD_BATCH = 16;

// Pooled statistics of the partials (n_i, mean_i, M2_i), i in [b0, b1):
//   N = sum n_i,  mean = sum n_i*mean_i / N,  M2 = sum [ M2_i + n_i*(mean_i - mean)^2 ]     (every term non-negative: stable)
template <class LoadFn>
__device__ __forceinline__ void pooled_stats(int b0, int b1, LoadFn load, float& N, float& mean, float& M2) {
    float n[FOLD_BATCH], mu[FOLD_BATCH], m2[FOLD_BATCH];
    if (b1 - b0 <= FOLD_BATCH) {          // the usual case: everything in registers, loaded once
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) {
            n[j] = 0.0f; mu[j] = 0.0f; m2[j] = 0.0f;
            if (b0 + j < b1) load(b0 + j, n[j], mu[j], m2[j]);
        }
        float S = 0.0f, Nn = 0.0f;
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) { S = fmaf(n[j], mu[j], S); Nn += n[j]; }
        const float mean_ = Nn > 0.0f ? S / Nn : 0.0f;
        float q = 0.0f;
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) { const float d = mu[j] - mean_; q += m2[j] + n[j] * d * d; }
        N = Nn; mean = mean_; M2 = q;
        return;
    }
    float S = 0.0f, Nn = 0.0f;
    for (int b = b0; b < b1; b += FOLD_BATCH) {
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) {
            n[j] = 0.0f; mu[j] = 0.0f; m2[j] = 0.0f;
            if (b + j < b1) load(b + j, n[j], mu[j], m2[j]);
        }
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) { S = fmaf(n[j], mu[j], S); Nn += n[j]; }
    }
    const float mean_ = Nn > 0.0f ? S / Nn : 0.0f;
    float q = 0.0f;
    for (int b = b0; b < b1; b += FOLD_BATCH) {
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) {
            n[j] = 0.0f; mu[j] = 0.0f; m2[j] = 0.0f;
            if (b + j < b1) load(b + j, n[j], mu[j], m2[j]);
        }
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) { const float d = mu[j] - mean_; q += m2[j] + n[j] * d * d; }
    }
    N = Nn; mean = mean_; M2 = q;
}

// out = sum over b in [b0, b1) of src[b * stride] in index order, loads batched
__device__ __forceinline__ float fold_sum(const float* src, long long stride, int b0, int b1) {
    float t = 0.0f;
    for (int b = b0; b < b1; b += FOLD_BATCH) {
        float v[FOLD_BATCH];
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) v[j] = (b + j < b1) ? __ldcg(src + (long long)(b + j) * stride) : 0.0f;
#pragma unroll
        for (int j = 0; j < FOLD_BATCH; ++j) t += v[j];
    }
    return t;
}

// Per-channel BatchNorm epilogue shared by the conv kernel's fused statistics and pv2_bn_stats_group.
struct BnFuseDev {
    pv2_bn_fuse f;
};
__device__ __forceinline__ void bn_write_channel(const pv2_bn_fuse& f, int c, float n, float mu, float M2) {
    for (int s = 0; s < f.nsegs; ++s) {
        const pv2_bn_seg& sg = f.seg[s];
        if (c < sg.c_begin || c >= sg.c_end) continue;
        const int cl = c - sg.c_begin;
        const float var = M2 / n;
        const float inv = rsqrtf(var + sg.eps);
        const float g = sg.gamma ? sg.gamma[cl] : 1.0f, b = sg.beta ? sg.beta[cl] : 0.0f;
        f.mean[c] = mu;
        f.invstd[c] = inv;
        f.scale[c] = g * inv;
        f.shift[c] = b - mu * g * inv;
        if (sg.running_mean) {
            sg.running_mean[cl] = (1.0f - sg.momentum) * sg.running_mean[cl] + sg.momentum * mu;
            sg.running_var[cl] = (1.0f - sg.momentum) * sg.running_var[cl] + sg.momentum * (n > 1.0f ? M2 / (n - 1.0f) : var);
        }
        if (cl == 0 && sg.num_batches_tracked) *sg.num_batches_tracked += 1;
        return;
    }
}

}  // namespace pv2
