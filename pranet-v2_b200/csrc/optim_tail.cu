// Optimizer tail of a training step: gradient scale (1/world) + element-wise clamp + Adam / AdamW, one flat stream.
//
//   binary_seg/utils/utils.py:7-17        clip_gradient: param.grad.data.clamp_(-clip, clip) for every parameter
//   binary_seg/MyTrain_med.py:85-86,148   clip_gradient(optimizer, opt.clip); optimizer.step()  with torch.optim.Adam(params, lr)
//   EMCAD/trainer.py:86,155-157           optim.AdamW(model.parameters(), lr, weight_decay=1e-4); optimizer.step()
//
// The reference walks ~950 parameter tensors twice (clamp, then the optimizer's multi-tensor lists).  Here every
// parameter, its gradient and both moments live in four flat fp32 buffers with ONE common layout (train.FlatParams), so
// the tail is a single pure stream: 16 B read (p, g, m, v) + 12 B written (p, m, v) per element, 16-byte accesses, no
// tables, no tails per tensor.  At 32.5 M parameters (PraNet-V2 Res2Net-50) that is 911 MB per step -- the
// largest-traffic launch of the whole step, bounded by HBM.
//
// The step counter lives in device memory (CUDA-graph replay must see it advance): every CTA reads it before any CTA
// can have finished, and the CTA that draws the last ticket writes the incremented value back.
#include "pv2_common.cuh"

namespace pv2 {
namespace {

constexpr int OT_THREADS = 256;
constexpr int OT_UNROLL = 4;   // float4 per thread per tile: 4 independent 16-byte loads per stream in flight

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, clip, grad_scale;
    int decoupled;       // 1: AdamW (p *= 1 - lr*wd), 0: Adam (g += wd*p)
    long long n4;        // number of float4 (n padded to a multiple of 4 by the caller's layout)
};

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void adam_elt(float& p, float g, float& m, float& v, const AdamArgs& a, float step_size, float inv_bc2_sqrt) {
    g = fminf(fmaxf(g * a.grad_scale, -a.clip), a.clip);          // allreduce mean + clip_gradient
    if (a.decoupled) p = p * (1.0f - a.lr * a.weight_decay);       // AdamW: param.mul_(1 - lr*wd)
    else g = fmaf(a.weight_decay, p, g);                           // Adam: grad.add(param, alpha=wd)
    m = m + (g - m) * (1.0f - a.beta1);                            // exp_avg.lerp_(grad, 1-beta1)
    v = fmaf(v, a.beta2, (1.0f - a.beta2) * g * g);                // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(v) * inv_bc2_sqrt + a.eps;           // (sqrt(v) / sqrt(bc2)).add_(eps)
    p = p - step_size * (m / denom);                               // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(OT_THREADS)
adam_clamp_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       long long* __restrict__ step, unsigned int* __restrict__ ticket, const AdamArgs a) {
    pdl_prologue();
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const double t = (double)(*step + 1);
        const double bc1 = 1.0 - pow((double)a.beta1, t), bc2 = 1.0 - pow((double)a.beta2, t);
        s_c[0] = (float)((double)a.lr / bc1);
        s_c[1] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const float step_size = s_c[0], inv_bc2_sqrt = s_c[1];
    const long long tile = (long long)OT_THREADS * OT_UNROLL;
    for (long long base = (long long)blockIdx.x * tile; base < a.n4; base += (long long)gridDim.x * tile) {
        float4 P[OT_UNROLL], G[OT_UNROLL], M[OT_UNROLL], V[OT_UNROLL];
#pragma unroll
        for (int u = 0; u < OT_UNROLL; ++u) {
            const long long i = base + (long long)u * OT_THREADS + threadIdx.x;
            if (i < a.n4) {
                P[u] = ld_f4(p + 4 * i); G[u] = ld_stream_f4(g + 4 * i); M[u] = ld_f4(m + 4 * i); V[u] = ld_f4(v + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < OT_UNROLL; ++u) {
            const long long i = base + (long long)u * OT_THREADS + threadIdx.x;
            if (i < a.n4) {
                adam_elt(P[u].x, G[u].x, M[u].x, V[u].x, a, step_size, inv_bc2_sqrt);
                adam_elt(P[u].y, G[u].y, M[u].y, V[u].y, a, step_size, inv_bc2_sqrt);
                adam_elt(P[u].z, G[u].z, M[u].z, V[u].z, a, step_size, inv_bc2_sqrt);
                adam_elt(P[u].w, G[u].w, M[u].w, V[u].w, a, step_size, inv_bc2_sqrt);
                *reinterpret_cast<float4*>(p + 4 * i) = P[u];     // parameters are re-read by the next forward: keep them cacheable
                st_stream_f4(m + 4 * i, M[u]);
                st_stream_f4(v + 4 * i, V[u]);
            }
        }
    }
    // advance the device-side step counter once every CTA has read it
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(ticket, 1u);
        s_last = (prev == gridDim.x - 1u) ? 1 : 0;
        if (s_last) { *ticket = 0u; *step += 1; }
    }
}

}  // namespace
}  // namespace pv2

using namespace pv2;

extern "C" int pv2_adam_clamp_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                   long long* step, unsigned int* ticket, float lr, float beta1, float beta2, float eps,
                                   float weight_decay, int decoupled, float clip, float grad_scale, void* stream) {
    PV2_CHECK(params && grads && exp_avg && exp_avg_sq && step && ticket, "adam_clamp_flat: null pointer");
    PV2_CHECK(n > 0 && n % 4 == 0, "adam_clamp_flat: n=%lld must be a positive multiple of 4 (pad the flat layout)", n);
    PV2_CHECK((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "adam_clamp_flat: buffers must be 16-byte aligned");
    PV2_CHECK(lr >= 0.0f && beta1 >= 0.0f && beta1 < 1.0f && beta2 >= 0.0f && beta2 < 1.0f && eps >= 0.0f && clip > 0.0f,
              "adam_clamp_flat: bad hyper-parameters");
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.clip = clip;
    a.grad_scale = grad_scale; a.decoupled = decoupled; a.n4 = n / 4;
    const long long tile = (long long)OT_THREADS * OT_UNROLL;
    const long long want = (a.n4 + tile - 1) / tile;
    static const int per_sm = [] {   // one resident wave, grid-stride inside it: no tail wave
        int nb = 0;
        return (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, adam_clamp_flat_kernel, OT_THREADS, 0) == cudaSuccess && nb > 0) ? nb : 2;
    }();
    const long long cap = (long long)kNumSMs * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    pv2::launch_streaming(adam_clamp_flat_kernel, dim3(grid), dim3(OT_THREADS), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, step, ticket, a);
    PV2_LAUNCH_CHECK("adam_clamp_flat");
    return 0;
}
