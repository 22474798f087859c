// Shared helpers for the pv2 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pv2.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pv2 kernels are written for sm_100a (B200) only"
#endif

namespace pv2 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PV2_CHECK(cond, ...)              \
    do {                                  \
        if (!(cond)) {                    \
            pv2::set_error(__VA_ARGS__);  \
            return 1;                     \
        }                                 \
    } while (0)

#define PV2_LAUNCH_CHECK(name)                                                   \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            pv2::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return 2;                                                            \
        }                                                                        \
        pv2::count_launch();                                                     \
    } while (0)

// Programmatic dependent launch (PDL): every pv2 kernel is launched with programmatic stream serialization, lets its
// successor start launching immediately (launch_dependents) and waits for its predecessor's memory (wait) before it
// touches global memory.  Inside a CUDA graph this removes most of the inter-kernel gap of the ~600 small launches a
// head step consists of.  PV2_PDL=0 in the environment falls back to plain stream-ordered launches.
// WHEN a kernel lets its dependents launch matters: a dependent that is launched early sits in griddepcontrol.wait holding its
// registers / shared memory / TMEM for as long as the primary runs.  With twelve chains of the head running concurrently those
// idle CTAs crowd out the CTAs that have work (a waiting 296-CTA apply kernel pins every SM's register file), so the default is
// LATE: a kernel signals when its own main work is done (pdl_done) and only the dependent's launch latency and prologue overlap
// the primary's tail.  -DPV2_PDL_EARLY=1 restores "signal first thing" (best for a single serial chain).
#ifndef PV2_PDL_EARLY
#define PV2_PDL_EARLY 0
#endif
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { if (PV2_PDL_EARLY) pdl_trigger(); pdl_wait(); }
__device__ __forceinline__ void pdl_done() { if (!PV2_PDL_EARLY) pdl_trigger(); }

bool pdl_enabled();
// integer tuning knob from the environment, read at every call (A/B measurements inside one process); `def` when unset
int tune_int(const char* name, int def);

void prefer_max_shared(const void* kernel, bool streaming = false);

template <bool STREAMING = false, typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    prefer_max_shared(reinterpret_cast<const void*>(kernel), STREAMING);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through PV2_LAUNCH_CHECK
}
// pure HBM streams with little or no shared memory (optimizer tail, loss, final upsamples): they keep the driver's default L1 split
template <typename... KArgs, typename... Args>
inline void launch_streaming(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    launch<true>(kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 16-byte load / store that do not pollute L1 (data touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4, for fp32 and bf16 storage (pointer must be 16 B / 8 B aligned)
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return ld_stream_f4(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { st_stream_f4(p, v); }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// ATen's bilinear source index (area_pixel_compute_source_index, cubic=false) and tap weights
struct Tap {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Tap bilinear_tap(int o, int in_size, float ratio, bool align_corners) {
    float src = align_corners ? ratio * (float)o : fmaxf(ratio * ((float)o + 0.5f) - 0.5f, 0.0f);
    int i0 = min((int)src, in_size - 1);
    Tap t;
    t.i0 = i0;
    t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    t.w1 = fminf(fmaxf(src - (float)i0, 0.0f), 1.0f);  // guard_index_and_lambda
    t.w0 = 1.0f - t.w1;
    return t;
}

}  // namespace pv2
