// common.cuh -- shared device/host helpers for libggufb200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ggufb200.h"

#define GGB_WARP 32
#define QK_K 256

// ---- error plumbing (thread-local message, negative return codes; no exceptions across the C-ABI)
void ggb_set_error(const char* fmt, ...);
#define GGB_FAIL(code, ...)          \
    do {                             \
        ggb_set_error(__VA_ARGS__);  \
        return (code);               \
    } while (0)
#define GGB_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) GGB_FAIL(GGB_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    } while (0)
#define GGB_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) GGB_FAIL(GGB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

static inline int ggb_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---- device helpers
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// streaming 128-bit load: weights are read exactly once per token, keep them out of L1
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_cached(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// u8 x s8 -> s32 dot of four byte lanes (weights are unsigned fields, activations signed int8)
__device__ __forceinline__ int dp4a_us(uint32_t a_u8, uint32_t b_s8, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8), "r"(b_s8), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_ss(uint32_t a_s8, uint32_t b_s8, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s8), "r"(b_s8), "r"(c));
    return d;
}

// exp() as a fixed sequence of IEEE operations, bit-identical with oracle/ggml_ref.c:gref_exp_ref
// (libm's and CUDA's expf differ by ulps, which would make softmax/SiLU parity a matter of luck).
__device__ __forceinline__ float exp_ref(float x) {
    if (x < -103.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = rintf(__fmul_rn(x, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.42860682030941723212e-6f, r);
    float p = 1.0f / 5040.0f;
    p = __fmaf_rn(p, r, 1.0f / 720.0f);
    p = __fmaf_rn(p, r, 1.0f / 120.0f);
    p = __fmaf_rn(p, r, 1.0f / 24.0f);
    p = __fmaf_rn(p, r, 1.0f / 6.0f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    int ni = (int)n;
    if (ni < -126) { p = __fmul_rn(p, 5.42101086242752217e-20f); ni += 64; }
    return __fmul_rn(p, __int_as_float((ni + 127) << 23));
}
__device__ __forceinline__ float silu_mul_ref(float g, float u) {
    return __fmul_rn(__fdiv_rn(g, __fadd_rn(1.0f, exp_ref(-g))), u);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Programmatic dependent launch (PDL): a kernel may start while its predecessor drains; it must not touch
// the predecessor's outputs before pdl_wait().  Both are no-ops when the launch carries no PDL attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
