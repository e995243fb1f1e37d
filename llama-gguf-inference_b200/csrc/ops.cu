// ops.cu -- standalone activation quantisation and the small element-wise / reduction ops
// (ggml quantize_row_q8_K, quantize_row_q8_0, rms_norm+mul, silu*mul, argmax [UPSTREAM-MEM]).
// The decode path uses the fused forms inside gemv.cu; these entry points serve the prefill path and the
// parity tests, and share the same device code (actquant.cuh).
#include <float.h>

#include "actquant.cuh"
#include "common.cuh"

// one warp per 256-block
__global__ void quantize_q8_K_kernel(const float* __restrict__ x, int8_t* __restrict__ qs, float* __restrict__ d,
                                     int16_t* __restrict__ bsums, int64_t nblocks) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const float4* p = reinterpret_cast<const float4*>(x + b * 256 + lane * 8);
    const float4 v0 = p[0], v1 = p[1];
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float dd;
    const Q8Codes c = warp_quantize_q8_K(v, lane, dd);
    *reinterpret_cast<uint2*>(qs + b * 256 + lane * 8) = c.q;
    const int s16 = c.sum8 + __shfl_xor_sync(0xffffffffu, c.sum8, 1);
    if (!(lane & 1)) bsums[b * 16 + (lane >> 1)] = (int16_t)s16;
    if (lane == 0) d[b] = dd;
}

__global__ void quantize_q8_0_kernel(const float* __restrict__ x, int8_t* __restrict__ qs, uint16_t* __restrict__ d, int64_t nchunks8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; /* one thread per 8 elements */
    const bool live = i < nchunks8;
    const int64_t ii = live ? i : nchunks8 - 1; /* keep the whole warp in the shuffles */
    const float4* p = reinterpret_cast<const float4*>(x + ii * 8);
    const float4 v0 = p[0], v1 = p[1];
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float df;
    uint16_t db;
    const Q8Codes c = warp_quantize_q8_0(v, df, db);
    if (!live) return;
    *reinterpret_cast<uint2*>(qs + i * 8) = c.q;
    if (!(i & 3)) d[i >> 2] = db;
}

// Activation operand of the prefill GEMM: x quantised EXACTLY as the CPU path quantises it (Q8_K per 256, or Q8_0 per 32)
// and dequantised again, d * q, as f16.  The tensor-core product then differs from ggml's integer dot only by the f16
// rounding of the two operands (2^-11 each) -- feeding the GEMM the unquantised activations instead makes it a DIFFERENT
// (more precise) computation than the reference's, 1-4e-2 away from it in the logits of the synthetic models.
// `up` != nullptr: x is the gate projection and the quantised vector is silu(x) * up (ggml silu + mul fused in front of the
// quantisation: the SwiGLU output never makes a round trip through HBM as floats)
__global__ void act_fakequant_f16_kernel(const float* __restrict__ x, const float* __restrict__ up, __half* __restrict__ y, int64_t nblocks, int q8_0) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   /* one warp per 256 elements */
    if (b >= nblocks) return;
    const float4* p = reinterpret_cast<const float4*>(x + b * 256 + lane * 8);
    const float4 v0 = p[0], v1 = p[1];
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (up) {
        const float4* pu = reinterpret_cast<const float4*>(up + b * 256 + lane * 8);
        const float4 u0 = pu[0], u1 = pu[1];
        const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = silu_mul_ref(v[i], u[i]);
    }
    float dd;
    Q8Codes c;
    if (q8_0) { uint16_t db; c = warp_quantize_q8_0(v, dd, db); }
    else c = warp_quantize_q8_K(v, lane, dd);
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t w = i < 2 ? c.q.x : c.q.y;
        const int q0 = (int)(int8_t)((w >> (16 * (i & 1))) & 0xFF), q1 = (int)(int8_t)((w >> (16 * (i & 1) + 8)) & 0xFF);
        const float a = fminf(fmaxf(__fmul_rn(dd, (float)q0), -65504.0f), 65504.0f), bb = fminf(fmaxf(__fmul_rn(dd, (float)q1), -65504.0f), 65504.0f);
        const __half2 h = __floats2half2_rn(a, bb);
        out[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(y + b * 256 + lane * 8) = make_uint4(out[0], out[1], out[2], out[3]);
}

extern "C" int ggb_act_fakequant_f16(const float* x, void* y_f16, int64_t k, int m, int q8_0, void* stream) {
    if (k < 0 || (k % 256) || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_act_fakequant_f16: k=%lld must be a multiple of 256", (long long)k);
    const int64_t nb = k / 256 * m;
    if (nb == 0) return GGB_OK;
    if (!x || !y_f16 || ((uintptr_t)x & 15) || ((uintptr_t)y_f16 & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_act_fakequant_f16: null or unaligned pointer");
    act_fakequant_f16_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, nullptr, (__half*)y_f16, nb, q8_0);
    GGB_CHECK_LAUNCH("ggb_act_fakequant_f16");
    return GGB_OK;
}

extern "C" int ggb_swiglu_fakequant_f16(const float* gate, const float* up, void* y_f16, int64_t k, int m, int q8_0, void* stream) {
    if (k < 0 || (k % 256) || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_swiglu_fakequant_f16: k=%lld must be a multiple of 256", (long long)k);
    const int64_t nb = k / 256 * m;
    if (nb == 0) return GGB_OK;
    if (!gate || !up || !y_f16 || ((uintptr_t)gate & 15) || ((uintptr_t)up & 15) || ((uintptr_t)y_f16 & 15))
        GGB_FAIL(GGB_ERR_ARG, "ggb_swiglu_fakequant_f16: null or unaligned pointer");
    act_fakequant_f16_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, (cudaStream_t)stream>>>(gate, up, (__half*)y_f16, nb, q8_0);
    GGB_CHECK_LAUNCH("ggb_swiglu_fakequant_f16");
    return GGB_OK;
}

extern "C" int ggb_quantize_q8_K(const float* x, int8_t* qs, float* d, int16_t* bsums, int64_t k, int m, void* stream) {
    if (k < 0 || (k % 256) || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_quantize_q8_K: k=%lld must be a multiple of 256", (long long)k);
    const int64_t nb = k / 256 * m;
    if (nb == 0) return GGB_OK;
    if (!x || !qs || !d || !bsums) GGB_FAIL(GGB_ERR_ARG, "ggb_quantize_q8_K: null pointer");
    quantize_q8_K_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, qs, d, bsums, nb);
    GGB_CHECK_LAUNCH("ggb_quantize_q8_K");
    return GGB_OK;
}

extern "C" int ggb_quantize_q8_0(const float* x, int8_t* qs, uint16_t* d, int64_t k, int m, void* stream) {
    if (k < 0 || (k % 32) || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_quantize_q8_0: k=%lld must be a multiple of 32", (long long)k);
    const int64_t n8 = k / 8 * m;
    if (n8 == 0) return GGB_OK;
    if (!x || !qs || !d) GGB_FAIL(GGB_ERR_ARG, "ggb_quantize_q8_0: null pointer");
    quantize_q8_0_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, qs, d, n8);
    GGB_CHECK_LAUNCH("ggb_quantize_q8_0");
    return GGB_OK;
}

// ------------------------------------------------------------------ rms_norm (+gain): one CTA per row
// squares in f32, sum in f64 (as ggml's ggml_float accumulator), scale = 1/sqrtf(mean+eps) in f32.
__global__ void rms_norm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int64_t k, float eps) {
    __shared__ double red[32];
    const float* xr = x + (int64_t)blockIdx.x * k;
    float* yr = y + (int64_t)blockIdx.x * k;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) { const float v = xr[i]; s += (double)__fmul_rn(v, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const float mean = (float)(red[0] / (double)k);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + eps));
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const float v = __fmul_rn(xr[i], scale);
        yr[i] = w ? __fmul_rn(v, w[i]) : v;
    }
}

extern "C" int ggb_rms_norm(const float* x, const float* w, float* y, int64_t k, int m, float eps, void* stream) {
    if (k <= 0 || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_rms_norm: bad shape");
    if (m == 0) return GGB_OK;
    if (!x || !y) GGB_FAIL(GGB_ERR_ARG, "ggb_rms_norm: null pointer");
    rms_norm_kernel<<<m, 512, 0, (cudaStream_t)stream>>>(x, w, y, k, eps);
    GGB_CHECK_LAUNCH("ggb_rms_norm");
    return GGB_OK;
}

// Prefill glue in one pass per token row: x += add (ggml_add of the previous projection, optional), rms_norm(x) * w,
// quantised as the CPU path quantises the next matmul's activation operand, written as f16 (d * q) for ggb_gemm.
// The arithmetic is rms_norm_kernel's + act_fakequant_f16_kernel's; three launches and two f32 round trips less per use.
__global__ void __launch_bounds__(256) add_rmsnorm_fakequant_f16_kernel(float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ w,
                                                                      __half* __restrict__ y, int64_t k, float eps, int q8_0) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* xr = x + (int64_t)blockIdx.x * k;
    const float* ar = add ? add + (int64_t)blockIdx.x * k : nullptr;
    __half* yr = y + (int64_t)blockIdx.x * k;
    const int nblk = (int)(k / 256);
    double s = 0.0;
    for (int b = warp; b < nblk; b += 8) {
        float4* p = reinterpret_cast<float4*>(xr + b * 256 + lane * 8);
        float4 v0 = p[0], v1 = p[1];
        if (ar) {
            const float4* q = reinterpret_cast<const float4*>(ar + b * 256 + lane * 8);
            const float4 a0 = q[0], a1 = q[1];
            v0 = make_float4(__fadd_rn(v0.x, a0.x), __fadd_rn(v0.y, a0.y), __fadd_rn(v0.z, a0.z), __fadd_rn(v0.w, a0.w));
            v1 = make_float4(__fadd_rn(v1.x, a1.x), __fadd_rn(v1.y, a1.y), __fadd_rn(v1.z, a1.z), __fadd_rn(v1.w, a1.w));
            p[0] = v0; p[1] = v1;
        }
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) s += (double)__fmul_rn(v[i], v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[i];
    const float mean = (float)(tot / (double)k);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + eps));
    for (int b = warp; b < nblk; b += 8) {
        const float4* p = reinterpret_cast<const float4*>(xr + b * 256 + lane * 8);   /* this thread's own stores: L1 hits */
        const float4* g = reinterpret_cast<const float4*>(w + b * 256 + lane * 8);
        const float4 v0 = p[0], v1 = p[1], g0 = g[0], g1 = g[1];
        const float xv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w}, gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(xv[i], scale), gv[i]);
        float dd;
        Q8Codes c;
        if (q8_0) { uint16_t db; c = warp_quantize_q8_0(v, dd, db); }
        else c = warp_quantize_q8_K(v, lane, dd);
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t wd = i < 2 ? c.q.x : c.q.y;
            const int q0 = (int)(int8_t)((wd >> (16 * (i & 1))) & 0xFF), q1 = (int)(int8_t)((wd >> (16 * (i & 1) + 8)) & 0xFF);
            const float a = fminf(fmaxf(__fmul_rn(dd, (float)q0), -65504.0f), 65504.0f), bb = fminf(fmaxf(__fmul_rn(dd, (float)q1), -65504.0f), 65504.0f);
            const __half2 h = __floats2half2_rn(a, bb);
            out[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(yr + b * 256 + lane * 8) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

extern "C" int ggb_add_rmsnorm_fakequant_f16(float* x, const float* add, const float* w, void* y_f16, int64_t k, int m, float eps, int q8_0, void* stream) {
    if (k <= 0 || (k % 256) || m < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_add_rmsnorm_fakequant_f16: k=%lld must be a positive multiple of 256", (long long)k);
    if (m == 0) return GGB_OK;
    if (!x || !w || !y_f16 || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)y_f16 & 15) || ((uintptr_t)add & 15))
        GGB_FAIL(GGB_ERR_ARG, "ggb_add_rmsnorm_fakequant_f16: null or unaligned pointer");
    add_rmsnorm_fakequant_f16_kernel<<<m, 256, 0, (cudaStream_t)stream>>>(x, add, w, (__half*)y_f16, k, eps, q8_0);
    GGB_CHECK_LAUNCH("ggb_add_rmsnorm_fakequant_f16");
    return GGB_OK;
}

__global__ void swiglu_kernel(const float* __restrict__ g, const float* __restrict__ u, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = silu_mul_ref(g[i], u[i]);
}
extern "C" int ggb_swiglu(const float* g, const float* u, float* out, int64_t n, void* stream) {
    if (n < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_swiglu: negative size");
    if (n == 0) return GGB_OK;
    if (!g || !u || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_swiglu: null pointer");
    swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, u, out, n);
    GGB_CHECK_LAUNCH("ggb_swiglu");
    return GGB_OK;
}

// ------------------------------------------------------------------ argmax (first index of the maximum), single CTA
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__global__ void argmax_kernel(const float* __restrict__ x, int64_t n, int32_t* __restrict__ out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) argmax_combine(v, idx, x[i], (int)i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -FLT_MAX;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) *out = idx;
    }
}
extern "C" int ggb_argmax(const float* x, int64_t n, int32_t* out_idx, void* stream) {
    if (n <= 0 || !x || !out_idx) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax: bad argument");
    argmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out_idx);
    GGB_CHECK_LAUNCH("ggb_argmax");
    return GGB_OK;
}

// ------------------------------------------------------------------ greedy sampler tail: reduce GEMV partials, emit token,
// advance position/step, gather next embedding row.  One CTA.
__global__ void argmax_next_kernel(const float* __restrict__ part_val, const int32_t* __restrict__ part_idx, int n_part,
                                   int32_t* tok, int32_t* pos, int32_t* step, int32_t* out_tokens, int32_t out_cap) {
    __shared__ float sv[32];
    __shared__ int si[32];
    pdl_wait();
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) argmax_combine(v, idx, part_val[i], part_idx[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -FLT_MAX;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) {
            *tok = idx;
            const int s = *step;
            if (s < out_cap) out_tokens[s] = idx;
            *step = s + 1;
            *pos = *pos + 1;
        }
    }
}

extern "C" int ggb_argmax_next(const float* part_val, const int32_t* part_idx, int n_part, int32_t* tok_dev,
                               int32_t* pos_dev, int32_t* step_dev, int32_t* out_tokens, int32_t out_cap,
                               int emb_type, const void* token_embd, int64_t k, float* x, void* stream) {
    if (!part_val || !part_idx || n_part <= 0 || !tok_dev || !pos_dev || !step_dev || !out_tokens)
        GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_next: bad argument");
    argmax_next_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(part_val, part_idx, n_part, tok_dev, pos_dev, step_dev, out_tokens, out_cap);
    GGB_CHECK_LAUNCH("ggb_argmax_next");
    if (token_embd) return ggb_embed_row(emb_type, token_embd, k, tok_dev, x, stream);
    return GGB_OK;
}


// ------------------------------------------------------------------ tensor-parallel glue
__global__ void residual_add_f64_kernel(float* __restrict__ x, const double* __restrict__ y, int64_t n) {
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fadd_rn(x[i], (float)y[i]);
}
extern "C" int ggb_residual_add_f64(float* x, const double* y64, int64_t n, int use_pdl, void* stream) {
    if (n < 0 || (n && (!x || !y64))) GGB_FAIL(GGB_ERR_ARG, "ggb_residual_add_f64: bad argument");
    if (n == 0) return GGB_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, residual_add_f64_kernel, x, y64, n));
    return GGB_OK;
}

// sortable key: larger value first, then smaller index.  Float bits are mapped to a monotone u32, the top bit of
// the u64 is flipped so that SIGNED 64-bit MAX (what the collective offers) orders the keys correctly.
__device__ __forceinline__ long long argmax_key(float v, int idx) {
    const unsigned b = __float_as_uint(v);
    const unsigned mono = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    const unsigned long long k = ((unsigned long long)mono << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
    return (long long)(k ^ 0x8000000000000000ull);
}
__global__ void argmax_pack_kernel(const float* __restrict__ part_val, const int32_t* __restrict__ part_idx, int n_part,
                                   int32_t row_offset, long long* __restrict__ key) {
    __shared__ float sv[32];
    __shared__ int si[32];
    pdl_wait();
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) argmax_combine(v, idx, part_val[i], part_idx[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -FLT_MAX;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) *key = (idx == 0x7fffffff) ? argmax_key(-FLT_MAX, 0x7ffffffe) : argmax_key(v, idx + row_offset);
    }
}
extern "C" int ggb_argmax_pack(const float* part_val, const int32_t* part_idx, int n_part, int32_t row_offset, int64_t* key, void* stream) {
    if (!part_val || !part_idx || n_part <= 0 || !key) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_pack: bad argument");
    argmax_pack_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(part_val, part_idx, n_part, row_offset, (long long*)key);
    GGB_CHECK_LAUNCH("ggb_argmax_pack");
    return GGB_OK;
}
__global__ void argmax_unpack_kernel(const long long* __restrict__ key, int32_t* tok, int32_t* pos, int32_t* step, int32_t* out_tokens, int32_t out_cap) {
    const unsigned long long k = (unsigned long long)(*key) ^ 0x8000000000000000ull;
    const int idx = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
    *tok = idx;
    const int s = *step;
    if (s < out_cap) out_tokens[s] = idx;
    *step = s + 1;
    *pos = *pos + 1;
}
extern "C" int ggb_argmax_unpack_next(const int64_t* key, int32_t* tok_dev, int32_t* pos_dev, int32_t* step_dev, int32_t* out_tokens,
                                      int32_t out_cap, int emb_type, const void* token_embd, int64_t k, float* x, void* stream) {
    if (!key || !tok_dev || !pos_dev || !step_dev || !out_tokens) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_unpack_next: bad argument");
    argmax_unpack_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const long long*)key, tok_dev, pos_dev, step_dev, out_tokens, out_cap);
    GGB_CHECK_LAUNCH("ggb_argmax_unpack_next");
    if (token_embd) return ggb_embed_row(emb_type, token_embd, k, tok_dev, x, stream);
    return GGB_OK;
}

// ------------------------------------------------------------------ sampler candidates: the k largest logits of every row
// A sampled request (temperature > 0, top-k) needs the k best logits of a 128 K row, not the row: reading it back and
// partitioning it on the host costs ~1.5 ms per token and sequence -- more than the whole decode step.  One CTA per row finds the
// k-th largest value EXACTLY by building its order-preserving 32-bit key two bits per pass (16 passes over the L2-resident row,
// three nested thresholds counted per pass), then emits every element >= that value (k of them, more on ties, at most `cap`).
// The host sorts the few candidates and runs the unchanged top-p / min-p / temperature / multinomial chain on them.
__device__ __forceinline__ uint32_t topk_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);      /* unsigned order == float order (-0 < +0, NaN above +Inf) */
}

#define TOPK_THREADS 1024
__global__ void __launch_bounds__(TOPK_THREADS) topk_rows_kernel(const float* __restrict__ x, int64_t n, int k, int cap, float* __restrict__ out_val,
                                                                 int32_t* __restrict__ out_idx, int32_t* __restrict__ out_cnt) {
    __shared__ int s_c[3][TOPK_THREADS / 32];
    __shared__ int s_tot[3];
    __shared__ int s_n;
    const float* xr = x + (int64_t)blockIdx.x * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n4 = ((uintptr_t)xr & 15) ? 0 : n / 4;     /* float4 body (rows of a [nb][n] matrix with n % 4 == 0 are aligned) */
    uint32_t T = 0;
    for (int bit = 30; bit >= 0; bit -= 2) {
        const uint32_t t1 = T | (1u << bit), t2 = T | (2u << bit), t3 = T | (3u << bit);
        int c1 = 0, c2 = 0, c3 = 0;
        for (int64_t i = tid; i < n4; i += TOPK_THREADS) {
            const float4 v = reinterpret_cast<const float4*>(xr)[i];
            const uint32_t ka = topk_key(v.x), kb = topk_key(v.y), kc = topk_key(v.z), kd = topk_key(v.w);
            c1 += (ka >= t1) + (kb >= t1) + (kc >= t1) + (kd >= t1);
            c2 += (ka >= t2) + (kb >= t2) + (kc >= t2) + (kd >= t2);
            c3 += (ka >= t3) + (kb >= t3) + (kc >= t3) + (kd >= t3);
        }
        for (int64_t i = 4 * n4 + tid; i < n; i += TOPK_THREADS) {
            const uint32_t ka = topk_key(xr[i]);
            c1 += ka >= t1; c2 += ka >= t2; c3 += ka >= t3;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
        }
        if (lane == 0) { s_c[0][warp] = c1; s_c[1][warp] = c2; s_c[2][warp] = c3; }
        __syncthreads();
        if (warp < 3) {
            int v = s_c[warp][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_tot[warp] = v;
        }
        __syncthreads();
        if (s_tot[2] >= k) T = t3; else if (s_tot[1] >= k) T = t2; else if (s_tot[0] >= k) T = t1;
        __syncthreads();
    }
    // T is the key of the k-th largest element (count(key >= T) >= k, count(key > T) < k): emit everything >= T
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int64_t i = tid; i < n; i += TOPK_THREADS) {
        const float v = xr[i];
        if (topk_key(v) >= T) {
            const int p = atomicAdd(&s_n, 1);
            if (p < cap) { out_val[(int64_t)blockIdx.x * cap + p] = v; out_idx[(int64_t)blockIdx.x * cap + p] = (int32_t)i; }
        }
    }
    __syncthreads();
    if (tid == 0) out_cnt[blockIdx.x] = s_n;      /* may exceed cap (ties at the k-th value): the caller then falls back to the whole row */
}

extern "C" int ggb_topk_rows(const float* x, int64_t n, int nb, int k, int cap, float* out_val, int32_t* out_idx, int32_t* out_cnt, void* stream) {
    if (n <= 0 || nb < 0 || k <= 0 || k > n || cap < k || (nb && (!x || !out_val || !out_idx || !out_cnt)))
        GGB_FAIL(GGB_ERR_ARG, "ggb_topk_rows: bad argument (n=%lld nb=%d k=%d cap=%d)", (long long)n, nb, k, cap);
    if (nb == 0) return GGB_OK;
    topk_rows_kernel<<<nb, TOPK_THREADS, 0, (cudaStream_t)stream>>>(x, n, k, cap, out_val, out_idx, out_cnt);
    GGB_CHECK_LAUNCH("ggb_topk_rows");
    return GGB_OK;
}

// x[b][idx[b][j]] for a few indices per row (the logits of the tokens in a request's penalty window); idx < 0 or >= n yields 0
__global__ void gather_rows_kernel(const float* __restrict__ x, int64_t n, const int32_t* __restrict__ idx, int m, float* __restrict__ out) {
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int32_t i = idx[(int64_t)b * m + j];
        out[(int64_t)b * m + j] = (i >= 0 && i < n) ? x[(int64_t)b * n + i] : 0.f;
    }
}

extern "C" int ggb_gather_rows(const float* x, int64_t n, int nb, const int32_t* idx, int m, float* out, void* stream) {
    if (n <= 0 || nb < 0 || m < 0 || (nb && m && (!x || !idx || !out))) GGB_FAIL(GGB_ERR_ARG, "ggb_gather_rows: bad argument");
    if (nb == 0 || m == 0) return GGB_OK;
    gather_rows_kernel<<<nb, 128, 0, (cudaStream_t)stream>>>(x, n, idx, m, out);
    GGB_CHECK_LAUNCH("ggb_gather_rows");
    return GGB_OK;
}
