// prefill.cu -- the batched (many-token) glue around the tcgen05 GEMM: embedding gather for a token list,
// RoPE + f16 KV-cache write for T tokens, causal GQA attention over the cache, and residual add.
// Stands in for the batch forms of ggml's get_rows / rope / cpy / flash_attn_ext / add [UPSTREAM-MEM].
// Numerics: f32 with an online softmax -- the tolerance-level path (the GEMM feeding it is bf16 x bf16).
#include <cuda_fp16.h>
#include <float.h>

#include <stdlib.h>

#include "common.cuh"

__global__ void add_f32_kernel(float* __restrict__ x, const float* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fadd_rn(x[i], y[i]);
}
extern "C" int ggb_add_f32(float* x, const float* y, int64_t n, void* stream) {
    if (n < 0 || (n && (!x || !y))) GGB_FAIL(GGB_ERR_ARG, "ggb_add_f32: bad argument");
    if (n == 0) return GGB_OK;
    add_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    GGB_CHECK_LAUNCH("ggb_add_f32");
    return GGB_OK;
}

// q [T][n_head*hd] rotated in place; k rotated and written as f16 to kcache[pos0+t]; v written as f16 to vcache[pos0+t].
// One thread per adjacent pair.
__global__ void rope_kv_prefill_kernel(float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int T, int pos0,
                                       int n_head, int n_kv, int hd, int n_rot, const float* __restrict__ tab,
                                       uint16_t* __restrict__ kc, uint16_t* __restrict__ vc) {
    const int qd = n_head * hd, kvd = n_kv * hd;
    const int pairs_per_tok = (qd + 2 * kvd) / 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)T * pairs_per_tok) return;
    const int t = (int)(idx / pairs_per_tok), p = (int)(idx - (int64_t)t * pairs_per_tok);
    const int e = 2 * p, pos = pos0 + t;
    const float* rt = tab + (int64_t)pos * n_rot;
    if (e < qd) {
        const int j = e % hd;
        float* x = q + (int64_t)t * qd + e;
        if (j < n_rot) {
            const float c = rt[j], s = rt[j + 1], a = x[0], b = x[1];
            x[0] = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s));
            x[1] = __fadd_rn(__fmul_rn(a, s), __fmul_rn(b, c));
        }
    } else if (e < qd + kvd) {
        const int r = e - qd, j = r % hd;
        float a = k[(int64_t)t * kvd + r], b = k[(int64_t)t * kvd + r + 1];
        if (j < n_rot) {
            const float c = rt[j], s = rt[j + 1], a0 = a, b0 = b;
            a = __fsub_rn(__fmul_rn(a0, c), __fmul_rn(b0, s));
            b = __fadd_rn(__fmul_rn(a0, s), __fmul_rn(b0, c));
        }
        *reinterpret_cast<uint32_t*>(kc + (int64_t)pos * kvd + r) = (uint32_t)f2h(a) | ((uint32_t)f2h(b) << 16);
    } else {
        const int r = e - qd - kvd;
        const float a = v[(int64_t)t * kvd + r], b = v[(int64_t)t * kvd + r + 1];
        *reinterpret_cast<uint32_t*>(vc + (int64_t)pos * kvd + r) = (uint32_t)f2h(a) | ((uint32_t)f2h(b) << 16);
    }
}

extern "C" int ggb_rope_kv_prefill(float* q, const float* k, const float* v, int tokens, int pos0, int n_head, int n_kv, int head_dim,
                                   int n_rot, const float* rope_tab, uint16_t* kcache, uint16_t* vcache, void* stream) {
    if (tokens < 0 || pos0 < 0 || n_head <= 0 || n_kv <= 0 || head_dim <= 0 || (head_dim & 1) || (n_rot & 1) || n_rot > head_dim)
        GGB_FAIL(GGB_ERR_ARG, "ggb_rope_kv_prefill: bad shape");
    if (tokens == 0) return GGB_OK;
    if (!q || !k || !v || !rope_tab || !kcache || !vcache) GGB_FAIL(GGB_ERR_ARG, "ggb_rope_kv_prefill: null pointer");
    const int64_t total = (int64_t)tokens * ((n_head + 2 * n_kv) * head_dim / 2);
    rope_kv_prefill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q, k, v, tokens, pos0, n_head, n_kv, head_dim,
                                                                                            n_rot, rope_tab, kcache, vcache);
    GGB_CHECK_LAUNCH("ggb_rope_kv_prefill");
    return GGB_OK;
}

// The same rotation for a decode batch: token b sits at position pos[b] of slot slot[b]'s cache (all slots in one
// allocation, slot_stride elements apart).  pos[b] < 0 = idle entry.
__global__ void rope_kv_batch_kernel(float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int nb,
                                     const int32_t* __restrict__ pos_dev, const int32_t* __restrict__ slot_dev, int64_t slot_stride,
                                     int n_head, int n_kv, int hd, int n_rot, const float* __restrict__ tab,
                                     uint16_t* __restrict__ kc, uint16_t* __restrict__ vc) {
    const int qd = n_head * hd, kvd = n_kv * hd;
    const int pairs_per_tok = (qd + 2 * kvd) / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();      /* programmatic launch (use_pdl): the attention behind may become resident; it waits for this grid */
    pdl_wait();                   /* q, k, v come from the launch in front */
    if (idx >= nb * pairs_per_tok) return;
    const int t = idx / pairs_per_tok, p = idx - t * pairs_per_tok;
    const int pos = pos_dev[t];
    if (pos < 0) return;
    const int64_t cbase = (int64_t)slot_dev[t] * slot_stride + (int64_t)pos * kvd;
    const int e = 2 * p;
    const float* rt = tab + (int64_t)pos * n_rot;
    if (e < qd) {
        const int j = e % hd;
        float* x = q + (int64_t)t * qd + e;
        if (j < n_rot) {
            const float c = rt[j], s = rt[j + 1], a = x[0], b = x[1];
            x[0] = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s));
            x[1] = __fadd_rn(__fmul_rn(a, s), __fmul_rn(b, c));
        }
    } else if (e < qd + kvd) {
        const int r = e - qd, j = r % hd;
        float a = k[(int64_t)t * kvd + r], b = k[(int64_t)t * kvd + r + 1];
        if (j < n_rot) {
            const float c = rt[j], s = rt[j + 1], a0 = a, b0 = b;
            a = __fsub_rn(__fmul_rn(a0, c), __fmul_rn(b0, s));
            b = __fadd_rn(__fmul_rn(a0, s), __fmul_rn(b0, c));
        }
        *reinterpret_cast<uint32_t*>(kc + cbase + r) = (uint32_t)f2h(a) | ((uint32_t)f2h(b) << 16);
    } else {
        const int r = e - qd - kvd;
        const float a = v[(int64_t)t * kvd + r], b = v[(int64_t)t * kvd + r + 1];
        *reinterpret_cast<uint32_t*>(vc + cbase + r) = (uint32_t)f2h(a) | ((uint32_t)f2h(b) << 16);
    }
}

extern "C" int ggb_rope_kv_batch(float* q, const float* k, const float* v, int nb, const int32_t* pos_dev, const int32_t* slot_dev,
                                 int64_t slot_stride, int n_head, int n_kv, int head_dim, int n_rot, const float* rope_tab,
                                 uint16_t* kcache, uint16_t* vcache, void* stream) {
    if (nb < 0 || n_head <= 0 || n_kv <= 0 || head_dim <= 0 || (head_dim & 1) || (n_rot & 1) || n_rot > head_dim || slot_stride < 0)
        GGB_FAIL(GGB_ERR_ARG, "ggb_rope_kv_batch: bad shape");
    if (nb == 0) return GGB_OK;
    if (!q || !k || !v || !pos_dev || !slot_dev || !rope_tab || !kcache || !vcache) GGB_FAIL(GGB_ERR_ARG, "ggb_rope_kv_batch: null pointer");
    const int64_t total = (int64_t)nb * ((n_head + 2 * n_kv) * head_dim / 2);
    // GGB_BATCH_PDL=1 (read per call; the launches are captured into graphs): programmatic dependent launch, like the kernels either side
    const char* pe = getenv("GGB_BATCH_PDL");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((total + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pe && *pe && atoi(pe) != 0) ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, rope_kv_batch_kernel, q, k, v, nb, pos_dev, slot_dev, slot_stride, n_head, n_kv, head_dim, n_rot, rope_tab, kcache, vcache));
    return GGB_OK;
}

// first index of the maximum of every row of x [nb][n]: one CTA per row
__global__ void argmax_rows_kernel(const float* __restrict__ x, int64_t n, int32_t* __restrict__ out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const float* xr = x + (int64_t)blockIdx.x * n;
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float o = xr[i];
        if (o > v || (o == v && (int)i < idx)) { v = o; idx = (int)i; }
    }
    auto comb = [&](float ov, int oi) { if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; } };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) comb(__shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -FLT_MAX;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) comb(__shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) out[blockIdx.x] = idx;
    }
}
extern "C" int ggb_argmax_rows(const float* x, int64_t n, int nb, int32_t* out_idx, void* stream) {
    if (n <= 0 || nb < 0 || (nb && (!x || !out_idx))) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_rows: bad argument");
    if (nb == 0) return GGB_OK;
    argmax_rows_kernel<<<nb, 1024, 0, (cudaStream_t)stream>>>(x, n, out_idx);
    GGB_CHECK_LAUNCH("ggb_argmax_rows");
    return GGB_OK;
}

// Vocabulary-sharded arg-max of a batch (tensor parallel): every rank packs (value, global index) of each row of ITS logits
// shard into one sortable signed 64-bit key (ops.cu: same key as the single-sequence path), the keys are MAX-all-reduced,
// and the index is unpacked.  First index wins ties, exactly as on one GPU.
__global__ void argmax_rows_key_kernel(const float* __restrict__ x, int64_t n, int32_t row_offset, long long* __restrict__ keys) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const float* xr = x + (int64_t)blockIdx.x * n;
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float o = xr[i];
        if (o > v || (o == v && (int)i < idx)) { v = o; idx = (int)i; }
    }
    auto comb = [&](float ov, int oi) { if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; } };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) comb(__shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -FLT_MAX;
        idx = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) comb(__shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) {
            if (idx == 0x7fffffff) { v = -FLT_MAX; idx = 0x7ffffffe - row_offset; }
            const unsigned b = __float_as_uint(v);
            const unsigned mono = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
            const unsigned long long k = ((unsigned long long)mono << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(idx + row_offset));
            keys[blockIdx.x] = (long long)(k ^ 0x8000000000000000ull);
        }
    }
}
__global__ void argmax_keys_unpack_kernel(const long long* __restrict__ keys, int nb, int32_t* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const unsigned long long k = (unsigned long long)keys[b] ^ 0x8000000000000000ull;
    out[b] = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
}
extern "C" int ggb_argmax_rows_key(const float* x, int64_t n, int nb, int32_t row_offset, int64_t* keys, void* stream) {
    if (n <= 0 || nb < 0 || (nb && (!x || !keys))) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_rows_key: bad argument");
    if (nb == 0) return GGB_OK;
    argmax_rows_key_kernel<<<nb, 1024, 0, (cudaStream_t)stream>>>(x, n, row_offset, (long long*)keys);
    GGB_CHECK_LAUNCH("ggb_argmax_rows_key");
    return GGB_OK;
}
extern "C" int ggb_argmax_keys_unpack(const int64_t* keys, int nb, int32_t* out_idx, void* stream) {
    if (nb < 0 || (nb && (!keys || !out_idx))) GGB_FAIL(GGB_ERR_ARG, "ggb_argmax_keys_unpack: bad argument");
    if (nb == 0) return GGB_OK;
    argmax_keys_unpack_kernel<<<(nb + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const long long*)keys, nb, out_idx);
    GGB_CHECK_LAUNCH("ggb_argmax_keys_unpack");
    return GGB_OK;
}

// Causal attention for T query tokens at positions pos0..pos0+T-1 over the f16 cache, flash-attention style on the
// tensor cores (mma.sync m16n8k16, f16 operands, f32 accumulate): CTA = one query head x 64 query tokens, 4 warps of
// 16 query rows each; K/V tiles of 64 positions are staged in shared memory by cp.async; S = Q.K^T and O += P.V are
// MMAs fed by ldmatrix (K non-transposed, V transposed); online softmax in registers.  q is rounded to f16 first (as
// the decode kernel and the CPU path do for the K.Q operand).  Tolerance-level numerics, like the GEMM feeding it.
#define AP_BM 64           /* query tokens per CTA */
#define AP_BN 64           /* cache positions per tile */
#define AP_WARPS 4

__device__ __forceinline__ void ap_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ap_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ap_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ap_pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int HD>
__global__ void __launch_bounds__(AP_WARPS * 32) attn_prefill_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kc,
                                                                   const uint16_t* __restrict__ vc, int T, int pos0, int n_head, int n_kv,
                                                                   float* __restrict__ out) {
    constexpr int LD = HD + 8;                   /* padded row (halves): 16-byte rows land in distinct banks for ldmatrix */
    constexpr int KS = HD / 16;                  /* k-steps of Q.K^T */
    constexpr int NT = HD / 8;                   /* n-tiles (8 dims) of the output */
    extern __shared__ __align__(16) uint16_t ap_sm[];
    uint16_t* sq = ap_sm;                        /* [AP_BM][LD] */
    uint16_t* sk = sq + AP_BM * LD;              /* [AP_BN][LD] */
    uint16_t* sv = sk + AP_BN * LD;              /* [AP_BN][LD] */
    const int head = blockIdx.x, qb = blockIdx.y;
    const int kvh = head / (n_head / n_kv);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int q0 = qb * AP_BM;
    const int64_t qd = (int64_t)n_head * HD, kvd = (int64_t)n_kv * HD;

    // Q tile -> shared memory as f16 (rows beyond T are zero)
    for (int i = tid; i < AP_BM * (HD / 4); i += AP_WARPS * 32) {
        const int r = i / (HD / 4), c4 = i % (HD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < T) v = *reinterpret_cast<const float4*>(q + (int64_t)(q0 + r) * qd + (int64_t)head * HD + c4 * 4);
        uint2 pk;
        pk.x = ap_pack_h2(v.x, v.y); pk.y = ap_pack_h2(v.z, v.w);
        *reinterpret_cast<uint2*>(sq + r * LD + c4 * 4) = pk;
    }
    __syncthreads();
    uint32_t qa[KS][4];
    {
        const int m = lane >> 3, rr = lane & 7;  /* matrix m of the x4 load: rows +8*(m&1), cols +8*(m>>1) */
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(sq + (warp * 16 + 8 * (m & 1) + rr) * LD + 8 * (m >> 1));
#pragma unroll
        for (int ks = 0; ks < KS; ks++) ap_ldsm_x4(base + ks * 32, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; n++) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int tq0 = q0 + warp * 16 + g, tq1 = tq0 + 8;                 /* this thread's two query tokens */
    const int lim0 = pos0 + tq0, lim1 = pos0 + tq1;                    /* last position each may attend */
    const int n_pos = pos0 + min(T, q0 + AP_BM);                       /* positions the CTA needs */

    for (int p0 = 0; p0 < n_pos; p0 += AP_BN) {
        __syncthreads();                                               /* previous tile fully consumed */
        for (int i = tid; i < AP_BN * (HD / 8); i += AP_WARPS * 32) {
            const int r = i / (HD / 8), c8 = i % (HD / 8);
            const bool live = p0 + r < n_pos;
            const int64_t off = (int64_t)(live ? p0 + r : 0) * kvd + (int64_t)kvh * HD + c8 * 8;
            const uint32_t dk = (uint32_t)__cvta_generic_to_shared(sk + r * LD + c8 * 8), dv = (uint32_t)__cvta_generic_to_shared(sv + r * LD + c8 * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(kc + off), "r"(live ? 16 : 0) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(vc + off), "r"(live ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (p0 > pos0 + q0 + warp * 16 + 15) continue;                 /* the whole tile is in this warp's future */

        // S = Q.K^T : 16 query rows x 64 positions per warp
        float sc[AP_BN / 8][4];
#pragma unroll
        for (int n = 0; n < AP_BN / 8; n++) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; }
        {
            const int m = lane >> 3, rr = lane & 7;                    /* matrix m: positions n*8 + rr, dims +8*m */
#pragma unroll
            for (int n = 0; n < AP_BN / 8; n++) {
                const uint32_t base = (uint32_t)__cvta_generic_to_shared(sk + (n * 8 + rr) * LD + 8 * m);
#pragma unroll
                for (int kp = 0; kp < KS / 2; kp++) {                  /* two k-steps per x4 load */
                    uint32_t b0, b1, b2, b3;
                    ap_ldsm_x4(base + kp * 64, b0, b1, b2, b3);
                    ap_mma(sc[n], qa[2 * kp][0], qa[2 * kp][1], qa[2 * kp][2], qa[2 * kp][3], b0, b1);
                    ap_mma(sc[n], qa[2 * kp + 1][0], qa[2 * kp + 1][1], qa[2 * kp + 1][2], qa[2 * kp + 1][3], b2, b3);
                }
            }
        }
        // scale, causal mask, online softmax (rows g and g+8; a row's 64 scores live in the 4 lanes of a quad)
        float mx0 = mrow[0], mx1 = mrow[1];
#pragma unroll
        for (int n = 0; n < AP_BN / 8; n++) {
            const int pc = p0 + n * 8 + 2 * t4;
            sc[n][0] = (pc <= lim0) ? sc[n][0] * scale : -INFINITY;
            sc[n][1] = (pc + 1 <= lim0) ? sc[n][1] * scale : -INFINITY;
            sc[n][2] = (pc <= lim1) ? sc[n][2] * scale : -INFINITY;
            sc[n][3] = (pc + 1 <= lim1) ? sc[n][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        // position 0 is visible to every query, so after the first tile the running maxima are finite
        const float c0 = __expf(mrow[0] - mx0), c1 = __expf(mrow[1] - mx1);
        mrow[0] = mx0; mrow[1] = mx1;
        lrow[0] *= c0; lrow[1] *= c1;
#pragma unroll
        for (int n = 0; n < NT; n++) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
        uint32_t pa[AP_BN / 16][4];                                    /* P as f16 A-fragments: k-step j = n-tiles 2j, 2j+1 */
#pragma unroll
        for (int n = 0; n < AP_BN / 8; n++) {
            const float e0 = __expf(sc[n][0] - mx0), e1 = __expf(sc[n][1] - mx0), e2 = __expf(sc[n][2] - mx1), e3 = __expf(sc[n][3] - mx1);
            lrow[0] += e0 + e1; lrow[1] += e2 + e3;
            pa[n >> 1][(n & 1) * 2 + 0] = ap_pack_h2(e0, e1);
            pa[n >> 1][(n & 1) * 2 + 1] = ap_pack_h2(e2, e3);
        }
        // O += P.V : V fragments by transposed ldmatrix (matrix m: positions +8*(m&1), dims +8*(m>>1))
        {
            const int m = lane >> 3, rr = lane & 7;
#pragma unroll
            for (int j = 0; j < AP_BN / 16; j++) {
                const uint32_t base = (uint32_t)__cvta_generic_to_shared(sv + (j * 16 + 8 * (m & 1) + rr) * LD + 8 * (m >> 1));
#pragma unroll
                for (int np = 0; np < NT / 2; np++) {                  /* two output n-tiles per x4 load */
                    uint32_t b0, b1, b2, b3;
                    ap_ldsm_x4_t(base + np * 32, b0, b1, b2, b3);
                    ap_mma(o[2 * np], pa[j][0], pa[j][1], pa[j][2], pa[j][3], b0, b1);
                    ap_mma(o[2 * np + 1], pa[j][0], pa[j][1], pa[j][2], pa[j][3], b2, b3);
                }
            }
        }
    }
    lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 1); lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 2);
    lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 1); lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 2);
    const float i0 = __fdiv_rn(1.0f, lrow[0]), i1 = __fdiv_rn(1.0f, lrow[1]);
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const int d = n * 8 + 2 * t4;
        if (tq0 < T) *reinterpret_cast<float2*>(out + (int64_t)tq0 * qd + (int64_t)head * HD + d) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (tq1 < T) *reinterpret_cast<float2*>(out + (int64_t)tq1 * qd + (int64_t)head * HD + d) = make_float2(o[n][2] * i1, o[n][3] * i1);
    }
}

template <int HD>
static int launch_attn_prefill(const float* q, const uint16_t* kc, const uint16_t* vc, int tokens, int pos0, int n_head, int n_kv, float* out,
                               cudaStream_t st) {
    const size_t smem = (size_t)(AP_BM + 2 * AP_BN) * (HD + 8) * sizeof(uint16_t);
    static bool attr = false;
    if (!attr) {
        GGB_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(n_head, (tokens + AP_BM - 1) / AP_BM);
    attn_prefill_kernel<HD><<<grid, AP_WARPS * 32, smem, st>>>(q, kc, vc, tokens, pos0, n_head, n_kv, out);
    GGB_CHECK_LAUNCH("ggb_attn_prefill");
    return GGB_OK;
}

// prefill_tc.cu: the tcgen05 / TMEM / TMA kernel (default); returns 1 when it does not apply
int ggb_attn_prefill_tc(const float* q, const uint16_t* kcache, const uint16_t* vcache, int tokens, int pos0, int n_head, int n_kv,
                        int head_dim, float* out, void* stream);

extern "C" int ggb_attn_prefill(const float* q, const uint16_t* kcache, const uint16_t* vcache, int tokens, int pos0, int n_head, int n_kv,
                                int head_dim, float* out, void* stream) {
    if (tokens < 0 || pos0 < 0 || n_head <= 0 || n_kv <= 0 || n_head % n_kv) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_prefill: bad shape");
    if (tokens == 0) return GGB_OK;
    if (!q || !kcache || !vcache || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_prefill: null pointer");
    {
        const int rc = ggb_attn_prefill_tc(q, kcache, vcache, tokens, pos0, n_head, n_kv, head_dim, out, stream);
        if (rc != 1) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return launch_attn_prefill<128>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out, st);
    if (head_dim == 64) return launch_attn_prefill<64>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out, st);
    GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_prefill: head_dim=%d (supported: 64, 128)", head_dim);
}
