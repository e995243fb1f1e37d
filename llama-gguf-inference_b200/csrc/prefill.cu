// prefill.cu -- the batched (many-token) glue around the tcgen05 GEMM: embedding gather for a token list,
// RoPE + f16 KV-cache write for T tokens, causal GQA attention over the cache, and residual add.
// Stands in for the batch forms of ggml's get_rows / rope / cpy / flash_attn_ext / add [UPSTREAM-MEM].
// Numerics: f32 with an online softmax -- the tolerance-level path (the GEMM feeding it is bf16 x bf16).
#include <float.h>

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
    rope_kv_batch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q, k, v, nb, pos_dev, slot_dev, slot_stride, n_head, n_kv,
                                                                                          head_dim, n_rot, rope_tab, kcache, vcache);
    GGB_CHECK_LAUNCH("ggb_rope_kv_batch");
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

// Causal attention for T query tokens at positions pos0..pos0+T-1 over the f16 cache.
// CTA = one KV head x a block of query tokens: 16 warps = G query heads x (16/G) tokens, so a K/V tile staged in
// shared memory is reused by every warp.  Warp = one (head, token) query: lanes own head_dim/32 dims, online softmax.
#define AP_WARPS 16
#define AP_TILE 64

template <int HD>
__global__ void __launch_bounds__(AP_WARPS * 32) attn_prefill_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kc,
                                                                   const uint16_t* __restrict__ vc, int T, int pos0, int n_head, int n_kv,
                                                                   float* __restrict__ out) {
    constexpr int DPL = HD / 32;                 /* dims per lane: 4 (hd 128) or 2 (hd 64) */
    __shared__ __align__(16) uint16_t sk[AP_TILE][HD];
    __shared__ __align__(16) uint16_t sv[AP_TILE][HD];
    const int G = n_head / n_kv;
    const int tpb = AP_WARPS / G;                /* query tokens per CTA */
    const int kvh = blockIdx.x, tb = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp % G, tq = tb * tpb + warp / G;
    const bool live = tq < T && warp < G * tpb;
    const int head = kvh * G + g;
    const int pq = pos0 + tq;                    /* this query attends positions 0..pq */
    const int64_t kvd = (int64_t)n_kv * HD;
    const int last_tok = min(T, (tb + 1) * tpb) - 1;
    const int n_pos = pos0 + last_tok + 1;       /* positions the CTA needs */

    float qr[DPL];
    if (live) {
#pragma unroll
        for (int i = 0; i < DPL; i++) qr[i] = h2f(f2h(q[(int64_t)tq * n_head * HD + (int64_t)head * HD + lane * DPL + i]));
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    float m = -FLT_MAX, l = 0.f, acc[DPL];
#pragma unroll
    for (int i = 0; i < DPL; i++) acc[i] = 0.f;

    for (int p0 = 0; p0 < n_pos; p0 += AP_TILE) {
        const int np = min(AP_TILE, n_pos - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < np * (HD / 8); i += AP_WARPS * 32) {
            const int r = i / (HD / 8), c = i % (HD / 8);
            *reinterpret_cast<uint4*>(&sk[r][c * 8]) = *reinterpret_cast<const uint4*>(kc + (int64_t)(p0 + r) * kvd + (int64_t)kvh * HD + c * 8);
            *reinterpret_cast<uint4*>(&sv[r][c * 8]) = *reinterpret_cast<const uint4*>(vc + (int64_t)(p0 + r) * kvd + (int64_t)kvh * HD + c * 8);
        }
        __syncthreads();
        if (live) {
            const int lim = min(np, pq - p0 + 1);
            for (int r = 0; r < lim; r++) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < DPL; i++) s += h2f(sk[r][lane * DPL + i]) * qr[i];
                s = warp_sum(s) * scale;
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn), pw = __expf(s - mn);
                l = l * corr + pw;
#pragma unroll
                for (int i = 0; i < DPL; i++) acc[i] = acc[i] * corr + pw * h2f(sv[r][lane * DPL + i]);
                m = mn;
            }
        }
    }
    if (live) {
        const float inv = __fdiv_rn(1.0f, l);
#pragma unroll
        for (int i = 0; i < DPL; i++) out[(int64_t)tq * n_head * HD + (int64_t)head * HD + lane * DPL + i] = acc[i] * inv;
    }
}

extern "C" int ggb_attn_prefill(const float* q, const uint16_t* kcache, const uint16_t* vcache, int tokens, int pos0, int n_head, int n_kv,
                                int head_dim, float* out, void* stream) {
    if (tokens < 0 || pos0 < 0 || n_head <= 0 || n_kv <= 0 || n_head % n_kv) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_prefill: bad shape");
    if (tokens == 0) return GGB_OK;
    if (!q || !kcache || !vcache || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_prefill: null pointer");
    const int G = n_head / n_kv;
    if (G > AP_WARPS || (AP_WARPS % G)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_prefill: GQA group %d not supported (1, 2, 4, 8, 16)", G);
    const int tpb = AP_WARPS / G;
    dim3 grid(n_kv, (tokens + tpb - 1) / tpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) attn_prefill_kernel<128><<<grid, AP_WARPS * 32, 0, st>>>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out);
    else if (head_dim == 64) attn_prefill_kernel<64><<<grid, AP_WARPS * 32, 0, st>>>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out);
    else GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_prefill: head_dim=%d (supported: 64, 128)", head_dim);
    GGB_CHECK_LAUNCH("ggb_attn_prefill");
    return GGB_OK;
}

// embedding gather for a list of tokens: out[t][k] = dequant(token_embd row ids[t]) -- one launch per token would do,
// but prompts are long; this reuses ggb_dequant's element decode through a per-token row pointer.
extern "C" int ggb_embed_row(int type, const void* token_embd, int64_t k, const int32_t* tok_dev, float* x, void* stream);
extern "C" int ggb_embed_rows(int type, const void* token_embd, int64_t k, const int32_t* ids_dev, int tokens, float* out, void* stream) {
    if (tokens < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_embed_rows: negative token count");
    for (int t = 0; t < tokens; t++) {
        const int rc = ggb_embed_row(type, token_embd, k, ids_dev + t, out + (int64_t)t * k, stream);
        if (rc) return rc;
    }
    return GGB_OK;
}
