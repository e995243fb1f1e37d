// gemv.cu -- K1: fused dequant-GEMV for batch-1 decode (the kernel that moves ~99.9 % of the bytes of a
// decoded token).  Stands in for ggml's mul_mat with one activation column:
//   quantize_row_q8_K / quantize_row_q8_0  +  ggml_vec_dot_{q4_K,q6_K}_q8_K / ggml_vec_dot_q8_0_q8_0
// [UPSTREAM-MEM: ggml-quants.c, ggml-cpu/quants.c], plus the element-wise ops either side of it
// (rms_norm*gain in front; residual add, SwiGLU, RoPE + KV-cache write, arg-max partials behind).
//
// Shape of the kernel (one launch = one "phase" of the layer, up to 3 weight matrices sharing the input):
//   grid  = one CTA per SM (x ctas_per_sm), 512 threads; CTA c owns rows [rows*c/G, rows*(c+1)/G) of every
//           segment, i.e. ONE contiguous byte range of each weight matrix -> HBM is streamed exactly once;
//   prologue: every CTA redundantly normalises + quantises the K-vector into shared memory (int8 codes in a
//           bank-swizzled order, per-block scales, per-16 sums).  K <= 28672 floats come from L2; doing it per
//           CTA removes a whole launch + grid-wide dependency per phase;
//   main loop: a warp takes (row, K-tile) items; lane u owns unit u of the tile (64 weights): 2-4 coalesced
//           128-bit streaming loads (ld.global.nc.L1::no_allocate) issued one item ahead, in-register
//           unpacking of the 6-bit scales/mins, dp4a integer dots against the int8 activations (identical
//           integers to the CPU path), one f32 term per unit, f64 warp-shuffle reduction per item;
//   epilogue: per-row f64 sum of the tile partials, ONE rounding to f32, then the fused op.
// Accumulating the f32 terms in f64 makes the result independent of summation order, so the kernel agrees
// bit-for-bit with the oracle's "canon" vec_dot (oracle/ggml_ref.c) -- greedy parity is then decided by
// logic, not by which int8 code a 1-ulp difference happens to flip.  Cost: 1 F2D + 6 DADD per 2048 weights.
// Programmatic dependent launch: the kernel calls griddepcontrol.wait only before it reads the previous
// phase's output, so its launch latency and setup overlap the tail of the previous phase.
#include <float.h>

#include "actquant.cuh"
#include "common.cuh"
#include "layout.cuh"

#define GEMV_NW 16
#define GEMV_THREADS (GEMV_NW * 32)

struct SegK {
    const uint8_t* w;
    float* y;
    int64_t stride;
    int type;
    int rows;
};

struct GemvK {
    SegK seg[GGB_MAX_SEG];
    int n_seg, k, T;
    int pro, epi;
    int act_q8_0;
    const float* x;
    const float* norm_w;
    float eps;
    const float* residual;
    const int32_t* pos_dev;
    const float* rope_tab;
    int n_rot, head_dim;
    uint16_t* kcache;
    uint16_t* vcache;
    float* part_val;
    int32_t* part_idx;
};

// bank swizzle of 16-byte activation chunks: conflict-free LDS.128 for both the Q4_K/Q8_0 unit pattern
// (chunks 4u+i) and the Q6_K pattern (chunks 16sb+8n+2r+t) -- see DESIGN.md "activation staging".
__device__ __forceinline__ int swz(int c) { return c ^ ((c >> 2) & 7); }

struct WReg {
    uint4 a, b, c, d;
    uint32_t e0, e1, e2;
};

template <int MASK>
__device__ __forceinline__ void load_item(WReg& w, int type, const uint8_t* tb, int U, int nsb, int lane) {
    const bool act = lane < U;
    w.a = w.b = w.c = w.d = make_uint4(0, 0, 0, 0);
    w.e0 = w.e1 = w.e2 = 0;
    if (!act) return;
    if ((MASK & 1) && type == GGB_TYPE_Q4_K) {
        w.a = ldg_stream(tb + 16 * lane);
        w.b = ldg_stream(tb + 16 * U + 16 * lane);
        w.c = ldg_cached(tb + 32 * U + 16 * (lane >> 2));
    } else if ((MASK & 2) && type == GGB_TYPE_Q6_K) {
        w.a = ldg_stream(tb + 16 * lane);
        w.b = ldg_stream(tb + 16 * U + 16 * lane);
        w.c = ldg_stream(tb + 32 * U + 16 * lane);
        const uint2 s = __ldg(reinterpret_cast<const uint2*>(tb + 48 * U + 16 * (lane >> 2) + 8 * ((lane >> 1) & 1)));
        w.e0 = s.x; w.e1 = s.y;
        w.e2 = __ldg(reinterpret_cast<const uint16_t*>(tb + 48 * U + 16 * nsb + 2 * (lane >> 2)));
    } else if ((MASK & 4) && type == GGB_TYPE_Q8_0) {
        w.a = ldg_stream(tb + 16 * lane);
        w.b = ldg_stream(tb + 16 * U + 16 * lane);
        w.c = ldg_stream(tb + 32 * U + 16 * lane);
        w.d = ldg_stream(tb + 48 * U + 16 * lane);
        w.e0 = __ldg(reinterpret_cast<const uint32_t*>(tb + 64 * U + 4 * lane));
    }
}

#define DP4_US(acc, wv, av) acc = dp4a_us((wv).x, (av).x, acc); acc = dp4a_us((wv).y, (av).y, acc); acc = dp4a_us((wv).z, (av).z, acc); acc = dp4a_us((wv).w, (av).w, acc)
__device__ __forceinline__ uint4 and4(uint4 v, uint32_t m) { return make_uint4(v.x & m, v.y & m, v.z & m, v.w & m); }

// gu = global unit index (tile*32 + lane); qs/bsums/dsc = the CTA's quantised activation vector in smem
__device__ __forceinline__ float unit_q4k(const WReg& w, int gu, const uint8_t* qs, const int16_t* bsums, const float* dsc) {
    const uint4 a0 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 0));
    const uint4 a1 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 1));
    const uint4 a2 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 2));
    const uint4 a3 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 3));
    int dlo = 0, dhi = 0;
    { const uint4 t = and4(w.a, 0x0F0F0F0Fu); DP4_US(dlo, t, a0); }
    { const uint4 t = and4(w.b, 0x0F0F0F0Fu); DP4_US(dlo, t, a1); }
    { const uint4 t = and4(w.a, 0xF0F0F0F0u); DP4_US(dhi, t, a2); }   // 16 x the high-nibble dot (exact)
    { const uint4 t = and4(w.b, 0xF0F0F0F0u); DP4_US(dhi, t, a3); }
    dhi >>= 4;
    // 6-bit scales/mins of sub-blocks 2g, 2g+1 (gguf/quants.py:478-502), two bytes at a time
    const int g = gu & 3;
    const int sh = 16 * (g & 1);
    const uint32_t p1 = (w.c.y >> sh) & 0xFFFFu, p2 = (w.c.z >> sh) & 0xFFFFu, p3 = (w.c.w >> sh) & 0xFFFFu;
    uint32_t sc2, mn2;
    if (g < 2) { sc2 = p1 & 0x3F3Fu; mn2 = p2 & 0x3F3Fu; }
    else { sc2 = (p3 & 0x0F0Fu) | ((p1 >> 2) & 0x3030u); mn2 = ((p3 >> 4) & 0x0F0Fu) | ((p2 >> 2) & 0x3030u); }
    const int isum = (int)(sc2 & 0xFF) * dlo + (int)(sc2 >> 8) * dhi;
    const uint2 bs = *reinterpret_cast<const uint2*>(bsums + 4 * gu); /* four per-16 sums */
    const int bs_lo = (int)(int16_t)(bs.x & 0xFFFF) + (int)(int16_t)(bs.x >> 16);
    const int bs_hi = (int)(int16_t)(bs.y & 0xFFFF) + (int)(int16_t)(bs.y >> 16);
    const int msum = (int)(mn2 & 0xFF) * bs_lo + (int)(mn2 >> 8) * bs_hi;
    const float dx = dsc[gu >> 2];
    const float d = h2f((uint16_t)(w.c.x & 0xFFFF)), dmin = h2f((uint16_t)(w.c.x >> 16));
    // one f32 term per unit, fixed operation order (oracle: gref_vec_dot_q4_K_q8_K_canon)
    return __fsub_rn(__fmul_rn(__fmul_rn(d, dx), (float)isum), __fmul_rn(__fmul_rn(dmin, dx), (float)msum));
}

__device__ __forceinline__ float unit_q6k(const WReg& w, int gu, const uint8_t* qs, const int16_t* bsums, const float* dsc) {
    const int tt = gu & 1;
    const int c0 = 4 * gu - 3 * tt; /* chunk of r = 0; r adds 2 */
    const uint4 a0 = *reinterpret_cast<const uint4*>(qs + 16 * swz(c0));
    const uint4 a1 = *reinterpret_cast<const uint4*>(qs + 16 * swz(c0 + 2));
    const uint4 a2 = *reinterpret_cast<const uint4*>(qs + 16 * swz(c0 + 4));
    const uint4 a3 = *reinterpret_cast<const uint4*>(qs + 16 * swz(c0 + 6));
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    { const uint4 t = and4(w.a, 0x0F0F0F0Fu); DP4_US(s0, t, a0); }
    { const uint4 t = and4(w.b, 0x0F0F0F0Fu); DP4_US(s1, t, a1); }
    { const uint4 t = and4(w.a, 0xF0F0F0F0u); DP4_US(s2, t, a2); }
    { const uint4 t = and4(w.b, 0xF0F0F0F0u); DP4_US(s3, t, a3); }
    { const uint4 t = and4(w.c, 0x03030303u); DP4_US(h0, t, a0); }
    { const uint4 t = and4(w.c, 0x0C0C0C0Cu); DP4_US(h1, t, a1); }
    { const uint4 t = and4(w.c, 0x30303030u); DP4_US(h2, t, a2); }
    { const uint4 t = and4(w.c, 0xC0C0C0C0u); DP4_US(h3, t, a3); }
    // sum (q-32)*a over each 16-group = nibble part + 16*high-bit part - 32*sum(a)
    const int v0 = s0 + (h0 << 4) - 32 * (int)bsums[c0];
    const int v1 = s1 + (h1 << 2) - 32 * (int)bsums[c0 + 2];
    const int v2 = (s2 >> 4) + h2 - 32 * (int)bsums[c0 + 4];
    const int v3 = (s3 >> 4) + (h3 >> 2) - 32 * (int)bsums[c0 + 6];
    // scales sc[8n + 2r + t]: e0,e1 hold the 8 scales of this half; byte (2r + t)
    const uint32_t lo = tt ? (w.e0 >> 8) : w.e0, hi = tt ? (w.e1 >> 8) : w.e1;
    const int isum = (int)(int8_t)(lo & 0xFF) * v0 + (int)(int8_t)((lo >> 16) & 0xFF) * v1 +
                     (int)(int8_t)(hi & 0xFF) * v2 + (int)(int8_t)((hi >> 16) & 0xFF) * v3;
    return __fmul_rn(__fmul_rn(h2f((uint16_t)w.e2), dsc[gu >> 2]), (float)isum);
}

#define DP4_SS(acc, wv, av) acc = dp4a_ss((wv).x, (av).x, acc); acc = dp4a_ss((wv).y, (av).y, acc); acc = dp4a_ss((wv).z, (av).z, acc); acc = dp4a_ss((wv).w, (av).w, acc)

__device__ __forceinline__ double unit_q8_0(const WReg& w, int gu, const uint8_t* qs, const float* dsc) {
    const uint4 a0 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 0));
    const uint4 a1 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 1));
    const uint4 a2 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 2));
    const uint4 a3 = *reinterpret_cast<const uint4*>(qs + 16 * swz(4 * gu + 3));
    int i0 = 0, i1 = 0;
    DP4_SS(i0, w.a, a0); DP4_SS(i0, w.b, a1);
    DP4_SS(i1, w.c, a2); DP4_SS(i1, w.d, a3);
    const float2 dx = *reinterpret_cast<const float2*>(dsc + 2 * gu);
    const float t0 = __fmul_rn((float)i0, __fmul_rn(h2f((uint16_t)(w.e0 & 0xFFFF)), dx.x));
    const float t1 = __fmul_rn((float)i1, __fmul_rn(h2f((uint16_t)(w.e0 >> 16)), dx.y));
    return (double)t0 + (double)t1; /* one f32 term per 32-block, added in f64 */
}

template <int MASK>
__device__ __forceinline__ double compute_item(const WReg& w, int type, int gu, bool act, const uint8_t* qs,
                                               const int16_t* bsums, const float* dsc) {
    double p = 0.0;
    if (act) {
        if ((MASK & 1) && type == GGB_TYPE_Q4_K) p = unit_q4k(w, gu, qs, bsums, dsc);
        else if ((MASK & 2) && type == GGB_TYPE_Q6_K) p = unit_q6k(w, gu, qs, bsums, dsc);
        else if ((MASK & 4) && type == GGB_TYPE_Q8_0) p = unit_q8_0(w, gu, qs, dsc);
    }
    return p;
}

__device__ __forceinline__ void argmax_comb(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// MASK: bit0 Q4_K, bit1 Q6_K, bit2 Q8_0 segments present (dead code elimination per launch shape)
template <int MASK>
__global__ void __launch_bounds__(GEMV_THREADS, 2) gemv_kernel(const __grid_constant__ GemvK P) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ double red[GEMV_NW];
    __shared__ float s_val[GEMV_NW];
    __shared__ int s_idx[GEMV_NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.k, T = P.T;
    uint8_t* qs = smem;                                                   // K bytes
    int16_t* bsums = reinterpret_cast<int16_t*>(smem + K);                // K/16 int16
    float* dsc = reinterpret_cast<float*>(smem + K + K / 8);              // K/32 floats
    double* partial = reinterpret_cast<double*>(smem + K + K / 8 + K / 8); // one f64 partial per (row, tile) item

    // ---- this CTA's row range in every segment (even-aligned so RoPE pairs stay together)
    int r0[GGB_MAX_SEG], cnt[GGB_MAX_SEG];
    int nloc = 0;
#pragma unroll
    for (int s = 0; s < GGB_MAX_SEG; s++) {
        r0[s] = 0; cnt[s] = 0;
        if (s < P.n_seg) {
            const int64_t rows = P.seg[s].rows;
            int a = (int)(rows * blockIdx.x / gridDim.x), b = (int)(rows * (blockIdx.x + 1) / gridDim.x);
            a &= ~1; if (blockIdx.x + 1 != gridDim.x) b &= ~1;
            r0[s] = a; cnt[s] = b - a;
            nloc += cnt[s];
        }
    }
    const int nitems = nloc * T;

    // item -> (segment, row, tile)
    auto locate = [&](int item, int& type, const uint8_t*& tb, int& U, int& nsb, int& t) {
        int lr = item / T;
        t = item - lr * T;
        int s = 0;
        if (lr >= cnt[0]) { lr -= cnt[0]; s = 1; if (lr >= cnt[1]) { lr -= cnt[1]; s = 2; } }
        type = P.seg[s].type;
        nsb = ggb_tile_nsb(K, t);
        U = 4 * nsb;
        tb = P.seg[s].w + (int64_t)(r0[s] + lr) * P.seg[s].stride + (int64_t)t * (ggb_sb_bytes(type) * GGB_TILE_SB);
    };

    // ---- weights do not depend on the previous phase: put the first item in flight before waiting for it
    WReg w0, w1;
    int type0 = 0, U0 = 0, nsb0 = 0, t0 = 0;
    const uint8_t* tb0 = nullptr;
    int item = warp;
    if (item < nitems) { locate(item, type0, tb0, U0, nsb0, t0); load_item<MASK>(w0, type0, tb0, U0, nsb0, lane); }

    pdl_launch_dependents();
    pdl_wait();

    // ---- prologue: (rms_norm * gain) and activation quantisation into shared memory
    float scale = 1.f;
    if (P.pro == GGB_PRO_RMSNORM) {
        double s = 0.0;
        for (int i = tid; i < K; i += GEMV_THREADS) { const float v = P.x[i]; s += (double)__fmul_rn(v, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < GEMV_NW; i++) tot += red[i];
        const float mean = (float)(tot / (double)K);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + P.eps));
    }
    for (int b = warp; b < K / 256; b += GEMV_NW) {
        const int e0 = b * 256 + lane * 8;
        const float4 x0 = *reinterpret_cast<const float4*>(P.x + e0), x1 = *reinterpret_cast<const float4*>(P.x + e0 + 4);
        float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        if (P.pro == GGB_PRO_RMSNORM) {
            const float4 g0 = *reinterpret_cast<const float4*>(P.norm_w + e0), g1 = *reinterpret_cast<const float4*>(P.norm_w + e0 + 4);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), g[i]);
        }
        const int chunk = e0 >> 4;
        uint2* dst = reinterpret_cast<uint2*>(qs + 16 * swz(chunk) + 8 * (lane & 1));
        if (P.act_q8_0) {
            float df; uint16_t db;
            const Q8Codes c = warp_quantize_q8_0(v, df, db);
            *dst = c.q;
            if (!(lane & 3)) dsc[b * 8 + (lane >> 2)] = df;
        } else {
            float dd;
            const Q8Codes c = warp_quantize_q8_K(v, lane, dd);
            *dst = c.q;
            const int s16 = c.sum8 + __shfl_xor_sync(0xffffffffu, c.sum8, 1);
            if (!(lane & 1)) bsums[chunk] = (int16_t)s16;
            if (lane == 0) dsc[b] = dd;
        }
    }
    __syncthreads();

    // ---- main loop: one (row, tile) item per warp per step, next item's loads in flight
    for (; item < nitems; item += 2 * GEMV_NW) {
        int type1 = 0, U1 = 0, nsb1 = 0, t1 = 0;
        const uint8_t* tb1 = nullptr;
        const int nx = item + GEMV_NW;
        if (nx < nitems) { locate(nx, type1, tb1, U1, nsb1, t1); load_item<MASK>(w1, type1, tb1, U1, nsb1, lane); }
        {
            double p = compute_item<MASK>(w0, type0, t0 * 32 + lane, lane < U0, qs, bsums, dsc);
            p = warp_sum_f64(p);
            if (lane == 0) partial[item] = p;
        }
        const int nx2 = item + 2 * GEMV_NW;
        if (nx2 < nitems) { locate(nx2, type0, tb0, U0, nsb0, t0); load_item<MASK>(w0, type0, tb0, U0, nsb0, lane); }
        if (nx < nitems) {
            double p = compute_item<MASK>(w1, type1, t1 * 32 + lane, lane < U1, qs, bsums, dsc);
            p = warp_sum_f64(p);
            if (lane == 0) partial[nx] = p;
        }
    }
    __syncthreads();

    // ---- epilogue
    auto rowval = [&](int lr) {
        double v = 0.0;
        for (int t = 0; t < T; t++) v += partial[lr * T + t];
        return (float)v; /* the only rounding of the accumulated sum */
    };
    if (P.epi == GGB_EPI_STORE) {
        for (int lr = tid; lr < nloc; lr += GEMV_THREADS) {
            int s = 0, l = lr;
            if (l >= cnt[0]) { l -= cnt[0]; s = 1; if (l >= cnt[1]) { l -= cnt[1]; s = 2; } }
            P.seg[s].y[r0[s] + l] = rowval(lr);
        }
    } else if (P.epi == GGB_EPI_RESIDUAL) {
        for (int lr = tid; lr < cnt[0]; lr += GEMV_THREADS) {
            const int r = r0[0] + lr;
            P.seg[0].y[r] = P.residual[r] + rowval(lr);
        }
    } else if (P.epi == GGB_EPI_SWIGLU) {
        for (int lr = tid; lr < cnt[0]; lr += GEMV_THREADS) {
            const float g = rowval(lr), u = rowval(cnt[0] + lr);
            P.seg[0].y[r0[0] + lr] = silu_mul_ref(g, u);
        }
    } else if (P.epi == GGB_EPI_ROPE_KV) {
        const int pos = *P.pos_dev;
        const float* tab = P.rope_tab + (int64_t)pos * P.n_rot; /* [n_rot/2][2] */
        const int npair = nloc >> 1;
        for (int pr = tid; pr < npair; pr += GEMV_THREADS) {
            int s = 0, l = 2 * pr;
            if (l >= cnt[0]) { l -= cnt[0]; s = 1; if (l >= cnt[1]) { l -= cnt[1]; s = 2; } }
            const int r = r0[s] + l;
            float v0 = rowval(2 * pr), v1 = rowval(2 * pr + 1);
            if (s < 2) {
                const int j = r % P.head_dim;
                if (j < P.n_rot) {
                    const float c = tab[j], sn = tab[j + 1];  /* pair index j/2 -> floats 2*(j/2) = j */
                    const float a = v0, b = v1;
                    v0 = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, sn));
                    v1 = __fadd_rn(__fmul_rn(a, sn), __fmul_rn(b, c));
                }
            }
            if (s == 0) { P.seg[0].y[r] = v0; P.seg[0].y[r + 1] = v1; }
            else {
                uint16_t* cache = (s == 1) ? P.kcache : P.vcache;
                const uint32_t packed = (uint32_t)f2h(v0) | ((uint32_t)f2h(v1) << 16);
                *reinterpret_cast<uint32_t*>(cache + (int64_t)pos * P.seg[s].rows + r) = packed;
            }
        }
    } else if (P.epi == GGB_EPI_ARGMAX) {
        float bv = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int lr = tid; lr < cnt[0]; lr += GEMV_THREADS) {
            const float v = rowval(lr);
            const int r = r0[0] + lr;
            if (P.seg[0].y) P.seg[0].y[r] = v;
            argmax_comb(bv, bi, v, r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) argmax_comb(bv, bi, __shfl_xor_sync(0xffffffffu, bv, o), __shfl_xor_sync(0xffffffffu, bi, o));
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < GEMV_NW; i++) argmax_comb(bv, bi, s_val[i], s_idx[i]);
            P.part_val[blockIdx.x] = bv;
            P.part_idx[blockIdx.x] = bi;
        }
    }
}

// ------------------------------------------------------------------ host side
static int act_class(int type) { return type == GGB_TYPE_Q8_0 ? 1 : 0; }

static int default_grid(const ggb_gemv_args* a) {
    if (a->grid > 0) return a->grid;
    return ggb_num_sms();
}

extern "C" int ggb_gemv_grid(const ggb_gemv_args* a) {
    if (!a) return GGB_ERR_ARG;
    return default_grid(a);
}

template <int MASK>
static int launch(const GemvK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    static bool attr_done = false;
    static size_t attr_smem = 0;
    if (!attr_done || smem > attr_smem) {
        GGB_CUDA(cudaFuncSetAttribute(gemv_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done = true;
        attr_smem = 160 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMV_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, gemv_kernel<MASK>, P));
    return GGB_OK;
}

extern "C" int ggb_gemv(const ggb_gemv_args* a, void* stream) {
    if (!a) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: null args");
    if (a->n_seg < 1 || a->n_seg > GGB_MAX_SEG) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: n_seg=%d out of range", a->n_seg);
    if (a->k <= 0 || (a->k % 256)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: k=%d must be a positive multiple of 256", a->k);
    if (!a->x) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: null input vector");
    if (((uintptr_t)a->x & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: x must be 16-byte aligned");
    if (a->prologue == GGB_PRO_RMSNORM && (!a->norm_w || ((uintptr_t)a->norm_w & 15))) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: RMSNORM prologue needs a 16-byte aligned norm_w");
    GemvK P = {};
    int mask = 0, cls = -1;
    int64_t total_rows = 0;
    for (int s = 0; s < a->n_seg; s++) {
        const ggb_gemv_seg& g = a->seg[s];
        int bit;
        switch (g.type) {
            case GGB_TYPE_Q4_K: bit = 1; break;
            case GGB_TYPE_Q6_K: bit = 2; break;
            case GGB_TYPE_Q8_0: bit = 4; break;
            default: GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: segment %d has unsupported weight type %d", s, g.type);
        }
        if (cls >= 0 && cls != act_class(g.type)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: Q8_0 and K-quant segments cannot share a launch");
        cls = act_class(g.type);
        mask |= bit;
        if (g.rows < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: negative row count");
        if (g.rows > 0 && (!g.w || ((uintptr_t)g.w & 15))) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: segment %d weights null or not 16-byte aligned", s);
        P.seg[s].w = (const uint8_t*)g.w;
        P.seg[s].y = g.y;
        P.seg[s].stride = ggb_row_stride(g.type, a->k);
        P.seg[s].type = g.type;
        P.seg[s].rows = g.rows;
        total_rows += g.rows;
    }
    switch (a->epilogue) {
        case GGB_EPI_STORE:
            for (int s = 0; s < a->n_seg; s++) if (a->seg[s].rows && !a->seg[s].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: STORE needs y for every segment");
            break;
        case GGB_EPI_RESIDUAL:
            if (a->n_seg != 1 || !a->residual || !a->seg[0].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: RESIDUAL needs one segment, y and residual");
            break;
        case GGB_EPI_SWIGLU:
            if (a->n_seg != 2 || a->seg[0].rows != a->seg[1].rows || !a->seg[0].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: SWIGLU needs gate/up segments of equal rows and y on segment 0");
            break;
        case GGB_EPI_ROPE_KV:
            if (a->n_seg != 3 || !a->pos_dev || !a->rope_tab || !a->kcache || !a->vcache || !a->seg[0].y || a->head_dim <= 0 ||
                a->n_rot <= 0 || (a->n_rot & 1) || a->n_rot > a->head_dim || (a->seg[0].rows & 1) || (a->seg[1].rows & 1) || (a->seg[2].rows & 1))
                GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: ROPE_KV needs q,k,v segments with even rows, pos_dev, rope_tab, caches, head_dim and an even n_rot");
            break;
        case GGB_EPI_ARGMAX:
            if (a->n_seg != 1 || !a->part_val || !a->part_idx) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: ARGMAX needs one segment and partial buffers");
            break;
        default: GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: unknown epilogue %d", a->epilogue);
    }
    if (total_rows == 0) return GGB_OK;
    P.n_seg = a->n_seg; P.k = a->k; P.T = ggb_tiles_per_row(a->k);
    P.pro = a->prologue; P.epi = a->epilogue; P.act_q8_0 = cls;
    P.x = a->x; P.norm_w = a->norm_w; P.eps = a->eps; P.residual = a->residual;
    P.pos_dev = a->pos_dev; P.rope_tab = a->rope_tab; P.n_rot = a->n_rot; P.head_dim = a->head_dim;
    P.kcache = a->kcache; P.vcache = a->vcache; P.part_val = a->part_val; P.part_idx = a->part_idx;
    const int grid = default_grid(a);
    int64_t max_local = 0;
    for (int s = 0; s < a->n_seg; s++) max_local += (a->seg[s].rows + grid - 1) / grid + 2;
    const size_t smem = (size_t)a->k + a->k / 8 + a->k / 8 + (size_t)max_local * P.T * sizeof(double);
    if (smem > 160 * 1024) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: k=%d rows=%lld needs %zu bytes of shared memory", a->k, (long long)total_rows, smem);
    cudaStream_t st = (cudaStream_t)stream;
    switch (mask) {
        case 1: return launch<1>(P, grid, smem, a->use_pdl, st);
        case 2: return launch<2>(P, grid, smem, a->use_pdl, st);
        case 3: return launch<3>(P, grid, smem, a->use_pdl, st);
        case 4: return launch<4>(P, grid, smem, a->use_pdl, st);
        default: GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: unsupported type mix (mask %d)", mask);
    }
}
