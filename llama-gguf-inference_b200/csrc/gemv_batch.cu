// gemv_batch.cu -- K1b: fused dequant-GEMV for SMALL-BATCH decode (several sequences, one new token each).
// Stands in for ggml's mul_mat with a few activation columns (llama-server's continuous batching of decode
// tokens from different slots) [UPSTREAM-MEM: ggml-cpu mul_mat -> ggml_vec_dot_q*_K_q8_K per column].
//
// Same arithmetic as gemv.cu, token by token: identical int8 activation codes (ggb_act_prep below restates the
// GEMV prologue), identical integer dots, identical f32 unit terms, summed in f64 -- so a sequence decoded inside a
// batch produces bit-identical logits to the same sequence decoded alone.  What changes is the data movement:
//   * weights stream from HBM ONCE per launch for up to 16 tokens (8 when the images of 16 do not fit shared memory) (the batch-1 kernel would read them once per token);
//   * the per-CTA quantisation prologue is replaced by a bulk copy of pre-quantised activation "images"
//     (ggb_act_prep, one small launch per phase) -- quantising 8 vectors redundantly in 148 CTAs would cost more
//     than the GEMV;
//   * a warp keeps the unpacked weights of a row pair in registers (nibbles / 6-bit codes as bytes, scale fields as
//     integers) and applies them to every token: 4 x LDS.128 of activations per token, 16 dp4a per (row, token).
// Per (warp, K-tile) the shared-memory traffic is NB x 2 KB of activations + 2 tiles of weights and the issue cost
// is ~36 instructions per (row, token): at NB = 8 both are ~1.5x the HBM time of the tile, so a batch of 8 costs
// about what 1.5 single tokens cost.
#include <float.h>
#include <stdlib.h>

#include "actquant.cuh"
#include "common.cuh"
#include "layout.cuh"
#include "gemv_common.cuh"
#include "../../include/ggufb200.h"

#define GB_NW 8
#define GB_THREADS (GB_NW * 32)
#define GB_MAX_R 4      /* rows per warp group: 2, or 4 for pure Q4_K launches (halves the activation loads per row) */
#define GB_STEPS 2      /* ring stages per warp */
#define GB_MAX_SMEM (226 * 1024)

struct SegB {
    const uint8_t* w;
    float* y;
    int64_t stride;
    int type;
    int rows;
};

struct GemvBK {
    SegB seg[GGB_MAX_SEG];
    int n_seg, k, T, epi, nb;
    int slot_bytes, ring_bytes;
    int image;                 /* bytes of one token's activation image (codes | per-16 sums | block scales) */
    int rowv_off;
    int rq[GGB_MAX_SEG], rr[GGB_MAX_SEG];
    const uint8_t* act;        /* [nb][image] */
    const float* residual;     /* [nb][rows0] */
};

// ------------------------------------------------------------------ unpacked weights of one (row, K-tile unit), kept in registers
struct W4 { uint4 q0l, q0h, q1l, q1h; int sc0, sc1, m0, m1; float d, dmin; };
struct W6 { uint4 w0, w1, w2, w3; int s0, s1, s2, s3; float d; };
struct W8 { uint4 w0, w1, w2, w3; float d0, d1; };

__device__ __forceinline__ uint4 shr4(uint4 v, int s) { return make_uint4(v.x >> s, v.y >> s, v.z >> s, v.w >> s); }
__device__ __forceinline__ uint4 shl4(uint4 v, int s) { return make_uint4(v.x << s, v.y << s, v.z << s, v.w << s); }
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }

__device__ __forceinline__ W4 prep_q4k(uint4 q0, uint4 q1, uint4 hdr, const LaneK& L) {
    W4 w;
    w.q0l = and4(q0, 0x0F0F0F0Fu); w.q0h = and4(shr4(q0, 4), 0x0F0F0F0Fu);
    w.q1l = and4(q1, 0x0F0F0F0Fu); w.q1h = and4(shr4(q1, 4), 0x0F0F0F0Fu);
    const uint32_t fa = __byte_perm(hdr.y, hdr.z, L.selA), fb = __byte_perm(hdr.z, hdr.w, L.selB);
    const uint32_t f = L.lowg ? fa : fb;
    w.sc0 = (int)(f & 63); w.sc1 = (int)((f >> 6) & 63); w.m0 = (int)((f >> 12) & 63); w.m1 = (int)((f >> 18) & 63);
    w.d = h2f((uint16_t)(hdr.x & 0xFFFF)); w.dmin = h2f((uint16_t)(hdr.x >> 16));
    return w;
}
// Q5_K: the 5-bit codes as bytes (q5k_codes, gemv_common.cuh); from there on it is a Q4_K unit
__device__ __forceinline__ W4 prep_q5k(uint4 q0, uint4 q1, uint2 qhu, uint4 hdr, const LaneK& L) {
    W4 w;
    q5k_codes(q0, q1, qhu, w.q0l, w.q1l, w.q0h, w.q1h);
    const uint32_t fa = __byte_perm(hdr.y, hdr.z, L.selA), fb = __byte_perm(hdr.z, hdr.w, L.selB);
    const uint32_t f = L.lowg ? fa : fb;
    w.sc0 = (int)(f & 63); w.sc1 = (int)((f >> 6) & 63); w.m0 = (int)((f >> 12) & 63); w.m1 = (int)((f >> 18) & 63);
    w.d = h2f((uint16_t)(hdr.x & 0xFFFF)); w.dmin = h2f((uint16_t)(hdr.x >> 16));
    return w;
}
// same integers and the same f32 operation order as term_q4k (gemv_common.cuh)
__device__ __forceinline__ float term_q4k_w(const W4& w, const Act& A) {
    const int dlo = dot16_us(w.q0l, A.a0) + dot16_us(w.q1l, A.a1);
    const int dhi = dot16_us(w.q0h, A.a2) + dot16_us(w.q1h, A.a3);
    const int isum = w.sc0 * dlo + w.sc1 * dhi;
    const int msum = w.m0 * A.b0 + w.m1 * A.b1;
    return __fsub_rn(__fmul_rn(__fmul_rn(w.d, A.dx0), (float)isum), __fmul_rn(__fmul_rn(w.dmin, A.dx0), (float)msum));
}

// Q6_K: the 6-bit codes are assembled once per row as bytes (nibble | 2 high bits << 4), so a 16-group costs 4 dp4a
// per token instead of 8; sum (q - 32) * a = dot(q, a) - 32 * sum(a) -- the same integer as term_q6k's.
__device__ __forceinline__ W6 prep_q6k(uint4 qla, uint4 qlb, uint4 qh, uint2 sc8, uint32_t dbits, int tt) {
    W6 w;
    w.w0 = or4(and4(qla, 0x0F0F0F0Fu), and4(shl4(qh, 4), 0x30303030u));
    w.w1 = or4(and4(qlb, 0x0F0F0F0Fu), and4(shl4(qh, 2), 0x30303030u));
    w.w2 = or4(and4(shr4(qla, 4), 0x0F0F0F0Fu), and4(qh, 0x30303030u));
    w.w3 = or4(and4(shr4(qlb, 4), 0x0F0F0F0Fu), and4(shr4(qh, 2), 0x30303030u));
    const uint32_t lo = tt ? (sc8.x >> 8) : sc8.x, hi = tt ? (sc8.y >> 8) : sc8.y;
    w.s0 = (int)(int8_t)(lo & 0xFF); w.s1 = (int)(int8_t)((lo >> 16) & 0xFF);
    w.s2 = (int)(int8_t)(hi & 0xFF); w.s3 = (int)(int8_t)((hi >> 16) & 0xFF);
    w.d = h2f((uint16_t)dbits);
    return w;
}
__device__ __forceinline__ float term_q6k_w(const W6& w, const Act& A) {
    const int v0 = dot16_us(w.w0, A.a0) - A.b0, v1 = dot16_us(w.w1, A.a1) - A.b1;
    const int v2 = dot16_us(w.w2, A.a2) - A.b2, v3 = dot16_us(w.w3, A.a3) - A.b3;
    const int isum = w.s0 * v0 + w.s1 * v1 + w.s2 * v2 + w.s3 * v3;
    return __fmul_rn(__fmul_rn(w.d, A.dx0), (float)isum);
}

__device__ __forceinline__ double term_q80_w(const W8& w, const Act& A) {
    const int i0 = dot16_ss(w.w0, A.a0) + dot16_ss(w.w1, A.a1);
    const int i1 = dot16_ss(w.w2, A.a2) + dot16_ss(w.w3, A.a3);
    const float t0 = __fmul_rn((float)i0, __fmul_rn(w.d0, A.dx0));
    const float t1 = __fmul_rn((float)i1, __fmul_rn(w.d1, A.dx1));
    return (double)t0 + (double)t1;
}

// butterfly reduction of N per-lane values across the warp: afterwards lane l with (l & (32/N - 1)) == 0 holds the
// full sum of value index l / (32/N) in v[0].
template <int N>
__device__ __forceinline__ void warp_reduce_many(double (&v)[N], int lane) {
    static_assert(N == 4 || N == 8 || N == 16 || N == 32, "value count");
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const double keep = up ? v[i + half] : v[i];
            const double send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        off >>= 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o <= off) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

// MASK: bit0 Q4_K, bit1 Q6_K, bit2 Q8_0, bit3 Q5_K segments present; NBT = tokens per launch (compile-time upper bound, the
// images of tokens >= nb are zero-filled by the prologue)
template <int MASK, int NBT, int R>
__global__ void __launch_bounds__(GB_THREADS, 1) ggb_dq_gemv_batch_kernel(const __grid_constant__ GemvBK P) {
    constexpr int STEPS = GB_STEPS;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[GB_NW][STEPS];
    __shared__ __align__(8) uint64_t s_abar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.k, T = P.T;
    uint8_t* act = smem + GB_NW * P.ring_bytes;
    double* rowv = reinterpret_cast<double*>(smem + P.rowv_off);   /* [local row][NBT] */
    const uint32_t SLOT = P.slot_bytes;

    const int G = gridDim.x, c = blockIdx.x;
    int r0_0, r0_1 = 0, r0_2 = 0, cnt0, cnt1 = 0, cnt2 = 0;
    {
        const bool last = (c + 1 == G);
        auto range = [&](int sg, int& r0, int& cnt) {
            const int q = P.rq[sg], r = P.rr[sg];
            const int a = (c * q + min(c, r)) & ~(R - 1);
            int b = (c + 1) * q + min(c + 1, r);
            if (!last) b &= ~(R - 1);
            r0 = a; cnt = b - a;
        };
        range(0, r0_0, cnt0);
        if (P.n_seg > 1) range(1, r0_1, cnt1);
        if (P.n_seg > 2) range(2, r0_2, cnt2);
    }
    const int nloc = cnt0 + cnt1 + cnt2;
    const int np0 = (cnt0 + R - 1) / R, np1 = (cnt1 + R - 1) / R, np2 = (cnt2 + R - 1) / R;
    const int npairs = np0 + np1 + np2;
    auto pair_info = [&](int p, int& s, int& row, int& nv, int& lr) {
        if (p < np0) { s = 0; row = r0_0 + R * p; nv = min(R, cnt0 - R * p); lr = R * p; }
        else if (p < np0 + np1) { p -= np0; s = 1; row = r0_1 + R * p; nv = min(R, cnt1 - R * p); lr = cnt0 + R * p; }
        else { p -= np0 + np1; s = 2; row = r0_2 + R * p; nv = min(R, cnt2 - R * p); lr = cnt0 + cnt1 + R * p; }
    };

    // ---- per-warp weight ring (same producer as gemv.cu: lane 0 issues bulk copies, one mbarrier per stage)
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]);
    const uint32_t abar = smem_u32(&s_abar);
    const uint32_t ring0 = smem_u32(smem) + warp * P.ring_bytes;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STEPS; i++) mbar_init(bar0 + 8 * i, 1);
        if (warp == 0) mbar_init(abar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nsb_last = ggb_tile_nsb(K, T - 1);
    int ip = warp, it = 0, istage = 0;
    const uint8_t* isrc = nullptr;
    int istride = 0, itile = 0, ilast = 0, inv = 0;
    auto issue_pair_setup = [&]() {
        int s, row, lr;
        pair_info(ip, s, row, inv, lr);
        const int sbb = ggb_sb_bytes(P.seg[s].type);
        istride = (int)P.seg[s].stride;
        itile = sbb * GGB_TILE_SB;
        ilast = (nsb_last * sbb + 15) & ~15;
        isrc = P.seg[s].w + (int64_t)row * istride;
    };
    if (ip < npairs) issue_pair_setup();
    auto issue_step = [&]() {
        const uint32_t bytes = (it == T - 1) ? (uint32_t)ilast : (uint32_t)itile;
        const uint32_t bar = bar0 + 8 * istage;
        const uint32_t dst = ring0 + istage * R * SLOT;
        mbar_expect_tx(bar, (uint32_t)inv * bytes);
#pragma unroll
        for (int r = 0; r < R; r++)
            if (r < inv) bulk_g2s(dst + r * SLOT, isrc + (int64_t)r * istride, bytes, bar);
        istage = (istage + 1 == STEPS) ? 0 : istage + 1;
        isrc += itile;
        if (++it == T) {
            it = 0;
            ip += GB_NW;
            if (ip < npairs) issue_pair_setup();
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STEPS; i++) if (ip < npairs) issue_step();
    }
    // images of absent tokens: zero (integer dots 0, scales 0 -> terms 0; nothing of them is stored)
    for (int i = P.nb * P.image + tid * 16; i < NBT * P.image; i += GB_THREADS * 16) *reinterpret_cast<uint4*>(act + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();   /* barrier inits visible to every thread */

    pdl_launch_dependents();
    pdl_wait();

    // ---- activation images of this launch's tokens: one bulk copy each
    const uint32_t act_s = smem_u32(act);
    if (tid == 0) {
        mbar_expect_tx(abar, (uint32_t)(P.nb * P.image));
        for (int b = 0; b < P.nb; b++) bulk_g2s(act_s + b * P.image, P.act + (int64_t)b * P.image, (uint32_t)P.image, abar);
    }
    mbar_wait(abar, 0);

    // ---- main loop
    const LaneK L = lane_consts(lane);
    const uint32_t bs_off = (uint32_t)K, dsc_off = (uint32_t)(K + K / 8);
    const int U_last = 4 * nsb_last;
    const bool last_full = (nsb_last == GGB_TILE_SB);
    int cstage = 0;
    uint32_t cphase = 0;
    for (int p = warp; p < npairs; p += GB_NW) {
        int s, row, lr, nv;
        pair_info(p, s, row, nv, lr);
        const int type = P.seg[s].type;
        double acc[R * NBT];
#pragma unroll
        for (int i = 0; i < R * NBT; i++) acc[i] = 0.0;
        for (int t = 0; t < T; t++) {
            const bool full = (t != T - 1) || last_full;
            const int U = full ? 32 : U_last;
            const uint32_t S = 16u * (uint32_t)U;
            const int nsb = full ? GGB_TILE_SB : nsb_last;
            const uint32_t slot0 = ring0 + cstage * R * SLOT;
            const int gu = t * 32 + lane;
            mbar_wait(bar0 + 8 * cstage, cphase);
            if (lane < U) {
                if ((MASK & 1) && (MASK == 1 || type == GGB_TYPE_Q4_K)) {
                    W4 w[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t sl = slot0 + r * SLOT;
                        w[r] = prep_q4k(lds128(sl + L.o_q), lds128(sl + L.o_q + S), lds128(sl + L.o_h + 2 * S), L);
                    }
#pragma unroll
                    for (int b = 0; b < NBT; b++) {
                        const uint32_t qb = act_s + b * P.image;
                        const Act A = load_act<MASK>(GGB_TYPE_Q4_K, gu, qb, qb + bs_off, qb + dsc_off);
#pragma unroll
                        for (int r = 0; r < R; r++) acc[r * NBT + b] += (double)term_q4k_w(w[r], A);
                    }
                } else if ((MASK & 2) && (MASK == 2 || type == GGB_TYPE_Q6_K)) {
                    W6 w[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t sl = slot0 + r * SLOT;
                        w[r] = prep_q6k(lds128(sl + L.o_q), lds128(sl + L.o_q + S), lds128(sl + L.o_q + 2 * S), lds64(sl + L.o_sc + 3 * S),
                                        lds16(sl + L.o_d + 3 * S + 16u * (uint32_t)nsb), lane & 1);
                    }
#pragma unroll
                    for (int b = 0; b < NBT; b++) {
                        const uint32_t qb = act_s + b * P.image;
                        const Act A = load_act<MASK>(GGB_TYPE_Q6_K, gu, qb, qb + bs_off, qb + dsc_off);
#pragma unroll
                        for (int r = 0; r < R; r++) acc[r * NBT + b] += (double)term_q6k_w(w[r], A);
                    }
                } else if ((MASK & 8) && type == GGB_TYPE_Q5_K) {
                    W4 w[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t sl = slot0 + r * SLOT;
                        w[r] = prep_q5k(lds128(sl + L.o_q), lds128(sl + L.o_q + S), lds64(sl + 2 * S + 8u * (uint32_t)lane),
                                        lds128(sl + L.o_h + 2 * S + S / 2), L);
                    }
#pragma unroll
                    for (int b = 0; b < NBT; b++) {
                        const uint32_t qb = act_s + b * P.image;
                        const Act A = load_act<MASK>(GGB_TYPE_Q5_K, gu, qb, qb + bs_off, qb + dsc_off);
#pragma unroll
                        for (int r = 0; r < R; r++) acc[r * NBT + b] += (double)term_q4k_w(w[r], A);
                    }
                } else if (MASK & 4) {
                    W8 w[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t sl = slot0 + r * SLOT;
                        w[r].w0 = lds128(sl + L.o_q); w[r].w1 = lds128(sl + L.o_q + S);
                        w[r].w2 = lds128(sl + L.o_q + 2 * S); w[r].w3 = lds128(sl + L.o_q + 3 * S);
                        const uint32_t dd = lds32(sl + 4 * S + 4u * lane);
                        w[r].d0 = h2f((uint16_t)(dd & 0xFFFF)); w[r].d1 = h2f((uint16_t)(dd >> 16));
                    }
#pragma unroll
                    for (int b = 0; b < NBT; b++) {
                        const uint32_t qb = act_s + b * P.image;
                        const Act A = load_act<MASK>(GGB_TYPE_Q8_0, gu, qb, qb + bs_off, qb + dsc_off);
#pragma unroll
                        for (int r = 0; r < R; r++) acc[r * NBT + b] += term_q80_w(w[r], A);
                    }
                }
            }
            if (++cstage == STEPS) { cstage = 0; cphase ^= 1; }
            __syncwarp();
            if (lane == 0 && ip < npairs) issue_step();
        }
        if constexpr (R * NBT <= 32) {
            warp_reduce_many<R * NBT>(acc, lane);
            constexpr int LPV = 32 / (R * NBT);   /* lanes per value */
            if ((lane & (LPV - 1)) == 0) {
                const int idx = lane / LPV, r = idx / NBT, b = idx % NBT;
                if (r < nv) rowv[(lr + r) * NBT + b] = acc[0];
            }
        } else {                                  /* 64 values: two rounds of 32 (rows 0,1 then rows 2,3) */
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                double v[32];
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = acc[hf * 32 + i];
                warp_reduce_many<32>(v, lane);
                const int idx = hf * 32 + lane, r = idx / NBT, b = idx % NBT;
                if (r < nv) rowv[(lr + r) * NBT + b] = v[0];
            }
        }
    }
    __syncthreads();

    // ---- epilogue ((float)rowv[..] is the only rounding of an accumulated sum)
    const int nb = P.nb;
    if (P.epi == GGB_EPI_STORE) {
        for (int i = tid; i < nloc * nb; i += GB_THREADS) {
            const int b = i / nloc, l = i - b * nloc;
            const float v = (float)rowv[l * NBT + b];
            if (l < cnt0) P.seg[0].y[(int64_t)b * P.seg[0].rows + r0_0 + l] = v;
            else if (l < cnt0 + cnt1) P.seg[1].y[(int64_t)b * P.seg[1].rows + r0_1 + l - cnt0] = v;
            else P.seg[2].y[(int64_t)b * P.seg[2].rows + r0_2 + l - cnt0 - cnt1] = v;
        }
    } else if (P.epi == GGB_EPI_RESIDUAL) {
        for (int i = tid; i < cnt0 * nb; i += GB_THREADS) {
            const int b = i / cnt0, l = i - b * cnt0;
            const int64_t o = (int64_t)b * P.seg[0].rows + r0_0 + l;
            P.seg[0].y[o] = __fadd_rn(P.residual[o], (float)rowv[l * NBT + b]);
        }
    } else if (P.epi == GGB_EPI_SWIGLU) {
        for (int i = tid; i < cnt0 * nb; i += GB_THREADS) {
            const int b = i / cnt0, l = i - b * cnt0;
            P.seg[0].y[(int64_t)b * P.seg[0].rows + r0_0 + l] = silu_mul_ref((float)rowv[l * NBT + b], (float)rowv[(cnt0 + l) * NBT + b]);
        }
    } else if (P.epi == GGB_EPI_STORE_F64) {   /* tensor-parallel partial: summed across ranks before the one rounding */
        double* y64 = reinterpret_cast<double*>(P.seg[0].y);
        for (int i = tid; i < cnt0 * nb; i += GB_THREADS) {
            const int b = i / cnt0, l = i - b * cnt0;
            y64[(int64_t)b * P.seg[0].rows + r0_0 + l] = rowv[l * NBT + b];
        }
    }
}

// ------------------------------------------------------------------ activation images
// One token's image = what the batch-1 GEMV prologue leaves in shared memory: K int8 codes in bank-swizzled 16-byte
// chunks | K/16 int16 group sums | block scales (Q8_K: K/256 floats; Q8_0: K/32 floats), padded to 1.25 K bytes.
// grid (nb, splits): the 256-blocks of a token are spread over `splits` CTAs; with RMSNorm each of them first takes the
// whole row's sum of squares (same order of additions in every CTA: same scale).
__global__ void __launch_bounds__(GB_THREADS) act_prep_kernel(const float* __restrict__ x, const float* __restrict__ norm_w, float eps, int K,
                                                            int q8_0, uint8_t* __restrict__ out, int image) {
    __shared__ double red[GB_NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xr = x + (int64_t)blockIdx.x * K;
    uint8_t* qs = out + (int64_t)blockIdx.x * image;
    int16_t* bsums = reinterpret_cast<int16_t*>(qs + K);
    float* dsc = reinterpret_cast<float*>(qs + K + K / 8);
    pdl_launch_dependents();
    pdl_wait();
    float scale = 1.f;
    if (norm_w) {
        double s = 0.0;
        for (int i = tid; i < K; i += GB_THREADS) { const float v = xr[i]; s += (double)__fmul_rn(v, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < GB_NW; i++) tot += red[i];
        const float mean = (float)(tot / (double)K);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + eps));
    }
    const int nblk = K / 256;
    for (int b = blockIdx.y * GB_NW + warp; b < nblk; b += gridDim.y * GB_NW) {
        const int e = b * 256 + lane * 8;
        const float4 xa = *reinterpret_cast<const float4*>(xr + e), xb = *reinterpret_cast<const float4*>(xr + e + 4);
        float v[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        if (norm_w) {
            const float4 ga = *reinterpret_cast<const float4*>(norm_w + e), gb = *reinterpret_cast<const float4*>(norm_w + e + 4);
            const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), g[i]);
        }
        const int chunk = e >> 4;
        uint2* dst = reinterpret_cast<uint2*>(qs + 16 * swz(chunk) + 8 * (lane & 1));
        if (q8_0) {
            float df; uint16_t db;
            const Q8Codes cq = warp_quantize_q8_0(v, df, db);
            *dst = cq.q;
            if (!(lane & 3)) dsc[b * 8 + (lane >> 2)] = df;
        } else {
            float dd;
            const Q8Codes cq = warp_quantize_q8_K(v, lane, dd);
            *dst = cq.q;
            const int s16 = cq.sum8 + __shfl_xor_sync(0xffffffffu, cq.sum8, 1);
            if (!(lane & 1)) bsums[chunk] = (int16_t)s16;
            if (lane == 0) dsc[b] = dd;
        }
    }
}

extern "C" int64_t ggb_act_image_bytes(int64_t k) { return (k > 0 && k % 256 == 0) ? k + k / 4 : -1; }

static int launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, dim3 grid, dim3 block, size_t smem, int use_pdl, cudaStream_t st) {
    cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return 0;
}

extern "C" int ggb_act_prep(const float* x, const float* norm_w, float eps, int64_t k, int nb, int q8_0, void* act, int use_pdl, void* stream) {
    if (k <= 0 || (k % 256) || nb < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_act_prep: k=%lld must be a positive multiple of 256, nb >= 0", (long long)k);
    if (nb == 0) return GGB_OK;
    if (!x || !act || ((uintptr_t)x & 15) || ((uintptr_t)act & 15) || (norm_w && ((uintptr_t)norm_w & 15)))
        GGB_FAIL(GGB_ERR_ARG, "ggb_act_prep: null or misaligned pointer");
    const int nblk = (int)(k / 256);
    /* with RMSNorm every CTA of a token takes the whole row's sum of squares itself (<= 57 KB from L2), so the 256-blocks can
     * still be spread: up to 4 CTAs per token */
    int splits = (nblk + GB_NW - 1) / GB_NW;
    if (norm_w && splits > 4) splits = 4;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    launch_cfg(cfg, at, dim3(nb, splits), dim3(GB_THREADS), 0, use_pdl, (cudaStream_t)stream);
    GGB_CUDA(cudaLaunchKernelEx(&cfg, act_prep_kernel, x, norm_w, eps, (int)k, q8_0, (uint8_t*)act, (int)(k + k / 4)));
    return GGB_OK;
}

// ------------------------------------------------------------------ host side of the batched GEMV
template <int MASK, int NBT, int R>
static int launch_b(const GemvBK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(ggb_dq_gemv_batch_kernel<MASK, NBT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_MAX_SMEM));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    launch_cfg(cfg, at, dim3(grid), dim3(GB_THREADS), smem, use_pdl, st);
    GGB_CUDA(cudaLaunchKernelEx(&cfg, ggb_dq_gemv_batch_kernel<MASK, NBT, R>, P));
    return GGB_OK;
}

template <int MASK>
static int launch_nbt(int nbt, int R, const GemvBK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    if constexpr (MASK == 1) {
        if (R == 4) {
            switch (nbt) {
                case 2: return launch_b<MASK, 2, 4>(P, grid, smem, use_pdl, st);
                case 4: return launch_b<MASK, 4, 4>(P, grid, smem, use_pdl, st);
                default: return launch_b<MASK, 8, 4>(P, grid, smem, use_pdl, st);   /* 16 tokens x 4 rows would spill */
            }
        }
    }
    switch (nbt) {
        case 2: return launch_b<MASK, 2, 2>(P, grid, smem, use_pdl, st);
        case 4: return launch_b<MASK, 4, 2>(P, grid, smem, use_pdl, st);
        case 8: return launch_b<MASK, 8, 2>(P, grid, smem, use_pdl, st);
        default: return launch_b<MASK, 16, 2>(P, grid, smem, use_pdl, st);
    }
}

static int env_int_b(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static size_t batch_smem(const GemvBK& P, int nbt, int64_t max_local) {
    size_t off = (size_t)GB_NW * P.ring_bytes + (size_t)nbt * P.image;
    off = (off + 15) & ~(size_t)15;
    return off + (size_t)max_local * nbt * sizeof(double);
}

int ggb_gemv_batch_mma(const ggb_gemv_batch_args* a, void* stream);   /* gemv_batch_mma.cu: 1 = shape not handled there */

extern "C" int ggb_gemv_batch(const ggb_gemv_batch_args* a, void* stream) {
    if (!a) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: null args");
    if (a->n_seg < 1 || a->n_seg > GGB_MAX_SEG) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: n_seg=%d out of range", a->n_seg);
    if (a->k <= 0 || (a->k % 256)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: k=%d must be a positive multiple of 256", a->k);
    if (a->nb < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: negative token count");
    if (!a->act || ((uintptr_t)a->act & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: activation images null or not 16-byte aligned");
    GemvBK P = {};
    int mask = 0, cls = -1, max_tile = 0;
    int64_t total_rows = 0;
    for (int s = 0; s < a->n_seg; s++) {
        const ggb_gemv_seg& g = a->seg[s];
        int bit;
        switch (g.type) {
            case GGB_TYPE_Q4_K: bit = 1; break;
            case GGB_TYPE_Q6_K: bit = 2; break;
            case GGB_TYPE_Q8_0: bit = 4; break;
            case GGB_TYPE_Q5_K: bit = 8; break;
            default: GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: segment %d has unsupported weight type %d", s, g.type);
        }
        const int c = g.type == GGB_TYPE_Q8_0 ? 1 : 0;
        if (cls >= 0 && cls != c) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: Q8_0 and K-quant segments cannot share a launch");
        cls = c;
        mask |= bit;
        if (g.rows < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: negative row count");
        if (g.rows > 0 && (!g.w || ((uintptr_t)g.w & 15))) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: segment %d weights null or not 16-byte aligned", s);
        P.seg[s].w = (const uint8_t*)g.w;
        P.seg[s].y = g.y;
        P.seg[s].stride = ggb_row_stride(g.type, a->k);
        P.seg[s].type = g.type;
        P.seg[s].rows = g.rows;
        total_rows += g.rows;
        const int tile = ggb_sb_bytes(g.type) * (a->k >= GGB_TILE_ELEMS ? GGB_TILE_SB : a->k / 256);
        if (tile > max_tile) max_tile = tile;
    }
    switch (a->epilogue) {
        case GGB_EPI_STORE:
            for (int s = 0; s < a->n_seg; s++) if (a->seg[s].rows && !a->seg[s].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: STORE needs y for every segment");
            break;
        case GGB_EPI_RESIDUAL:
            if (a->n_seg != 1 || !a->residual || !a->seg[0].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: RESIDUAL needs one segment, y and residual");
            break;
        case GGB_EPI_SWIGLU:
            if (a->n_seg != 2 || a->seg[0].rows != a->seg[1].rows || !a->seg[0].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: SWIGLU needs gate/up segments of equal rows and y on segment 0");
            break;
        case GGB_EPI_STORE_F64:
            if (a->n_seg != 1 || !a->seg[0].y || ((uintptr_t)a->seg[0].y & 7)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: STORE_F64 needs one segment and an 8-byte aligned f64 output [nb][rows]");
            break;
        default: GGB_FAIL(GGB_ERR_ARG, "ggb_gemv_batch: epilogue %d is not available in the batched kernel (STORE, STORE_F64, RESIDUAL, SWIGLU)", a->epilogue);
    }
    if (total_rows == 0 || a->nb == 0) return GGB_OK;
    // from 5 tokens up the integer dots of pure Q4_K / Q6_K launches go to the tensor cores (same arithmetic, half the
    // instructions); everything else stays on the dp4a kernel below
    static const int use_mma = env_int_b("GGB_BATCH_MMA", 1);
    if (a->act_tiled) {   /* tiled activation images exist for the tensor-core kernel only */
        if (mask != 1 && mask != 2 && mask != 3) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: tiled activation images need Q4_K / Q6_K segments");
        const int rc = ggb_gemv_batch_mma(a, stream);
        if (rc == 1) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: shape k=%d does not take tiled activation images", a->k);
        return rc;
    }
    if (use_mma && a->nb >= 5 && (mask == 1 || mask == 2 || mask == 3)) {
        const int rc = ggb_gemv_batch_mma(a, stream);
        if (rc != 1) return rc;
    }
    if (mask & 8) { if (mask & 4) mask = 0; else mask = 11; }   /* any mix with Q5_K runs the generic K-quant kernel */
    if (mask != 1 && mask != 2 && mask != 3 && mask != 4 && mask != 11) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: unsupported type mix (mask %d)", mask);
    P.n_seg = a->n_seg; P.k = a->k; P.T = ggb_tiles_per_row(a->k);
    P.epi = a->epilogue; P.residual = a->residual;
    P.slot_bytes = (max_tile + 15) & ~15;
    static const int r_q4k = env_int_b("GGB_BATCH_R_Q4K", 2), cap_max = env_int_b("GGB_BATCH_CAP", 16);
    const int R = (mask == 1 && r_q4k == 4) ? 4 : 2;
    P.ring_bytes = R * GB_STEPS * P.slot_bytes;
    P.image = a->k + a->k / 4;
    const int grid = a->grid > 0 ? a->grid : ggb_num_sms();
    int64_t max_local = 0;
    for (int s = 0; s < a->n_seg; s++) {
        P.rq[s] = a->seg[s].rows / grid; P.rr[s] = a->seg[s].rows % grid;
        max_local += (a->seg[s].rows + grid - 1) / grid + 2 * GB_MAX_R;
    }
    // tokens per pass: what fits in shared memory next to the ring, at most 8
    int cap = (cap_max >= 16 && R == 2) ? 16 : 8;
    while (cap > 1 && batch_smem(P, cap, max_local) > GB_MAX_SMEM) cap >>= 1;
    if (cap < 2 || batch_smem(P, 2, max_local) > GB_MAX_SMEM)
        GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: k=%d rows=%lld does not fit shared memory", a->k, (long long)total_rows);
    cudaStream_t st = (cudaStream_t)stream;
    for (int b0 = 0; b0 < a->nb; b0 += cap) {
        const int nb = a->nb - b0 < cap ? a->nb - b0 : cap;
        const int nbt = nb <= 2 ? 2 : (nb <= 4 ? 4 : (nb <= 8 ? 8 : 16));
        GemvBK Q = P;
        Q.nb = nb;
        Q.act = (const uint8_t*)a->act + (int64_t)b0 * P.image;
        for (int s = 0; s < a->n_seg; s++) if (Q.seg[s].y) Q.seg[s].y += (int64_t)b0 * Q.seg[s].rows * (a->epilogue == GGB_EPI_STORE_F64 ? 2 : 1);
        if (Q.residual) Q.residual += (int64_t)b0 * Q.seg[0].rows;
        size_t off = (size_t)GB_NW * Q.ring_bytes + (size_t)nbt * Q.image;
        off = (off + 15) & ~(size_t)15;
        Q.rowv_off = (int)off;
        const size_t smem = batch_smem(Q, nbt, max_local);
        int rc;
        switch (mask) {
            case 1: rc = launch_nbt<1>(nbt, R, Q, grid, smem, a->use_pdl, st); break;
            case 2: rc = launch_nbt<2>(nbt, R, Q, grid, smem, a->use_pdl, st); break;
            case 3: rc = launch_nbt<3>(nbt, R, Q, grid, smem, a->use_pdl, st); break;
            case 11: rc = launch_nbt<11>(nbt, R, Q, grid, smem, a->use_pdl, st); break;
            default: rc = launch_nbt<4>(nbt, R, Q, grid, smem, a->use_pdl, st); break;
        }
        if (rc) return rc;
    }
    return GGB_OK;
}
