// mega.cu -- K7: one persistent kernel per decoded token (batch-1): every GEMV phase of every layer, the attention
// between them and the lm-head run inside ONE cooperative launch of one CTA per SM, separated by grid barriers
// instead of kernel boundaries.
//
// Why.  A decoded token is ~130 dependent GEMV phases of 9..66 MB each (431 MB for the head).  As separate launches
// every phase pays a fixed 4..8 us during which HBM idles (dependency release, per-CTA quantisation prologue, ramp-up,
// tail) -- more than the streaming time of the small phases (profiles/README.md).  Weights do not depend on
// activations, so here each warp's private shared-memory ring simply keeps running ACROSS phase boundaries: while a
// CTA sits in a grid barrier or quantises the next input vector, the first tiles of the next phase are already in
// (or on their way to) its ring -- up to ~190 KB per SM, i.e. the whole O or QKV projection.
//
//   ring      per warp, byte-granular FIFO of "steps" (one K-tile of the R = 4 rows of a row group), filled by
//             cp.async.bulk + mbarrier exactly as in gemv.cu; the producer cursor walks the static sequence
//             (phase, row group, tile) of its warp and never stops at a phase boundary;
//   phases    described by a table in global memory (one MgPhase per GEMV launch of the multi-kernel path, same
//             segments / prologue / epilogue semantics, same per-CTA row partition);
//   barrier   monotone counter in global memory (zeroed by a memset node before the launch), one arrive per CTA;
//   attention CTA h computes head h alone (two-pass softmax, f64 sums: the arithmetic of attn.cu without the cluster
//             exchange); the other CTAs go straight to the next barrier with their rings full.
// Arithmetic is gemv.cu's / attn.cu's ("canon"), so the result is bit-identical to the multi-kernel path.
// Data written by other CTAs during the launch (x, q, attention output, h, the new K/V row) is read with ld.global.cg.
#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "actquant.cuh"
#include "common.cuh"
#include "layout.cuh"
#include "gemv_common.cuh"

#define MG_NW 8
#define MG_THREADS (MG_NW * 32)
#define MG_R 4
#define MG_NBAR 8            /* mbarriers per warp = steps in flight */
#define MG_SMEM_LIMIT (227 * 1024 - 2048)
#define MG_MAX_CTX 4096      /* attention scores of one head live in shared memory */

struct MgSeg {
    const uint8_t* w;
    float* y;
    int64_t stride;
    int type;
    int rows;
};

struct MgPhase {
    MgSeg seg[GGB_MAX_SEG];
    int n_seg, k, T, pro, epi, act_q8_0;
    int rq[GGB_MAX_SEG], rr[GGB_MAX_SEG];
    const float* x;
    const float* norm_w;
    const float* residual;
    uint16_t* kcache;
    uint16_t* vcache;
    float* attn_out;          /* non-null: attention over (q = seg[0].y, kcache, vcache) follows this phase */
    float eps;
    int pad;
};

struct MgParams {
    const MgPhase* ph;
    int n_ph;
    unsigned* bar;
    const int32_t* pos_dev;
    const float* rope_tab;
    int n_rot, head_dim, n_head, n_kv;
    float* part_val;
    int32_t* part_idx;
    int ring_bytes;            /* per warp */
    int act_off, rowv_off;     /* byte offsets in dynamic shared memory (attention scratch aliases act_off..) */
    unsigned long long* tl;    /* optional [n_ph][grid][6] %globaltimer stamps (tools/mega_timeline.py) */
};

__device__ __forceinline__ float ldcg_f32(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f32x4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint4 ldcg_u32x4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

struct MgCtx { int r0_0, r0_1, r0_2, cnt0, cnt1, cnt2, np0, np1, np2, npairs; };

__device__ __forceinline__ MgCtx mg_ctx(const MgPhase* P, int c, int G) {
    MgCtx X;
    const bool last = (c + 1 == G);
    auto range = [&](int sg, int& r0, int& cnt) {
        const int q = P->rq[sg], r = P->rr[sg];
        const int a = (c * q + min(c, r)) & ~(MG_R - 1);
        int b = (c + 1) * q + min(c + 1, r);
        if (!last) b &= ~(MG_R - 1);
        r0 = a; cnt = b - a;
    };
    const int ns = P->n_seg;
    range(0, X.r0_0, X.cnt0);
    X.r0_1 = X.r0_2 = X.cnt1 = X.cnt2 = 0;
    if (ns > 1) range(1, X.r0_1, X.cnt1);
    if (ns > 2) range(2, X.r0_2, X.cnt2);
    X.np0 = (X.cnt0 + MG_R - 1) / MG_R; X.np1 = (X.cnt1 + MG_R - 1) / MG_R; X.np2 = (X.cnt2 + MG_R - 1) / MG_R;
    X.npairs = X.np0 + X.np1 + X.np2;
    return X;
}
__device__ __forceinline__ void mg_group(const MgCtx& X, int p, int& s, int& row, int& nv, int& lr) {
    if (p < X.np0) { s = 0; row = X.r0_0 + MG_R * p; nv = min(MG_R, X.cnt0 - MG_R * p); lr = MG_R * p; }
    else if (p < X.np0 + X.np1) { p -= X.np0; s = 1; row = X.r0_1 + MG_R * p; nv = min(MG_R, X.cnt1 - MG_R * p); lr = X.cnt0 + MG_R * p; }
    else { p -= X.np0 + X.np1; s = 2; row = X.r0_2 + MG_R * p; nv = min(MG_R, X.cnt2 - MG_R * p); lr = X.cnt0 + X.cnt1 + MG_R * p; }
}

// FIFO placement of a step of `sz` bytes: at the running offset, or at 0 when it does not fit before the end of the
// ring.  (vh, off) = virtual head (monotone, counts the skipped bytes) and its physical offset.  Returns the virtual
// start; the physical start is the updated off - sz.
__device__ __forceinline__ unsigned mg_place(unsigned& vh, unsigned& off, unsigned sz, unsigned ring) {
    if (off + sz > ring) { vh += ring - off; off = 0; }
    const unsigned vs = vh;
    vh += sz; off += sz;
    return vs;
}

__device__ __forceinline__ void mg_grid_sync(unsigned* bar, unsigned& target, unsigned G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += G;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

// ------------------------------------------------------------------ attention of one head by one CTA (attn.cu's arithmetic)
template <int HD>
__device__ __forceinline__ void mg_attention(const float* __restrict__ q, const uint16_t* __restrict__ kc, const uint16_t* __restrict__ vc,
                                             int n, int head, int n_head, int n_kv, float* __restrict__ out, uint8_t* scratch) {
    constexpr int LPG = HD / 8, PPW = 32 / LPG, SLOTS = MG_NW * PPW;
    float* s_scores = reinterpret_cast<float*>(scratch);                                  /* [n] */
    double* sm_acc = reinterpret_cast<double*>(scratch + (size_t)MG_MAX_CTX * 4);         /* [SLOTS][HD] */
    double* sm_sum = sm_acc + SLOTS * HD;                                                 /* [SLOTS] */
    float* sm_max = reinterpret_cast<float*>(sm_sum + SLOTS);                             /* [MG_NW] */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPG, li = lane % LPG;
    const int kvh = head / (n_head / n_kv);
    const int64_t kv_stride = (int64_t)n_kv * HD;

    float qr[8];
    {
        const float4 a = ldcg_f32x4(q + (int64_t)head * HD + li * 8), b = ldcg_f32x4(q + (int64_t)head * HD + li * 8 + 4);
        const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) qr[i] = h2f(f2h(t[i]));
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    float mx = -INFINITY;
    for (int p0 = warp * PPW; p0 < n; p0 += SLOTS) {
        const int p = p0 + sub;
        const bool live = p < n;
        const uint4 kraw = ldcg_u32x4(kc + (live ? p : 0) * kv_stride + (int64_t)kvh * HD + li * 8);
        const uint32_t kw[4] = {kraw.x, kraw.y, kraw.z, kraw.w};
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            s += (double)__fmul_rn(h2f((uint16_t)(kw[i] & 0xFFFF)), qr[2 * i]);
            s += (double)__fmul_rn(h2f((uint16_t)(kw[i] >> 16)), qr[2 * i + 1]);
        }
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float sf = __fmul_rn((float)s, scale);
        if (live) {
            if (li == 0) s_scores[p] = sf;
            mx = fmaxf(mx, sf);
        }
    }
    mx = warp_max(mx);
    if (lane == 0) sm_max[warp] = mx;
    __syncthreads();
    float M = sm_max[0];
#pragma unroll
    for (int w = 1; w < MG_NW; w++) M = fmaxf(M, sm_max[w]);

    double acc[8], sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0;
    for (int p0 = warp * PPW; p0 < n; p0 += SLOTS) {
        const int p = p0 + sub;
        if (p < n) {
            const float e = exp_ref(__fsub_rn(s_scores[p], M));
            const uint4 vraw = ldcg_u32x4(vc + p * kv_stride + (int64_t)kvh * HD + li * 8);
            const uint32_t vw[4] = {vraw.x, vraw.y, vraw.z, vraw.w};
            sum += (double)e;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[2 * i] += (double)__fmul_rn(e, h2f((uint16_t)(vw[i] & 0xFFFF)));
                acc[2 * i + 1] += (double)__fmul_rn(e, h2f((uint16_t)(vw[i] >> 16)));
            }
        }
    }
    const int slot = warp * PPW + sub;
    if (li == 0) sm_sum[slot] = sum;
#pragma unroll
    for (int i = 0; i < 8; i++) sm_acc[slot * HD + li * 8 + i] = acc[i];
    __syncthreads();
    double S = 0.0;
#pragma unroll 8
    for (int s2 = 0; s2 < SLOTS; s2++) S += sm_sum[s2];
    for (int d = threadIdx.x; d < HD; d += MG_THREADS) {
        double a = 0.0;
#pragma unroll 8
        for (int s2 = 0; s2 < SLOTS; s2++) a += sm_acc[s2 * HD + d];
        out[(int64_t)head * HD + d] = (float)(a / S);
    }
    __syncthreads();   /* scratch is reused by the next head / the next phase */
}

__device__ __forceinline__ void mg_argmax_comb(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// MASK: bit0 Q4_K, bit1 Q6_K, bit2 Q8_0 weight types present in the model
template <int MASK>
__global__ void __launch_bounds__(MG_THREADS, 1) mega_kernel(const __grid_constant__ MgParams M) {
    constexpr int R = MG_R;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ double red[MG_NW];
    __shared__ float s_val[MG_NW];
    __shared__ int s_idx[MG_NW];
    __shared__ __align__(8) uint64_t s_bar[MG_NW][MG_NBAR];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, c = blockIdx.x;
    const unsigned RING = (unsigned)M.ring_bytes;
    uint8_t* actb = smem + M.act_off;
    double* rowv = reinterpret_cast<double*>(smem + M.rowv_off);

    const uint32_t bar0 = smem_u32(&s_bar[warp][0]);
    const uint32_t ring0 = smem_u32(smem) + warp * RING;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < MG_NBAR; i++) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---- producer cursor: the next step (K-tile `it` of the rows of group `ip` of phase `iph`) this warp will fetch.
    // Kept uniform across the warp (every lane tracks it; lane 0 issues the copies).
    int iph = 0, ip = warp, it = 0, iT = 1;
    bool idone = false;
    MgCtx ictx = mg_ctx(M.ph, c, G);
    const uint8_t* isrc = nullptr;
    int istride = 0, islot = 0, ilast = 0, inv = 0;
    unsigned ivh = 0, ioff = 0, in = 0;        /* virtual head, physical offset, steps issued */
    unsigned cv = 0, cn = 0;                   /* virtual end of the last consumed step, steps consumed */
    auto igroup = [&]() {                      /* make (iph, ip) a valid group or set idone */
        while (ip >= ictx.npairs) {
            if (++iph >= M.n_ph) { idone = true; return; }
            ictx = mg_ctx(M.ph + iph, c, G);
            ip = warp;
        }
        const MgPhase* P = M.ph + iph;
        int s, row, lr;
        mg_group(ictx, ip, s, row, inv, lr);
        const int sbb = ggb_sb_bytes(P->seg[s].type);
        iT = P->T;
        istride = (int)P->seg[s].stride;
        islot = sbb * GGB_TILE_SB;
        ilast = (ggb_tile_nsb(P->k, iT - 1) * sbb + 15) & ~15;
        isrc = P->seg[s].w + (int64_t)row * istride;
        it = 0;
    };
    igroup();
    auto can_issue = [&]() -> bool {
        if (idone || in - cn >= MG_NBAR) return false;
        const unsigned sz = (unsigned)(R * islot);
        unsigned vh = ivh, off = ioff;
        const unsigned vs = mg_place(vh, off, sz, RING);
        return vs + sz - cv <= RING;
    };
    auto issue = [&]() {
        const unsigned sz = (unsigned)(R * islot);
        mg_place(ivh, ioff, sz, RING);
        const unsigned pos = ioff - sz;
        if (lane == 0) {
            const uint32_t bytes = (it == iT - 1) ? (uint32_t)ilast : (uint32_t)islot;
            const uint32_t bar = bar0 + 8 * (in % MG_NBAR);
            mbar_expect_tx(bar, (uint32_t)inv * bytes);
#pragma unroll
            for (int r = 0; r < R; r++)
                if (r < inv) bulk_g2s(ring0 + pos + r * islot, isrc + (int64_t)r * istride, bytes, bar);
        }
        in++;
        isrc += islot;
        if (++it == iT) { ip += MG_NW; igroup(); }
    };
    while (can_issue()) issue();   /* fill the ring before anything else */

    unsigned cvh = 0, coff = 0;    /* the consumer replays the same placement */
    unsigned bar_target = 0;
    const LaneK L = lane_consts(lane);
    const int pos_tok = *M.pos_dev;

    auto stamp = [&](int ph, int k) {
        if (M.tl && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            M.tl[((size_t)ph * G + c) * 6 + k] = t;
        }
    };
    for (int ph = 0; ph < M.n_ph; ph++) {
        const MgPhase* P = M.ph + ph;
        stamp(ph, 0);
        if (ph > 0) mg_grid_sync(M.bar, bar_target, (unsigned)G);
        stamp(ph, 1);
        const int K = P->k, T = P->T;
        uint8_t* qs = actb;
        int16_t* bsums = reinterpret_cast<int16_t*>(qs + K);
        float* dsc = reinterpret_cast<float*>(qs + K + K / 8);
        const float* xin = P->x;

        // ---- prologue: (rms_norm * gain) and activation quantisation into shared memory (gemv.cu's, x read from L2)
        float scale = 1.f;
        const bool norm = (P->pro == GGB_PRO_RMSNORM);
        if (norm) {
            double s = 0.0;
            for (int i = tid * 4; i < K; i += MG_THREADS * 4) {
                const float4 v = ldcg_f32x4(xin + i);
                s += (double)__fmul_rn(v.x, v.x); s += (double)__fmul_rn(v.y, v.y);
                s += (double)__fmul_rn(v.z, v.z); s += (double)__fmul_rn(v.w, v.w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < MG_NW; i++) tot += red[i];
            const float mean = (float)(tot / (double)K);
            scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + P->eps));
        }
        {
            const int nblk = K / 256;
            constexpr int PB = 4;
            for (int b = warp; b < nblk; b += PB * MG_NW) {
                float4 xa[PB], xb[PB], ga[PB], gb[PB];
#pragma unroll
                for (int j = 0; j < PB; j++) {
                    const int bj = b + j * MG_NW;
                    const int e = (bj < nblk ? bj : b) * 256 + lane * 8;
                    xa[j] = ldcg_f32x4(xin + e);
                    xb[j] = ldcg_f32x4(xin + e + 4);
                    if (norm) { ga[j] = *reinterpret_cast<const float4*>(P->norm_w + e); gb[j] = *reinterpret_cast<const float4*>(P->norm_w + e + 4); }
                }
#pragma unroll
                for (int j = 0; j < PB; j++) {
                    const int bj = b + j * MG_NW;
                    if (bj < nblk) {
                        float v[8] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w, xb[j].x, xb[j].y, xb[j].z, xb[j].w};
                        if (norm) {
                            const float g[8] = {ga[j].x, ga[j].y, ga[j].z, ga[j].w, gb[j].x, gb[j].y, gb[j].z, gb[j].w};
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), g[i]);
                        }
                        const int chunk = (bj * 256 + lane * 8) >> 4;
                        uint2* dst = reinterpret_cast<uint2*>(qs + 16 * swz(chunk) + 8 * (lane & 1));
                        if (P->act_q8_0) {
                            float df; uint16_t db;
                            const Q8Codes cq = warp_quantize_q8_0(v, df, db);
                            *dst = cq.q;
                            if (!(lane & 3)) dsc[bj * 8 + (lane >> 2)] = df;
                        } else {
                            float dd;
                            const Q8Codes cq = warp_quantize_q8_K(v, lane, dd);
                            *dst = cq.q;
                            const int s16 = cq.sum8 + __shfl_xor_sync(0xffffffffu, cq.sum8, 1);
                            if (!(lane & 1)) bsums[chunk] = (int16_t)s16;
                            if (lane == 0) dsc[bj] = dd;
                        }
                    }
                }
            }
        }
        __syncthreads();
        stamp(ph, 2);

        // ---- main loop over this warp's row groups of the phase
        const MgCtx X = mg_ctx(P, c, G);
        const uint32_t qs_s = smem_u32(qs), bs_s = smem_u32(bsums), dsc_s = smem_u32(dsc);
        const int nsb_last = ggb_tile_nsb(K, T - 1);
        const int U_last = 4 * nsb_last;
        const bool last_full = (nsb_last == GGB_TILE_SB);
        for (int p = warp; p < X.npairs; p += MG_NW) {
            int s, row, lr, nv;
            mg_group(X, p, s, row, nv, lr);
            const int type = P->seg[s].type;
            const unsigned slot = (unsigned)(ggb_sb_bytes(type) * GGB_TILE_SB), sz = R * slot;
            double acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0;
            for (int t = 0; t < T; t++) {
                const bool full = (t != T - 1) || last_full;
                const int U = full ? 32 : U_last;
                Act A;
                if (lane < U) A = load_act<MASK>(type, t * 32 + lane, qs_s, bs_s, dsc_s);
                const unsigned vs = mg_place(cvh, coff, sz, RING);
                const uint32_t slot0 = ring0 + (coff - sz);
                mbar_wait(bar0 + 8 * (cn % MG_NBAR), (cn / MG_NBAR) & 1);
                if (full) {
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if (r < nv) acc[r] += consume<MASK, true>(type, slot0 + r * slot, lane, 32, GGB_TILE_SB, A, L);
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if (r < nv) acc[r] += consume<MASK, false>(type, slot0 + r * slot, lane, U, nsb_last, A, L);
                }
                cv = vs + sz;
                cn++;
                __syncwarp();
                while (can_issue()) issue();
            }
            {
                const bool up16 = lane & 16, up8 = lane & 8;
                double k0 = up16 ? acc[2] : acc[0], k1 = up16 ? acc[3] : acc[1];
                const double s0 = up16 ? acc[0] : acc[2], s1 = up16 ? acc[1] : acc[3];
                k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                double keep = up8 ? k1 : k0;
                const double send = up8 ? k0 : k1;
                keep += __shfl_xor_sync(0xffffffffu, send, 8);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
                const int r = lane >> 3;
                if ((lane & 7) == 0 && r < nv) rowv[lr + r] = keep;
            }
        }
        __syncthreads();
        stamp(ph, 3);

        // ---- epilogue (gemv.cu's)
        const int cnt0 = X.cnt0, cnt1 = X.cnt1, r0_0 = X.r0_0, r0_1 = X.r0_1, r0_2 = X.r0_2;
        const int nloc = X.cnt0 + X.cnt1 + X.cnt2;
        if (P->epi == GGB_EPI_STORE) {
            for (int l = tid; l < nloc; l += MG_THREADS) {
                if (l < cnt0) P->seg[0].y[r0_0 + l] = (float)rowv[l];
                else if (l < cnt0 + cnt1) P->seg[1].y[r0_1 + l - cnt0] = (float)rowv[l];
                else P->seg[2].y[r0_2 + l - cnt0 - cnt1] = (float)rowv[l];
            }
        } else if (P->epi == GGB_EPI_RESIDUAL) {
            for (int l = tid; l < cnt0; l += MG_THREADS) {
                const int r = r0_0 + l;
                P->seg[0].y[r] = __fadd_rn(ldcg_f32(P->residual + r), (float)rowv[l]);
            }
        } else if (P->epi == GGB_EPI_SWIGLU) {
            for (int l = tid; l < cnt0; l += MG_THREADS) P->seg[0].y[r0_0 + l] = silu_mul_ref((float)rowv[l], (float)rowv[cnt0 + l]);
        } else if (P->epi == GGB_EPI_ROPE_KV) {
            const float* tab = M.rope_tab + (int64_t)pos_tok * M.n_rot;
            const int npair = nloc >> 1;
            for (int pr = tid; pr < npair; pr += MG_THREADS) {
                int s = 0, l = 2 * pr, r;
                if (l < cnt0) r = r0_0 + l;
                else if (l < cnt0 + cnt1) { s = 1; r = r0_1 + l - cnt0; }
                else { s = 2; r = r0_2 + l - cnt0 - cnt1; }
                float v0 = (float)rowv[2 * pr], v1 = (float)rowv[2 * pr + 1];
                if (s < 2) {
                    const int j = r % M.head_dim;
                    if (j < M.n_rot) {
                        const float cs = tab[j], sn = tab[j + 1];
                        const float a = v0, b = v1;
                        v0 = __fsub_rn(__fmul_rn(a, cs), __fmul_rn(b, sn));
                        v1 = __fadd_rn(__fmul_rn(a, sn), __fmul_rn(b, cs));
                    }
                }
                if (s == 0) { P->seg[0].y[r] = v0; P->seg[0].y[r + 1] = v1; }
                else {
                    uint16_t* cache = (s == 1) ? P->kcache : P->vcache;
                    const uint32_t packed = (uint32_t)f2h(v0) | ((uint32_t)f2h(v1) << 16);
                    *reinterpret_cast<uint32_t*>(cache + (int64_t)pos_tok * P->seg[s].rows + r) = packed;
                }
            }
        } else if (P->epi == GGB_EPI_ARGMAX) {
            float bv = -FLT_MAX;
            int bi = 0x7fffffff;
            for (int l = tid; l < cnt0; l += MG_THREADS) {
                const float v = (float)rowv[l];
                const int r = r0_0 + l;
                if (P->seg[0].y) P->seg[0].y[r] = v;
                mg_argmax_comb(bv, bi, v, r);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mg_argmax_comb(bv, bi, __shfl_xor_sync(0xffffffffu, bv, o), __shfl_xor_sync(0xffffffffu, bi, o));
            if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int i = 1; i < MG_NW; i++) mg_argmax_comb(bv, bi, s_val[i], s_idx[i]);
                M.part_val[blockIdx.x] = bv;
                M.part_idx[blockIdx.x] = bi;
            }
        }

        // ---- attention between the QKV phase and the output projection: CTA h computes head h
        stamp(ph, 4);
        if (P->attn_out) {
            mg_grid_sync(M.bar, bar_target, (unsigned)G);
            stamp(ph, 5);
            for (int head = c; head < M.n_head; head += G) {
                if (M.head_dim == 128) mg_attention<128>(P->seg[0].y, P->kcache, P->vcache, pos_tok + 1, head, M.n_head, M.n_kv, P->attn_out, actb);
                else mg_attention<64>(P->seg[0].y, P->kcache, P->vcache, pos_tok + 1, head, M.n_head, M.n_kv, P->attn_out, actb);
            }
        }
    }
}

// ------------------------------------------------------------------ host side
static int mg_bit(int type) {
    return type == GGB_TYPE_Q4_K ? 1 : (type == GGB_TYPE_Q6_K ? 2 : (type == GGB_TYPE_Q8_0 ? 4 : (type == GGB_TYPE_Q5_K ? 8 : 0)));
}

extern "C" int64_t ggb_mega_plan_bytes(int n_phases) { return n_phases > 0 ? (int64_t)n_phases * (int64_t)sizeof(MgPhase) : -1; }

// Build the device-side phase table from the same launch descriptions the multi-kernel path uses.  attn_out[i] != NULL
// marks a phase that is followed by attention (its epilogue must be ROPE_KV).  Synchronous (set-up time only).
extern "C" int ggb_mega_plan(const ggb_gemv_args* phases, const float* const* attn_out, int n_phases, void* plan_dev, int* mask_out) {
    if (!phases || n_phases <= 0 || !plan_dev || !mask_out) GGB_FAIL(GGB_ERR_ARG, "ggb_mega_plan: bad argument");
    const int grid = ggb_num_sms();
    MgPhase* h = (MgPhase*)calloc((size_t)n_phases, sizeof(MgPhase));
    if (!h) GGB_FAIL(GGB_ERR_ARG, "ggb_mega_plan: out of host memory");
    int mask = 0, rc = GGB_OK;
    for (int i = 0; i < n_phases && rc == GGB_OK; i++) {
        const ggb_gemv_args* a = phases + i;
        MgPhase& P = h[i];
        if (a->n_seg < 1 || a->n_seg > GGB_MAX_SEG || a->k <= 0 || (a->k % 256) || !a->x) { rc = GGB_ERR_ARG; break; }
        if (a->epilogue == GGB_EPI_STORE_F64 || a->epilogue == GGB_EPI_PEER_F64) { rc = GGB_ERR_UNSUPPORTED; break; }
        int cls = -1;
        for (int s = 0; s < a->n_seg; s++) {
            const int b = mg_bit(a->seg[s].type);
            const int cl = a->seg[s].type == GGB_TYPE_Q8_0 ? 1 : 0;
            if (!b || (cls >= 0 && cls != cl) || !a->seg[s].w) { rc = GGB_ERR_UNSUPPORTED; break; }
            cls = cl;
            mask |= b;
            P.seg[s].w = (const uint8_t*)a->seg[s].w; P.seg[s].y = a->seg[s].y;
            P.seg[s].stride = ggb_row_stride(a->seg[s].type, a->k);
            P.seg[s].type = a->seg[s].type; P.seg[s].rows = a->seg[s].rows;
            P.rq[s] = a->seg[s].rows / grid; P.rr[s] = a->seg[s].rows % grid;
        }
        P.n_seg = a->n_seg; P.k = a->k; P.T = ggb_tiles_per_row(a->k);
        P.pro = a->prologue; P.epi = a->epilogue; P.act_q8_0 = cls;
        P.x = a->x; P.norm_w = a->norm_w; P.residual = a->residual; P.eps = a->eps;
        P.kcache = a->kcache; P.vcache = a->vcache;
        P.attn_out = attn_out ? (float*)attn_out[i] : nullptr;
        if (P.attn_out && a->epilogue != GGB_EPI_ROPE_KV) rc = GGB_ERR_ARG;
    }
    if (rc == GGB_OK && cudaMemcpy(plan_dev, h, (size_t)n_phases * sizeof(MgPhase), cudaMemcpyHostToDevice) != cudaSuccess) rc = GGB_ERR_CUDA;
    free(h);
    if (rc != GGB_OK) GGB_FAIL(rc, "ggb_mega_plan: phase table rejected (unsupported type mix, epilogue or null pointer)");
    *mask_out = mask;
    return GGB_OK;
}

template <int MASK>
static int mg_launch(const MgParams& M, size_t smem, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(mega_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM_LIMIT));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ggb_num_sms());
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;     /* all CTAs co-resident: the grid barrier cannot deadlock */
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, mega_kernel<MASK>, M));
    return GGB_OK;
}

extern "C" int ggb_mega_run(const ggb_mega_args* a, void* stream) {
    if (!a || !a->plan || a->n_phases <= 0 || !a->barrier || !a->pos_dev || !a->rope_tab || !a->part_val || !a->part_idx)
        GGB_FAIL(GGB_ERR_ARG, "ggb_mega_run: null argument");
    if (a->n_ctx > MG_MAX_CTX) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_mega_run: n_ctx=%d exceeds the single-CTA attention limit %d", a->n_ctx, MG_MAX_CTX);
    if (a->head_dim != 64 && a->head_dim != 128) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_mega_run: head_dim=%d (supported: 64, 128)", a->head_dim);
    if (a->n_head <= 0 || a->n_kv <= 0 || a->n_head % a->n_kv || a->k_max <= 0 || (a->k_max % 256) || a->rows_max <= 0)
        GGB_FAIL(GGB_ERR_ARG, "ggb_mega_run: bad shape");
    MgParams M = {};
    M.ph = (const MgPhase*)a->plan; M.n_ph = a->n_phases; M.bar = (unsigned*)a->barrier;
    M.pos_dev = a->pos_dev; M.rope_tab = a->rope_tab; M.n_rot = a->n_rot; M.head_dim = a->head_dim;
    M.n_head = a->n_head; M.n_kv = a->n_kv; M.part_val = a->part_val; M.part_idx = a->part_idx;
    M.tl = (unsigned long long*)a->timeline;
    // shared memory: [8 rings][union(activations + row results, attention scratch)]
    const int grid = ggb_num_sms();
    const size_t act = (size_t)a->k_max + a->k_max / 4;
    const size_t rowv = ((size_t)(a->rows_max + grid - 1) / grid + 2 * MG_R * GGB_MAX_SEG) * sizeof(double);
    const int slots = 8 * 2 * (a->head_dim == 128 ? 1 : 2);   /* MG_NW * PPW */
    const size_t attn = (size_t)MG_MAX_CTX * 4 + (size_t)slots * a->head_dim * 8 + (size_t)slots * 8 + 64;
    size_t uni = ((act + 15) & ~(size_t)15) + rowv;
    if (attn > uni) uni = attn;
    uni = (uni + 127) & ~(size_t)127;
    if (uni + MG_NW * 2 * MG_R * 2176 > MG_SMEM_LIMIT)   /* a ring must hold two of the largest steps */ GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_mega_run: k=%d does not leave room for the weight rings", a->k_max);
    size_t ring = ((MG_SMEM_LIMIT - uni) / MG_NW) & ~(size_t)127;
    M.ring_bytes = (int)ring;
    M.act_off = (int)(ring * MG_NW);
    M.rowv_off = (int)(M.act_off + ((act + 15) & ~(size_t)15));
    const size_t smem = ring * MG_NW + uni;
    cudaStream_t st = (cudaStream_t)stream;
    GGB_CUDA(cudaMemsetAsync(a->barrier, 0, 4, st));
    switch (a->type_mask) {
        case 1: case 2: case 3: return mg_launch<3>(M, smem, st);
        case 4: return mg_launch<4>(M, smem, st);
        default:
            if ((a->type_mask & 8) && !(a->type_mask & 4)) return mg_launch<11>(M, smem, st);
            GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_mega_run: unsupported weight type mix (mask %d)", a->type_mask);
    }
}
