// attn.cu -- K6: KV-cache attention for one query token (decode), grouped-query heads, f16 cache.
// Stands in for ggml's flash_attn_ext / soft_max + mul_mat path on the CPU backend [UPSTREAM-MEM]:
//   q is rounded to f16 (the CPU path converts the K.Q operand to the cache type), scores = q.k / sqrt(hd),
//   softmax in f32, out = P.V normalised at the end.
//
// One thread-block CLUSTER of 8 CTAs per query head (B200: distributed shared memory + cluster barrier):
//   pass 1  every CTA scores its slice of positions (f16 x f16 products are exact in f32; summed in f64) and
//           keeps them in shared memory; every CTA PUSHES its slice maximum into all peers' shared memory (DSMEM
//           stores), one cluster barrier -> global max M from a local array;
//   pass 2  e = exp_ref(s - M); per-CTA f64 partial sums of e and of e*v, pushed into rank 0's shared memory; after the
//           second (last) cluster barrier rank 0 adds them in rank order and writes out = (float)(sum_ev / sum_e).
// A true two-pass softmax (one global max) in a single launch: no split-KV merge kernel, no running-max
// rescaling, and -- because every sum is an f64 sum of f32 terms -- the result does not depend on how the
// positions are split, so it is bit-identical with the oracle's gref_attn_decode_canon.
// The grid does not depend on the position (read from device memory): the launch replays inside a CUDA graph.
#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#ifdef GGB_TIMELINE   /* debug builds: stamps of the LAST launch [cta][entry, wait done, max exchanged, sums exchanged, exit, pass-1 loop done,
                       * pass-2 loop done, partials pushed]: %globaltimer (256 ns ticks, common origin) and the SM's clock64 (durations) */
__device__ unsigned long long ggb_tl_attn[2 * 512 * 8];
#define ATL(i) do { if (threadIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); \
    ggb_tl_attn[(blockIdx.x & 511) * 8 + (i)] = t_; ggb_tl_attn[512 * 8 + (blockIdx.x & 511) * 8 + (i)] = (unsigned long long)clock64(); } } while (0)
extern "C" int ggb_debug_timeline_attn(unsigned long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, ggb_tl_attn, sizeof(ggb_tl_attn)) == cudaSuccess ? 0 : -2;
}
#else
#define ATL(i) do { } while (0)
#endif

// f16 -> f64 in one conversion (SASS F2F.F64.F16).  Both dot products of the attention are sums of EXACT products in f64:
// f16 x f16 (K.Q) and f32 x f16 (e.V) fit 53 bits, so one DFMA per element replaces h2f + FMUL + F2F + DADD, and a K / V element
// converted once serves every query head that shares it (oracle: gref_attn_decode_canon, the same definition).
__device__ __forceinline__ double h2d(uint32_t h16) {
    double d;
    asm("cvt.f64.f16 %0, %1;" : "=d"(d) : "h"((uint16_t)h16));
    return d;
}

#define ATTN_CL 8      /* CTAs per cluster = position slices per head (batch-1 decode) */
#define ATTN_CL_BATCH 2 /* batched decode: many (entry, head) clusters are in flight, fewer CTAs each is cheaper (measured) */
#define ATTN_WARPS 8        /* batch-1 decode: warps per CTA (4 / 8 / 16 measured: 516 / 527 / 456 tok/s) */
#define ATTN_WARPS_BATCH 4  /* batched decode: leaner CTAs win when hundreds of clusters are in flight (measured) */

template <int HD, int CL, int NW>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NW * 32)
attn_decode_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kc, const uint16_t* __restrict__ vc,
                   const int32_t* __restrict__ pos_dev, int n_head, int n_kv, float* __restrict__ out,
                   const int32_t* __restrict__ slot_dev, int64_t slot_stride, int trigger) {
    constexpr int LPG = HD / 8;        // lanes per position (each lane owns 8 consecutive dims = one 16-byte load)
    constexpr int PPW = 32 / LPG;      // positions per warp step
    constexpr int SLOTS = NW * PPW;
    extern __shared__ __align__(16) float s_scores[];   // this CTA's slice of scores
    __shared__ double sm_acc[SLOTS][HD];
    __shared__ double sm_sum[SLOTS];
    __shared__ float sm_max[NW];
    // exchange areas, WRITTEN by the peers through distributed shared memory (pushing costs one cluster barrier;
    // pulling after the barrier would add a remote-read round trip on the critical path)
    __shared__ float cl_max[CL];        // slice maxima of all ranks (every CTA holds a full copy)
    __shared__ double cl_acc[CL][HD];   // rank 0 only: the partial e*v sums of all ranks
    __shared__ double cl_sum[CL];       // rank 0 only: the partial e sums

    ATL(0);
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int head = blockIdx.x / CL;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so that the loops below (whose bounds depend on
    // it) count as convergent and their shuffles are plain SHFL instead of WARPSYNC / ENDCOLLECTIVE sequences
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int sub = lane / LPG, li = lane % LPG;
    const int kvh = head / (n_head / n_kv);
    const int64_t kv_stride = (int64_t)n_kv * HD;

    // batch entry blockIdx.y: its own query row, position and (through slot_dev) cache.  The position is stable for the
    // whole token (the sampler tail of the previous token wrote it and has completed before this token's first launch).
    const int be = blockIdx.y;
    const int pos = pos_dev[be];
    if (pos < 0) return;   /* idle batch entry: the whole cluster leaves before any cluster barrier */
    if (slot_dev) { const int64_t o = (int64_t)slot_dev[be] * slot_stride; kc += o; vc += o; }
    q += (int64_t)be * n_head * HD;
    out += (int64_t)be * n_head * HD;
    const int n = pos + 1;
    int chunk = (n + CL - 1) / CL;
    chunk = (chunk + 7) & ~7;
    const int p_begin = min(n, crank * chunk), p_end = min(n, p_begin + chunk);

    // K and V rows of positions < pos were written by earlier tokens: request the first batch of them BEFORE waiting for
    // the QKV launch that is producing q and row `pos` -- their L2 / HBM latency then hides behind that launch
#ifndef GGB_ATTN_AU
#define GGB_ATTN_AU 4
#endif
    constexpr int AU = GGB_ATTN_AU;
    uint4 kpre[AU], vpre[AU];
#pragma unroll
    for (int u = 0; u < AU; u++) {
        const int p = p_begin + warp * PPW + u * SLOTS + sub;
        const bool old_row = p < p_end && p < pos;
        const int64_t off = (old_row ? p : 0) * kv_stride + (int64_t)kvh * HD + li * 8;
        kpre[u] = *reinterpret_cast<const uint4*>(kc + off);
        vpre[u] = *reinterpret_cast<const uint4*>(vc + off);
    }
    pdl_wait();
    ATL(1);
    // Releasing the output projection here lets it become resident (and fill its weight ring) while the attention runs:
    // 543 -> 555 tok/s -- but ONLY if the projection's CTAs cannot land twice on one SM (ggb_gemv_args.min_smem).
    // Otherwise they are placed unevenly around this kernel's small CTAs -- two here, none there -- the launch behind
    // them inherits the imbalance, and the step gets 10 % slower (485 tok/s; tools/step_timeline.py shows entry times
    // spread over 20 us).  Hence a flag, set by the host when it has padded the projection.
    if (trigger) pdl_launch_dependents();

    double qd[8];
    {
        const float4 a = *reinterpret_cast<const float4*>(q + (int64_t)head * HD + li * 8);
        const float4 b = *reinterpret_cast<const float4*>(q + (int64_t)head * HD + li * 8 + 4);
        const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) qd[i] = h2d(f2h(t[i]));
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));

    // ---- pass 1: scores of my slice + slice maximum.  AU positions per lane group are in flight at once (the loop is
    // a chain of L2 round trips otherwise: one 16-byte load, then its dependent arithmetic)
    float mx = -INFINITY;
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        const bool first = (p0 == p_begin + warp * PPW);
        uint4 kraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            if (first && p < pos) kraw[u] = kpre[u];           /* requested before the dependency wait */
            else kraw[u] = *reinterpret_cast<const uint4*>(kc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            const bool live = p < p_end;
            const uint32_t kw[4] = {kraw[u].x, kraw[u].y, kraw[u].z, kraw[u].w};
            double s = 0.0, s1 = 0.0;                                   // exact products, two chains
#pragma unroll
            for (int i = 0; i < 4; i++) {
                s = __fma_rn(h2d(kw[i] & 0xFFFF), qd[2 * i], s);
                s1 = __fma_rn(h2d(kw[i] >> 16), qd[2 * i + 1], s1);
            }
            s += s1;
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float sf = __fmul_rn((float)s, scale);
            if (live) {
                if (li == 0) s_scores[p - p_begin] = sf;
                mx = fmaxf(mx, sf);
            }
        }
    }
    ATL(5);
    mx = warp_max(mx);
    if (lane == 0) sm_max[warp] = mx;
    __syncthreads();
    if (threadIdx.x < CL) {                     /* thread r hands this slice's maximum to rank r */
        float m = sm_max[0];
#pragma unroll
        for (int w = 1; w < NW; w++) m = fmaxf(m, sm_max[w]);
        cluster.map_shared_rank(cl_max, threadIdx.x)[crank] = m;
    }
    cluster.sync();
    ATL(2);
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < CL; r++) M = fmaxf(M, cl_max[r]);

    // ---- pass 2: e = exp_ref(s - M), f64 partial sums
    double acc[8], sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0;
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        const bool first = (p0 == p_begin + warp * PPW);
        uint4 vraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            if (first && p < pos) vraw[u] = vpre[u];
            else vraw[u] = *reinterpret_cast<const uint4*>(vc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            if (p < p_end) {
                const double e = (double)exp_ref(__fsub_rn(s_scores[p - p_begin], M));
                const uint32_t vw[4] = {vraw[u].x, vraw[u].y, vraw[u].z, vraw[u].w};
                sum += e;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    acc[2 * i] = __fma_rn(e, h2d(vw[i] & 0xFFFF), acc[2 * i]);
                    acc[2 * i + 1] = __fma_rn(e, h2d(vw[i] >> 16), acc[2 * i + 1]);
                }
            }
        }
    }
    ATL(6);
    const int slot = warp * PPW + sub;
    if (li == 0) sm_sum[slot] = sum;
#pragma unroll
    for (int i = 0; i < 8; i++) sm_acc[slot][li * 8 + i] = acc[i];
    __syncthreads();
    {   /* this slice's partial sums go straight into rank 0's shared memory */
        double* acc0 = &cluster.map_shared_rank(&cl_acc[0][0], 0)[crank * HD];
        for (int d = threadIdx.x; d < HD; d += blockDim.x) {
            double a = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < SLOTS; s2++) a += sm_acc[s2][d];
            acc0[d] = a;
        }
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < SLOTS; s2++) t += sm_sum[s2];
            cluster.map_shared_rank(cl_sum, 0)[crank] = t;
        }
    }
    ATL(7);
    cluster.sync();
    ATL(3);
    if (crank == 0) {                           /* rank order, as before: the same f64 sums */
        double S = 0.0;
#pragma unroll
        for (int r = 0; r < CL; r++) S += cl_sum[r];
        for (int d = threadIdx.x; d < HD; d += blockDim.x) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < CL; r++) a += cl_acc[r][d];
            out[(int64_t)head * HD + d] = (float)(a / S);
        }
    }
    ATL(4);
}

// ------------------------------------------------------------------ grouped-query variant (batched decode, long contexts)
// One cluster per (batch entry, GROUP of GQ query heads that share a KV head): a K / V row is loaded ONCE and applied to all
// GQ query heads, so the cache is read n_head / n_kv times less often than with one cluster per query head (4x on
// Llama-3-8B, 4x of the 8x on 70B).  Same arithmetic, same two-pass softmax, same f64 sums: bit-identical results.
//   pass 1  per position one 16-byte load per lane, GQ dots; the GQ partials of the LPG lanes of a position are reduced with
//           a transposing butterfly (log2 GQ levels halve the values a lane holds, the rest are plain levels): lane li ends up
//           with the score of query head li / (LPG / GQ);
//   pass 2  per position one V load, GQ exponentials, GQ x 8 f64 accumulators per lane.
template <int HD, int CL, int NW, int GQ, int AU>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NW * 32)
attn_decode_gqa_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kc, const uint16_t* __restrict__ vc,
                       const int32_t* __restrict__ pos_dev, int n_head, int n_kv, float* __restrict__ out,
                       const int32_t* __restrict__ slot_dev, int64_t slot_stride, int chunk_max) {
    constexpr int LPG = HD / 8, PPW = 32 / LPG, SLOTS = NW * PPW, LQ = LPG / GQ;   /* LQ lanes end up holding one head's score */
    static_assert(GQ == 4 && LPG % GQ == 0, "group of four query heads");
    extern __shared__ __align__(16) uint8_t sm_raw[];
    double* sm_acc = reinterpret_cast<double*>(sm_raw);                 /* [SLOTS][GQ][HD] */
    double* sm_sum = sm_acc + SLOTS * GQ * HD;                          /* [SLOTS][GQ] */
    double* cl_acc = sm_sum + SLOTS * GQ;                               /* [CL][GQ][HD], rank 0 only, written by the peers */
    double* cl_sum = cl_acc + CL * GQ * HD;                             /* [CL][GQ] */
    float* sm_max = reinterpret_cast<float*>(cl_sum + CL * GQ);         /* [NW][GQ] */
    float* cl_max = sm_max + NW * GQ;                                   /* [CL][GQ], every rank holds a full copy */
    float* s_scores = cl_max + CL * GQ;                                 /* [GQ][chunk_max] */

    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int hg = blockIdx.x / CL, head0 = hg * GQ;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   /* warp-uniform for the compiler */
    const int sub = lane / LPG, li = lane % LPG;
    const int kvh = head0 / (n_head / n_kv);
    const int64_t kv_stride = (int64_t)n_kv * HD;
    const int be = blockIdx.y;
    const int pos = pos_dev[be];
    if (pos < 0) return;   /* idle batch entry: the whole cluster leaves before any cluster barrier */
    if (slot_dev) { const int64_t o = (int64_t)slot_dev[be] * slot_stride; kc += o; vc += o; }
    q += (int64_t)be * n_head * HD;
    out += (int64_t)be * n_head * HD;
    const int n = pos + 1;
    int chunk = (n + CL - 1) / CL;
    chunk = (chunk + 7) & ~7;
    const int p_begin = min(n, crank * chunk), p_end = min(n, p_begin + chunk);
    pdl_wait();

    double qd[GQ][8];
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        const float4 a = *reinterpret_cast<const float4*>(q + (int64_t)(head0 + g) * HD + li * 8);
        const float4 b = *reinterpret_cast<const float4*>(q + (int64_t)(head0 + g) * HD + li * 8 + 4);
        const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) qd[g][i] = h2d(f2h(t[i]));
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    const int myg = li / LQ;                       /* the query head whose reduced score this lane holds */

    // ---- pass 1 (AU positions per lane group in flight: the loop is bound by the latency of the cache reads, not by their volume)
    float mx = -INFINITY;
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        uint4 kraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            kraw[u] = *reinterpret_cast<const uint4*>(kc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            const uint32_t kw[4] = {kraw[u].x, kraw[u].y, kraw[u].z, kraw[u].w};
            double kd[8];                                                  // converted ONCE for the four query heads
#pragma unroll
            for (int i = 0; i < 4; i++) { kd[2 * i] = h2d(kw[i] & 0xFFFF); kd[2 * i + 1] = h2d(kw[i] >> 16); }
            double s[GQ];
#pragma unroll
            for (int g = 0; g < GQ; g++) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) { a = __fma_rn(kd[2 * i], qd[g][2 * i], a); b = __fma_rn(kd[2 * i + 1], qd[g][2 * i + 1], b); }   // exact products
                s[g] = a + b;
            }
            // transposing butterfly: 4 values -> 2 -> 1 per lane, then the plain levels below LQ
            {
                const bool up = li & (LPG / 2);
                const double k0 = up ? s[2] : s[0], k1 = up ? s[3] : s[1], d0 = up ? s[0] : s[2], d1 = up ? s[1] : s[3];
                s[0] = k0 + __shfl_xor_sync(0xffffffffu, d0, LPG / 2);
                s[1] = k1 + __shfl_xor_sync(0xffffffffu, d1, LPG / 2);
                const bool up2 = li & (LPG / 4);
                const double k = up2 ? s[1] : s[0], d = up2 ? s[0] : s[1];
                s[0] = k + __shfl_xor_sync(0xffffffffu, d, LPG / 4);
#pragma unroll
                for (int o = LQ / 2; o > 0; o >>= 1) s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
            }
            const float sf = __fmul_rn((float)s[0], scale);
            if (p < p_end) {
                if (li % LQ == 0) s_scores[myg * chunk_max + (p - p_begin)] = sf;
                mx = fmaxf(mx, sf);
            }
        }
    }
    // maximum per query head: over all lanes of the warp that hold the same head (every lane bit except the two head bits)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < LQ || o >= LPG) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (sub == 0 && li % LQ == 0) sm_max[warp * GQ + myg] = mx;
    __syncthreads();
    if (threadIdx.x < CL * GQ) {                 /* thread (r, g) hands this slice's maximum of head g to rank r */
        const int r = threadIdx.x / GQ, g = threadIdx.x % GQ;
        float m = sm_max[g];
#pragma unroll
        for (int w = 1; w < NW; w++) m = fmaxf(m, sm_max[w * GQ + g]);
        cluster.map_shared_rank(cl_max, r)[crank * GQ + g] = m;
    }
    cluster.sync();
    float M[GQ];
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        float m = -INFINITY;
#pragma unroll
        for (int r = 0; r < CL; r++) m = fmaxf(m, cl_max[r * GQ + g]);
        M[g] = m;
    }

    const float M_mine = (li & 3) == 0 ? M[0] : ((li & 3) == 1 ? M[1] : ((li & 3) == 2 ? M[2] : M[3]));
    // ---- pass 2
    double acc[GQ][8], sum[GQ];
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        sum[g] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = 0.0;
    }
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        uint4 vraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            vraw[u] = *reinterpret_cast<const uint4*>(vc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            const bool live = p < p_end;                 /* the two positions of a warp step may differ: shuffles stay outside the branch */
            const uint32_t vw[4] = {vraw[u].x, vraw[u].y, vraw[u].z, vraw[u].w};
            double vd[8];
#pragma unroll
            for (int i = 0; i < 4; i++) { vd[2 * i] = h2d(vw[i] & 0xFFFF); vd[2 * i + 1] = h2d(vw[i] >> 16); }
            // one exponential per lane (lane li takes head li & 3), handed round the position's lanes by four shuffles
            const float e_mine = live ? exp_ref(__fsub_rn(s_scores[(li & 3) * chunk_max + (p - p_begin)], M_mine)) : 0.f;
#pragma unroll
            for (int g = 0; g < GQ; g++) {
                const double e = (double)__shfl_sync(0xffffffffu, e_mine, (lane & ~(LPG - 1)) | g);
                if (live) {
                    sum[g] += e;
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[g][i] = __fma_rn(e, vd[i], acc[g][i]);
                }
            }
        }
    }
    const int slot = warp * PPW + sub;
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        if (li == 0) sm_sum[slot * GQ + g] = sum[g];
#pragma unroll
        for (int i = 0; i < 8; i++) sm_acc[(slot * GQ + g) * HD + li * 8 + i] = acc[g][i];
    }
    __syncthreads();
    {   /* this slice's partial sums go straight into rank 0's shared memory */
        double* acc0 = cluster.map_shared_rank(cl_acc, 0) + crank * GQ * HD;
        for (int o = threadIdx.x; o < GQ * HD; o += NW * 32) {
            double a = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < SLOTS; s2++) a += sm_acc[s2 * GQ * HD + o];
            acc0[o] = a;
        }
        if (threadIdx.x < GQ) {
            double t = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < SLOTS; s2++) t += sm_sum[s2 * GQ + threadIdx.x];
            cluster.map_shared_rank(cl_sum, 0)[crank * GQ + threadIdx.x] = t;
        }
    }
    cluster.sync();
    if (crank == 0) {                           /* rank order: the same f64 sums for any split */
        for (int o = threadIdx.x; o < GQ * HD; o += NW * 32) {
            const int g = o / HD;
            double a = 0.0, S = 0.0;
#pragma unroll
            for (int r = 0; r < CL; r++) { a += cl_acc[r * GQ * HD + o]; S += cl_sum[r * GQ + g]; }
            out[(int64_t)head0 * HD + o] = (float)(a / S);
        }
    }
}

template <int HD, int CL, int NW, int GQ, int AU>
static int launch_attn_gqa(const float* q, const uint16_t* kc, const uint16_t* vc, const int32_t* pos_dev, int n_head, int n_kv,
                           int n_ctx, float* out, int use_pdl, cudaStream_t st, const int32_t* slot_dev, int64_t slot_stride, int nb) {
    constexpr int SLOTS = NW * (32 / (HD / 8));
    int chunk_max = (n_ctx + CL - 1) / CL;
    chunk_max = (chunk_max + 7) & ~7;
    const size_t smem = (size_t)(SLOTS * GQ * HD + SLOTS * GQ + CL * GQ * HD + CL * GQ) * sizeof(double) +
                        (size_t)(NW * GQ + CL * GQ + GQ * chunk_max) * sizeof(float);
    if (smem > 200 * 1024) return 1;            /* context too long for this variant: the caller falls back */
    static bool attr = false;
    if (!attr) {
        GGB_CUDA(cudaFuncSetAttribute(attn_decode_gqa_kernel<HD, CL, NW, GQ, AU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        GGB_CUDA(cudaFuncSetAttribute(attn_decode_gqa_kernel<HD, CL, NW, GQ, AU>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_head / GQ * CL, nb);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (use_pdl & 1) ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, attn_decode_gqa_kernel<HD, CL, NW, GQ, AU>, q, kc, vc, pos_dev, n_head, n_kv, out, slot_dev, slot_stride, chunk_max));
    return GGB_OK;
}

// ------------------------------------------------------------------ long contexts, one sequence: split over the whole GPU
// A cluster of 8 CTAs per head reads a long cache at ~0.6 TB/s (8 000 positions: 52 us per layer, 3.1 ms per token on
// Llama-3-8B): too few warps, one conversion per element per QUERY head.  For contexts beyond ATTN_SPLIT_MIN_CTX the two passes
// of the same two-pass softmax become two launches over (position slices x groups of four query heads) -- every SM busy, each
// K / V element loaded and converted once for the four heads of its group -- with the exchange through a global workspace
// (Llama-3-8B, one sequence: 8 000 positions 3.15 -> 2.33 ms per token, 30 000 positions 7.65 -> 4.08; at 300 positions the
// three launches cost 1.70 vs 1.45 ms, hence the caller's per-step choice):
//   scores   slice of positions x 4 heads: scores -> ws, slice maxima -> ws
//   values   global maximum per head from the slice maxima; e = exp_ref(s - M); f64 partial sums of e and e.V per slice -> ws
//   merge    per head: partials added in slice order, out = (float)(sum_ev / sum_e)
// Same arithmetic (exact products, f64 sums, one rounding): bit-identical with the cluster kernel and the oracle.
#define ATTN_SPLIT_MIN_CTX 2048   /* caches that can hold sequences long enough for the split kernels to pay */
#define ATTN_SPLIT_NW 8
#define ATTN_SPLIT_MAX_S 64

struct SplitWs { float* scores; float* smax; double* part; double* psum; };

template <int HD, int GQ, int AU>
__global__ void __launch_bounds__(ATTN_SPLIT_NW * 32)
attn_split_scores_kernel(const float* __restrict__ q, const uint16_t* __restrict__ kc, const int32_t* __restrict__ pos_dev,
                         int n_head, int n_kv, SplitWs W, int n_ctx_pad, int S) {
    constexpr int NW = ATTN_SPLIT_NW, LPG = HD / 8, PPW = 32 / LPG, SLOTS = NW * PPW, LQ = LPG / GQ;
    static_assert(GQ == 4, "group of four query heads");
    __shared__ float sm_max[NW * GQ];
    const int split = blockIdx.x, head0 = blockIdx.y * GQ;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int sub = lane / LPG, li = lane % LPG;
    const int kvh = head0 / (n_head / n_kv);
    const int64_t kv_stride = (int64_t)n_kv * HD;
    const int n = pos_dev[0] + 1;
    int chunk = (n + S - 1) / S;
    chunk = (chunk + SLOTS - 1) / SLOTS * SLOTS;
    const int p_begin = min(n, split * chunk), p_end = min(n, p_begin + chunk);
    pdl_wait();
    pdl_launch_dependents();
    double qd[GQ][8];
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        const float4 a = *reinterpret_cast<const float4*>(q + (int64_t)(head0 + g) * HD + li * 8);
        const float4 b = *reinterpret_cast<const float4*>(q + (int64_t)(head0 + g) * HD + li * 8 + 4);
        const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) qd[g][i] = h2d(f2h(t[i]));
    }
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn((float)HD));
    const int myg = li / LQ;
    float mx = -INFINITY;
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        uint4 kraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            kraw[u] = *reinterpret_cast<const uint4*>(kc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            const uint32_t kw[4] = {kraw[u].x, kraw[u].y, kraw[u].z, kraw[u].w};
            double kd[8];
#pragma unroll
            for (int i = 0; i < 4; i++) { kd[2 * i] = h2d(kw[i] & 0xFFFF); kd[2 * i + 1] = h2d(kw[i] >> 16); }
            double s[GQ];
#pragma unroll
            for (int g = 0; g < GQ; g++) {
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) { a = __fma_rn(kd[2 * i], qd[g][2 * i], a); b = __fma_rn(kd[2 * i + 1], qd[g][2 * i + 1], b); }
                s[g] = a + b;
            }
            {   /* transposing butterfly, as in attn_decode_gqa_kernel: lane li ends up with head li / LQ */
                const bool up = li & (LPG / 2);
                const double k0 = up ? s[2] : s[0], k1 = up ? s[3] : s[1], d0 = up ? s[0] : s[2], d1 = up ? s[1] : s[3];
                s[0] = k0 + __shfl_xor_sync(0xffffffffu, d0, LPG / 2);
                s[1] = k1 + __shfl_xor_sync(0xffffffffu, d1, LPG / 2);
                const bool up2 = li & (LPG / 4);
                const double k = up2 ? s[1] : s[0], d = up2 ? s[0] : s[1];
                s[0] = k + __shfl_xor_sync(0xffffffffu, d, LPG / 4);
#pragma unroll
                for (int o = LQ / 2; o > 0; o >>= 1) s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
            }
            const float sf = __fmul_rn((float)s[0], scale);
            if (p < p_end) {
                if (li % LQ == 0) W.scores[(int64_t)(head0 + myg) * n_ctx_pad + p] = sf;
                mx = fmaxf(mx, sf);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < LQ || o >= LPG) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (sub == 0 && li % LQ == 0) sm_max[warp * GQ + myg] = mx;
    __syncthreads();
    if (threadIdx.x < GQ) {
        float m = sm_max[threadIdx.x];
#pragma unroll
        for (int w = 1; w < NW; w++) m = fmaxf(m, sm_max[w * GQ + threadIdx.x]);
        W.smax[(head0 + threadIdx.x) * S + split] = m;      /* -inf for an empty slice */
    }
}

template <int HD, int GQ, int AU>
__global__ void __launch_bounds__(ATTN_SPLIT_NW * 32)
attn_split_values_kernel(const uint16_t* __restrict__ vc, const int32_t* __restrict__ pos_dev, int n_head, int n_kv, SplitWs W,
                         int n_ctx_pad, int S) {
    constexpr int NW = ATTN_SPLIT_NW, LPG = HD / 8, PPW = 32 / LPG, SLOTS = NW * PPW;
    __shared__ double sm_acc[NW][GQ][HD];      /* 32 KB */
    __shared__ double sm_sum[NW][GQ];
    __shared__ float sm_M[GQ];
    const int split = blockIdx.x, head0 = blockIdx.y * GQ;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int sub = lane / LPG, li = lane % LPG;
    const int kvh = head0 / (n_head / n_kv);
    const int64_t kv_stride = (int64_t)n_kv * HD;
    const int n = pos_dev[0] + 1;
    int chunk = (n + S - 1) / S;
    chunk = (chunk + SLOTS - 1) / SLOTS * SLOTS;
    const int p_begin = min(n, split * chunk), p_end = min(n, p_begin + chunk);
    // V rows of my first iteration do not depend on the scores launch: request them before the dependency wait
    uint4 vpre[AU];
#pragma unroll
    for (int u = 0; u < AU; u++) {
        const int p = p_begin + warp * PPW + u * SLOTS + sub;
        vpre[u] = *reinterpret_cast<const uint4*>(vc + (p < p_end && p < n - 1 ? p : 0) * kv_stride + (int64_t)kvh * HD + li * 8);
    }
    pdl_wait();
    pdl_launch_dependents();
    if (warp < GQ) {                                  /* global maximum of head head0 + warp over the slices */
        float m = -INFINITY;
        for (int s2 = lane; s2 < S; s2 += 32) m = fmaxf(m, W.smax[(head0 + warp) * S + s2]);
        m = warp_max(m);
        if (lane == 0) sm_M[warp] = m;
    }
    __syncthreads();
    const float M_mine = sm_M[li & 3];
    const float* sc_mine = W.scores + (int64_t)(head0 + (li & 3)) * n_ctx_pad;
    double acc[GQ][8], sum[GQ];
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        sum[g] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = 0.0;
    }
    for (int p0 = p_begin + warp * PPW; p0 < p_end; p0 += AU * SLOTS) {
        const bool first = (p0 == p_begin + warp * PPW);
        uint4 vraw[AU];
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            if (first && p < n - 1) vraw[u] = vpre[u];
            else vraw[u] = *reinterpret_cast<const uint4*>(vc + (p < p_end ? p : p_begin) * kv_stride + (int64_t)kvh * HD + li * 8);
        }
#pragma unroll
        for (int u = 0; u < AU; u++) {
            const int p = p0 + u * SLOTS + sub;
            const bool live = p < p_end;
            const uint32_t vw[4] = {vraw[u].x, vraw[u].y, vraw[u].z, vraw[u].w};
            double vd[8];
#pragma unroll
            for (int i = 0; i < 4; i++) { vd[2 * i] = h2d(vw[i] & 0xFFFF); vd[2 * i + 1] = h2d(vw[i] >> 16); }
            const float e_mine = live ? exp_ref(__fsub_rn(sc_mine[p], M_mine)) : 0.f;
#pragma unroll
            for (int g = 0; g < GQ; g++) {
                const double e = (double)__shfl_sync(0xffffffffu, e_mine, (lane & ~(LPG - 1)) | g);
                if (live) {
                    sum[g] += e;
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[g][i] = __fma_rn(e, vd[i], acc[g][i]);
                }
            }
        }
    }
    // the two positions of a warp step meet by one shuffle level, the warps in shared memory, the slices in the workspace
#pragma unroll
    for (int g = 0; g < GQ; g++) {
        if (PPW == 2) {
            sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], LPG);
#pragma unroll
            for (int i = 0; i < 8; i++) acc[g][i] += __shfl_xor_sync(0xffffffffu, acc[g][i], LPG);
        }
        if (lane == 0) sm_sum[warp][g] = sum[g];
        if (sub == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) sm_acc[warp][g][li * 8 + i] = acc[g][i];
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < GQ * HD; o += NW * 32) {
        const int g = o / HD, d = o % HD;
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) a += sm_acc[w][g][d];
        W.part[((int64_t)(head0 + g) * S + split) * HD + d] = a;
    }
    if (threadIdx.x < GQ) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) t += sm_sum[w][threadIdx.x];
        W.psum[(head0 + threadIdx.x) * S + split] = t;
    }
}

template <int HD>
__global__ void __launch_bounds__(HD) attn_split_merge_kernel(SplitWs W, int S, float* __restrict__ out, int trigger) {
    const int head = blockIdx.x, d = threadIdx.x;
    pdl_wait();
    if (trigger) pdl_launch_dependents();
    double a = 0.0, t = 0.0;
    for (int s2 = 0; s2 < S; s2++) {                  /* slice order: the same f64 sums for any split */
        a += W.part[((int64_t)head * S + s2) * HD + d];
        t += W.psum[head * S + s2];
    }
    out[(int64_t)head * HD + d] = (float)(a / t);
}

static int split_count(int n_head) {
    static const int per_sm = []() { const char* v = getenv("GGB_ATTN_SPLIT_CTAS_PER_SM"); return v && *v ? atoi(v) : 2; }();
    int s = per_sm * ggb_num_sms() / (n_head / 4);
    return s < 1 ? 1 : (s > ATTN_SPLIT_MAX_S ? ATTN_SPLIT_MAX_S : s);
}
static size_t split_ws_layout(int n_head, int head_dim, int n_ctx, SplitWs* W, void* base) {
    const int S = split_count(n_head);
    const size_t n_ctx_pad = ((size_t)n_ctx + 63) & ~(size_t)63;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_part = take((size_t)n_head * S * head_dim * sizeof(double)), o_psum = take((size_t)n_head * S * sizeof(double));
    const size_t o_sc = take((size_t)n_head * n_ctx_pad * sizeof(float)), o_max = take((size_t)n_head * S * sizeof(float));
    if (W && base) {
        uint8_t* b = (uint8_t*)base;
        W->part = (double*)(b + o_part); W->psum = (double*)(b + o_psum); W->scores = (float*)(b + o_sc); W->smax = (float*)(b + o_max);
    }
    return off;
}
// geometry the split kernels cover; whether a launch USES them is the caller's choice per step (use_pdl bit 2), because the
// crossover is a matter of the sequence's current length (~2 000 positions on Llama-3-8B), not of the cache's allocated size
static bool split_geometry(int n_head, int n_kv, int head_dim, int n_ctx) {
    const char* e = getenv("GGB_ATTN_SPLIT");         /* 0: never */
    if (e && *e && atoi(e) == 0) return false;
    return head_dim == 128 && n_head % n_kv == 0 && (n_head / n_kv) % 4 == 0 && n_ctx > ATTN_SPLIT_MIN_CTX;
}
extern "C" size_t ggb_attn_decode_ws_bytes_ctx(int n_head, int n_kv, int head_dim, int n_ctx) {
    if (n_head <= 0 || n_kv <= 0 || n_head % n_kv || n_ctx <= 0) return 16;
    if (!split_geometry(n_head, n_kv, head_dim, n_ctx)) return 16;
    return split_ws_layout(n_head, head_dim, n_ctx, nullptr, nullptr);
}
static int launch_attn_split(const float* q, const uint16_t* kc, const uint16_t* vc, const int32_t* pos_dev, int n_head, int n_kv, int n_ctx,
                             void* ws, float* out, int use_pdl, cudaStream_t st) {
    constexpr int HD = 128, GQ = 4, AU = 2;          /* measured: 2 positions per lane group in flight, 2 CTAs per SM (AU 1/4/8, 1/3/4 per SM slower) */
    SplitWs W;
    split_ws_layout(n_head, HD, n_ctx, &W, ws);
    const int S = split_count(n_head), n_ctx_pad = (n_ctx + 63) & ~63;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = st; cfg.attrs = at; cfg.numAttrs = (use_pdl & 1) ? 1 : 0;
    cfg.gridDim = dim3(S, n_head / GQ); cfg.blockDim = dim3(ATTN_SPLIT_NW * 32);
    GGB_CUDA(cudaLaunchKernelEx(&cfg, attn_split_scores_kernel<HD, GQ, AU>, q, kc, pos_dev, n_head, n_kv, W, n_ctx_pad, S));
    GGB_CUDA(cudaLaunchKernelEx(&cfg, attn_split_values_kernel<HD, GQ, AU>, vc, pos_dev, n_head, n_kv, W, n_ctx_pad, S));
    cfg.gridDim = dim3(n_head); cfg.blockDim = dim3(HD);
    GGB_CUDA(cudaLaunchKernelEx(&cfg, attn_split_merge_kernel<HD>, W, S, out, (use_pdl & 2) ? 1 : 0));
    return GGB_OK;
}

extern "C" size_t ggb_attn_decode_ws_bytes(int n_head, int head_dim) {
    (void)n_head; (void)head_dim;
    return 16; /* the cluster kernel needs no global workspace; kept in the ABI for split-KV variants */
}

template <int HD, int CL, int NW = ATTN_WARPS_BATCH>
static int launch_attn(const float* q, const uint16_t* kc, const uint16_t* vc, const int32_t* pos_dev, int n_head, int n_kv,
                       int n_ctx, float* out, int use_pdl, cudaStream_t st, const int32_t* slot_dev = nullptr, int64_t slot_stride = 0,
                       int nb = 1) {
    const int trigger = (use_pdl & 2) ? 1 : 0;
    int chunk_max = (n_ctx + CL - 1) / CL;
    chunk_max = (chunk_max + 7) & ~7;
    const size_t smem = (size_t)chunk_max * sizeof(float);
    if (smem > 96 * 1024) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_decode: n_ctx=%d too large for the cluster kernel", n_ctx);
    static size_t attr = 0;
    if (smem > attr) {
        GGB_CUDA(cudaFuncSetAttribute(attn_decode_kernel<HD, CL, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        GGB_CUDA(cudaFuncSetAttribute(attn_decode_kernel<HD, CL, NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   /* see gemv.cu */
        attr = 96 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_head * CL, nb);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (use_pdl & 1) ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, attn_decode_kernel<HD, CL, NW>, q, kc, vc, pos_dev, n_head, n_kv, out, slot_dev, slot_stride, trigger));
    return GGB_OK;
}

extern "C" int ggb_attn_decode(const float* q, const uint16_t* kcache, const uint16_t* vcache, const int32_t* pos_dev,
                               int n_head, int n_kv, int head_dim, int n_ctx, void* ws, float* out, int use_pdl, void* stream) {
    if (!q || !kcache || !vcache || !pos_dev || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode: null pointer");
    if (n_head <= 0 || n_kv <= 0 || n_head % n_kv) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode: n_head=%d must be a multiple of n_kv=%d", n_head, n_kv);
    if (n_ctx <= 0) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode: n_ctx must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if ((use_pdl & 4) && split_geometry(n_head, n_kv, head_dim, n_ctx)) {   /* the caller says the sequence is long: three launches over the whole GPU */
        if (!ws || ((uintptr_t)ws & 255)) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode: the long-sequence path needs the 256-byte aligned workspace of ggb_attn_decode_ws_bytes_ctx()");
        return launch_attn_split(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, ws, out, use_pdl, st);
    }
    /* experiment knobs, read per call (tools/knob_sweep.py switches them): CTAs per head, warps per CTA */
    const char* cl_s = getenv("GGB_ATTN_CL");
    const char* nw_s = getenv("GGB_ATTN_NW");
    const int cl1 = cl_s && *cl_s ? atoi(cl_s) : 0, nw1 = nw_s && *nw_s ? atoi(nw_s) : 0;
    if (head_dim == 128 && nw1 == 16 && n_ctx <= 16384) {
        if (cl1 == 2) return launch_attn<128, 2, 16>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
        if (cl1 == 4) return launch_attn<128, 4, 16>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
    }
    if (head_dim == 128 && cl1 == 4) return launch_attn<128, 4, ATTN_WARPS>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
    if (head_dim == 128 && cl1 == 2) return launch_attn<128, 2, ATTN_WARPS>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
    if (head_dim == 128) return launch_attn<128, ATTN_CL, ATTN_WARPS>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
    if (head_dim == 64) return launch_attn<64, ATTN_CL, ATTN_WARPS>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st);
    GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_decode: head_dim=%d (supported: 64, 128)", head_dim);
}

extern "C" int ggb_attn_decode_batch(const float* q, const uint16_t* kcache, const uint16_t* vcache, const int32_t* pos_dev,
                                     const int32_t* slot_dev, int64_t slot_stride, int nb, int n_head, int n_kv, int head_dim,
                                     int n_ctx, float* out, int use_pdl, void* stream) {
    if (nb < 0 || nb > 65535) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode_batch: nb=%d out of range", nb);
    if (nb == 0) return GGB_OK;
    if (!q || !kcache || !vcache || !pos_dev || !slot_dev || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode_batch: null pointer");
    if (n_head <= 0 || n_kv <= 0 || n_head % n_kv) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode_batch: n_head=%d must be a multiple of n_kv=%d", n_head, n_kv);
    if (n_ctx <= 0 || slot_stride < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_attn_decode_batch: n_ctx must be positive, slot_stride non-negative");
    cudaStream_t st = (cudaStream_t)stream;
    // cluster size per (entry, head): with many clusters in flight short contexts are cheaper with fewer CTAs each; long
    // contexts still want the 8-way position split (GGB_ATTN_BATCH_CL overrides: 2, 4 or 8)
    static const int cl_env = []() { const char* v = getenv("GGB_ATTN_BATCH_CL"); return v && *v ? atoi(v) : 0; }();
    const int cl = cl_env ? cl_env : (n_ctx <= 4096 ? ATTN_CL_BATCH : (n_ctx <= 16384 ? 4 : 8));
    // grouped-query variant: four query heads of a KV head per cluster (each K / V row loaded AND converted once for the four).
    // Measured on Llama-3-8B, 16 sequences (tools/batch_bench.py), per-head kernel -> this one: ~50 positions 4.58 -> 4.5 ms per
    // step, ~930 positions 7.24 -> 5.34 (4 positions per lane group in flight), ~8000 positions 23.4 -> 15.4 (8 in flight: with 8
    // warps per SM the loop is bound by the latency of its cache reads).  GGB_ATTN_GQA=0 switches it off, 2 forces it.
    const char* gqa_s = getenv("GGB_ATTN_GQA");            /* read per call (launches are captured into graphs): tests force it on / off */
    const int gqa_env = gqa_s && *gqa_s ? atoi(gqa_s) : 1;
    // (it needs enough clusters to fill the GPU: with few entries -- 2 sequences x 8 KV groups x 4 CTAs = 64 CTAs measured 20.9 ms
    // per step at 30 K positions against 2 x 6.2 ms -- the per-head kernel below keeps four times as many CTAs busy)
    if (gqa_env && head_dim == 128 && (n_head / n_kv) % 4 == 0 && (gqa_env == 2 || nb * (n_head / 4) * 4 >= 2 * ggb_num_sms())) {
        int rc;
        if (n_ctx <= 2048) rc = launch_attn_gqa<128, 4, ATTN_WARPS_BATCH, 4, 4>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        else rc = launch_attn_gqa<128, 4, ATTN_WARPS_BATCH, 4, 8>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        if (rc != 1) return rc;      /* 1: the context does not fit this variant's shared memory */
    }
    if (head_dim == 128) {
        if (cl == 1) return launch_attn<128, 1>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        if (cl == 2) return launch_attn<128, 2>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        if (cl == 4) return launch_attn<128, 4>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        return launch_attn<128, 8>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
    }
    if (head_dim == 64) {
        if (cl == 2) return launch_attn<64, 2>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        if (cl == 4) return launch_attn<64, 4>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
        return launch_attn<64, 8>(q, kcache, vcache, pos_dev, n_head, n_kv, n_ctx, out, use_pdl, st, slot_dev, slot_stride, nb);
    }
    GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_attn_decode_batch: head_dim=%d (supported: 64, 128)", head_dim);
}
