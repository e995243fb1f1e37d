// gemv.cu -- K1: fused dequant-GEMV for batch-1 decode (the kernel that moves ~99.9 % of the bytes of a
// decoded token).  Stands in for ggml's mul_mat with one activation column:
//   quantize_row_q8_K / quantize_row_q8_0  +  ggml_vec_dot_{q4_K,q6_K}_q8_K / ggml_vec_dot_q8_0_q8_0
// [UPSTREAM-MEM: ggml-quants.c, ggml-cpu/quants.c], plus the element-wise ops either side of it
// (rms_norm*gain in front; residual add, SwiGLU, RoPE + KV-cache write, arg-max partials behind).
//
// Shape of the kernel (one launch = one "phase" of a layer, up to 3 weight matrices sharing the input):
//   grid   one CTA per SM, 16 warps.  CTA c owns rows [rows*c/G, rows*(c+1)/G) of every segment, i.e. ONE
//          contiguous byte range of each weight matrix: HBM is streamed exactly once, in order.
//   ring   every warp owns a private ring of tile slots in shared memory and is its own producer: one lane
//          issues cp.async.bulk (TMA bulk copy, global -> shared, mbarrier complete_tx) for the tiles the
//          warp will consume, several tiles ahead.  Weights do not depend on the previous phase, so the
//          ring is filled BEFORE griddepcontrol.wait: with programmatic dependent launch the next phase's
//          CTA is co-resident (<= 110 KB of shared memory, 64 registers) and its first ~80 KB per SM stream
//          from HBM while the current phase is still computing -- HBM never idles across phase boundaries.
//   prologue  every CTA redundantly normalises + quantises the K-vector into shared memory (int8 codes in a
//          bank-swizzled order, per-block scales, per-16 sums); K <= 28672 floats come from L2.  Doing it
//          per CTA removes a launch and a grid-wide dependency per phase.
//   main loop  a warp walks PAIRS of rows; for each K-tile it loads the 64 int8 activations of its unit once
//          (4 x LDS.128) and applies them to both rows: per row 3 x LDS.128 of packed weights from the ring,
//          in-register unpacking of the 6-bit scales/mins, 16 dp4a, one f32 term, added to an f64 lane sum.
//   reduce one butterfly per row pair (f64 shuffles), ONE rounding to f32 per output.
//   epilogue  the fused element-wise op on the CTA's rows.
// The f64 accumulation makes the result independent of summation order, so the kernel agrees bit-for-bit
// with the oracle's "canon" vec_dot (oracle/ggml_ref.c): greedy parity is decided by logic, not rounding.
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "actquant.cuh"
#include "common.cuh"
#include "layout.cuh"
#include "gemv_common.cuh"
#include "peer.cuh"

#ifndef GEMV_NW
#define GEMV_NW 8                          /* warps per CTA */
#endif
#define GEMV_THREADS (GEMV_NW * 32)
#define GEMV_MAX_R 4
// Per launch shape: R = rows a warp processes together (they share the activation registers), STEPS = ring stages
// per warp (a stage = K-tile t of the R rows of a group, one mbarrier).  Chosen so that the ring of one CTA stays
// <= ~80 KB and two CTAs (this phase and the next, launched early through PDL) fit one SM:
//   Q4_K only        R=4 STEPS=2   8 warps x 8 x 1152 B = 72 KB
//   Q6_K only        R=4 STEPS=2   8 warps x 8 x 1680 B = 105 KB (measured: the wider group beats co-residency here)
//   Q4_K + Q6_K      R=2 STEPS=3   8 warps x 6 x 1680 B = 79 KB
//   Q8_0             R=2 STEPS=2   8 warps x 4 x 2176 B = 68 KB
#define GEMV_MIN_CTAS (GEMV_THREADS <= 256 ? 2 : (GEMV_THREADS <= 512 ? 2 : 1))
#define RING_MAX_SLOTS 8

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
    /* ring geometry (host-computed).  A stage holds K-tile t of the rows of ONE group; a group has 1 << lg[s] rows of
     * segment s (<= R, the kernel's template maximum), each in a slot of slot_b[s] bytes (the segment's tile size). */
    int slot_b[GGB_MAX_SEG], lg[GGB_MAX_SEG];
    int stage_bytes;           /* max over segments of (rows per group * slot) */
    int ring_bytes;            /* per warp = STEPS * stage_bytes: 9 KB (Q4_K) .. 13 KB (Q6_K) -> 72..105 KB per CTA */
    int rowv_off;              /* byte offset of the per-row results in dynamic shared memory */
    unsigned tl_slot;
    int rq[GGB_MAX_SEG], rr[GGB_MAX_SEG];   /* rows = rq*grid + rr: CTA c starts at c*rq + min(c, rr) (no division on the device) */
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
    int peer_n, peer_rank;
    int64_t peer_d_cap;
    uint64_t peer_base[GGB_PEER_MAX];
};

// ------------------------------------------------------------------ optional in-kernel timeline (debug builds only)
#ifdef GGB_TIMELINE
__device__ unsigned long long ggb_tl[8 * 1024 * 8];   /* [launch slot][cta][stamp] */
__device__ unsigned int ggb_tl_launch;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TL_STAMP(i) do { if (threadIdx.x == 0) ggb_tl[((tl_slot & 7) * 1024 + (blockIdx.x & 1023)) * 8 + (i)] = gtime(); } while (0)
#else
#define TL_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ void argmax_comb(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// MASK: bit0 Q4_K, bit1 Q6_K, bit2 Q8_0, bit3 Q5_K segments present (dead-code elimination per launch shape)
// FULLK: every K-tile of the launch is a full tile (k % 2048 == 0) -- the instance carries no partial-tile code at all.  The
// kernel is instruction-fetch bound before it is anything else (ncu: stall_no_instruction is the top stall reason, 31-44 %
// of the samples), so code that cannot run is still worth leaving out: +2.2 % tokens/s on Llama-3-8B.
template <int MASK, int R, int STEPS, bool FULLK = false>
__global__ void __launch_bounds__(GEMV_THREADS, GEMV_MIN_CTAS) ggb_dq_gemv_kernel(const __grid_constant__ GemvK P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ double red[GEMV_NW];
    __shared__ float s_val[GEMV_NW];
    __shared__ int s_idx[GEMV_NW];
    __shared__ __align__(8) uint64_t s_bar[GEMV_NW][STEPS];
    static_assert(R == 2 || R == 4, "row group size");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.k, T = P.T;
#ifdef GGB_TIMELINE
    const unsigned tl_slot = P.tl_slot;
#endif
    TL_STAMP(0);
    uint8_t* qs = smem + GEMV_NW * P.ring_bytes;                  // K bytes
    int16_t* bsums = reinterpret_cast<int16_t*>(qs + K);          // K/16 int16
    float* dsc = reinterpret_cast<float*>(qs + K + K / 8);        // K/32 floats
    double* rowv = reinterpret_cast<double*>(smem + P.rowv_off);  // one f64 row sum per local row (rounded once, in the epilogue)
    const uint32_t STAGE = P.stage_bytes;

    // ---- this CTA's row range in every segment (even-aligned: row pairs and RoPE pairs stay together)
    const int G = gridDim.x, c = blockIdx.x;
    int r0_0, r0_1 = 0, r0_2 = 0, cnt0, cnt1 = 0, cnt2 = 0;
    {
        const bool last = (c + 1 == G);
        auto range = [&](int sg, int& r0, int& cnt) {
            const int q = P.rq[sg], r = P.rr[sg], m = ~((1 << P.lg[sg]) - 1);
            const int a = (c * q + min(c, r)) & m;
            int b = (c + 1) * q + min(c + 1, r);
            if (!last) b &= m;
            r0 = a; cnt = b - a;
        };
        range(0, r0_0, cnt0);
        if (P.n_seg > 1) range(1, r0_1, cnt1);
        if (P.n_seg > 2) range(2, r0_2, cnt2);
    }
    const int nloc = cnt0 + cnt1 + cnt2;
    // local rows are laid out [seg0 | seg1 | seg2]; a group (2 or 4 rows, per segment) never straddles a segment; a count
    // that is not a multiple of the group size (possible only in the last CTA) leaves a short last group.
    const int lg0 = P.lg[0], lg1 = P.lg[1], lg2 = P.lg[2];
    const int np0 = (cnt0 + (1 << lg0) - 1) >> lg0, np1 = (cnt1 + (1 << lg1) - 1) >> lg1, np2 = (cnt2 + (1 << lg2) - 1) >> lg2;
    const int npairs = np0 + np1 + np2;   /* number of row groups of this CTA */

    // group index -> segment, first row, number of rows present (1..rows per group), first local row index
    auto pair_info = [&](int p, int& s, int& row, int& nv, int& lr) {
        if (p < np0) { s = 0; const int o = p << lg0; row = r0_0 + o; nv = min(1 << lg0, cnt0 - o); lr = o; }
        else if (p < np0 + np1) { p -= np0; s = 1; const int o = p << lg1; row = r0_1 + o; nv = min(1 << lg1, cnt1 - o); lr = cnt0 + o; }
        else { p -= np0 + np1; s = 2; const int o = p << lg2; row = r0_2 + o; nv = min(1 << lg2, cnt2 - o); lr = cnt0 + cnt1 + o; }
    };

    // ---- producer (lane 0 of each warp).  One "step" = K-tile t of the R rows of a group = one ring stage
    // (R tile slots, one mbarrier).  The cursor lives in registers: pointer of the next tile of the group's
    // first row, tile index, group index.
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]);
    const uint32_t ring0 = smem_u32(smem) + warp * P.ring_bytes;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STEPS; i++) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nsb_last = ggb_tile_nsb(K, T - 1);
    const uint64_t wpolicy = l2_policy_evict_first();   /* streamed weights must not flush code / activations out of L2 */
    int ip = warp, it = 0, istage = 0;
    const uint8_t* isrc = nullptr;   // tile `it` of the first row of group `ip`
    int istride = 0, itile = 0, ilast = 0;   // row stride, full-tile bytes, last-tile bytes (16-byte rounded)
    int inv = 0, islot = 0;
    auto issue_pair_setup = [&]() {
        int s, row, lr;
        pair_info(ip, s, row, inv, lr);
        const int sbb = ggb_sb_bytes(P.seg[s].type);
        istride = (int)P.seg[s].stride;
        itile = sbb * GGB_TILE_SB;
        ilast = (nsb_last * sbb + 15) & ~15;
        isrc = P.seg[s].w + (int64_t)row * istride;
        islot = P.slot_b[s];
    };
    if (ip < npairs) issue_pair_setup();
    auto issue_step = [&]() {  // lane 0 only: tile `it` of every row of the group into stage `istage`, then advance
        const uint32_t bytes = (it == T - 1) ? (uint32_t)ilast : (uint32_t)itile;
        const uint32_t bar = bar0 + 8 * istage;
        const uint32_t dst = ring0 + istage * STAGE;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)inv * bytes);
        if (lane < inv) bulk_g2s_hint(dst + lane * islot, isrc + (int64_t)lane * istride, bytes, bar, wpolicy);
        istage = (istage + 1 == STEPS) ? 0 : istage + 1;
        isrc += itile;
        if (++it == T) {
            it = 0;
            ip += GEMV_NW;
            if (ip < npairs) issue_pair_setup();
        }
    };
    // fill the ring before waiting for the previous phase (weights are independent of it)
#pragma unroll 1
    for (int i = 0; i < STEPS; i++) if (ip < npairs) issue_step();

    // ---- activation prologue, part 1 (before the dependency wait).  Three shapes, by the number of 256-blocks:
    //   PBR = 2 / 4   every warp keeps its (up to PBR) blocks in registers: ONE trip to L2 for x serves both the RMSNorm
    //                 sum of squares and the quantisation; the RMSNorm gains are weights and are fetched here already
    //   PBR = 0       more than 4 blocks per warp (ffn_down): batches of 4 (2 with RMSNorm), the next batch's loads
    //                 (all 7 blocks of ffn_down resident at once: 124 registers instead of 113 and 667 vs 678 tok/s)
    //                 issued before the current batch is quantised
    constexpr bool Q80 = (MASK == 4);
    const int nblk = K / 256;
    const int PBR = nblk <= 2 * GEMV_NW ? 2 : (nblk <= 4 * GEMV_NW ? 4 : 0);
    const bool norm = (P.pro == GGB_PRO_RMSNORM);
    // The dependency wait sits INSIDE the prologue shapes below (the resident ones fetch their gains first).
    // The dependent launch is triggered right after the wait, not before it: at most the next launch is then resident
    // while this one works, filling its ring.  Triggering before the wait lets launches pile up three deep once
    // co-residency really works (equal shared-memory carveouts, see the host side), and their ring fills then slow the
    // latency-bound prologues of the running launch: 505 vs 543 tok/s.  Triggering after the prologue: no difference.
    float res_pre = 0.f;
    auto wait_dep = [&]() {
        TL_STAMP(1);
        pdl_wait();
#if !defined(GGB_TRIGGER_LATE) && !defined(GGB_TRIGGER_NEVER)
        pdl_launch_dependents();
#endif
        TL_STAMP(2);
        // epilogue operands that only depend on the previous phase: request them now, use them at the end
        if (P.epi == GGB_EPI_RESIDUAL && tid < cnt0) res_pre = P.residual[r0_0 + tid];
    };

    // ---- activation prologue, part 2: (rms_norm * gain) and quantisation into shared memory
    auto load_block = [&](const float* src, int b, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(src + b * 256 + lane * 8), c4 = *reinterpret_cast<const float4*>(src + b * 256 + lane * 8 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
    };
    auto rms_scale = [&](double s) {   /* s = this lane's share of the sum of squares */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < GEMV_NW; i++) tot += red[i];
        const float mean = (float)(tot / (double)K);
        return __fdiv_rn(1.0f, __fsqrt_rn(mean + P.eps));
    };
    // PB blocks (b0, b0 + NW, ...) in registers -> codes, sums, scales in shared memory.  Straight-line: blocks past the end
    // are computed on whatever the clamped load returned and only their stores are predicated, so the PB chains interleave.
    bool redo = false;      /* some block's scale fell outside the exact range of the inline division (never in practice) */
    float rms_used = 1.f;   /* the RMSNorm factor the prologue applied */
    auto quant_store = [&](auto pbc, int b0, float (&v)[decltype(pbc)::value][8], const float (&g)[decltype(pbc)::value][8], float scale) {
        constexpr int PB = decltype(pbc)::value;
        rms_used = scale;
        Q8Codes cq[PB];
        float dd[PB];
        bool all_ok = true;
#pragma unroll
        for (int j = 0; j < PB; j++) {
            if (norm) {
#pragma unroll
                for (int i = 0; i < 8; i++) v[j][i] = __fmul_rn(__fmul_rn(v[j][i], scale), g[j][i]);
            }
            if (Q80) { uint16_t db; cq[j] = warp_quantize_q8_0(v[j], dd[j], db); }
            else { bool ok; cq[j] = warp_quantize_q8_K_sl(v[j], lane, dd[j], ok); all_ok = all_ok && (ok || b0 + j * GEMV_NW >= nblk); }
        }
#pragma unroll
        for (int j = 0; j < PB; j++) {
            const int b = b0 + j * GEMV_NW;
            const bool valid = b < nblk;
            const int chunk = (b * 256 + lane * 8) >> 4;
            uint2* dst = reinterpret_cast<uint2*>(qs + 16 * swz(chunk) + 8 * (lane & 1));
            if (valid) *dst = cq[j].q;
            if (Q80) {
                if (valid && !(lane & 3)) dsc[b * 8 + (lane >> 2)] = dd[j];
            } else {
                const int s16 = cq[j].sum8 + __shfl_xor_sync(0xffffffffu, cq[j].sum8, 1);
                if (valid && !(lane & 1)) bsums[chunk] = (int16_t)s16;
                if (valid && lane == 0) dsc[b] = dd[j];
            }
        }
        if (!Q80 && !all_ok) redo = true;   /* see below: one out-of-range scale redoes the warp's blocks in the reference form */
    };
    auto resident = [&](auto pbc) {
        constexpr int PB = decltype(pbc)::value;
        float v[PB][8], g[PB][8];
        if (norm) {   /* the gains are weights: on their way before the dependency wait */
#pragma unroll
            for (int j = 0; j < PB; j++) { const int bj = warp + j * GEMV_NW; load_block(P.norm_w, bj < nblk ? bj : 0, g[j]); }
        }
        wait_dep();
#pragma unroll
        for (int j = 0; j < PB; j++) { const int bj = warp + j * GEMV_NW; load_block(P.x, bj < nblk ? bj : 0, v[j]); }
        float scale = 1.f;
        if (norm) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < PB; j++) {
                const bool valid = warp + j * GEMV_NW < nblk;
#pragma unroll
                for (int i = 0; i < 8; i++) s += valid ? (double)__fmul_rn(v[j][i], v[j][i]) : 0.0;
            }
            scale = rms_scale(s);
        }
        quant_store(pbc, warp, v, g, scale);
    };
    auto pipelined = [&](auto pbc, float scale) {
        constexpr int PB = decltype(pbc)::value;
        float cur[PB][8], nxt[PB][8], gc[PB][8], gn[PB][8];
        auto fetch = [&](int b0, float (&xv)[PB][8], float (&gg)[PB][8]) {
#pragma unroll
            for (int j = 0; j < PB; j++) {
                const int bj = b0 + j * GEMV_NW;
                load_block(P.x, bj < nblk ? bj : b0, xv[j]);
                if (norm) load_block(P.norm_w, bj < nblk ? bj : b0, gg[j]);
            }
        };
        if (warp < nblk) fetch(warp, cur, gc);
        for (int b = warp; b < nblk; b += PB * GEMV_NW) {
            const int bn = b + PB * GEMV_NW;
            if (bn < nblk) fetch(bn, nxt, gn);
            quant_store(pbc, b, cur, gc, scale);
#pragma unroll
            for (int j = 0; j < PB; j++)
#pragma unroll
                for (int i = 0; i < 8; i++) { cur[j][i] = nxt[j][i]; if (norm) gc[j][i] = gn[j][i]; }
        }
    };
    if (PBR == 2) resident(std::integral_constant<int, 2>());
    else if (PBR == 4) resident(std::integral_constant<int, 4>());
    else if (!norm) { wait_dep(); pipelined(std::integral_constant<int, 4>(), 1.f); }
    else {
        wait_dep();
        double s = 0.0;
        for (int i = tid; i < K; i += GEMV_THREADS) { const float v = P.x[i]; s += (double)__fmul_rn(v, v); }
        pipelined(std::integral_constant<int, 2>(), rms_scale(s));
    }
    // The reference form of the quantiser (__fdiv_rn) for a warp that met an out-of-range scale: ONE rolled copy behind
    // all prologue shapes.  (Inlined into the straight-line quantiser it was 12 unrolled copies that never ran and still
    // cost 3.6 % tokens/s in instruction fetch; as a __noinline__ routine its call cost 15 registers kernel-wide.)
    if (redo) {
#pragma unroll 1
        for (int b = warp; b < nblk; b += GEMV_NW) {
            float w[8], gw[8];
            load_block(P.x, b, w);
            if (norm) {
                load_block(P.norm_w, b, gw);
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = __fmul_rn(__fmul_rn(w[i], rms_used), gw[i]);
            }
            float d;
            const Q8Codes cq = warp_quantize_q8_K(w, lane, d);
            const int chunk = (b * 256 + lane * 8) >> 4;
            *reinterpret_cast<uint2*>(qs + 16 * swz(chunk) + 8 * (lane & 1)) = cq.q;
            const int s16 = cq.sum8 + __shfl_xor_sync(0xffffffffu, cq.sum8, 1);
            if (!(lane & 1)) bsums[chunk] = (int16_t)s16;
            if (lane == 0) dsc[b] = d;
        }
    }
    __syncthreads();
#if defined(GGB_TRIGGER_LATE)
    pdl_launch_dependents();
#endif
    TL_STAMP(3);

    // ---- main loop over this warp's row groups
    const LaneK L = lane_consts(lane);
    const uint32_t qs_s = smem_u32(qs), bs_s = smem_u32(bsums), dsc_s = smem_u32(dsc);
    const int U_last = 4 * nsb_last;
    const bool last_full = (nsb_last == GGB_TILE_SB);
    int cstage = 0;
    uint32_t cphase = 0;   // parity of the current pass over the ring
    for (int p = warp; p < npairs; p += GEMV_NW) {
        int s, row, lr, nv;
        pair_info(p, s, row, nv, lr);
        const int type = P.seg[s].type;
        const int rg = 1 << (s == 0 ? lg0 : (s == 1 ? lg1 : lg2));   /* rows per group of this segment */
        // Rows beyond nv (the short last group of the last CTA) read a duplicate of the last present row's slot and their
        // sums are discarded below: no separate code path for short groups, and the common path stays branch-free so the
        // rows' dependency chains interleave.
        uint32_t ro[R];
#pragma unroll
        for (int r = 0; r < R; r++) ro[r] = (uint32_t)min(r, nv - 1) * (uint32_t)P.slot_b[s];
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        for (int t = 0; t < T; t++) {
            const bool full = FULLK || (t != T - 1) || last_full;
            const int U = full ? 32 : U_last;
            Act A;
            if (lane < U) A = load_act<MASK>(type, t * 32 + lane, qs_s, bs_s, dsc_s);
            const uint32_t slot0 = ring0 + cstage * STAGE;
            mbar_wait(bar0 + 8 * cstage, cphase);
            // TM = the weight format(s) the unrolled rows may hold (one format = no branch between the rows), NR = rows
            // the stage is free again once every lane's reads of it have been issued: refill it.  (Fetching all rows' packed
            // bytes into registers first and refilling BEFORE the arithmetic was measured: 609 vs 643 tok/s, 128 registers.)
            auto refill = [&]() {
                if (++cstage == STEPS) { cstage = 0; cphase ^= 1; }
                __syncwarp();
                if (ip < npairs) issue_step();
            };
            // TM = the weight format(s) the unrolled rows may hold (one format = no branch between the rows), NR = rows
            auto rows = [&](auto tm_c, auto nr_c) {
                constexpr int TM = decltype(tm_c)::value, NR = decltype(nr_c)::value;
                if (full) {
#pragma unroll
                    for (int r = 0; r < NR; r++) acc[r] += consume<TM, true>(type, slot0 + ro[r], lane, 32, GGB_TILE_SB, A, L);
                } else {
#pragma unroll
                    for (int r = 0; r < NR; r++) acc[r] += consume<TM, false>(type, slot0 + ro[r], lane, U, nsb_last, A, L);
                }
                refill();
            };
            if constexpr (MASK == 3) {   /* the unified K-quant instance: Q4_K groups have R rows, Q6_K groups R or R / 2 */
                if (type == GGB_TYPE_Q4_K) rows(std::integral_constant<int, 1>(), std::integral_constant<int, R>());
                else if (rg == R) rows(std::integral_constant<int, 2>(), std::integral_constant<int, R>());
                else rows(std::integral_constant<int, 2>(), std::integral_constant<int, R / 2>());
            } else {
                rows(std::integral_constant<int, MASK>(), std::integral_constant<int, R>());
            }
        }
        // butterfly reduction of the group: after log2(R) exchange levels each lane holds ONE row's partial,
        // then the remaining levels finish all R rows at once.  Row r ends up in lane r * (32 / R).
        {
            if (R == 4 && rg == 4) {
            const bool up16 = lane & 16, up8 = lane & 8;
            double k0 = up16 ? acc[R - 2] : acc[0], k1 = up16 ? acc[R - 1] : acc[1];
            const double s0 = up16 ? acc[0] : acc[R - 2], s1 = up16 ? acc[1] : acc[R - 1];
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            double keep = up8 ? k1 : k0;
            const double send = up8 ? k0 : k1;
            keep += __shfl_xor_sync(0xffffffffu, send, 8);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
            const int r = lane >> 3;
            if ((lane & 7) == 0 && r < nv) rowv[lr + r] = keep;
            } else {
            const bool up = lane & 16;
            const double send = up ? acc[0] : acc[1];
            double keep = up ? acc[1] : acc[0];
            keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
            const int r = lane >> 4;
            if ((lane & 15) == 0 && r < nv) rowv[lr + r] = keep;
            }
        }
    }
    __syncthreads();
    TL_STAMP(4);

    // ---- epilogue
    // (float)rowv[..] below is the only rounding of an accumulated row sum
    switch (P.epi) {
    case GGB_EPI_STORE: {
        for (int lr = tid; lr < nloc; lr += GEMV_THREADS) {
            if (lr < cnt0) P.seg[0].y[r0_0 + lr] = (float)rowv[lr];
            else if (lr < cnt0 + cnt1) P.seg[1].y[r0_1 + lr - cnt0] = (float)rowv[lr];
            else P.seg[2].y[r0_2 + lr - cnt0 - cnt1] = (float)rowv[lr];
        }
    } break;
    case GGB_EPI_STORE_F64: {
        double* y64 = reinterpret_cast<double*>(P.seg[0].y);   /* tensor-parallel partial: summed across ranks before rounding */
        for (int lr = tid; lr < cnt0; lr += GEMV_THREADS) y64[r0_0 + lr] = rowv[lr];
    } break;
    case GGB_EPI_PEER_F64: {
        // tensor-parallel exchange fused into the GEMV: {partial, epoch} words go to every rank's region, slot = my
        // rank (peer.cuh); the data is its own arrival flag, so nothing else is needed on this side
        const int n = P.peer_n;
        const uint8_t* own = reinterpret_cast<const uint8_t*>(P.peer_base[P.peer_rank]);
        const uint32_t e = (uint32_t)(*reinterpret_cast<const volatile int*>(own + ggb_peer_state_off(n, P.peer_d_cap) + 4) + 1);
        const size_t slot = ((size_t)(e & 1) * n + P.peer_rank) * (size_t)P.peer_d_cap;
        for (int i = tid; i < cnt0 * n; i += GEMV_THREADS) {
            const int p = i / cnt0, lr = i - p * cnt0;
            st_ll_f64(reinterpret_cast<uint8_t*>(P.peer_base[p]) + (slot + r0_0 + lr) * 16, rowv[lr], e);
        }
    } break;
    case GGB_EPI_RESIDUAL: {
        for (int lr = tid; lr < cnt0; lr += GEMV_THREADS) {
            const int r = r0_0 + lr;
            P.seg[0].y[r] = __fadd_rn(lr == tid ? res_pre : P.residual[r], (float)rowv[lr]);
        }
    } break;
    case GGB_EPI_SWIGLU: {
        for (int lr = tid; lr < cnt0; lr += GEMV_THREADS) P.seg[0].y[r0_0 + lr] = silu_mul_ref((float)rowv[lr], (float)rowv[cnt0 + lr]);
    } break;
    case GGB_EPI_ROPE_KV: {
        const int pos = *P.pos_dev;
        const float* tab = P.rope_tab + (int64_t)pos * P.n_rot; /* [n_rot/2][2] */
        const int npair = nloc >> 1;
        for (int pr = tid; pr < npair; pr += GEMV_THREADS) {
            int s = 0, l = 2 * pr, r;
            if (l < cnt0) r = r0_0 + l;
            else if (l < cnt0 + cnt1) { s = 1; r = r0_1 + l - cnt0; }
            else { s = 2; r = r0_2 + l - cnt0 - cnt1; }
            float v0 = (float)rowv[2 * pr], v1 = (float)rowv[2 * pr + 1];
            if (s < 2) {
                const int j = r % P.head_dim;
                if (j < P.n_rot) {
                    const float cs = tab[j], sn = tab[j + 1];  /* pair index j/2 -> floats 2*(j/2) = j */
                    const float a = v0, b = v1;
                    v0 = __fsub_rn(__fmul_rn(a, cs), __fmul_rn(b, sn));
                    v1 = __fadd_rn(__fmul_rn(a, sn), __fmul_rn(b, cs));
                }
            }
            if (s == 0) { P.seg[0].y[r] = v0; P.seg[0].y[r + 1] = v1; }
            else {
                uint16_t* cache = (s == 1) ? P.kcache : P.vcache;
                const uint32_t packed = (uint32_t)f2h(v0) | ((uint32_t)f2h(v1) << 16);
                *reinterpret_cast<uint32_t*>(cache + (int64_t)pos * P.seg[s].rows + r) = packed;
            }
        }
    } break;
    case GGB_EPI_ARGMAX: {
        float bv = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int lr = tid; lr < cnt0; lr += GEMV_THREADS) {
            const float v = (float)rowv[lr];
            const int r = r0_0 + lr;
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
    break;
    default: break;
    }
    TL_STAMP(5);
}

#ifdef GGB_TIMELINE
static unsigned g_tl_counter = 0;
extern "C" int ggb_debug_timeline(unsigned long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, ggb_tl, sizeof(unsigned long long) * 8 * 1024 * 8) == cudaSuccess ? 0 : -2;
}
#endif

// ------------------------------------------------------------------ host side
static int act_class(int type) { return type == GGB_TYPE_Q8_0 ? 1 : 0; }

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static int default_grid(const ggb_gemv_args* a) {
    if (a->grid > 0) return a->grid;
    static int per_sm = env_int("GGB_GEMV_CTAS_PER_SM", 1);
    /* experiment: GGB_GEMV_GRID=n CTAs for every launch but the lm-head (read per call: tools/knob_sweep.py switches it) */
    const int forced = env_int("GGB_GEMV_GRID", 0);
    if (forced > 0 && forced <= 2 * ggb_num_sms()) {
        int64_t rows = 0;
        for (int s = 0; s < a->n_seg && s < GGB_MAX_SEG; s++) rows += a->seg[s].rows;
        if (rows < 100000) return forced;
    }
    return ggb_num_sms() * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
}

extern "C" int ggb_gemv_grid(const ggb_gemv_args* a) {
    if (!a) return GGB_ERR_ARG;
    return default_grid(a);
}

#define GEMV_MAX_SMEM (200 * 1024)

template <int MASK, int R, int STEPS, bool FULLK = false>
static int launch(const GemvK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(ggb_dq_gemv_kernel<MASK, R, STEPS, FULLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMV_MAX_SMEM));
        // every decode kernel asks for the SAME (maximal) shared-memory carveout: an SM whose L1/shared split differs from
        // what the next launch prefers must drain before it is reconfigured, which silently defeats the PDL co-residency
        // (in-situ timeline: only launches with equal footprints overlapped)
        GGB_CUDA(cudaFuncSetAttribute(ggb_dq_gemv_kernel<MASK, R, STEPS, FULLK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_done = true;
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
    GGB_CUDA(cudaLaunchKernelEx(&cfg, ggb_dq_gemv_kernel<MASK, R, STEPS, FULLK>, P));
    return GGB_OK;
}

static thread_local int64_t* g_smem_query = nullptr;   /* ggb_gemv_smem_bytes: plan only, report the shared memory, do not launch */

extern "C" int ggb_gemv(const ggb_gemv_args* a, void* stream) {
    if (!a) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: null args");
    if (a->n_seg < 1 || a->n_seg > GGB_MAX_SEG) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: n_seg=%d out of range", a->n_seg);
    if (a->k <= 0 || (a->k % 256)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: k=%d must be a positive multiple of 256", a->k);
    if (!a->x) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: null input vector");
    if (((uintptr_t)a->x & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: x must be 16-byte aligned");
    if (a->prologue == GGB_PRO_RMSNORM && (!a->norm_w || ((uintptr_t)a->norm_w & 15))) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: RMSNORM prologue needs a 16-byte aligned norm_w");
    GemvK P = {};
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
        const int tile = ggb_sb_bytes(g.type) * (a->k >= GGB_TILE_ELEMS ? GGB_TILE_SB : a->k / 256);
        if (tile > max_tile) max_tile = tile;
    }
    switch (a->epilogue) {
        case GGB_EPI_STORE:
            for (int s = 0; s < a->n_seg; s++) if (a->seg[s].rows && !a->seg[s].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: STORE needs y for every segment");
            break;
        case GGB_EPI_STORE_F64:
            if (a->n_seg != 1 || !a->seg[0].y || ((uintptr_t)a->seg[0].y & 7)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: STORE_F64 needs one segment and an 8-byte aligned f64 output");
            break;
        case GGB_EPI_RESIDUAL:
            if (a->n_seg != 1 || !a->residual || !a->seg[0].y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: RESIDUAL needs one segment, y and residual");
            break;
        case GGB_EPI_PEER_F64:
            if (a->n_seg != 1 || a->peer_n < 1 || a->peer_n > GGB_PEER_MAX || a->peer_rank < 0 || a->peer_rank >= a->peer_n ||
                a->peer_d_cap < a->seg[0].rows)
                GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: PEER_F64 needs one segment, 1..%d ranks and a region sized for the rows", GGB_PEER_MAX);
            for (int p = 0; p < a->peer_n; p++) if (!a->peer_base[p]) GGB_FAIL(GGB_ERR_ARG, "ggb_gemv: PEER_F64: region of rank %d is not mapped", p);
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
    P.peer_n = a->peer_n; P.peer_rank = a->peer_rank; P.peer_d_cap = a->peer_d_cap;
    for (int p = 0; p < GGB_PEER_MAX; p++) P.peer_base[p] = a->peer_base[p];
    // Kernel instance and ring geometry.  ONE instance serves every launch of a Q4_K / Q6_K model (mask 1, 2 and 3 all map
    // to <3, 4, 2>): switching between instances costs 2.5-4.5 us per switch -- the instruction caches of the SMs hold one
    // instance's hot code, not two (tools/gemv_bench.py, sequences of shapes) -- and a layer would switch four times.
    //   rows per group  4, except the Q6_K segments of a launch that also has Q4_K segments (q, k, v): 2, so that a stage
    //                   stays 4 x 1152 B and two CTAs (this launch and the next, resident early through PDL) fit one SM
    //   STEPS           2 stages per warp
    // Q5_K mixes use the generic instance <11, 2, 2>, Q8_0 models <4, 2, 2> (one instance per model as well).
    static const int force_generic = env_int("GGB_GEMV_GENERIC", 0);   /* experiment: the generic instance for everything */
    if (force_generic && !(mask & 4)) mask = 11;
    const bool unified = (mask == 1 || mask == 2 || mask == 3);
    const int R = unified ? 4 : 2, STEPS = 2;
    (void)max_tile;
    P.stage_bytes = 0;
    for (int s = 0; s < GGB_MAX_SEG; s++) {
        P.lg[s] = unified ? 2 : 1;
        P.slot_b[s] = 0;
        if (s >= a->n_seg) continue;
        if (mask == 3 && a->seg[s].type == GGB_TYPE_Q6_K) P.lg[s] = 1;
        const int tile = ggb_sb_bytes(a->seg[s].type) * (a->k >= GGB_TILE_ELEMS ? GGB_TILE_SB : a->k / 256);
        P.slot_b[s] = (tile + 15) & ~15;
        const int st = (1 << P.lg[s]) * P.slot_b[s];
        if (st > P.stage_bytes) P.stage_bytes = st;
    }
    P.ring_bytes = STEPS * P.stage_bytes;
#ifdef GGB_TIMELINE
    P.tl_slot = g_tl_counter++;
#endif
    const int grid = default_grid(a);
    for (int s = 0; s < a->n_seg; s++) { P.rq[s] = a->seg[s].rows / grid; P.rr[s] = a->seg[s].rows % grid; }
    int64_t max_local = 0;
    for (int s = 0; s < a->n_seg; s++) max_local += (a->seg[s].rows + grid - 1) / grid + 2 * GEMV_MAX_R;
    size_t off = (size_t)GEMV_NW * P.ring_bytes + (size_t)a->k + a->k / 8 + a->k / 8;
    off = (off + 15) & ~(size_t)15;
    P.rowv_off = (int)off;
    size_t smem = off + (size_t)max_local * sizeof(double);
    // min_smem: the caller may ask for MORE shared memory than the launch needs, so that two CTAs of it can never land on
    // one SM.  A launch that becomes resident while small CTAs of another kernel (the attention) still occupy some SMs
    // is otherwise placed unevenly -- two CTAs here, none there -- and so is everything launched behind it.
    if (a->min_smem > 0 && smem < (size_t)a->min_smem && (size_t)a->min_smem <= GEMV_MAX_SMEM) smem = (size_t)a->min_smem;
    if (smem > GEMV_MAX_SMEM) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: k=%d rows=%lld needs %zu bytes of shared memory", a->k, (long long)total_rows, smem);
    if (g_smem_query) { *g_smem_query = (int64_t)smem; return GGB_OK; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (mask) {
        case 1: case 2: case 3:
            // the caller vouches (full_k_model) that EVERY launch of the model has k % 2048 == 0, so that all of them share the
            // lean instance; a model with one ragged k keeps the generic instance everywhere (two instances evict each other)
            if (a->full_k_model && a->k % GGB_TILE_ELEMS == 0) return launch<3, 4, 2, true>(P, grid, smem, a->use_pdl, st);
            return launch<3, 4, 2>(P, grid, smem, a->use_pdl, st);
        case 4: return launch<4, 2, 2>(P, grid, smem, a->use_pdl, st);
        default:
            if ((mask & 8) && !(mask & 4)) return launch<11, 2, 2>(P, grid, smem, a->use_pdl, st);   /* Q5_K alone or mixed with Q4_K / Q6_K */
            GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv: unsupported type mix (mask %d)", mask);
    }
}

extern "C" int64_t ggb_gemv_smem_bytes(const ggb_gemv_args* a) {
    int64_t v = 0;
    g_smem_query = &v;
    const int rc = ggb_gemv(a, nullptr);
    g_smem_query = nullptr;
    return rc == GGB_OK ? v : (int64_t)rc;
}
