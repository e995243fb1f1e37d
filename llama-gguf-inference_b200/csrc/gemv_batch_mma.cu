// gemv_batch_mma.cu -- K1c: the small-batch decode GEMV with the integer dots on the tensor cores (mma.sync, u8 x s8).
//
// gemv_batch.cu is issue-bound from 8 tokens up: 16 dp4a + ~20 other instructions per (row, token, 64-weight unit).
// Here one mma.sync.m16n8k32 (Q4_K) / m16n8k16 (Q6_K) yields the 32- / 16-element sub-block dots of 16 rows x 8 tokens
// at once -- exactly the integers the scalar path sums with dp4a -- and the C fragment leaves 4 (row, token) results
// per lane, so the scale arithmetic stays perfectly distributed.  What follows the integer dot is unchanged: the
// same per-unit f32 term, added into f64 ("canon"), so the result is bit-identical to gemv.cu / gemv_batch.cu.
//
//   CTA      8 consumer warps + 1 producer warp; a row group = 16 rows; a ring stage = K-tile t of the 16 rows
//            (cp.async.bulk + full/empty mbarriers, filled before griddepcontrol.wait like every GEMV here);
//   K split  a K-tile holds 8 super-blocks: consumer warp w owns super-block w of every tile, for all 16 rows and all
//            tokens; the 8 partial sums of a (row, token) meet in shared memory once per row group;
//   A        the packed weights are read straight from the ring as 32-bit words (rows padded by 16 B so the 8 rows of
//            a fragment land in distinct banks), masked / shifted into u8 fragments in registers;
//   B        int8 activations from the same "activation images" gemv_batch.cu uses (token stride padded likewise).
// Launches of pure Q4_K, pure Q6_K or their mix use it when >= 5 tokens are batched (GGB_BATCH_MMA=0 disables);
// Q5_K / Q8_0 stay on the dp4a kernel.
#include <float.h>
#include <stdlib.h>

#include "actquant.cuh"
#include "common.cuh"
#include "layout.cuh"
#include "gemv_common.cuh"
#include "../../include/ggufb200.h"

#define GBM_NW 8                      /* consumer warps */
#define GBM_THREADS ((GBM_NW + 1) * 32)
#define GBM_ROWS 16
#define GBM_RS 18                     /* doubles per (super-block, token) row of the partial-sum buffer: 16 rows + 2, so that the 16 lanes of a
                                       * half-warp -- rows g = 0..3, tokens 2*t4 + q, i.e. words 36 t4 + g -- hit 16 different 8-byte banks */
#define GBM_MAX_STAGES 6
#define GBM_MAX_SMEM (226 * 1024)

struct SegM {
    const uint8_t* w;
    float* y;
    int64_t stride;
    int type;
    int rows;
};

struct GbmK {
    SegM seg[GGB_MAX_SEG];
    int n_seg, k, T, epi, nb;
    int slotp, stage_bytes, nstage;     /* padded tile slot, 16 slots, ring depth */
    int image, imgp;                    /* activation image bytes / padded token stride */
    int act_off, red_off, rowv_off, rowv_ld;   /* rowv_ld: local rows per token of the SWIGLU row buffer */
    int rq[GGB_MAX_SEG], rr[GGB_MAX_SEG];
    const uint8_t* act;
    const float* residual;
};

__device__ __forceinline__ void mma_u8s8_k32(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8s8_k16(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mbar_arrive1(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GBM_NW * 32) : "memory"); }

struct GbmCtx { int r0[GGB_MAX_SEG], cnt[GGB_MAX_SEG], ng[GGB_MAX_SEG], ngroups; };

// One ring stage = K-tile t of a 16-row group: consumer warp `warp` applies super-block `warp` of the tile to every token.
// The activation side is described by (act_s, imgp, bs_off, dsc_off, gsb0): whole-vector images (gsb0 = 8 t, offsets K and
// K + K/8) or the K-tile slice of the tiled layout (gsb0 = 0, fixed offsets) -- the arithmetic is the same.
template <int NT>
__device__ __forceinline__ void gbm_tile_step(int type, uint32_t stage, int slotp, int nsb, int warp, int lane, uint32_t act_s, int imgp,
                                              uint32_t bs_off, uint32_t dsc_off, int gsb0, double (&acc)[NT][4]) {
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t S = 64u * (uint32_t)nsb;                     /* bytes of a 16-byte-per-unit section */
    if (warp < nsb) {
        const int sb = warp;
        const uint32_t rowA = stage + g * slotp, rowB = rowA + 8 * slotp;
        const int gsb = gsb0 + sb;                              /* super-block index within the activation image */
        // block scales of my C-fragment tokens (2*t4, 2*t4+1 of every n-tile)
        float dx[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            dx[nt][0] = __uint_as_float(lds32(act_s + (nt * 8 + 2 * t4) * imgp + dsc_off + 4 * gsb));
            dx[nt][1] = __uint_as_float(lds32(act_s + (nt * 8 + 2 * t4 + 1) * imgp + dsc_off + 4 * gsb));
        }
        // Activation addresses without per-access arithmetic.  Chunk c of the image sits at swz(c) = c ^ ((c >> 2) & 7); inside my
        // super-block c = 16 gsb + l (l = 0..15), so swz(c) = 16 gsb + (l ^ x ^ 4 par) with x = (l >> 2) + (constant of the unit) known
        // at compile time and par = gsb & 1: the parity moves the chunk by +-64 bytes, the sign given by bit 2 of l ^ x.  Two base
        // registers per n-tile (plus / minus) and immediates replace a LOP3 + IMAD / LEA per load.
        const int par = gsb & 1;
        uint32_t bp[NT], bm[NT], bq[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const uint32_t tb = act_s + (nt * 8 + g) * imgp + 4 * t4 + 256 * gsb;     /* B fragment: token nt*8 + g */
            bp[nt] = tb + 64 * par; bm[nt] = tb - 64 * par;
#pragma unroll
            for (int q = 0; q < 2; q++) bq[nt][q] = act_s + (nt * 8 + 2 * t4 + q) * imgp + bs_off + 32 * gsb;   /* per-16 sums of C token 2*t4 + q */
        }
#define GBM_BADDR(nt, l, x) ((((((l) ^ (x)) >> 2) & 1) ? bm[nt] : bp[nt]) + 16 * ((l) ^ (x)))
        if (type == GGB_TYPE_Q4_K) {
            const uint4 hA = lds128(rowA + 2 * S + 16 * sb), hB = lds128(rowB + 2 * S + 16 * sb);
            const float dA = h2f((uint16_t)(hA.x & 0xFFFF)), mA = h2f((uint16_t)(hA.x >> 16));
            const float dB = h2f((uint16_t)(hB.x & 0xFFFF)), mB = h2f((uint16_t)(hB.x >> 16));
            const uint32_t hwA[4] = {hA.x, hA.y, hA.z, hA.w}, hwB[4] = {hB.x, hB.y, hB.z, hB.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int u = 4 * sb + j;
                const uint32_t wA0 = lds32(rowA + 16 * u + 4 * t4), wA1 = lds32(rowA + S + 16 * u + 4 * t4);
                const uint32_t wB0 = lds32(rowB + 16 * u + 4 * t4), wB1 = lds32(rowB + S + 16 * u + 4 * t4);
                // field j (24 bits at header byte 4 + 3j): sc[2j] | sc[2j+1] << 6 | min[2j] << 12 | min[2j+1] << 18
                const int bo = 4 + 3 * j;
                const uint32_t fA = (uint32_t)((((uint64_t)hwA[bo >> 2] | ((uint64_t)hwA[(bo >> 2) + ((bo >> 2) < 3 ? 1 : 0)] << 32)) >> (8 * (bo & 3))) & 0xFFFFFFu);
                const uint32_t fB = (uint32_t)((((uint64_t)hwB[bo >> 2] | ((uint64_t)hwB[(bo >> 2) + ((bo >> 2) < 3 ? 1 : 0)] << 32)) >> (8 * (bo & 3))) & 0xFFFFFFu);
                const int mmA0 = (int)(((fA >> 12) & 63) * 0x0101u), mmA1 = (int)(((fA >> 18) & 63) * 0x0101u);   /* (m, m) byte pairs */
                const int mmB0 = (int)(((fB >> 12) & 63) * 0x0101u), mmB1 = (int)(((fB >> 18) & 63) * 0x0101u);
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    int clo[4] = {0, 0, 0, 0}, chi[4] = {0, 0, 0, 0};              /* chunks 4j .. 4j+3 of the super-block: l = 4j + c, x = j */
                    mma_u8s8_k32(clo, wA0 & 0x0F0F0F0Fu, wB0 & 0x0F0F0F0Fu, wA1 & 0x0F0F0F0Fu, wB1 & 0x0F0F0F0Fu,
                                 lds32(GBM_BADDR(nt, 4 * j, j)), lds32(GBM_BADDR(nt, 4 * j + 1, j)));
                    mma_u8s8_k32(chi, (wA0 >> 4) & 0x0F0F0F0Fu, (wB0 >> 4) & 0x0F0F0F0Fu, (wA1 >> 4) & 0x0F0F0F0Fu, (wB1 >> 4) & 0x0F0F0F0Fu,
                                 lds32(GBM_BADDR(nt, 4 * j + 2, j)), lds32(GBM_BADDR(nt, 4 * j + 3, j)));
#pragma unroll
                    for (int q = 0; q < 2; q++) {                                     /* C-fragment tokens 2*t4 + q */
                        const uint2 bs = lds64(bq[nt][q] + 8 * j);                    /* four per-16 sums (int16): sub-block 2j in .x, 2j+1 in .y */
                        const float x = dx[nt][q];
                        {   /* min[2j] * (s0 + s1) + min[2j+1] * (s2 + s3) as two 2-way dots of the int16 pairs with the byte pairs (m, m) */
                            const int isum = (int)(fA & 63) * clo[q] + (int)((fA >> 6) & 63) * chi[q];
                            const int msum = __dp2a_lo((int)bs.y, mmA1, __dp2a_lo((int)bs.x, mmA0, 0));
                            acc[nt][q] += (double)__fsub_rn(__fmul_rn(__fmul_rn(dA, x), (float)isum), __fmul_rn(__fmul_rn(mA, x), (float)msum));
                        }
                        {
                            const int isum = (int)(fB & 63) * clo[2 + q] + (int)((fB >> 6) & 63) * chi[2 + q];
                            const int msum = __dp2a_lo((int)bs.y, mmB1, __dp2a_lo((int)bs.x, mmB0, 0));
                            acc[nt][2 + q] += (double)__fsub_rn(__fmul_rn(__fmul_rn(dB, x), (float)isum), __fmul_rn(__fmul_rn(mB, x), (float)msum));
                        }
                    }
                }
            }
        } else {   /* Q6_K: unit (half n, column tt) = four 16-element groups r with their own int8 scale */
            const uint4 scA = lds128(rowA + 3 * S + 16 * sb), scB = lds128(rowB + 3 * S + 16 * sb);
            const float dA = h2f((uint16_t)lds16(rowA + 3 * S + 16 * nsb + 2 * sb)), dB = h2f((uint16_t)lds16(rowB + 3 * S + 16 * nsb + 2 * sb));
            const uint32_t swA[4] = {scA.x, scA.y, scA.z, scA.w}, swB[4] = {scB.x, scB.y, scB.z, scB.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = j >> 1, tt = j & 1, u = 4 * sb + j;
                const uint32_t laA = lds32(rowA + 16 * u + 4 * t4), lbA = lds32(rowA + S + 16 * u + 4 * t4), hqA = lds32(rowA + 2 * S + 16 * u + 4 * t4);
                const uint32_t laB = lds32(rowB + 16 * u + 4 * t4), lbB = lds32(rowB + S + 16 * u + 4 * t4), hqB = lds32(rowB + 2 * S + 16 * u + 4 * t4);
                const uint32_t cA[4] = {(laA & 0x0F0F0F0Fu) | ((hqA << 4) & 0x30303030u), (lbA & 0x0F0F0F0Fu) | ((hqA << 2) & 0x30303030u),
                                        ((laA >> 4) & 0x0F0F0F0Fu) | (hqA & 0x30303030u), ((lbA >> 4) & 0x0F0F0F0Fu) | ((hqA >> 2) & 0x30303030u)};
                const uint32_t cB[4] = {(laB & 0x0F0F0F0Fu) | ((hqB << 4) & 0x30303030u), (lbB & 0x0F0F0F0Fu) | ((hqB << 2) & 0x30303030u),
                                        ((laB >> 4) & 0x0F0F0F0Fu) | (hqB & 0x30303030u), ((lbB >> 4) & 0x0F0F0F0Fu) | ((hqB >> 2) & 0x30303030u)};
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    int isum[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int l = 8 * n + 2 * r + tt;                             /* chunk within the super-block; x = l >> 2 */
                        int cc[4] = {0, 0, 0, 0};
                        mma_u8s8_k16(cc, cA[r], cB[r], lds32(GBM_BADDR(nt, l, l >> 2)));
                        const int si = 8 * n + 2 * r + tt;                           /* scale byte index */
                        const int sA = (int)(int8_t)((swA[si >> 2] >> (8 * (si & 3))) & 0xFF), sB = (int)(int8_t)((swB[si >> 2] >> (8 * (si & 3))) & 0xFF);
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int b16 = 32 * (int)(int16_t)lds16(bq[nt][q] + 2 * l);
                            isum[q] += sA * (cc[q] - b16);
                            isum[2 + q] += sB * (cc[2 + q] - b16);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        acc[nt][q] += (double)__fmul_rn(__fmul_rn(dA, dx[nt][q]), (float)isum[q]);
                        acc[nt][2 + q] += (double)__fmul_rn(__fmul_rn(dB, dx[nt][q]), (float)isum[2 + q]);
                    }
                }
            }
        }
#undef GBM_BADDR
    }
}

// NT = 8-token n-tiles per launch (1: <= 8 tokens, 2: <= 16)
template <int NT>
__global__ void __launch_bounds__(GBM_THREADS, 1) ggb_dq_gemv_batch_mma_kernel(const __grid_constant__ GbmK P) {
    constexpr int NBT = 8 * NT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[GBM_MAX_STAGES], s_empty[GBM_MAX_STAGES], s_abar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.k, T = P.T, NS = P.nstage;
    const int G = gridDim.x, c = blockIdx.x;

    // this CTA's rows in every segment (even-aligned), cut into groups of 16
    GbmCtx X;
    X.ngroups = 0;
#pragma unroll
    for (int s = 0; s < GGB_MAX_SEG; s++) {
        X.r0[s] = X.cnt[s] = X.ng[s] = 0;
        if (s < P.n_seg) {
            const int q = P.rq[s], r = P.rr[s];
            const int a = (c * q + min(c, r)) & ~1;
            int b = (c + 1) * q + min(c + 1, r);
            if (c + 1 != G) b &= ~1;
            X.r0[s] = a; X.cnt[s] = b - a; X.ng[s] = (b - a + GBM_ROWS - 1) / GBM_ROWS;
            X.ngroups += X.ng[s];
        }
    }
    auto group = [&](int p, int& s, int& row, int& nv, int& lr) {
        int base = 0;
        s = 0;
        while (p >= X.ng[s]) { p -= X.ng[s]; base += X.cnt[s]; s++; }
        row = X.r0[s] + GBM_ROWS * p; nv = min(GBM_ROWS, X.cnt[s] - GBM_ROWS * p); lr = base + GBM_ROWS * p;
    };

    if (tid == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), GBM_NW); }
        mbar_init(smem_u32(&s_abar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ring = smem_u32(smem);
    const int nsb_last = ggb_tile_nsb(K, T - 1);

    if (warp == GBM_NW) {
        // ===== producer: walks (group, tile) and keeps the ring full; weights do not depend on the previous launch
        if (lane == 0) {
            int step = 0;
            for (int p = 0; p < X.ngroups; p++) {
                int s, row, nv, lr;
                group(p, s, row, nv, lr);
                const int sbb = ggb_sb_bytes(P.seg[s].type);
                const int tile = sbb * GGB_TILE_SB, last = (nsb_last * sbb + 15) & ~15;
                const uint8_t* src = P.seg[s].w + (int64_t)row * P.seg[s].stride;
                for (int t = 0; t < T; t++, step++) {
                    const int st = step % NS;
                    if (step >= NS) mbar_wait(smem_u32(&s_empty[st]), ((step / NS) - 1) & 1);
                    const uint32_t bytes = (t == T - 1) ? (uint32_t)last : (uint32_t)tile;
                    const uint32_t bar = smem_u32(&s_full[st]);
                    mbar_expect_tx(bar, (uint32_t)nv * bytes);
                    for (int r = 0; r < nv; r++)
                        bulk_g2s(ring + st * P.stage_bytes + r * P.slotp, src + (int64_t)r * P.seg[s].stride + (int64_t)t * tile, bytes, bar);
                }
            }
        }
        return;
    }

    // ===== consumers
    uint8_t* act = smem + P.act_off;
    double* red = reinterpret_cast<double*>(smem + P.red_off);     /* [2][GBM_NW][NBT][GBM_RS] */
    double* rowv = reinterpret_cast<double*>(smem + P.rowv_off);   /* SWIGLU only: [token][rowv_ld local rows] */
    // rows of absent tokens: zero images (dx = 0 -> every term 0)
    for (int i = tid * 16; i < (NBT - P.nb) * P.imgp; i += GBM_NW * 32 * 16) *reinterpret_cast<uint4*>(act + P.nb * P.imgp + i) = make_uint4(0, 0, 0, 0);
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t act_s = smem_u32(act);
    if (tid == 0) {
        mbar_expect_tx(smem_u32(&s_abar), (uint32_t)(P.nb * P.image));
        for (int b = 0; b < P.nb; b++) bulk_g2s(act_s + b * P.imgp, P.act + (int64_t)b * P.image, (uint32_t)P.image, smem_u32(&s_abar));
    }
    mbar_wait(smem_u32(&s_abar), 0);

    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t bs_off = (uint32_t)K, dsc_off = (uint32_t)(K + K / 8);
    int step = 0;
    for (int p = 0; p < X.ngroups; p++) {
        int s, row0, nv, lr;
        group(p, s, row0, nv, lr);
        const int type = P.seg[s].type;
        double acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0; }
        for (int t = 0; t < T; t++, step++) {
            const int st = step % NS;
            const int nsb = (t == T - 1) ? nsb_last : GGB_TILE_SB;
            mbar_wait(smem_u32(&s_full[st]), (step / NS) & 1);
            gbm_tile_step<NT>(type, ring + st * P.stage_bytes, P.slotp, nsb, warp, lane, act_s, P.imgp, bs_off, dsc_off, t * GGB_TILE_SB, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive1(smem_u32(&s_empty[st]));
        }
        // ---- the 8 K-slices of every (row, token) meet here
        double* rb = red + (size_t)(p & 1) * GBM_NW * GBM_RS * NBT;   /* [super-block][token][GBM_RS rows] */
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                rb[(warp * NBT + nt * 8 + 2 * t4 + q) * GBM_RS + g] = acc[nt][q];
                rb[(warp * NBT + nt * 8 + 2 * t4 + q) * GBM_RS + g + 8] = acc[nt][2 + q];
            }
        consumer_sync();
        for (int i = tid; i < GBM_ROWS * NBT; i += GBM_NW * 32) {
            const int rl = i & (GBM_ROWS - 1), b = i / GBM_ROWS;         /* consecutive threads: consecutive rows (banks, and the stores below) */
            if (rl < nv && b < P.nb) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < GBM_NW; w++) v += rb[(w * NBT + b) * GBM_RS + rl];
                const int row = row0 + rl;
                if (P.epi == GGB_EPI_STORE) P.seg[s].y[(int64_t)b * P.seg[s].rows + row] = (float)v;
                else if (P.epi == GGB_EPI_RESIDUAL) {
                    const int64_t o = (int64_t)b * P.seg[0].rows + row;
                    P.seg[0].y[o] = __fadd_rn(P.residual[o], (float)v);
                } else if (P.epi == GGB_EPI_STORE_F64) reinterpret_cast<double*>(P.seg[0].y)[(int64_t)b * P.seg[0].rows + row] = v;
                else rowv[b * P.rowv_ld + lr + rl] = v;
            }
        }
        /* the other half of `red` is used by the next group; the sync of the group after that orders its reuse */
    }
    if (P.epi == GGB_EPI_SWIGLU) {
        consumer_sync();
        const int cnt0 = X.cnt[0];
        for (int i = tid; i < cnt0 * P.nb; i += GBM_NW * 32) {
            const int b = i / cnt0, l = i - b * cnt0;
            P.seg[0].y[(int64_t)b * P.seg[0].rows + X.r0[0] + l] = silu_mul_ref((float)rowv[b * P.rowv_ld + l], (float)rowv[b * P.rowv_ld + cnt0 + l]);
        }
    }
}

// ------------------------------------------------------------------ tiled form: activation images streamed K-tile by K-tile
// Sixteen images of a long vector (ffn_down: K = 14336 -> 16 x 17.9 KB) do not fit shared memory next to the ring, and two
// 8-token passes stream the weights twice.  Here the CTA walks TILE-major -- for every K-tile, all of its (<= 4) row groups --
// so that only the K-tile slice of the images is needed at a time: ggb_act_prep_tiled leaves the images as
// [tile][token][GGB_ACT_TILE_STRIDE] (codes 2048 | per-16 sums 256 | block scales 32, padded so that the token stride is
// 4 words mod 32 banks), one bulk copy per tile into one of two slice buffers.  The partial sums of a (group, warp, row,
// token) live in shared memory between tiles (f64, each lane adds into its own words); what is summed is unchanged, so the
// result is bit-identical with the one- and two-pass forms.
#define GBM_TILED_MAX_GROUPS 4

template <int NT>
__global__ void __launch_bounds__(GBM_THREADS, 1) ggb_dq_gemv_batch_mma_tiled_kernel(const __grid_constant__ GbmK P) {
    constexpr int NBT = 8 * NT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[GBM_MAX_STAGES], s_empty[GBM_MAX_STAGES], s_afull[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.k, T = P.T, NS = P.nstage;
    const int G = gridDim.x, c = blockIdx.x;

    // one segment: this CTA's rows (even-aligned), cut into groups of 16
    const int q_ = P.rq[0], r_ = P.rr[0];
    const int ra = (c * q_ + min(c, r_)) & ~1;
    int rb_ = (c + 1) * q_ + min(c + 1, r_);
    if (c + 1 != G) rb_ &= ~1;
    const int cnt = rb_ - ra, ngroups = (cnt + GBM_ROWS - 1) / GBM_ROWS;
    const int type = P.seg[0].type;

    if (tid == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), GBM_NW); }
        mbar_init(smem_u32(&s_afull[0]), 1); mbar_init(smem_u32(&s_afull[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ring = smem_u32(smem);
    const int nsb_last = ggb_tile_nsb(K, T - 1);

    if (warp == GBM_NW) {
        // ===== producer: (tile, group) order; weights do not depend on the previous launch
        if (lane == 0) {
            const int sbb = ggb_sb_bytes(type);
            const int tile = sbb * GGB_TILE_SB, last = (nsb_last * sbb + 15) & ~15;
            int step = 0;
            for (int t = 0; t < T; t++) {
                const uint32_t bytes = (t == T - 1) ? (uint32_t)last : (uint32_t)tile;
                for (int p = 0; p < ngroups; p++, step++) {
                    const int nv = min(GBM_ROWS, cnt - GBM_ROWS * p);
                    const uint8_t* src = P.seg[0].w + (int64_t)(ra + GBM_ROWS * p) * P.seg[0].stride + (int64_t)t * tile;
                    const int st = step % NS;
                    if (step >= NS) mbar_wait(smem_u32(&s_empty[st]), ((step / NS) - 1) & 1);
                    const uint32_t bar = smem_u32(&s_full[st]);
                    mbar_expect_tx(bar, (uint32_t)nv * bytes);
                    for (int r = 0; r < nv; r++) bulk_g2s(ring + st * P.stage_bytes + r * P.slotp, src + (int64_t)r * P.seg[0].stride, bytes, bar);
                }
            }
        }
        return;
    }

    // ===== consumers
    uint8_t* abuf = smem + P.act_off;                               /* two slice buffers of NBT x imgp bytes */
    double* red = reinterpret_cast<double*>(smem + P.red_off);     /* [group][GBM_NW][NBT][GBM_RS] */
    const int abytes = NBT * P.imgp;
    // rows of absent tokens: zero slices in both buffers (dx = 0 -> every term 0; the bulk copies never touch them)
    for (int bsel = 0; bsel < 2; bsel++)
        for (int i = tid * 16; i < (NBT - P.nb) * P.imgp; i += GBM_NW * 32 * 16)
            *reinterpret_cast<uint4*>(abuf + bsel * abytes + P.nb * P.imgp + i) = make_uint4(0, 0, 0, 0);
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t abuf_s = smem_u32(abuf);
    const uint32_t slice = (uint32_t)(P.nb * P.imgp);              /* bytes of one tile's slice in global memory: [nb][imgp] */
    auto fetch_slice = [&](int t) {
        const uint32_t bar = smem_u32(&s_afull[t & 1]);
        mbar_expect_tx(bar, slice);
        bulk_g2s(abuf_s + (t & 1) * abytes, P.act + (int64_t)t * slice, slice, bar);
    };
    if (tid == 0) {
        fetch_slice(0);
        if (T > 1) fetch_slice(1);
    }

    const int g = lane >> 2, t4 = lane & 3;
    int step = 0;
    for (int t = 0; t < T; t++) {
        const int nsb = (t == T - 1) ? nsb_last : GGB_TILE_SB;
        const uint32_t act_s = abuf_s + (t & 1) * abytes;
        mbar_wait(smem_u32(&s_afull[t & 1]), (t >> 1) & 1);
        for (int p = 0; p < ngroups; p++, step++) {
            const int st = step % NS;
            double acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0; }
            mbar_wait(smem_u32(&s_full[st]), (step / NS) & 1);
            gbm_tile_step<NT>(type, ring + st * P.stage_bytes, P.slotp, nsb, warp, lane, act_s, P.imgp, GGB_ACT_TILE_BS_OFF, GGB_ACT_TILE_DSC_OFF, 0, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive1(smem_u32(&s_empty[st]));
            double* rb = red + (size_t)p * GBM_NW * GBM_RS * NBT;
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    double* a0 = rb + (warp * NBT + nt * 8 + 2 * t4 + q) * GBM_RS + g;
                    double* a1 = a0 + 8;
                    if (t == 0) { *a0 = acc[nt][q]; *a1 = acc[nt][2 + q]; }
                    else { *a0 += acc[nt][q]; *a1 += acc[nt][2 + q]; }
                }
        }
        consumer_sync();                       /* every consumer is done with slice buffer t & 1 (and, after the last tile, with red) */
        if (tid == 0 && t + 2 < T) fetch_slice(t + 2);
    }
    // ---- the 8 K-slices of every (row, token) meet here
    for (int p = 0; p < ngroups; p++) {
        const double* rb = red + (size_t)p * GBM_NW * GBM_RS * NBT;
        const int nv = min(GBM_ROWS, cnt - GBM_ROWS * p), row0 = ra + GBM_ROWS * p;
        for (int i = tid; i < GBM_ROWS * NBT; i += GBM_NW * 32) {
            const int rl = i & (GBM_ROWS - 1), b = i / GBM_ROWS;
            if (rl < nv && b < P.nb) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < GBM_NW; w++) v += rb[(w * NBT + b) * GBM_RS + rl];
                const int64_t o = (int64_t)b * P.seg[0].rows + row0 + rl;
                if (P.epi == GGB_EPI_STORE) P.seg[0].y[o] = (float)v;
                else if (P.epi == GGB_EPI_RESIDUAL) P.seg[0].y[o] = __fadd_rn(P.residual[o], (float)v);
                else reinterpret_cast<double*>(P.seg[0].y)[o] = v;
            }
        }
    }
}

// ggb_act_prep for the tiled form: token `blockIdx.x`, 256-blocks spread over gridDim.y CTAs; with RMSNorm every CTA first
// takes the whole row's sum of squares (16-57 KB from L2: cheaper than a second launch).  Same quantiser, same codes.
__global__ void __launch_bounds__(256) act_prep_tiled_kernel(const float* __restrict__ x, const float* __restrict__ norm_w, float eps, int K, int nb,
                                                             uint8_t* __restrict__ out) {
    __shared__ double red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xr = x + (int64_t)blockIdx.x * K;
    pdl_launch_dependents();
    pdl_wait();
    float scale = 1.f;
    if (norm_w) {
        double s = 0.0;
        for (int i = tid; i < K; i += 256) { const float v = xr[i]; s += (double)__fmul_rn(v, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) tot += red[i];
        const float mean = (float)(tot / (double)K);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(mean + eps));
    }
    const int nblk = K / 256;
    for (int b = blockIdx.y * 8 + warp; b < nblk; b += gridDim.y * 8) {
        const int e = b * 256 + lane * 8;
        const float4 xa = *reinterpret_cast<const float4*>(xr + e), xb = *reinterpret_cast<const float4*>(xr + e + 4);
        float v[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        if (norm_w) {
            const float4 ga = *reinterpret_cast<const float4*>(norm_w + e), gb = *reinterpret_cast<const float4*>(norm_w + e + 4);
            const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), gg[i]);
        }
        const int t = b / GGB_TILE_SB, sbr = b - t * GGB_TILE_SB;
        uint8_t* base = out + ((int64_t)t * nb + blockIdx.x) * GGB_ACT_TILE_STRIDE;
        const int chunk = (sbr * 256 + lane * 8) >> 4;                         /* chunk within the tile */
        float dd;
        const Q8Codes cq = warp_quantize_q8_K(v, lane, dd);
        *reinterpret_cast<uint2*>(base + 16 * swz(chunk) + 8 * (lane & 1)) = cq.q;
        const int s16 = cq.sum8 + __shfl_xor_sync(0xffffffffu, cq.sum8, 1);
        if (!(lane & 1)) *reinterpret_cast<int16_t*>(base + GGB_ACT_TILE_BS_OFF + 2 * chunk) = (int16_t)s16;
        if (lane == 0) *reinterpret_cast<float*>(base + GGB_ACT_TILE_DSC_OFF + 4 * sbr) = dd;
    }
}

extern "C" int64_t ggb_act_tiled_bytes(int64_t k, int nb) {
    return (k > 0 && k % 256 == 0 && nb >= 0) ? (int64_t)ggb_tiles_per_row((int)k) * nb * GGB_ACT_TILE_STRIDE : -1;
}

extern "C" int ggb_act_prep_tiled(const float* x, const float* norm_w, float eps, int64_t k, int nb, void* act, int use_pdl, void* stream) {
    if (k <= 0 || (k % 256) || nb < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_act_prep_tiled: k=%lld must be a positive multiple of 256, nb >= 0", (long long)k);
    if (nb == 0) return GGB_OK;
    if (!x || !act || ((uintptr_t)x & 15) || ((uintptr_t)act & 15) || (norm_w && ((uintptr_t)norm_w & 15)))
        GGB_FAIL(GGB_ERR_ARG, "ggb_act_prep_tiled: null or misaligned pointer");
    const int nblk = (int)(k / 256);
    int splits = (nblk + 7) / 8;
    if (norm_w && splits > 4) splits = 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nb, splits); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, act_prep_tiled_kernel, x, norm_w, eps, (int)k, nb, (uint8_t*)act));
    return GGB_OK;
}

// ------------------------------------------------------------------ host side (called from ggb_gemv_batch)
template <int NT>
static int gbm_launch_nt(const GbmK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(ggb_dq_gemv_batch_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBM_MAX_SMEM));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GBM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, ggb_dq_gemv_batch_mma_kernel<NT>, P));
    return GGB_OK;
}

template <int NT>
static int gbm_launch_tiled_nt(const GbmK& P, int grid, size_t smem, int use_pdl, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(ggb_dq_gemv_batch_mma_tiled_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GBM_MAX_SMEM));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GBM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, ggb_dq_gemv_batch_mma_tiled_kernel<NT>, P));
    return GGB_OK;
}

// whole-vector images: ring depth that fits next to `nbt` images (0 = does not fit)
static int gbm_plan_whole(const ggb_gemv_batch_args* a, int nbt, int stage_bytes, int imgp, int64_t max_local) {
    const size_t img = (size_t)nbt * imgp, red = (size_t)2 * GBM_NW * GBM_RS * nbt * 8;
    const size_t rowv = a->epilogue == GGB_EPI_SWIGLU ? (size_t)max_local * nbt * 8 : 0;
    const size_t fixed = ((img + 127) & ~(size_t)127) + red + rowv + 256;
    int nstage = fixed < GBM_MAX_SMEM ? (int)((GBM_MAX_SMEM - fixed) / stage_bytes) : 0;
    return nstage > GBM_MAX_STAGES ? GBM_MAX_STAGES : nstage;
}

// tiled images: one segment, no SWIGLU, at most GBM_TILED_MAX_GROUPS row groups per CTA; returns the ring depth (0 = no)
static int gbm_plan_tiled(const ggb_gemv_batch_args* a, int nbt, int stage_bytes, int grid, int* groups_out) {
    if (a->n_seg != 1 || a->epilogue == GGB_EPI_SWIGLU || a->k < GGB_TILE_ELEMS) return 0;
    const int groups = (a->seg[0].rows / grid + 2 + GBM_ROWS - 1) / GBM_ROWS;
    if (groups > GBM_TILED_MAX_GROUPS) return 0;
    const size_t fixed = (size_t)2 * nbt * GGB_ACT_TILE_STRIDE + (size_t)groups * GBM_NW * GBM_RS * nbt * 8 + 256;
    int nstage = fixed < GBM_MAX_SMEM ? (int)((GBM_MAX_SMEM - fixed) / stage_bytes) : 0;
    if (groups_out) *groups_out = groups;
    return nstage > GBM_MAX_STAGES ? GBM_MAX_STAGES : nstage;
}

static bool gbm_types_ok(const ggb_gemv_batch_args* a, int* max_tile) {
    *max_tile = 0;
    for (int s = 0; s < a->n_seg; s++) {
        const ggb_gemv_seg& g = a->seg[s];
        if (g.type != GGB_TYPE_Q4_K && g.type != GGB_TYPE_Q6_K) return false;
        const int tile = ggb_sb_bytes(g.type) * GGB_TILE_SB;
        if (tile > *max_tile) *max_tile = tile;
    }
    return a->k >= GGB_TILE_ELEMS && a->k % 256 == 0;
}

// 1 when a launch of this shape should be fed tiled images (ggb_act_prep_tiled + act_tiled = 1): more than 8 tokens whose
// whole-vector images do not fit shared memory (the launch would otherwise run as two 8-token passes over the weights)
extern "C" int ggb_gemv_batch_prefers_tiled(const ggb_gemv_batch_args* a) {
    const char* ev = getenv("GGB_BATCH_TILED");   /* read per call: the tests switch it */
    const int enabled = ev && *ev ? atoi(ev) : 1;
    static const int use_mma = []() { const char* v = getenv("GGB_BATCH_MMA"); return v && *v ? atoi(v) : 1; }();
    if (!a || !enabled || !use_mma || a->nb <= 8 || a->nb > 16 || a->n_seg < 1 || a->n_seg > GGB_MAX_SEG) return 0;
    int max_tile;
    if (!gbm_types_ok(a, &max_tile)) return 0;
    const int grid = a->grid > 0 ? a->grid : ggb_num_sms();
    const int stage_bytes = GBM_ROWS * (max_tile + (16 - max_tile % 128 + 128) % 128);
    const int image = a->k + a->k / 4, imgp = image + 4 * ((4 - (image / 4) % 32 + 32) % 32);
    int64_t max_local = 0;
    for (int s = 0; s < a->n_seg; s++) max_local += (a->seg[s].rows + grid - 1) / grid + 4;
    if (enabled != 2 && gbm_plan_whole(a, 16, stage_bytes, imgp, max_local) >= 2) return 0;   /* GGB_BATCH_TILED=2: whenever possible (tests) */
    return gbm_plan_tiled(a, 16, stage_bytes, grid, nullptr) >= 2 ? 1 : 0;
}

static int gbm_launch_tiled(const ggb_gemv_batch_args* a, GbmK P, int grid, void* stream) {
    if (a->nb > 16) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: tiled activation images hold at most 16 tokens (nb=%d)", a->nb);
    const int nbt = a->nb > 8 ? 16 : 8;
    int groups = 0;
    const int nstage = gbm_plan_tiled(a, nbt, P.stage_bytes, grid, &groups);
    if (nstage < 2) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: shape k=%d rows=%d does not take tiled activation images", a->k, a->seg[0].rows);
    P.nb = a->nb; P.act = (const uint8_t*)a->act; P.nstage = nstage;
    P.image = P.imgp = GGB_ACT_TILE_STRIDE;
    P.act_off = (nstage * P.stage_bytes + 127) & ~127;
    P.red_off = P.act_off + 2 * nbt * GGB_ACT_TILE_STRIDE;
    P.rowv_off = 0;
    const size_t smem = (size_t)P.red_off + (size_t)groups * GBM_NW * GBM_RS * nbt * 8;
    if (smem > GBM_MAX_SMEM) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemv_batch: tiled plan needs %zu bytes of shared memory", smem);
    cudaStream_t st = (cudaStream_t)stream;
    return nbt == 16 ? gbm_launch_tiled_nt<2>(P, grid, smem, a->use_pdl, st) : gbm_launch_tiled_nt<1>(P, grid, smem, a->use_pdl, st);
}

// returns GGB_OK after launching, or 1 when the shape does not fit (the caller then uses the dp4a kernel)
int ggb_gemv_batch_mma(const ggb_gemv_batch_args* a, void* stream) {
    GbmK P = {};
    int max_tile = 0;
    for (int s = 0; s < a->n_seg; s++) {
        const ggb_gemv_seg& g = a->seg[s];
        if (g.type != GGB_TYPE_Q4_K && g.type != GGB_TYPE_Q6_K) return 1;
        P.seg[s].w = (const uint8_t*)g.w; P.seg[s].y = g.y; P.seg[s].stride = ggb_row_stride(g.type, a->k);
        P.seg[s].type = g.type; P.seg[s].rows = g.rows;
        const int tile = ggb_sb_bytes(g.type) * GGB_TILE_SB;
        if (tile > max_tile) max_tile = tile;
    }
    if (a->k < GGB_TILE_ELEMS) return 1;          /* tiny K: not worth a second code path */
    const int grid = a->grid > 0 ? a->grid : ggb_num_sms();
    int64_t max_local = 0;
    for (int s = 0; s < a->n_seg; s++) {
        P.rq[s] = a->seg[s].rows / grid; P.rr[s] = a->seg[s].rows % grid;
        max_local += (a->seg[s].rows + grid - 1) / grid + 4;
    }
    P.n_seg = a->n_seg; P.k = a->k; P.T = ggb_tiles_per_row(a->k); P.epi = a->epilogue; P.residual = a->residual;
    P.slotp = max_tile + (16 - max_tile % 128 + 128) % 128;   /* slot stride == 16 B mod 128: rows g = 0..7 of a fragment fall into distinct banks (Q4_K: + 16, Q6_K: + 0) */
    P.stage_bytes = GBM_ROWS * P.slotp;
    if (a->act_tiled) return gbm_launch_tiled(a, P, grid, stream);
    P.image = a->k + a->k / 4;
    P.imgp = P.image + 4 * ((4 - (P.image / 4) % 32 + 32) % 32);   /* token stride == 4 words (mod 32 banks) */
    cudaStream_t st = (cudaStream_t)stream;
    for (int b0 = 0; b0 < a->nb;) {
        int nbt = (a->nb - b0 > 8) ? 16 : 8;
        size_t smem = 0;
        int nstage = 0;
        for (;;) {   /* shared memory: [ring][images][red x2][rowv (SWIGLU)] */
            const size_t img = (size_t)nbt * P.imgp, red = (size_t)2 * GBM_NW * GBM_RS * nbt * 8;
            const size_t rowv = a->epilogue == GGB_EPI_SWIGLU ? (size_t)max_local * nbt * 8 : 0;
            const size_t fixed = ((img + 127) & ~(size_t)127) + red + rowv + 256;
            nstage = fixed < GBM_MAX_SMEM ? (int)((GBM_MAX_SMEM - fixed) / P.stage_bytes) : 0;
            if (nstage > GBM_MAX_STAGES) nstage = GBM_MAX_STAGES;
            if (nstage >= 2) {
                P.nstage = nstage;
                P.act_off = nstage * P.stage_bytes;
                P.act_off = (P.act_off + 127) & ~127;
                P.red_off = P.act_off + (int)((img + 127) & ~(size_t)127);
                P.rowv_off = P.red_off + (int)red;
                P.rowv_ld = (int)max_local;
                smem = (size_t)P.rowv_off + rowv;
                break;
            }
            if (nbt == 16) { nbt = 8; continue; }
            return b0 == 0 ? 1 : GGB_ERR_UNSUPPORTED;
        }
        const int nb = a->nb - b0 < nbt ? a->nb - b0 : nbt;
        GbmK Q = P;
        Q.nb = nb;
        Q.act = (const uint8_t*)a->act + (int64_t)b0 * P.image;
        for (int s = 0; s < a->n_seg; s++) if (Q.seg[s].y) Q.seg[s].y += (int64_t)b0 * Q.seg[s].rows * (a->epilogue == GGB_EPI_STORE_F64 ? 2 : 1);
        if (Q.residual) Q.residual += (int64_t)b0 * Q.seg[0].rows;
        const int rc = (nbt == 16) ? gbm_launch_nt<2>(Q, grid, smem, a->use_pdl, st) : gbm_launch_nt<1>(Q, grid, smem, a->use_pdl, st);
        if (rc) return rc;
        b0 += nb;
    }
    return GGB_OK;
}
