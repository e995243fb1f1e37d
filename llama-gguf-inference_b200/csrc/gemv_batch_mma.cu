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

#include "common.cuh"
#include "layout.cuh"
#include "gemv_common.cuh"

#define GBM_NW 8                      /* consumer warps */
#define GBM_THREADS ((GBM_NW + 1) * 32)
#define GBM_ROWS 16
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
    int act_off, red_off, rowv_off;
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
    double* red = reinterpret_cast<double*>(smem + P.red_off);     /* [2][GBM_NW][16][NBT] */
    double* rowv = reinterpret_cast<double*>(smem + P.rowv_off);   /* SWIGLU only: [local row][NBT] */
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
            const uint32_t S = 64u * (uint32_t)nsb;                     /* bytes of a 16-byte-per-unit section */
            mbar_wait(smem_u32(&s_full[st]), (step / NS) & 1);
            if (warp < nsb) {
                const int sb = warp;
                const uint32_t rowA = ring + st * P.stage_bytes + g * P.slotp, rowB = rowA + 8 * P.slotp;
                const int gsb = t * GGB_TILE_SB + sb;                   /* super-block index along K */
                const int chunk_sb = gsb * 16;                          /* first 16-element chunk of the super-block */
                // block scales of my C-fragment tokens (2*t4, 2*t4+1 of every n-tile)
                float dx[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    dx[nt][0] = __uint_as_float(lds32(act_s + (nt * 8 + 2 * t4) * P.imgp + dsc_off + 4 * gsb));
                    dx[nt][1] = __uint_as_float(lds32(act_s + (nt * 8 + 2 * t4 + 1) * P.imgp + dsc_off + 4 * gsb));
                }
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
                        const int chunk0 = chunk_sb + 4 * j;
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            const uint32_t tb = act_s + (nt * 8 + g) * P.imgp + 4 * t4;      /* B fragment: token nt*8 + g */
                            int clo[4] = {0, 0, 0, 0}, chi[4] = {0, 0, 0, 0};
                            mma_u8s8_k32(clo, wA0 & 0x0F0F0F0Fu, wB0 & 0x0F0F0F0Fu, wA1 & 0x0F0F0F0Fu, wB1 & 0x0F0F0F0Fu,
                                         lds32(tb + 16 * swz(chunk0)), lds32(tb + 16 * swz(chunk0 + 1)));
                            mma_u8s8_k32(chi, (wA0 >> 4) & 0x0F0F0F0Fu, (wB0 >> 4) & 0x0F0F0F0Fu, (wA1 >> 4) & 0x0F0F0F0Fu, (wB1 >> 4) & 0x0F0F0F0Fu,
                                         lds32(tb + 16 * swz(chunk0 + 2)), lds32(tb + 16 * swz(chunk0 + 3)));
#pragma unroll
                            for (int q = 0; q < 2; q++) {                                     /* C-fragment tokens 2*t4 + q */
                                const uint2 bs = lds64(act_s + (nt * 8 + 2 * t4 + q) * P.imgp + bs_off + 2 * chunk0);   /* four per-16 sums */
                                const int blo = (int)(int16_t)(bs.x & 0xFFFF) + ((int)bs.x >> 16);
                                const int bhi = (int)(int16_t)(bs.y & 0xFFFF) + ((int)bs.y >> 16);
                                const float x = dx[nt][q];
                                {
                                    const int isum = (int)(fA & 63) * clo[q] + (int)((fA >> 6) & 63) * chi[q];
                                    const int msum = (int)((fA >> 12) & 63) * blo + (int)((fA >> 18) & 63) * bhi;
                                    acc[nt][q] += (double)__fsub_rn(__fmul_rn(__fmul_rn(dA, x), (float)isum), __fmul_rn(__fmul_rn(mA, x), (float)msum));
                                }
                                {
                                    const int isum = (int)(fB & 63) * clo[2 + q] + (int)((fB >> 6) & 63) * chi[2 + q];
                                    const int msum = (int)((fB >> 12) & 63) * blo + (int)((fB >> 18) & 63) * bhi;
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
                            const uint32_t tb = act_s + (nt * 8 + g) * P.imgp + 4 * t4;
                            int isum[4] = {0, 0, 0, 0};
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                const int chunk = chunk_sb + 8 * n + 2 * r + tt;
                                int cc[4] = {0, 0, 0, 0};
                                mma_u8s8_k16(cc, cA[r], cB[r], lds32(tb + 16 * swz(chunk)));
                                const int si = 8 * n + 2 * r + tt;                           /* scale byte index */
                                const int sA = (int)(int8_t)((swA[si >> 2] >> (8 * (si & 3))) & 0xFF), sB = (int)(int8_t)((swB[si >> 2] >> (8 * (si & 3))) & 0xFF);
#pragma unroll
                                for (int q = 0; q < 2; q++) {
                                    const int b16 = 32 * (int)(int16_t)lds16(act_s + (nt * 8 + 2 * t4 + q) * P.imgp + bs_off + 2 * chunk);
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive1(smem_u32(&s_empty[st]));
        }
        // ---- the 8 K-slices of every (row, token) meet here
        double* rb = red + (size_t)(p & 1) * GBM_NW * GBM_ROWS * NBT;
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                rb[(warp * GBM_ROWS + g) * NBT + nt * 8 + 2 * t4 + q] = acc[nt][q];
                rb[(warp * GBM_ROWS + g + 8) * NBT + nt * 8 + 2 * t4 + q] = acc[nt][2 + q];
            }
        consumer_sync();
        for (int i = tid; i < GBM_ROWS * NBT; i += GBM_NW * 32) {
            const int rl = i / NBT, b = i - rl * NBT;
            if (rl < nv && b < P.nb) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < GBM_NW; w++) v += rb[(w * GBM_ROWS + rl) * NBT + b];
                const int row = row0 + rl;
                if (P.epi == GGB_EPI_STORE) P.seg[s].y[(int64_t)b * P.seg[s].rows + row] = (float)v;
                else if (P.epi == GGB_EPI_RESIDUAL) {
                    const int64_t o = (int64_t)b * P.seg[0].rows + row;
                    P.seg[0].y[o] = __fadd_rn(P.residual[o], (float)v);
                } else if (P.epi == GGB_EPI_STORE_F64) reinterpret_cast<double*>(P.seg[0].y)[(int64_t)b * P.seg[0].rows + row] = v;
                else rowv[(lr + rl) * NBT + b] = v;
            }
        }
        /* the other half of `red` is used by the next group; the sync of the group after that orders its reuse */
    }
    if (P.epi == GGB_EPI_SWIGLU) {
        consumer_sync();
        const int cnt0 = X.cnt[0];
        for (int i = tid; i < cnt0 * P.nb; i += GBM_NW * 32) {
            const int b = i / cnt0, l = i - b * cnt0;
            P.seg[0].y[(int64_t)b * P.seg[0].rows + X.r0[0] + l] = silu_mul_ref((float)rowv[l * NBT + b], (float)rowv[(cnt0 + l) * NBT + b]);
        }
    }
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
    P.slotp = max_tile + 16;                        /* +16 B: rows g = 0..7 of a fragment fall into distinct banks */
    P.stage_bytes = GBM_ROWS * P.slotp;
    P.image = a->k + a->k / 4;
    P.imgp = P.image + 4 * ((4 - (P.image / 4) % 32 + 32) % 32);   /* token stride == 4 words (mod 32 banks) */
    cudaStream_t st = (cudaStream_t)stream;
    for (int b0 = 0; b0 < a->nb;) {
        int nbt = (a->nb - b0 > 8) ? 16 : 8;
        size_t smem = 0;
        int nstage = 0;
        for (;;) {   /* shared memory: [ring][images][red x2][rowv (SWIGLU)] */
            const size_t img = (size_t)nbt * P.imgp, red = (size_t)2 * GBM_NW * GBM_ROWS * nbt * 8;
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
