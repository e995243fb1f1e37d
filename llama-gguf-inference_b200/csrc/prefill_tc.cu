// prefill_tc.cu -- causal prefill attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
// Stands in for the batched flash_attn_ext of the reference's backend [UPSTREAM-MEM]; same contract as the mma.sync kernel in
// prefill.cu (ggb_attn_prefill: T query tokens at positions pos0.. over the f16 cache of one slot and layer), which stays as
// the fallback (GGB_ATTN_PREFILL_TC=0).
//
// One CTA = one query head x 128 query tokens: 8 softmax warps -- two threads per query row (= TMEM lane), each owning half of
// the row's 128 positions and half of the output dims; the only cross-thread step of the softmax is one exchange of the
// halves' maxima per tile -- plus one warp whose elected lane issues every TMA copy and every MMA, so that
// S of the NEXT tile is computed (into the second of two S buffers in TMEM) while the softmax of the current one runs.
// Per tile of 128 cache positions:
//   K tile   two cp.async.bulk.tensor.2d (TMA, SWIZZLE_128B) from a tensor map over the cache -> K-major UMMA layout;
//   S = Q.K^T   tcgen05.mma kind::f16, M128 N128 K16 x (HD/16), A = Q (f16, written once, K-major SW128), D = TMEM columns [0,128);
//   softmax  every thread reads ITS row of S with tcgen05.ld (32x32b.x32), scales, masks (causal), keeps a running maximum
//            and sum, writes P = exp(s - m) as f16 into shared memory in the K-major SW128 layout (its own 256-byte row);
//   O_j = P.V   tcgen05.mma, A = P, B = V tile in the MN-major SW128 layout (exactly what TMA delivers for a [positions][dims]
//            box: no transposition anywhere), D = TMEM columns [128,128+HD) -- a fresh product per tile;
//   acc = acc * exp(m_old - m_new) + O_j   in registers (HD floats per thread), so TMEM is never rescaled in place.
// The next K tile is requested as soon as S has been computed, the next V tile as soon as P.V has: the copies overlap the
// softmax and the other product.  Tolerance-level numerics (f16 operands, f32 accumulation, __expf), as the GEMM feeding it.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

#define PT_BM 128          /* query tokens per CTA (UMMA M) */
#define PT_BN 128          /* cache positions per tile */
#define PT_SOFTMAX_WARPS 8  /* two threads per query row: warp w owns TMEM lanes 32 (w & 3) .., column half w >> 2 */
#define PT_THREADS ((PT_SOFTMAX_WARPS + 1) * 32)   /* + the warp that issues the TMA copies and the MMAs */

__device__ __forceinline__ uint32_t pt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PT_DONE;\n\t"
        "bra PT_WAIT;\n\t"
        "PT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1
// [46,48), SWIZZLE_128B = 2 at [61,64).
//   K-major  (Q, K, P): rows of 128 B (64 f16 of K), 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major (V): rows of 128 B (64 f16 of N) = one position each, 8-position groups 1024 B apart (SBO), the second 64-wide
//            half of N `lbo` bytes further on (LBO)
__device__ __forceinline__ uint64_t pt_desc_k(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t pt_desc_mn(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = f16, b_major (bit 16) = 1 for an MN-major B, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t pt_idesc(int n, bool b_mn) {
    return (1u << 4) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(PT_BM >> 4) << 24);
}
__device__ __forceinline__ void pt_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void pt_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pt_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pt_pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a [rows][128 B] SWIZZLE_128B atom
__device__ __forceinline__ uint32_t pt_sw(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void pt_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ float pt_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(PT_THREADS, 1) ggb_attn_prefill_tc_kernel(const float* __restrict__ q, const __grid_constant__ CUtensorMap tmapK,
                                                                              const __grid_constant__ CUtensorMap tmapV, int T, int pos0,
                                                                              int n_head, int n_kv, float* __restrict__ out) {
    constexpr int NA = HD / 64;                       /* 64-wide swizzle atoms along the head dimension */
    constexpr int ATOM = PT_BM * 128;                 /* bytes of one [128 rows][128 B] atom */
    constexpr uint32_t KV_BYTES = (uint32_t)(PT_BN * HD * 2);
    extern __shared__ __align__(1024) uint8_t pt_sm_raw[];
    __shared__ __align__(8) uint64_t bar_k, bar_v, bar_s[2], bar_p, bar_o;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_half[2][2][PT_BM];          /* [tile parity][column half][row]: the two threads of a row exchange maxima / sums */
    uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)pt_sm_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = sm;                                 /* NA atoms: [128 tokens][64 dims] f16, K-major */
    uint8_t* sK = sQ + NA * ATOM;                     /* NA atoms: [128 positions][64 dims], K-major (N = positions) */
    uint8_t* sV = sK + NA * ATOM;                     /* NA atoms: [128 positions][64 dims], MN-major (N = dims, K = positions) */
    uint8_t* sP = sV + NA * ATOM;                     /* 2 atoms:  [128 tokens][64 positions] f16, K-major */

    const int tid = threadIdx.x, warp = tid >> 5;
    const int head = blockIdx.x, qb = (int)gridDim.y - 1 - (int)blockIdx.y;      /* the longest tiles (last query tokens) start first */
    const int kvh = head / (n_head / n_kv);
    const int q0 = qb * PT_BM;
    const int64_t qd = (int64_t)n_head * HD;
    const int n_pos = pos0 + min(T, q0 + PT_BM);      /* positions this CTA needs */
    const int n_tiles = (n_pos + PT_BN - 1) / PT_BN;

    if (tid == 0) {
        pt_mbar_init(pt_smem_u32(&bar_k), 1); pt_mbar_init(pt_smem_u32(&bar_v), 1);
        pt_mbar_init(pt_smem_u32(&bar_s[0]), 1); pt_mbar_init(pt_smem_u32(&bar_s[1]), 1);
        pt_mbar_init(pt_smem_u32(&bar_p), PT_SOFTMAX_WARPS * 32); pt_mbar_init(pt_smem_u32(&bar_o), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PT_SOFTMAX_WARPS) {                   /* the issuing warp owns the TMEM allocation: S0 | S1 | O_j */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pt_smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Q tile -> f16, K-major SW128 (rows beyond T are zero)
    for (int i = tid; i < PT_BM * (HD / 8); i += PT_THREADS) {
        const int r = i / (HD / 8), c = i % (HD / 8);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (q0 + r < T) {
            const float* src = q + (int64_t)(q0 + r) * qd + (int64_t)head * HD + c * 8;
            a = *reinterpret_cast<const float4*>(src);
            b = *reinterpret_cast<const float4*>(src + 4);
        }
        *reinterpret_cast<uint4*>(sQ + (c >> 3) * ATOM + pt_sw(r, c & 7)) =
            make_uint4(pt_pack_h2(a.x, a.y), pt_pack_h2(a.z, a.w), pt_pack_h2(b.x, b.y), pt_pack_h2(b.z, b.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s, tmem_o = tmem_base + 256;

    if (warp == PT_SOFTMAX_WARPS) {
        // ===== issuing warp (one elected lane): TMA copies and both products of every tile =====
        if ((tid & 31) == 0) {
            auto load_tile = [&](const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int p0) {
                const uint32_t b = pt_smem_u32(bar);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(KV_BYTES) : "memory");
#pragma unroll
                for (int h = 0; h < NA; h++)   /* rows beyond the cache's valid positions are zero-filled (and masked by the softmax) */
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(pt_smem_u32(dst + h * ATOM)), "l"(tm), "r"(kvh * HD + h * 64), "r"(p0), "r"(b) : "memory");
            };
            const uint32_t id_s = pt_idesc(PT_BN, false), id_o = pt_idesc(HD, true);
            auto issue_s = [&](int j) {                /* S_j = Q.K_j^T into S buffer j & 1 */
                pt_mbar_wait(pt_smem_u32(&bar_k), (uint32_t)(j & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < HD / 16; k++) {
                    const uint32_t off = (uint32_t)((k >> 2) * ATOM + (k & 3) * 32);      /* atom, then 32 B per K = 16 step */
                    pt_mma(tmem_base + (uint32_t)((j & 1) * 128), pt_desc_k(pt_smem_u32(sQ) + off), pt_desc_k(pt_smem_u32(sK) + off), id_s, k > 0 ? 1u : 0u);
                }
                pt_commit(pt_smem_u32(&bar_s[j & 1]));
            };
            load_tile(&tmapK, sK, &bar_k, 0);
            load_tile(&tmapV, sV, &bar_v, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; j++) {
                // K_j has been read once S_j is complete: fetch K_{j+1} and compute S_{j+1} while the softmax of tile j runs
                pt_mbar_wait(pt_smem_u32(&bar_s[j & 1]), (uint32_t)((j >> 1) & 1));
                if (j + 1 < n_tiles) {
                    load_tile(&tmapK, sK, &bar_k, (j + 1) * PT_BN);
                    issue_s(j + 1);    /* its S buffer was last read by the softmax of tile j-1, which bar_p(j-1) below has seen finished */
                }
                // O_j = P_j.V_j once every row of P_j is written (which also means every thread has finished with O_{j-1})
                pt_mbar_wait(pt_smem_u32(&bar_p), (uint32_t)(j & 1));
                pt_mbar_wait(pt_smem_u32(&bar_v), (uint32_t)(j & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < PT_BN / 16; k++) {
                    const uint64_t da = pt_desc_k(pt_smem_u32(sP) + (uint32_t)((k >> 2) * ATOM + (k & 3) * 32));
                    const uint64_t db = pt_desc_mn(pt_smem_u32(sV) + (uint32_t)(k * 16 * 128), (uint32_t)ATOM);   /* 16 positions = 2048 B per step */
                    pt_mma(tmem_o, da, db, id_o, k > 0 ? 1u : 0u);
                }
                pt_commit(pt_smem_u32(&bar_o));
                if (j + 1 < n_tiles) {
                    pt_mbar_wait(pt_smem_u32(&bar_o), (uint32_t)(j & 1));                 /* V_j has been read */
                    load_tile(&tmapV, sV, &bar_v, (j + 1) * PT_BN);
                }
            }
        }
    } else {
        // ===== softmax warps: threads (row, half) own 64 of the 128 positions of a row of S and HD / 2 of the dims of O =====
        constexpr int HC = PT_BN / 2, HDH = HD / 2;
        const int row = tid & (PT_BM - 1), half = tid >> 7;
        const float sl2 = __fdiv_rn(1.0f, __fsqrt_rn((float)HD)) * 1.4426950408889634f;   /* scale * log2(e): exponentials as ex2 */
        const int lim = pos0 + q0 + row;              /* the last position query row `row` may attend */
        float acc[HDH];
#pragma unroll
        for (int d = 0; d < HDH; d++) acc[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;          /* running maximum (scaled log2 domain, whole row), running sum (my half) */
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        uint8_t* prow = sP + half * ATOM + (uint32_t)row * 128u;   /* my 64 positions = one 128-byte row of atom `half` */
        const uint32_t x7 = (uint32_t)(row & 7) << 4;
        auto pair_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(PT_SOFTMAX_WARPS * 32) : "memory"); };

        for (int j = 0; j < n_tiles; j++) {
            const int p0 = j * PT_BN + half * HC;     /* first position of my columns */
            const uint32_t tmem_s = tmem_base + (uint32_t)((j & 1) * 128 + half * HC) + lane_base;
            const bool diag = j * PT_BN + PT_BN - 1 > pos0 + q0;   /* the tile reaches past the first row's limit: mask (CTA-uniform) */
            pt_mbar_wait(pt_smem_u32(&bar_s[j & 1]), (uint32_t)((j >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // pass 1: maximum of the raw scores of my half (scale > 0 commutes with max), exchanged with the other half
            float mraw = -INFINITY;
#pragma unroll 1
            for (int cb = 0; cb < HC / 32; cb++) {
                uint32_t v[32];
                pt_ld32(tmem_s + (uint32_t)(cb * 32), v);
                if (!diag) {
#pragma unroll
                    for (int i = 0; i < 32; i++) mraw = fmaxf(mraw, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++) mraw = fmaxf(mraw, (p0 + cb * 32 + i <= lim) ? __uint_as_float(v[i]) : -INFINITY);
                }
            }
            s_half[j & 1][half][row] = mraw;
            pair_sync();
            mraw = fmaxf(mraw, s_half[j & 1][half ^ 1][row]);
            const float m_new = fmaxf(m_run, mraw * sl2);   /* position 0 is visible to every query: finite from the first tile on */
            const float alpha = pt_ex2(m_run - m_new);
            // pass 2: P = 2^(s * sl2 - m) as f16 into my 128-byte row of the K-major tile
            float l_add = 0.f;
#pragma unroll 1
            for (int cb = 0; cb < HC / 32; cb++) {
                uint32_t v[32];
                pt_ld32(tmem_s + (uint32_t)(cb * 32), v);
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    float e0 = pt_ex2(fmaf(__uint_as_float(v[2 * i]), sl2, -m_new)), e1 = pt_ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m_new));
                    if (diag) {
                        const int c0 = p0 + cb * 32 + 2 * i;
                        e0 = (c0 <= lim) ? e0 : 0.f;
                        e1 = (c0 + 1 <= lim) ? e1 : 0.f;
                    }
                    l_add += e0 + e1;
                    pk[i] = pt_pack_h2(e0, e1);
                }
#pragma unroll
                for (int c = 0; c < 4; c++)       /* 32 positions = chunks 4 cb .. 4 cb + 3 of my row */
                    *reinterpret_cast<uint4*>(prow + ((((uint32_t)(4 * cb + c)) << 4) ^ x7)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            l_run = fmaf(l_run, alpha, l_add);
            m_run = m_new;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     /* P: generic-proxy writes -> visible to the tensor core */
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            pt_mbar_arrive(pt_smem_u32(&bar_p));
            // acc = acc * alpha + O_j (my half of the dims)
            pt_mbar_wait(pt_smem_u32(&bar_o), (uint32_t)(j & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int cb = 0; cb < HDH / 32; cb++) {
                uint32_t v[32];
                pt_ld32(tmem_o + lane_base + (uint32_t)(half * HDH + cb * 32), v);
#pragma unroll
                for (int i = 0; i < 32; i++) acc[cb * 32 + i] = fmaf(acc[cb * 32 + i], alpha, __uint_as_float(v[i]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        // the row's sum = both halves' sums (same running maximum on both sides)
        s_half[n_tiles & 1][half][row] = l_run;
        pair_sync();
        const float l_row = l_run + s_half[n_tiles & 1][half ^ 1][row];
        const int tq = q0 + row;
        if (tq < T) {
            const float inv = __fdiv_rn(1.0f, l_row);
            float* dst = out + (int64_t)tq * qd + (int64_t)head * HD + half * HDH;
#pragma unroll
            for (int d = 0; d < HDH; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
        }
    }
    __syncthreads();
    if (warp == PT_SOFTMAX_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(512) : "memory");
}

// ------------------------------------------------------------------ host side
typedef CUresult (*pt_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pt_encode_tiled_fn pt_encode_tiled() {
    static pt_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (pt_encode_tiled_fn)p;
    }
    return fn;
}

template <int HD>
static int launch_tc(const float* q, const uint16_t* kc, const uint16_t* vc, int tokens, int pos0, int n_head, int n_kv, float* out, cudaStream_t st) {
    pt_encode_tiled_fn enc = pt_encode_tiled();
    if (!enc) return 1;
    // tensor maps over the VALID rows of this slot's cache, [pos0 + tokens][n_kv * HD] f16: box = 64 dims (one 128-byte swizzle
    // row) x 128 positions; the shared-memory image is the UMMA SW128 layout (K-major for K, MN-major for V)
    CUtensorMap tk, tv;
    const cuuint64_t dims[2] = {(cuuint64_t)n_kv * HD, (cuuint64_t)(pos0 + tokens)}, strides[1] = {(cuuint64_t)n_kv * HD * 2};
    const cuuint32_t box[2] = {64, PT_BN}, estr[2] = {1, 1};
    for (int i = 0; i < 2; i++) {
        const CUresult cr = enc(i ? &tv : &tk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(i ? vc : kc), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) GGB_FAIL(GGB_ERR_CUDA, "ggb_attn_prefill: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    }
    const size_t smem = (size_t)(3 * (HD / 64) + 2) * PT_BM * 128 + 1024;
    static bool attr = false;
    if (!attr) {
        GGB_CUDA(cudaFuncSetAttribute(ggb_attn_prefill_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(n_head, (tokens + PT_BM - 1) / PT_BM);
    ggb_attn_prefill_tc_kernel<HD><<<grid, PT_THREADS, smem, st>>>(q, tk, tv, tokens, pos0, n_head, n_kv, out);
    GGB_CHECK_LAUNCH("ggb_attn_prefill (tcgen05)");
    return GGB_OK;
}

// returns GGB_OK after launching, 1 when this path does not apply (the caller uses the mma.sync kernel), < 0 on error
int ggb_attn_prefill_tc(const float* q, const uint16_t* kcache, const uint16_t* vcache, int tokens, int pos0, int n_head, int n_kv,
                        int head_dim, float* out, void* stream) {
    const char* env = getenv("GGB_ATTN_PREFILL_TC");      /* read per call: the tests run both kernels in one process */
    if (env && *env && atoi(env) == 0) return 1;
    if (((uintptr_t)kcache & 15) || ((uintptr_t)vcache & 15) || ((uintptr_t)q & 15) || ((n_kv * head_dim * 2) & 15)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return launch_tc<128>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out, st);
    if (head_dim == 64) return launch_tc<64>(q, kcache, vcache, tokens, pos0, n_head, n_kv, out, st);
    return 1;
}
