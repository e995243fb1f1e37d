// gemm.cu -- K2: dequant-GEMM for prefill / large batches on the 5th-generation tensor cores.
//   Y[tokens][rows] = X[tokens][K] . W[rows][K]^T      W quantised (tile-SoA Q4_K / Q6_K / Q8_0), X f16
// Stands in for ggml-cuda's mul_mat_q / dequantise+cuBLAS path of the reference's backend [UPSTREAM-MEM].
//
// One CTA computes a 128 (weight rows) x 512 (tokens) output tile as TWO accumulators of UMMA N = 256: every dequantised
// weight tile feeds two MMAs (round 2; with one N = 256 accumulator the producers, not the tensor pipe, set the pace):
//   * 8 producer warps unpack the packed weights of a 64-wide K block straight into shared memory as f16 in the K-major
//     SWIZZLE_128B layout the tensor core reads (no TMA path exists for K-quants: the "B operand must be produced by a
//     dequant stage", SURVEY.md section 7 hard part 5).  Warps 0-3 own the even K blocks and stage 0, warps 4-7 the odd
//     ones and stage 1: the two halves run one block apart, which is the whole software pipeline.  The matching f16
//     activation block (512 tokens x 64 K) arrives next to it by TMA -- two cp.async.bulk.tensor.2d (SASS UTMALDG) from a
//     SWIZZLE_128B tensor map over X, issued the moment the stage is free, completing on the stage's mbarrier; weight
//     tiles are prefetched into L2 one tile ahead; fence.proxy.async + mbarrier hand the stage to the MMA warp;
//   * 1 MMA warp: a single elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) four K steps x
//     two token halves per block from shared-memory descriptors; the f32 accumulators (128 lanes x 2 x 256 columns = all
//     of TMEM) ; tcgen05.commit releases the stage back to its producers and, after the last block, signals the epilogue;
//   * epilogue: the 8 producer warps read the accumulators with tcgen05.ld (32x32b.x32) and store Y.
// Two shared-memory stages (2 x 80 KB); TMEM allocation = 512 columns.
// Numerics: weights are dequantised exactly as ggml does (f32) and rounded to f16, activations are f16, products
// are accumulated in f32 by the tensor core -- the tolerance-level path (like upstream's CUDA backend for batches),
// not the bit-exact integer path of the decode GEMV.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "layout.cuh"

#define GM_BM 128      /* weight rows per CTA  (UMMA M) */
#define GM_BN 512      /* tokens per CTA: TWO accumulators of UMMA N = 256 -- every dequantised weight tile feeds two MMAs */
#define GM_NH 256      /* UMMA N */
#define GM_BK 64       /* K elements per stage (one 64-element swizzle atom, four K=16 steps x two token halves) */
#define GM_STAGES 2
#define GM_PRODUCER_WARPS 8              /* warps 0-3 produce the even K blocks (stage 0), warps 4-7 the odd ones (stage 1) */
#define GM_THREADS ((GM_PRODUCER_WARPS + 1) * 32)
#define GM_A_BYTES (GM_BM * GM_BK * 2)   /* 16 KB */
#define GM_B_BYTES (GM_BN * GM_BK * 2)   /* 64 KB */
#define GM_STAGE_BYTES (GM_A_BYTES + GM_B_BYTES)
#define GM_TMEM_COLS 512

__device__ __forceinline__ uint32_t gm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gm_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gm_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gm_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GM_DONE;\n\t"
        "bra GM_WAIT;\n\t"
        "GM_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=2 (SWIZZLE_128B) [61,64))
__device__ __forceinline__ uint64_t gm_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    /* leading byte offset (unused for swizzled K-major) */
    d |= (uint64_t)(1024 >> 4) << 32;          /* stride byte offset: next 8-row group */
    d |= (uint64_t)1 << 46;                    /* descriptor version (Blackwell) */
    d |= (uint64_t)2 << 61;                    /* SWIZZLE_128B */
    return d;
}
// instruction descriptor, kind::f16: D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0, K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29).  fp16 operands, not f16: the same UTCHMMA rate with 11 significand bits instead of
// 8 -- the prefill's end-to-end error against the integer path drops 8x (tests/test_gpu_engine.py).
__device__ __forceinline__ uint32_t gm_idesc() {
    return (1u << 4) | ((uint32_t)(GM_NH >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
}

// byte offset of the 16-byte chunk `c` (0..7 inside a 128-byte swizzle row) of row `r` in a [rows][64 f16] atom
__device__ __forceinline__ uint32_t gm_sw(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {   /* two dequantised weights -> f16x2 (|w| << 65504) */
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// ---- packed-half arithmetic of the operand production (round 2).  The producers were the GEMM's bottleneck: ~12 f32
// instructions per weight (extract, int->float, mul, sub, round, pack) kept the tensor pipe at 31 %.  Now two weights travel
// per instruction: a byte b becomes the half 1024 + b by ONE byte-permute (0x6400 | b: halves in [1024, 2048) are spaced 1
// apart), HADD2 removes the bias exactly, and one HFMA2 / HMUL2 applies the sub-block scale (and the Q4_K/Q5_K offset).
//   Q8_0   d * q with d already an f16: the product is rounded once -- identical to "dequantise exactly, round to f16";
//   K-quants  the sub-block scale d*sc (and dmin*m) is rounded to f16 BEFORE the multiply: each weight carries up to ~3
//          f16 roundings instead of one (still ~1e-3 of ggml's integer dot per GEMM, tests/test_gpu_gemm.py).
__device__ __forceinline__ uint32_t h2_of_bytes(uint32_t x, uint32_t sel) { return __byte_perm(x, 0x64646464u, sel); }   /* sel 0x4140: bytes 0,1; 0x4342: bytes 2,3 */
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
// (byte - bias) * scale + offset for the two selected bytes of x; bias as the f16 pattern of 1024 + bias
__device__ __forceinline__ uint32_t h2_scale(uint32_t x, uint32_t sel, uint32_t bias1024, __half2 scale, __half2 offset) {
    return as_u32(__hfma2(__hsub2(as_h2(h2_of_bytes(x, sel)), as_h2(bias1024)), scale, offset));
}
__device__ __forceinline__ uint32_t h2_scale0(uint32_t x, uint32_t sel, uint32_t bias1024, __half2 scale) {
    return as_u32(__hmul2(__hsub2(as_h2(h2_of_bytes(x, sel)), as_h2(bias1024)), scale));
}

// small unsigned integer -> float without the (quarter-rate) I2F unit: 2^23 + n as a bit pattern, minus the bias (exact)
__device__ __forceinline__ float u2f(uint32_t n) { return __fsub_rn(__uint_as_float(0x4B000000u | n), 8388608.0f); }
__device__ __forceinline__ float u2f_bias(uint32_t n, float bias) { return __fsub_rn(__uint_as_float(0x4B000000u | n), bias); }   /* n - (bias - 2^23) */

// dequantise 64 consecutive elements (one "unit") of a weight row into 8 chunks of 8 f16
struct Chunk8 { uint4 c[8]; };

// The dequantiser is split into "fetch the packed bytes of my 64 elements" (issued one K block ahead, so the L2 /
// HBM latency hides behind the conversion of the current block) and "convert them to 8 chunks of 8 f16".
template <int TYPE> struct Raw;
template <> struct Raw<GGB_TYPE_Q4_K> { uint4 q0, q1, h; int g; };
template <> struct Raw<GGB_TYPE_Q5_K> { uint4 q0, q1, h; uint2 qhu; int g; };
template <> struct Raw<GGB_TYPE_Q6_K> { uint4 ql[4], qh[2]; uint2 sc8; uint32_t d16; int rp; };
template <> struct Raw<GGB_TYPE_Q8_0> { uint4 w[4]; uint32_t dd; };

// ek = element offset of the 64-element group inside its 2048-element tile
__device__ __forceinline__ void dq_fetch(Raw<GGB_TYPE_Q4_K>& R, const uint8_t* tile, int U, int nsb, int ek) {
    const int u = ek >> 6;
    R.q0 = ldg_stream(tile + 16 * u);
    R.q1 = ldg_stream(tile + 16 * U + 16 * u);
    R.h = ldg_cached(tile + 32 * U + 16 * (u >> 2));
    R.g = u & 3;
}
__device__ __forceinline__ void dq_convert(const Raw<GGB_TYPE_Q4_K>& R, Chunk8& o) {
    int s0, m0, s1, m1;
    {
        const uint32_t hw[4] = {R.h.x, R.h.y, R.h.z, R.h.w};
        const int bo = 4 + 3 * R.g;                              /* byte offset of the 24-bit field inside the header */
        const uint64_t two = (uint64_t)hw[bo >> 2] | ((uint64_t)hw[min((bo >> 2) + 1, 3)] << 32);
        const uint32_t f = (uint32_t)(two >> (8 * (bo & 3))) & 0xFFFFFFu;
        s0 = f & 63; s1 = (f >> 6) & 63; m0 = (f >> 12) & 63; m1 = (f >> 18) & 63;
    }
    const float d = h2f((uint16_t)(R.h.x & 0xFFFF)), dmin = h2f((uint16_t)(R.h.x >> 16));
    const __half2 dd[2] = {__float2half2_rn(__fmul_rn(d, (float)s0)), __float2half2_rn(__fmul_rn(d, (float)s1))};
    const __half2 nn[2] = {__float2half2_rn(-__fmul_rn(dmin, (float)m0)), __float2half2_rn(-__fmul_rn(dmin, (float)m1))};
    const uint32_t w[8] = {R.q0.x, R.q0.y, R.q0.z, R.q0.w, R.q1.x, R.q1.y, R.q1.z, R.q1.w};   /* bytes 0..31 of the 32-byte group */
    // elements 0..31 = low nibbles (sub-block 2g), 32..63 = high nibbles (sub-block 2g+1)
#pragma unroll
    for (int half = 0; half < 2; half++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {       /* 8 elements = bytes 8c..8c+7 = words 2c, 2c+1 */
            const uint32_t x0 = (half ? (w[2 * c] >> 4) : w[2 * c]) & 0x0F0F0F0Fu, x1 = (half ? (w[2 * c + 1] >> 4) : w[2 * c + 1]) & 0x0F0F0F0Fu;
            o.c[4 * half + c] = make_uint4(h2_scale(x0, 0x4140u, 0x64006400u, dd[half], nn[half]), h2_scale(x0, 0x4342u, 0x64006400u, dd[half], nn[half]),
                                           h2_scale(x1, 0x4140u, 0x64006400u, dd[half], nn[half]), h2_scale(x1, 0x4342u, 0x64006400u, dd[half], nn[half]));
        }
    }
}

// Q5_K: Q4_K plus the fifth bits (QHU[u]: bit l of word 0 / 1 = element l of sub-block 2g / 2g+1)
__device__ __forceinline__ void dq_fetch(Raw<GGB_TYPE_Q5_K>& R, const uint8_t* tile, int U, int nsb, int ek) {
    const int u = ek >> 6;
    R.q0 = ldg_stream(tile + 16 * u);
    R.q1 = ldg_stream(tile + 16 * U + 16 * u);
    R.qhu = __ldg(reinterpret_cast<const uint2*>(tile + 32 * U + 8 * u));
    R.h = ldg_cached(tile + 40 * U + 16 * (u >> 2));
    R.g = u & 3;
}
__device__ __forceinline__ void dq_convert(const Raw<GGB_TYPE_Q5_K>& R, Chunk8& o) {
    int s0, m0, s1, m1;
    {
        const uint32_t hw[4] = {R.h.x, R.h.y, R.h.z, R.h.w};
        const int bo = 4 + 3 * R.g;
        const uint64_t two = (uint64_t)hw[bo >> 2] | ((uint64_t)hw[min((bo >> 2) + 1, 3)] << 32);
        const uint32_t f = (uint32_t)(two >> (8 * (bo & 3))) & 0xFFFFFFu;
        s0 = f & 63; s1 = (f >> 6) & 63; m0 = (f >> 12) & 63; m1 = (f >> 18) & 63;
    }
    const float d = h2f((uint16_t)(R.h.x & 0xFFFF)), dmin = h2f((uint16_t)(R.h.x >> 16));
    const __half2 dd[2] = {__float2half2_rn(__fmul_rn(d, (float)s0)), __float2half2_rn(__fmul_rn(d, (float)s1))};
    const __half2 nn[2] = {__float2half2_rn(-__fmul_rn(dmin, (float)m0)), __float2half2_rn(-__fmul_rn(dmin, (float)m1))};
    const uint32_t w[8] = {R.q0.x, R.q0.y, R.q0.z, R.q0.w, R.q1.x, R.q1.y, R.q1.z, R.q1.w};
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t bits = half ? R.qhu.y : R.qhu.x;
#pragma unroll
        for (int c = 0; c < 4; c++) {       /* elements 8c..8c+7 of the sub-block: nibble | fifth bit << 4, four bits spread per word */
            const uint32_t f0 = (((bits >> (8 * c)) & 0xFu) * 0x02040810u) & 0x10101010u, f1 = (((bits >> (8 * c + 4)) & 0xFu) * 0x02040810u) & 0x10101010u;
            const uint32_t x0 = ((half ? (w[2 * c] >> 4) : w[2 * c]) & 0x0F0F0F0Fu) | f0, x1 = ((half ? (w[2 * c + 1] >> 4) : w[2 * c + 1]) & 0x0F0F0F0Fu) | f1;
            o.c[4 * half + c] = make_uint4(h2_scale(x0, 0x4140u, 0x64006400u, dd[half], nn[half]), h2_scale(x0, 0x4342u, 0x64006400u, dd[half], nn[half]),
                                           h2_scale(x1, 0x4140u, 0x64006400u, dd[half], nn[half]), h2_scale(x1, 0x4342u, 0x64006400u, dd[half], nn[half]));
        }
    }
}

// Q6_K: a 64-element K block j (0..3) of a super-block = elements 64j..64j+63 = half n=j>>1, r in {2*(j&1), 2*(j&1)+1}
__device__ __forceinline__ void dq_fetch(Raw<GGB_TYPE_Q6_K>& R, const uint8_t* tile, int U, int nsb, int ek) {
    const int sb = ek >> 8, j = (ek >> 6) & 3, n = j >> 1;
    R.rp = (j & 1) * 2;
    R.d16 = __ldg(reinterpret_cast<const uint16_t*>(tile + 48 * U + 16 * nsb + 2 * sb));
    R.sc8 = __ldg(reinterpret_cast<const uint2*>(tile + 48 * U + 16 * sb + 8 * n));
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int u = 4 * sb + 2 * n + t;
        R.qh[t] = ldg_stream(tile + 32 * U + 16 * u);
#pragma unroll
        for (int rr = 0; rr < 2; rr++) R.ql[2 * rr + t] = ldg_stream(tile + (((R.rp + rr) & 1) ? 16 * U : 0) + 16 * u);
    }
}
__device__ __forceinline__ void dq_convert(const Raw<GGB_TYPE_Q6_K>& R, Chunk8& o) {
    const float d = h2f((uint16_t)R.d16);
    const uint64_t sc64 = (uint64_t)R.sc8.x | ((uint64_t)R.sc8.y << 32);
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int r = R.rp + rr;
#pragma unroll
        for (int t = 0; t < 2; t++) {       /* 16 elements l = 16t..16t+15 of row-group r */
            const uint4 ql = R.ql[2 * rr + t], qh = R.qh[t];
            const __half2 ds = __float2half2_rn(__fmul_rn(d, (float)(int8_t)((sc64 >> (8 * (2 * r + t))) & 0xFF)));
            const uint32_t lw[4] = {ql.x, ql.y, ql.z, ql.w}, hw[4] = {qh.x, qh.y, qh.z, qh.w};
            uint32_t q[4];                  /* the 6-bit codes of 4 bytes per word: low or high nibble | two bits of qh << 4 */
#pragma unroll
            for (int i = 0; i < 4; i++) q[i] = (((r & 2) ? (lw[i] >> 4) : lw[i]) & 0x0F0F0F0Fu) | (((hw[i] >> (2 * r)) & 0x03030303u) << 4);
#pragma unroll
            for (int c = 0; c < 2; c++)     /* (q - 32) * (d * sc): bias pattern 1024 + 32 = 0x6420 */
                o.c[4 * rr + 2 * t + c] = make_uint4(h2_scale0(q[2 * c], 0x4140u, 0x64206420u, ds), h2_scale0(q[2 * c], 0x4342u, 0x64206420u, ds),
                                                     h2_scale0(q[2 * c + 1], 0x4140u, 0x64206420u, ds), h2_scale0(q[2 * c + 1], 0x4342u, 0x64206420u, ds));
        }
    }
}

__device__ __forceinline__ void dq_fetch(Raw<GGB_TYPE_Q8_0>& R, const uint8_t* tile, int U, int nsb, int ek) {
    const int u = ek >> 6;
    R.dd = __ldg(reinterpret_cast<const uint32_t*>(tile + 64 * U + 4 * u));
#pragma unroll
    for (int i = 0; i < 4; i++) R.w[i] = ldg_stream(tile + i * 16 * U + 16 * u);
}
__device__ __forceinline__ void dq_convert(const Raw<GGB_TYPE_Q8_0>& R, Chunk8& o) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t dbits = (i >> 1) ? (R.dd >> 16) : (R.dd & 0xFFFF);
        const __half2 d2 = as_h2(dbits | (dbits << 16));              /* the block's f16 scale, as stored */
        const uint32_t w[4] = {R.w[i].x ^ 0x80808080u, R.w[i].y ^ 0x80808080u, R.w[i].z ^ 0x80808080u, R.w[i].w ^ 0x80808080u};   /* q + 128 */
#pragma unroll
        for (int c = 0; c < 2; c++)         /* (byte - 128) * d: bias pattern 1024 + 128 = 0x6480; one rounding, as exact-then-round */
            o.c[2 * i + c] = make_uint4(h2_scale0(w[2 * c], 0x4140u, 0x64806480u, d2), h2_scale0(w[2 * c], 0x4342u, 0x64806480u, d2),
                                        h2_scale0(w[2 * c + 1], 0x4140u, 0x64806480u, d2), h2_scale0(w[2 * c + 1], 0x4342u, 0x64806480u, d2));
    }
}

template <int TYPE>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_kernel(const uint8_t* __restrict__ W, int64_t w_stride, int rows, int K, const __grid_constant__ CUtensorMap tmapX, int tokens,
            float* __restrict__ Y, int64_t y_stride, const uint8_t* __restrict__ W2, float* __restrict__ Y2) {
    extern __shared__ __align__(1024) uint8_t gsm[];
    __shared__ __align__(8) uint64_t bar_full[GM_STAGES], bar_empty[GM_STAGES], bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Two weight matrices of the same shape and format may share a launch (ffn_gate + ffn_up: 2 x 448 tiles = 6.05 waves on 148
    // SMs instead of two launches of 3.03 waves, each with a nearly empty fourth wave): row tiles beyond the first matrix's
    // belong to the second one.
    const int rtiles = (rows + GM_BM - 1) / GM_BM;
    int rt = blockIdx.x;
    if (rt >= rtiles) { rt -= rtiles; W = W2; Y = Y2; }
    const int row0 = rt * GM_BM, tok0 = blockIdx.y * GM_BN;
    const int nkb = K / GM_BK;
    uint8_t* stage_base = (uint8_t*)(((uintptr_t)gsm + 1023) & ~(uintptr_t)1023);

    if (tid == 0) {
        /* full: one arrival per producer warp (A stored) + the expect_tx arrival of the thread that issued the TMA copies of B */
        for (int s = 0; s < GM_STAGES; s++) { gm_mbar_init(gm_smem_u32(&bar_full[s]), GM_PRODUCER_WARPS / 2 + 1); gm_mbar_init(gm_smem_u32(&bar_empty[s]), 1); }
        gm_mbar_init(gm_smem_u32(&bar_acc), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GM_PRODUCER_WARPS) {   // the MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(&tmem_base_s)), "n"(GM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_s;

    if (warp < GM_PRODUCER_WARPS) {
        // ===== producers: dequantise A, copy B =====
        // The activation block travels global -> shared with cp.async (no registers, swizzle applied per 16-byte
        // chunk) one K block AHEAD, so its latency hides behind the dequantisation of the current block; the weight
        // tiles (2048 elements per row) are pulled into L2 one tile ahead so the dequantiser's loads are L2 hits.
        const int pt = tid;                                   /* 0..255 */
        const int tile_bytes = ggb_sb_bytes(TYPE) * GGB_TILE_SB;
        const int n_half = (tokens - tok0 > GM_NH) ? 2 : 1;   /* token halves this CTA really has */
        auto issue_b = [&](int kb) {   /* ONE thread: the 512-token x 64-K activation block, one 256-token box per half */
            const int st = kb % GM_STAGES;
            const uint32_t sB = gm_smem_u32(stage_base + st * GM_STAGE_BYTES + GM_A_BYTES), bar = gm_smem_u32(&bar_full[st]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(n_half * GM_NH * GM_BK * 2)) : "memory");
            for (int h = 0; h < n_half; h++)   /* rows beyond `tokens` are zero-filled by the copy engine */
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(sB + h * (GM_NH * 128)), "l"(&tmapX), "r"(kb * GM_BK), "r"(tok0 + h * GM_NH), "r"(bar) : "memory");
        };
        auto prefetch_tile = [&](int t) {                      /* one thread per row pulls tile t of its row into L2 */
            if (pt < GM_BM && row0 + pt < rows && t * GGB_TILE_ELEMS < K) {
                const uint8_t* p = W + (int64_t)(row0 + pt) * w_stride + (int64_t)t * tile_bytes;
                for (int o = 0; o < tile_bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
            }
        };
        prefetch_tile(0);
        prefetch_tile(1);
        // thread -> (row, parity): the half `par` of the producers owns the K blocks kb = par, par + 2, ... and with them
        // stage `par` -- the two halves run one block apart, which is the whole software pipeline
        const int r = pt & 127, par = pt >> 7;
        const int grow = row0 + r;
        const bool arow = grow < rows;
        Raw<TYPE> raw;
        auto fetch_a = [&](int kb) {
            const int k0 = kb * GM_BK;
            const int t = k0 / GGB_TILE_ELEMS, ek = k0 - t * GGB_TILE_ELEMS;
            const int nsb = ggb_tile_nsb(K, t);
            dq_fetch(raw, W + (int64_t)grow * w_stride + (int64_t)t * tile_bytes, 4 * nsb, nsb, ek);
        };
        if (arow && par < nkb) fetch_a(par);
        uint8_t* sA = stage_base + par * GM_STAGE_BYTES;
        for (int kb = par; kb < nkb; kb += 2) {
            // my stage is free once the MMAs of block kb - 2 have read it; its activation block starts travelling at once
            if (kb >= 2) gm_mbar_wait(gm_smem_u32(&bar_empty[par]), ((kb >> 1) & 1) ^ 1);
            if (r == 0) issue_b(kb);
            if (par == 0 && (kb * GM_BK) % GGB_TILE_ELEMS == 0) prefetch_tile(kb * GM_BK / GGB_TILE_ELEMS + 2);
            Chunk8 ch;
            // Where the next block's packed bytes are requested is a measured choice per format: the loads of two
            // iterations share a scoreboard, so requesting early makes the first use of the CURRENT bytes wait for the
            // new loads too (ncu: stall_long_sb).  Q4_K (3 loads, short conversion) is better off requesting after the
            // conversion; Q6_K / Q8_0 (long conversions) before it.
            constexpr bool FETCH_EARLY = (TYPE != GGB_TYPE_Q4_K);
            if (arow) {
                const Raw<TYPE> cur = raw;
                if (FETCH_EARLY && kb + 2 < nkb) fetch_a(kb + 2);
                dq_convert(cur, ch);
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++) ch.c[c] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(sA + gm_sw(r, c)) = ch.c[c];
            if (!FETCH_EARLY && arow && kb + 2 < nkb) fetch_a(kb + 2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* generic-proxy writes -> visible to the tensor core */
            __syncwarp();
            if (lane == 0) gm_mbar_arrive(gm_smem_u32(&bar_full[par]));
        }
        // ===== epilogue: TMEM -> registers -> Y =====
        gm_mbar_wait(gm_smem_u32(&bar_acc), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, colh = warp >> 2;              /* TMEM lane quarter, column half */
        const int erow = row0 + 32 * q + lane;
#pragma unroll 1
        for (int cb = 0; cb < GM_NH / 32; cb++) {              /* 8 x 32 columns = this warp's 256 tokens */
            const int col = colh * GM_NH + cb * 32;
            if (tok0 + col >= tokens) break;
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)col;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                  "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                  "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (erow < rows) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int tk = tok0 + col + j;
                    if (tk < tokens) Y[(int64_t)tk * y_stride + erow] = __uint_as_float(v[j]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===== MMA warp: one elected lane issues =====
        const uint32_t idesc = gm_idesc();
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % GM_STAGES;
            const uint32_t ph = (kb / GM_STAGES) & 1;
            gm_mbar_wait(gm_smem_u32(&bar_full[s]), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a0 = gm_smem_u32(stage_base + s * GM_STAGE_BYTES), b0 = a0 + GM_A_BYTES;
                const int n_half = (tokens - tok0 > GM_NH) ? 2 : 1;
#pragma unroll
                for (int k = 0; k < GM_BK / 16; k++) {
                    const uint64_t da = gm_desc(a0 + k * 32);                              /* 32 B per K=16 step inside the atom */
                    const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                    for (int h = 0; h < n_half; h++) {                                     /* the same weight tile against both token halves */
                        const uint64_t db = gm_desc(b0 + h * (GM_NH * 128) + k * 32);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem_acc + (uint32_t)(h * GM_NH)), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    }
                }
                // tcgen05.commit implies fence::before_thread_sync; frees the smem stage when the MMAs have read it
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gm_smem_u32(&bar_empty[s])) : "memory");
                if (kb == nkb - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gm_smem_u32(&bar_acc)) : "memory");
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == GM_PRODUCER_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(GM_TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ f32 -> f16 activations (saturating: no Inf from an outlier)
__device__ __forceinline__ __half f2h_sat(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }
__global__ void f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        *reinterpret_cast<__half2*>(y + i) = __halves2half2(f2h_sat(x[i]), f2h_sat(x[i + 1]));
    } else if (i < n) {
        y[i] = f2h_sat(x[i]);
    }
}

extern "C" int ggb_f32_to_f16(const float* x, void* y_f16, int64_t n, void* stream) {
    if (n < 0 || (n && (!x || !y_f16))) GGB_FAIL(GGB_ERR_ARG, "ggb_f32_to_f16: bad argument");
    if (n == 0) return GGB_OK;
    f32_to_f16_kernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)y_f16, n);
    GGB_CHECK_LAUNCH("ggb_f32_to_f16");
    return GGB_OK;
}

// tensor map over X [tokens][k] f16: box = 64 K-elements (128 B, one swizzle atom row) x GM_BN tokens, SWIZZLE_128B -- the
// shared-memory image is exactly the K-major layout the UMMA descriptors describe (16-byte chunk c of row r at c ^ (r & 7))
typedef CUresult (*ggb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ggb_encode_tiled_fn encode_tiled() {
    static ggb_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (ggb_encode_tiled_fn)p;
    }
    return fn;
}

template <int TYPE>
static int launch_gemm(const void* w, int rows, int k, const void* x, int tokens, float* y, int64_t y_stride, cudaStream_t st,
                       const void* w2 = nullptr, float* y2 = nullptr) {
    ggb_encode_tiled_fn enc = encode_tiled();
    if (!enc) GGB_FAIL(GGB_ERR_CUDA, "ggb_gemm: the driver does not export cuTensorMapEncodeTiled");
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)tokens}, strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {64, GM_NH}, estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) GGB_FAIL(GGB_ERR_CUDA, "ggb_gemm: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    static bool attr = false;
    const size_t smem = GM_STAGES * GM_STAGE_BYTES + 1024;
    if (!attr) {
        GGB_CUDA(cudaFuncSetAttribute(gemm_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid((rows + GM_BM - 1) / GM_BM * (w2 ? 2 : 1), (tokens + GM_BN - 1) / GM_BN);
    gemm_kernel<TYPE><<<grid, GM_THREADS, smem, st>>>((const uint8_t*)w, ggb_row_stride(TYPE, k), rows, k, tmap, tokens, y, y_stride,
                                                      (const uint8_t*)w2, y2);
    GGB_CHECK_LAUNCH("ggb_gemm");
    return GGB_OK;
}

extern "C" int ggb_gemm(int type, const void* w, int rows, int k, const void* x_f16, int tokens, float* y, int64_t y_stride, void* stream) {
    if (rows < 0 || tokens < 0 || k <= 0 || (k % GM_BK)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm: k=%d must be a positive multiple of %d", k, GM_BK);
    if (rows == 0 || tokens == 0) return GGB_OK;
    if (!w || !x_f16 || !y) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm: null pointer");
    if (((uintptr_t)w & 15) || ((uintptr_t)x_f16 & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm: operands must be 16-byte aligned");
    if (y_stride < rows) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm: y_stride smaller than rows");
    cudaStream_t st = (cudaStream_t)stream;
    switch (type) {
        case GGB_TYPE_Q4_K: return launch_gemm<GGB_TYPE_Q4_K>(w, rows, k, x_f16, tokens, y, y_stride, st);
        case GGB_TYPE_Q5_K: return launch_gemm<GGB_TYPE_Q5_K>(w, rows, k, x_f16, tokens, y, y_stride, st);
        case GGB_TYPE_Q6_K: return launch_gemm<GGB_TYPE_Q6_K>(w, rows, k, x_f16, tokens, y, y_stride, st);
        case GGB_TYPE_Q8_0: return launch_gemm<GGB_TYPE_Q8_0>(w, rows, k, x_f16, tokens, y, y_stride, st);
        default: GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemm: unsupported weight type %d", type);
    }
}

// two weight matrices of the same format and shape against the same activations in ONE launch (ffn_gate + ffn_up)
extern "C" int ggb_gemm2(int type, const void* w0, const void* w1, int rows, int k, const void* x_f16, int tokens, float* y0, float* y1,
                         int64_t y_stride, void* stream) {
    if (rows < 0 || tokens < 0 || k <= 0 || (k % GM_BK)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm2: k=%d must be a positive multiple of %d", k, GM_BK);
    if (rows == 0 || tokens == 0) return GGB_OK;
    if (!w0 || !w1 || !x_f16 || !y0 || !y1) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm2: null pointer");
    if (((uintptr_t)w0 & 15) || ((uintptr_t)w1 & 15) || ((uintptr_t)x_f16 & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm2: operands must be 16-byte aligned");
    if (y_stride < rows) GGB_FAIL(GGB_ERR_ARG, "ggb_gemm2: y_stride smaller than rows");
    cudaStream_t st = (cudaStream_t)stream;
    switch (type) {
        case GGB_TYPE_Q4_K: return launch_gemm<GGB_TYPE_Q4_K>(w0, rows, k, x_f16, tokens, y0, y_stride, st, w1, y1);
        case GGB_TYPE_Q5_K: return launch_gemm<GGB_TYPE_Q5_K>(w0, rows, k, x_f16, tokens, y0, y_stride, st, w1, y1);
        case GGB_TYPE_Q6_K: return launch_gemm<GGB_TYPE_Q6_K>(w0, rows, k, x_f16, tokens, y0, y_stride, st, w1, y1);
        case GGB_TYPE_Q8_0: return launch_gemm<GGB_TYPE_Q8_0>(w0, rows, k, x_f16, tokens, y0, y_stride, st, w1, y1);
        default: GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_gemm2: unsupported weight type %d", type);
    }
}
