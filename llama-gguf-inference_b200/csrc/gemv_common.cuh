// gemv_common.cuh -- device helpers shared by the batch-1 GEMV (gemv.cu) and the small-batch GEMV (gemv_batch.cu):
// mbarrier / bulk-copy PTX, the bank swizzle of the activation image, shared-memory loads by 32-bit address, and
// the per-unit integer dot + f32 term of every weight format ("canon" arithmetic, oracle/ggml_ref.c *_canon).
#pragma once
#include "common.cuh"
#include "layout.cuh"

// ------------------------------------------------------------------ mbarrier / bulk-copy primitives (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP;\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared::cta, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// The same copy with an L2 eviction-priority hint.  Weights are read exactly once per token and 4.6 GB of them stream
// through the 126 MB L2 per token; with the default policy they flush everything else out of it -- including the
// INSTRUCTIONS of the kernels that are not running at the moment, so that every switch between kernel instances re-fetched
// its code from HBM (measured: +2.5..4.5 us per switch, tools/gemv_bench.py sequences).  evict_first makes the streamed
// lines the first victims and leaves code, activations and the KV cache resident.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// bank swizzle of 16-byte activation chunks: conflict-free LDS.128 for both the Q4_K/Q8_0 unit pattern
// (chunks 4u+i) and the Q6_K pattern (chunks 16sb+8n+2r+t) -- see DESIGN.md "activation staging".
__device__ __forceinline__ int swz(int c) { return c ^ ((c >> 2) & 7); }

// shared-memory loads by 32-bit shared address (+ immediate): no generic-address arithmetic in the hot loop
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
    return r;
}

__device__ __forceinline__ uint4 and4(uint4 v, uint32_t m) { return make_uint4(v.x & m, v.y & m, v.z & m, v.w & m); }
__device__ __forceinline__ int dot16_us(uint4 w, uint4 a) {  /* two independent dp4a chains */
    const int s0 = dp4a_us(w.y, a.y, dp4a_us(w.x, a.x, 0));
    const int s1 = dp4a_us(w.w, a.w, dp4a_us(w.z, a.z, 0));
    return s0 + s1;
}
__device__ __forceinline__ int dot16_ss(uint4 w, uint4 a) {
    const int s0 = dp4a_ss(w.y, a.y, dp4a_ss(w.x, a.x, 0));
    const int s1 = dp4a_ss(w.w, a.w, dp4a_ss(w.z, a.z, 0));
    return s0 + s1;
}

// the activations one lane needs for its unit of a K-tile (shared by both rows of a pair)
struct Act {
    uint4 a0, a1, a2, a3;
    int b0, b1, b2, b3;   /* Q4_K: b0 = sum of sub-block 2g, b1 = sum of 2g+1; Q6_K: 32 * per-16 sums */
    float dx0, dx1;       /* activation block scale(s) */
};

// qs_s / bs_s / dsc_s = 32-bit shared addresses of the int8 codes, per-16 sums (int16), block scales (f32)
template <int MASK>
__device__ __forceinline__ Act load_act(int type, int gu, uint32_t qs_s, uint32_t bs_s, uint32_t dsc_s) {
    Act A;
    A.b0 = A.b1 = A.b2 = A.b3 = 0;
    A.dx1 = 0.f;
    if ((MASK & 2) && (MASK == 2 || type == GGB_TYPE_Q6_K)) {
        const int c0 = 4 * gu - 3 * (gu & 1); /* chunk of r = 0; r adds 2 */
        A.a0 = lds128(qs_s + 16 * swz(c0));
        A.a1 = lds128(qs_s + 16 * swz(c0 + 2));
        A.a2 = lds128(qs_s + 16 * swz(c0 + 4));
        A.a3 = lds128(qs_s + 16 * swz(c0 + 6));
        A.b0 = 32 * (int)(int16_t)lds16(bs_s + 2 * c0);
        A.b1 = 32 * (int)(int16_t)lds16(bs_s + 2 * c0 + 4);
        A.b2 = 32 * (int)(int16_t)lds16(bs_s + 2 * c0 + 8);
        A.b3 = 32 * (int)(int16_t)lds16(bs_s + 2 * c0 + 12);
        A.dx0 = __uint_as_float(lds32(dsc_s + 4 * (gu >> 2)));
        return A;
    }
    A.a0 = lds128(qs_s + 16 * swz(4 * gu + 0));
    A.a1 = lds128(qs_s + 16 * swz(4 * gu + 1));
    A.a2 = lds128(qs_s + 16 * swz(4 * gu + 2));
    A.a3 = lds128(qs_s + 16 * swz(4 * gu + 3));
    if (MASK & 4) {
        const uint2 dx = lds64(dsc_s + 8 * gu);
        A.dx0 = __uint_as_float(dx.x); A.dx1 = __uint_as_float(dx.y);
    } else {
        const uint2 bs = lds64(bs_s + 8 * gu); /* four per-16 sums */
        A.b0 = (int)(int16_t)(bs.x & 0xFFFF) + ((int)bs.x >> 16);
        A.b1 = (int)(int16_t)(bs.y & 0xFFFF) + ((int)bs.y >> 16);
        A.dx0 = __uint_as_float(lds32(dsc_s + 4 * (gu >> 2)));
    }
    return A;
}

// per-lane constants: byte-permute selectors that pull the lane's 24-bit scale field out of the header
// (layout.cuh: field g sits at header bytes 4+3g .. 6+3g)
struct LaneK { uint32_t selA, selB; bool lowg; uint32_t o_q, o_h, o_sc, o_d; };
__device__ __forceinline__ LaneK lane_consts(int lane) {
    LaneK L;
    const int g = lane & 3;
    L.lowg = g < 2;
    L.selA = (g == 0) ? 0x3210u : 0x0543u;   /* on (hdr.y, hdr.z) */
    L.selB = (g == 2) ? 0x0432u : 0x0765u;   /* on (hdr.z, hdr.w) */
    L.o_q = 16u * lane;                        /* the lane's 16-byte chunk inside a section */
    L.o_h = 16u * (lane >> 2);                 /* its super-block header / scale row */
    L.o_sc = 16u * (lane >> 2) + 8u * ((lane >> 1) & 1);  /* Q6_K: the 8 scales of half n */
    L.o_d = 2u * (lane >> 2);
    return L;
}

// one f32 term of a Q4_K unit: sub-blocks 2g (low nibbles) and 2g+1 (high nibbles) of one super-block
// (oracle: gref_vec_dot_q4_K_q8_K_canon -- same integers, same f32 operation order)
__device__ __forceinline__ float term_q4k(uint4 q0, uint4 q1, uint4 hdr, const Act& A, const LaneK& L) {
    const int dlo = dot16_us(and4(q0, 0x0F0F0F0Fu), A.a0) + dot16_us(and4(q1, 0x0F0F0F0Fu), A.a1);
    const int dhi = (dot16_us(and4(q0, 0xF0F0F0F0u), A.a2) + dot16_us(and4(q1, 0xF0F0F0F0u), A.a3)) >> 4; /* exact */
    const uint32_t fa = __byte_perm(hdr.y, hdr.z, L.selA), fb = __byte_perm(hdr.z, hdr.w, L.selB);
    const uint32_t f = L.lowg ? fa : fb;       /* sc[2g] | sc[2g+1]<<6 | min[2g]<<12 | min[2g+1]<<18 */
    const int isum = (int)(f & 63) * dlo + (int)((f >> 6) & 63) * dhi;
    const int msum = (int)((f >> 12) & 63) * A.b0 + (int)((f >> 18) & 63) * A.b1;
    const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
    return __fsub_rn(__fmul_rn(__fmul_rn(d, A.dx0), (float)isum), __fmul_rn(__fmul_rn(dmin, A.dx0), (float)msum));
}

// Q5_K unit: Q4_K's layout plus one "fifth bit" per element (QHU[u] = two u32, bit l = element l of sub-blocks 2g / 2g+1).
// The 5-bit codes are assembled as bytes (nibble | bit << 4; four bits are spread to four bytes with one multiply), then
// the arithmetic is term_q4k's (oracle: vec_dot_q5_K_q8_K_f64).
__device__ __forceinline__ uint32_t spread5(uint32_t bits, int sh) { return (((bits >> sh) & 0xFu) * 0x02040810u) & 0x10101010u; }
__device__ __forceinline__ void q5k_codes(uint4 q0, uint4 q1, uint2 qhu, uint4& l0, uint4& l1, uint4& h0, uint4& h1) {
    const uint32_t bl = qhu.x, bh = qhu.y;
    l0 = make_uint4((q0.x & 0x0F0F0F0Fu) | spread5(bl, 0), (q0.y & 0x0F0F0F0Fu) | spread5(bl, 4),
                    (q0.z & 0x0F0F0F0Fu) | spread5(bl, 8), (q0.w & 0x0F0F0F0Fu) | spread5(bl, 12));
    l1 = make_uint4((q1.x & 0x0F0F0F0Fu) | spread5(bl, 16), (q1.y & 0x0F0F0F0Fu) | spread5(bl, 20),
                    (q1.z & 0x0F0F0F0Fu) | spread5(bl, 24), (q1.w & 0x0F0F0F0Fu) | spread5(bl, 28));
    h0 = make_uint4(((q0.x >> 4) & 0x0F0F0F0Fu) | spread5(bh, 0), ((q0.y >> 4) & 0x0F0F0F0Fu) | spread5(bh, 4),
                    ((q0.z >> 4) & 0x0F0F0F0Fu) | spread5(bh, 8), ((q0.w >> 4) & 0x0F0F0F0Fu) | spread5(bh, 12));
    h1 = make_uint4(((q1.x >> 4) & 0x0F0F0F0Fu) | spread5(bh, 16), ((q1.y >> 4) & 0x0F0F0F0Fu) | spread5(bh, 20),
                    ((q1.z >> 4) & 0x0F0F0F0Fu) | spread5(bh, 24), ((q1.w >> 4) & 0x0F0F0F0Fu) | spread5(bh, 28));
}
__device__ __forceinline__ float term_q5k(uint4 q0, uint4 q1, uint2 qhu, uint4 hdr, const Act& A, const LaneK& L) {
    uint4 l0, l1, h0, h1;
    q5k_codes(q0, q1, qhu, l0, l1, h0, h1);
    const int dlo = dot16_us(l0, A.a0) + dot16_us(l1, A.a1);
    const int dhi = dot16_us(h0, A.a2) + dot16_us(h1, A.a3);
    const uint32_t fa = __byte_perm(hdr.y, hdr.z, L.selA), fb = __byte_perm(hdr.z, hdr.w, L.selB);
    const uint32_t f = L.lowg ? fa : fb;
    const int isum = (int)(f & 63) * dlo + (int)((f >> 6) & 63) * dhi;
    const int msum = (int)((f >> 12) & 63) * A.b0 + (int)((f >> 18) & 63) * A.b1;
    const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
    return __fsub_rn(__fmul_rn(__fmul_rn(d, A.dx0), (float)isum), __fmul_rn(__fmul_rn(dmin, A.dx0), (float)msum));
}

// Q6_K unit (half n, column t): elements 128n + 32r + 16t + (0..15), r = 0..3; sc8 = the 8 scales of half n
__device__ __forceinline__ float term_q6k(uint4 qla, uint4 qlb, uint4 qh, uint2 sc8, uint32_t dbits, const Act& A, int tt) {
    // sum (q-32)*a over a 16-group = nibble dot + 16*(2-bit dot) - 32*sum(a); masked bytes keep their position,
    // the power-of-two factor is removed by an exact shift
    const int v0 = dot16_us(and4(qla, 0x0F0F0F0Fu), A.a0) + (dot16_us(and4(qh, 0x03030303u), A.a0) << 4) - A.b0;
    const int v1 = dot16_us(and4(qlb, 0x0F0F0F0Fu), A.a1) + (dot16_us(and4(qh, 0x0C0C0C0Cu), A.a1) << 2) - A.b1;
    const int v2 = (dot16_us(and4(qla, 0xF0F0F0F0u), A.a2) >> 4) + dot16_us(and4(qh, 0x30303030u), A.a2) - A.b2;
    const int v3 = (dot16_us(and4(qlb, 0xF0F0F0F0u), A.a3) >> 4) + (dot16_us(and4(qh, 0xC0C0C0C0u), A.a3) >> 2) - A.b3;
    const uint32_t lo = tt ? (sc8.x >> 8) : sc8.x, hi = tt ? (sc8.y >> 8) : sc8.y; /* scale byte 2r + t */
    const int isum = (int)(int8_t)(lo & 0xFF) * v0 + (int)(int8_t)((lo >> 16) & 0xFF) * v1 +
                     (int)(int8_t)(hi & 0xFF) * v2 + (int)(int8_t)((hi >> 16) & 0xFF) * v3;
    return __fmul_rn(__fmul_rn(h2f((uint16_t)dbits), A.dx0), (float)isum);
}

// one (row, tile) item read from its ring slot (32-bit shared address); U16 = 16*U = section size in bytes.
// FULL tiles (U = 32) get compile-time offsets; returns the lane's contribution as f64.
template <int MASK, bool FULL>
__device__ __forceinline__ double consume(int type, uint32_t slot, int lane, int U, int nsb, const Act& A, const LaneK& L) {
    if (!FULL && lane >= U) return 0.0;
    const uint32_t S = FULL ? 512u : 16u * (uint32_t)U;   /* bytes per 16-byte-per-unit section */
    if ((MASK & 1) && (MASK == 1 || type == GGB_TYPE_Q4_K)) {
        const uint4 q0 = lds128(slot + L.o_q);
        const uint4 q1 = lds128(slot + L.o_q + S);
        const uint4 hd = lds128(slot + L.o_h + 2 * S);
        return (double)term_q4k(q0, q1, hd, A, L);
    } else if ((MASK & 2) && (MASK == 2 || type == GGB_TYPE_Q6_K)) {
        const uint4 qla = lds128(slot + L.o_q);
        const uint4 qlb = lds128(slot + L.o_q + S);
        const uint4 qh = lds128(slot + L.o_q + 2 * S);
        const uint2 sc = lds64(slot + L.o_sc + 3 * S);
        const uint32_t db = lds16(slot + L.o_d + 3 * S + (FULL ? 128u : 16u * (uint32_t)nsb));
        return (double)term_q6k(qla, qlb, qh, sc, db, A, lane & 1);
    } else if ((MASK & 8) && type == GGB_TYPE_Q5_K) {
        const uint4 q0 = lds128(slot + L.o_q);
        const uint4 q1 = lds128(slot + L.o_q + S);
        const uint2 qhu = lds64(slot + 2 * S + 8u * (uint32_t)lane);
        const uint4 hd = lds128(slot + L.o_h + 2 * S + S / 2);
        return (double)term_q5k(q0, q1, qhu, hd, A, L);
    } else if (MASK & 4) {
        const uint4 w0 = lds128(slot + L.o_q);
        const uint4 w1 = lds128(slot + L.o_q + S);
        const uint4 w2 = lds128(slot + L.o_q + 2 * S);
        const uint4 w3 = lds128(slot + L.o_q + 3 * S);
        const uint32_t dd = lds32(slot + 4 * S + 4u * lane);
        const int i0 = dot16_ss(w0, A.a0) + dot16_ss(w1, A.a1);
        const int i1 = dot16_ss(w2, A.a2) + dot16_ss(w3, A.a3);
        const float t0 = __fmul_rn((float)i0, __fmul_rn(h2f((uint16_t)(dd & 0xFFFF)), A.dx0));
        const float t1 = __fmul_rn((float)i1, __fmul_rn(h2f((uint16_t)(dd >> 16)), A.dx1));
        return (double)t0 + (double)t1; /* one f32 term per 32-block, added in f64 */
    }
    return 0.0;
}

