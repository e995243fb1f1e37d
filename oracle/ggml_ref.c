/*
 * oracle/ggml_ref.c -- CPU restatement of the ggml CPU arithmetic that the reference's backend
 * (`/app/llama-server`, NGL=0; call sites /root/reference/scripts/start.sh:473-480,516 and
 * /root/reference/Dockerfile.cpu:11,84-89) runs for a llama-architecture GGUF model.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product path
 * (llama-gguf-inference_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY PINNING STATUS
 *   - dequantize_row_{q8_0,q4_K,q5_K,q6_K,q4_0,q5_0}: PINNED bit-exact against gguf-py 0.19.0
 *     (`gguf.quants.dequantize`, site-packages/gguf/quants.py:396-401, 475-522, 525-549, 552-572; Q4_0 / Q5_0
 *     classes of the same file -- the python package published from the llama.cpp tree) through
 *     tests/golden/dequant_*.npz.
 *   - quantize_row_q8_0: PINNED bit-exact against gguf-py `Q8_0.quantize_blocks` (quants.py:378-394).
 *   - quantize_row_q8_K, vec_dot_*_q8_K, vec_dot_{q4_0,q5_0}_q8_0, rms_norm, rope, soft_max, silu, graph order:
 *     PARITY UNPINNED. The upstream sources (ggml/src/ggml-quants.c, ggml/src/ggml-cpu/quants.c,
 *     ggml/src/ggml-cpu/ops.cpp, src/llama-model.cpp) are not vendored by the reference (it pulls a
 *     prebuilt, unpinned `ghcr.io/ggml-org/llama.cpp:server` image) and are not on this machine, and
 *     the reference's own tests hold no numeric vector for this path.  These
 *     functions restate the published generic-C algorithms from knowledge of upstream; they are
 *     bounded (not pinned) by a float32 matvec over gguf-py-dequantised weights in tests/.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
 * -ffp-contract=off matters: ggml's dequantisation is "two f32 products then a subtract" and must not
 * be contracted into an FMA, or bit-exactness against gguf-py is lost.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QK_K 256
#define QK8_0 32

/* ggml tensor type ids, gguf/constants.py:4059-4093 */
enum { GREF_F32 = 0, GREF_F16 = 1, GREF_Q4_0 = 2, GREF_Q5_0 = 6, GREF_Q8_0 = 8, GREF_Q4_K = 12, GREF_Q5_K = 13, GREF_Q6_K = 14 };

/* ---- block layouts (gguf/quants.py sizes: Q8_0 (32,34)  Q4_K (256,144)  Q5_K (256,176)  Q6_K (256,210)) ---- */
#pragma pack(push, 1)
typedef struct { uint16_t d; int8_t qs[QK8_0]; } blk_q8_0;                                     /* 34 */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                                        /* 18: 32 x 4 bit, value d * (q - 8) */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } blk_q5_0;                         /* 22: 32 x 5 bit, value d * (q - 16) */
#pragma pack(pop)
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qs[QK_K / 2]; } blk_q4_K; /* 144 */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qh[QK_K / 8]; uint8_t qs[QK_K / 2]; } blk_q5_K; /* 176 */
typedef struct { uint8_t ql[QK_K / 2]; uint8_t qh[QK_K / 4]; int8_t scales[QK_K / 16]; uint16_t d; } blk_q6_K; /* 210 */
/* activation blocks used by the CPU matmul (upstream: block_q8_K) */
typedef struct { float d; int8_t qs[QK_K]; int16_t bsums[QK_K / 16]; } blk_q8_K;               /* 292 */
#pragma pack(pop)

/* ------------------------------------------------------------------ fp16 */
float gref_fp16_to_fp32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* round-to-nearest-even, like F16C / ggml's software path */
uint16_t gref_fp32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7FFFFFFFu;
    if (absx >= 0x7F800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7C00u | ((absx > 0x7F800000u) ? 0x200u : 0u));
    }
    if (absx >= 0x477FF000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7C00u);
    }
    if (absx < 0x33000001u) { /* < 2^-25 (or exactly 2^-25 -> ties to even 0) */
        return (uint16_t)sign;
    }
    int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t hexp;
    if (e < -14) { shift = 13 + (-14 - e); hexp = 0; }
    else         { shift = 13;            hexp = (uint32_t)(e + 15); }
    uint32_t keep = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (keep & 1u))) keep++;
    uint32_t h;
    if (hexp == 0) h = keep;                 /* subnormal; a carry into bit 10 lands on exp=1 correctly */
    else           h = ((hexp - 1) << 10) + keep; /* keep carries the implicit 1 (bit 10) */
    return (uint16_t)(sign | h);
}

void gref_fp16_to_fp32_row(const uint16_t *x, float *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = gref_fp16_to_fp32(x[i]); }
void gref_fp32_to_fp16_row(const float *x, uint16_t *y, int64_t n) { for (int64_t i = 0; i < n; i++) y[i] = gref_fp32_to_fp16(x[i]); }

/* ------------------------------------------------------------------ sizes */
int64_t gref_block_elems(int type) {
    switch (type) { case GREF_F32: case GREF_F16: return 1; case GREF_Q8_0: case GREF_Q4_0: case GREF_Q5_0: return QK8_0; default: return QK_K; }
}
int64_t gref_block_bytes(int type) {
    switch (type) {
        case GREF_F32: return 4; case GREF_F16: return 2; case GREF_Q8_0: return 34; case GREF_Q4_0: return 18; case GREF_Q5_0: return 22;
        case GREF_Q4_K: return 144; case GREF_Q5_K: return 176; case GREF_Q6_K: return 210; default: return -1;
    }
}
int64_t gref_row_bytes(int type, int64_t k) { return k / gref_block_elems(type) * gref_block_bytes(type); }

/* ------------------------------------------------------------------ dequantisation
 * [UPSTREAM: ggml-quants.c dequantize_row_*; layouts per gguf/quants.py as cited in the header] */

/* the 6-bit (scale, min) pair j of a Q4_K/Q5_K super-block, gguf/quants.py:478-502 */
static inline void k4_scale_min(int j, const uint8_t *s, uint8_t *sc, uint8_t *mn) {
    if (j < 4) { *sc = s[j] & 63; *mn = s[j + 4] & 63; }
    else {
        *sc = (uint8_t)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4));
        *mn = (uint8_t)((s[j + 4] >> 4) | ((s[j] >> 6) << 4));
    }
}

/* legacy 32-element blocks [UPSTREAM: ggml-quants.c dequantize_row_q4_0 / q5_0; gguf/quants.py Q4_0 / Q5_0]: element j of the
 * block is the low nibble of qs[j], element j + 16 the high nibble; Q5_0 adds bit j (resp. j + 16) of the 32-bit qh as bit 4 */
static inline int q4_0_code(const blk_q4_0 *b, int j) { return (j < 16 ? (b->qs[j] & 0x0F) : (b->qs[j - 16] >> 4)) - 8; }
static inline int q5_0_code(const blk_q5_0 *b, int j) {
    uint32_t qh;
    memcpy(&qh, b->qh, 4);
    const int lo = j < 16 ? (b->qs[j] & 0x0F) : (b->qs[j - 16] >> 4);
    return (lo | (int)(((qh >> j) & 1u) << 4)) - 16;
}
void gref_dequantize_row_q4_0(const void *vx, float *y, int64_t k) {
    const blk_q4_0 *x = (const blk_q4_0 *)vx;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        const float d = gref_fp16_to_fp32(x[b].d);
        for (int j = 0; j < QK8_0; j++) y[b * QK8_0 + j] = q4_0_code(&x[b], j) * d;
    }
}
void gref_dequantize_row_q5_0(const void *vx, float *y, int64_t k) {
    const blk_q5_0 *x = (const blk_q5_0 *)vx;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        const float d = gref_fp16_to_fp32(x[b].d);
        for (int j = 0; j < QK8_0; j++) y[b * QK8_0 + j] = q5_0_code(&x[b], j) * d;
    }
}

void gref_dequantize_row_q8_0(const void *vx, float *y, int64_t k) {
    const blk_q8_0 *x = (const blk_q8_0 *)vx;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        const float d = gref_fp16_to_fp32(x[b].d);
        for (int j = 0; j < QK8_0; j++) y[b * QK8_0 + j] = x[b].qs[j] * d;
    }
}

void gref_dequantize_row_q4_K(const void *vx, float *y, int64_t k) {
    const blk_q4_K *x = (const blk_q4_K *)vx;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(x[b].d), dmin = gref_fp16_to_fp32(x[b].dmin);
        const uint8_t *q = x[b].qs;
        float *out = y + b * QK_K;
        for (int pair = 0; pair < 4; pair++) { /* 32 bytes -> sub-block 2p in low nibbles, 2p+1 in high */
            uint8_t sc, mn;
            k4_scale_min(2 * pair, x[b].scales, &sc, &mn);
            const float d1 = d * sc, m1 = dmin * mn;
            k4_scale_min(2 * pair + 1, x[b].scales, &sc, &mn);
            const float d2 = d * sc, m2 = dmin * mn;
            for (int l = 0; l < 32; l++) out[l] = d1 * (q[l] & 0xF) - m1;
            for (int l = 0; l < 32; l++) out[32 + l] = d2 * (q[l] >> 4) - m2;
            q += 32; out += 64;
        }
    }
}

void gref_dequantize_row_q5_K(const void *vx, float *y, int64_t k) {
    const blk_q5_K *x = (const blk_q5_K *)vx;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(x[b].d), dmin = gref_fp16_to_fp32(x[b].dmin);
        const uint8_t *ql = x[b].qs, *qh = x[b].qh;
        float *out = y + b * QK_K;
        for (int pair = 0; pair < 4; pair++) {
            uint8_t sc, mn;
            k4_scale_min(2 * pair, x[b].scales, &sc, &mn);
            const float d1 = d * sc, m1 = dmin * mn;
            k4_scale_min(2 * pair + 1, x[b].scales, &sc, &mn);
            const float d2 = d * sc, m2 = dmin * mn;
            const int b1 = 2 * pair, b2 = 2 * pair + 1; /* bit of qh[l] carrying the 5th bit */
            for (int l = 0; l < 32; l++) out[l] = d1 * ((ql[l] & 0xF) + (((qh[l] >> b1) & 1) << 4)) - m1;
            for (int l = 0; l < 32; l++) out[32 + l] = d2 * ((ql[l] >> 4) + (((qh[l] >> b2) & 1) << 4)) - m2;
            ql += 32; out += 64;
        }
    }
}

void gref_dequantize_row_q6_K(const void *vx, float *y, int64_t k) {
    const blk_q6_K *x = (const blk_q6_K *)vx;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(x[b].d);
        const uint8_t *ql = x[b].ql, *qh = x[b].qh;
        const int8_t *sc = x[b].scales;
        float *out = y + b * QK_K;
        for (int half = 0; half < 2; half++) {
            for (int l = 0; l < 32; l++) {
                const int is = l / 16;
                const int8_t q1 = (int8_t)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                const int8_t q3 = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                out[l] = d * sc[is + 0] * q1;
                out[l + 32] = d * sc[is + 2] * q2;
                out[l + 64] = d * sc[is + 4] * q3;
                out[l + 96] = d * sc[is + 6] * q4;
            }
            out += 128; ql += 64; qh += 32; sc += 8;
        }
    }
}

int gref_dequantize_row(int type, const void *x, float *y, int64_t k) {
    switch (type) {
        case GREF_F32: memcpy(y, x, (size_t)k * 4); return 0;
        case GREF_F16: gref_fp16_to_fp32_row((const uint16_t *)x, y, k); return 0;
        case GREF_Q8_0: gref_dequantize_row_q8_0(x, y, k); return 0;
        case GREF_Q4_0: gref_dequantize_row_q4_0(x, y, k); return 0;
        case GREF_Q5_0: gref_dequantize_row_q5_0(x, y, k); return 0;
        case GREF_Q4_K: gref_dequantize_row_q4_K(x, y, k); return 0;
        case GREF_Q5_K: gref_dequantize_row_q5_K(x, y, k); return 0;
        case GREF_Q6_K: gref_dequantize_row_q6_K(x, y, k); return 0;
        default: return -1;
    }
}

/* ------------------------------------------------------------------ activation quantisation
 * [UPSTREAM: ggml-quants.c quantize_row_q8_K_ref / quantize_row_q8_0_ref] */

/* round-half-to-even through the 1.5*2^23 magic constant, as upstream's nearest_int() */
static inline int nearest_int(float f) {
    float v = f + 12582912.f;
    int i;
    memcpy(&i, &v, 4);
    return (i & 0x007fffff) - 0x00400000;
}

void gref_quantize_row_q8_K(const float *x, void *vy, int64_t k) {
    blk_q8_K *y = (blk_q8_K *)vy;
    for (int64_t b = 0; b < k / QK_K; b++, x += QK_K) {
        float vmax = 0.f, amax = 0.f; /* signed value of the FIRST element of largest magnitude */
        for (int j = 0; j < QK_K; j++) {
            const float ax = fabsf(x[j]);
            if (ax > amax) { amax = ax; vmax = x[j]; }
        }
        if (amax == 0.f) { y[b].d = 0.f; memset(y[b].qs, 0, QK_K); memset(y[b].bsums, 0, sizeof y[b].bsums); continue; }
        const float iscale = -127.f / vmax;
        for (int j = 0; j < QK_K; j++) {
            int v = nearest_int(iscale * x[j]);
            y[b].qs[j] = (int8_t)(v > 127 ? 127 : v);
        }
        for (int g = 0; g < QK_K / 16; g++) {
            int s = 0;
            for (int j = 0; j < 16; j++) s += y[b].qs[g * 16 + j];
            y[b].bsums[g] = (int16_t)s;
        }
        y[b].d = 1.f / iscale;
    }
}

void gref_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    blk_q8_0 *y = (blk_q8_0 *)vy;
    for (int64_t b = 0; b < k / QK8_0; b++, x += QK8_0) {
        float amax = 0.f;
        for (int j = 0; j < QK8_0; j++) { const float ax = fabsf(x[j]); if (ax > amax) amax = ax; }
        const float d = amax / 127.f;
        const float id = d ? 1.f / d : 0.f;
        y[b].d = gref_fp32_to_fp16(d);
        for (int j = 0; j < QK8_0; j++) y[b].qs[j] = (int8_t)roundf(x[j] * id);
    }
}

/* ------------------------------------------------------------------ integer dot products
 * [UPSTREAM: ggml-cpu/quants.c ggml_vec_dot_{q4_K,q5_K,q6_K}_q8_K, ggml_vec_dot_q8_0_q8_0, generic C paths]
 * Float accumulation follows the generic path: eight f32 lanes `lane[l] += d * isum[l]` per super-block,
 * the min term subtracted from a scalar, lanes folded in index order at the end. */

float gref_vec_dot_q4_K_q8_K(int64_t k, const void *vw, const void *va) {
    const blk_q4_K *w = (const blk_q4_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    float lane[8] = {0}, sumf = 0.f;
    int8_t u[QK_K];
    for (int64_t b = 0; b < k / QK_K; b++) {
        const uint8_t *q = w[b].qs;
        for (int p = 0; p < 4; p++, q += 32) {
            for (int l = 0; l < 32; l++) { u[64 * p + l] = (int8_t)(q[l] & 0xF); u[64 * p + 32 + l] = (int8_t)(q[l] >> 4); }
        }
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) k4_scale_min(j, w[b].scales, &sc[j], &mn[j]);
        int32_t msum = 0;
        for (int g = 0; g < QK_K / 16; g++) msum += a[b].bsums[g] * mn[g / 2];
        int32_t isum[8] = {0};
        for (int j = 0; j < QK_K / 32; j++)
            for (int i = 0; i < 32; i++) isum[i & 7] += (int32_t)sc[j] * (int16_t)(a[b].qs[32 * j + i] * u[32 * j + i]);
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d;
        for (int l = 0; l < 8; l++) lane[l] += d * isum[l];
        const float dmin = gref_fp16_to_fp32(w[b].dmin) * a[b].d;
        sumf -= dmin * msum;
    }
    for (int l = 0; l < 8; l++) sumf += lane[l];
    return sumf;
}

float gref_vec_dot_q5_K_q8_K(int64_t k, const void *vw, const void *va) {
    const blk_q5_K *w = (const blk_q5_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    float lane[8] = {0}, sumf = 0.f;
    int8_t u[QK_K];
    for (int64_t b = 0; b < k / QK_K; b++) {
        const uint8_t *q = w[b].qs, *qh = w[b].qh;
        for (int p = 0; p < 4; p++, q += 32) {
            for (int l = 0; l < 32; l++) {
                u[64 * p + l] = (int8_t)((q[l] & 0xF) + (((qh[l] >> (2 * p)) & 1) << 4));
                u[64 * p + 32 + l] = (int8_t)((q[l] >> 4) + (((qh[l] >> (2 * p + 1)) & 1) << 4));
            }
        }
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) k4_scale_min(j, w[b].scales, &sc[j], &mn[j]);
        int32_t msum = 0;
        for (int g = 0; g < QK_K / 16; g++) msum += a[b].bsums[g] * mn[g / 2];
        int32_t isum[8] = {0};
        for (int j = 0; j < QK_K / 32; j++)
            for (int i = 0; i < 32; i++) isum[i & 7] += (int32_t)sc[j] * (int16_t)(a[b].qs[32 * j + i] * u[32 * j + i]);
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d;
        for (int l = 0; l < 8; l++) lane[l] += d * isum[l];
        const float dmin = gref_fp16_to_fp32(w[b].dmin) * a[b].d;
        sumf -= dmin * msum;
    }
    for (int l = 0; l < 8; l++) sumf += lane[l];
    return sumf;
}

float gref_vec_dot_q6_K_q8_K(int64_t k, const void *vw, const void *va) {
    const blk_q6_K *w = (const blk_q6_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    float lane[8] = {0}, sumf = 0.f;
    int8_t u[QK_K];
    for (int64_t b = 0; b < k / QK_K; b++) {
        const uint8_t *ql = w[b].ql, *qh = w[b].qh;
        for (int half = 0; half < 2; half++, ql += 64, qh += 32) {
            int8_t *o = u + 128 * half;
            for (int l = 0; l < 32; l++) {
                o[l] = (int8_t)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                o[l + 32] = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                o[l + 64] = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                o[l + 96] = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
            }
        }
        int32_t isum[8] = {0};
        for (int g = 0; g < QK_K / 16; g++) {
            const int32_t s = w[b].scales[g];
            for (int i = 0; i < 16; i++) isum[i & 7] += s * (int16_t)(a[b].qs[16 * g + i] * u[16 * g + i]);
        }
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d;
        for (int l = 0; l < 8; l++) lane[l] += d * isum[l];
    }
    for (int l = 0; l < 8; l++) sumf += lane[l];
    return sumf;
}

float gref_vec_dot_q8_0_q8_0(int64_t k, const void *vw, const void *va) {
    const blk_q8_0 *w = (const blk_q8_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    float sumf = 0.f;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += w[b].qs[j] * a[b].qs[j];
        sumf += s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d));
    }
    return sumf;
}

/* [UPSTREAM: ggml-cpu/quants.c ggml_vec_dot_q4_0_q8_0 / q5_0_q8_0, generic form]: integer dot of the 32 codes (offset removed)
 * with the Q8_0 activation block, one f32 term per block */
float gref_vec_dot_q4_0_q8_0(int64_t k, const void *vw, const void *va) {
    const blk_q4_0 *w = (const blk_q4_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    float sumf = 0.f;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += q4_0_code(&w[b], j) * a[b].qs[j];
        sumf += s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d));
    }
    return sumf;
}
float gref_vec_dot_q5_0_q8_0(int64_t k, const void *vw, const void *va) {
    const blk_q5_0 *w = (const blk_q5_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    float sumf = 0.f;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += q5_0_code(&w[b], j) * a[b].qs[j];
        sumf += s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d));
    }
    return sumf;
}

/* bytes of the quantised-activation row that pairs with weight type `type` */
int64_t gref_act_row_bytes(int type, int64_t k) {
    if (type == GREF_Q8_0 || type == GREF_Q4_0 || type == GREF_Q5_0) return k / QK8_0 * (int64_t)sizeof(blk_q8_0);
    if (type == GREF_Q4_K || type == GREF_Q5_K || type == GREF_Q6_K) return k / QK_K * (int64_t)sizeof(blk_q8_K);
    return k * 4;
}

int gref_quantize_act(int wtype, const float *x, void *out, int64_t k) {
    switch (wtype) {
        case GREF_Q8_0: case GREF_Q4_0: case GREF_Q5_0: gref_quantize_row_q8_0(x, out, k); return 0;
        case GREF_Q4_K: case GREF_Q5_K: case GREF_Q6_K: gref_quantize_row_q8_K(x, out, k); return 0;
        case GREF_F32: case GREF_F16: memcpy(out, x, (size_t)k * 4); return 0;
        default: return -1;
    }
}

static float dot_f32_w(int type, int64_t k, const void *w, const void *a) {
    /* F32/F16 weights (norm vectors are never matmul operands; kept for completeness): f32 products, double sum */
    const float *x = (const float *)a;
    double s = 0.0;
    if (type == GREF_F32) { const float *ww = (const float *)w; for (int64_t i = 0; i < k; i++) s += (double)(ww[i] * x[i]); }
    else { const uint16_t *ww = (const uint16_t *)w; for (int64_t i = 0; i < k; i++) s += (double)(gref_fp16_to_fp32(ww[i]) * x[i]); }
    return (float)s;
}

/* Y[m][rows] = W[rows x k] . X[m][k]  -- the CPU mul_mat: quantise each activation row once, then
 * one integer vec_dot per (row, column).  OpenMP over output rows. nthreads<=0 -> all cores. */
int gref_matmul(int type, const void *W, int64_t rows, int64_t k, const float *X, int64_t m, float *Y, int nthreads) {
    const int64_t wrow = gref_row_bytes(type, k);
    const int64_t arow = gref_act_row_bytes(type, k);
    if (wrow <= 0) return -1;
    uint8_t *act = (uint8_t *)malloc((size_t)(arow * m));
    if (!act) return -2;
    for (int64_t j = 0; j < m; j++)
        if (gref_quantize_act(type, X + j * k, act + j * arow, k)) { free(act); return -1; }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const uint8_t *wr = (const uint8_t *)W + r * wrow;
        for (int64_t j = 0; j < m; j++) {
            const void *ar = act + j * arow;
            float v;
            switch (type) {
                case GREF_Q4_K: v = gref_vec_dot_q4_K_q8_K(k, wr, ar); break;
                case GREF_Q5_K: v = gref_vec_dot_q5_K_q8_K(k, wr, ar); break;
                case GREF_Q6_K: v = gref_vec_dot_q6_K_q8_K(k, wr, ar); break;
                case GREF_Q8_0: v = gref_vec_dot_q8_0_q8_0(k, wr, ar); break;
                case GREF_Q4_0: v = gref_vec_dot_q4_0_q8_0(k, wr, ar); break;
                case GREF_Q5_0: v = gref_vec_dot_q5_0_q8_0(k, wr, ar); break;
                default: v = dot_f32_w(type, k, wr, ar); break;
            }
            Y[j * rows + r] = v;
        }
    }
    free(act);
    return 0;
}

int gref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ block glue
 * [UPSTREAM: ggml-cpu/ops.cpp rms_norm / rope / soft_max, ggml-cpu/vec.h silu] */

/* y = x * (1/sqrt(mean(x^2)+eps)) [* w] ; squares in f32, sum in double */
void gref_rms_norm(const float *x, const float *w, float *y, int64_t n, float eps) {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += (double)(x[i] * x[i]);
    const float mean = (float)(s / (double)n);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = 0; i < n; i++) { float v = x[i] * scale; y[i] = w ? v * w[i] : v; }
}

/* cos/sin table of one position: theta_0 = pos, theta_{i+1} = theta_i * base^(-2/n_rot), all in f32;
 * freq_factors (rope_freqs.weight) divide theta when present. out = [n_rot/2][2] (cos, sin). */
void gref_rope_table(int32_t pos, int n_rot, float freq_base, const float *freq_factors, float *out) {
    const float theta_scale = powf(freq_base, -2.0f / (float)n_rot);
    float theta = (float)pos;
    for (int i = 0; i < n_rot / 2; i++) {
        const float ff = freq_factors ? freq_factors[i] : 1.0f;
        const float t = theta / ff;
        out[2 * i] = cosf(t);
        out[2 * i + 1] = sinf(t);
        theta *= theta_scale;
    }
}

/* "NORM" mode: rotate adjacent pairs (x[2i], x[2i+1]) of the first n_rot dims of every head, in place */
void gref_rope_norm(float *x, int n_heads, int head_dim, int n_rot, int32_t pos, float freq_base, const float *freq_factors) {
    float *tab = (float *)malloc(sizeof(float) * (size_t)n_rot);
    gref_rope_table(pos, n_rot, freq_base, freq_factors, tab);
    for (int h = 0; h < n_heads; h++) {
        float *v = x + (int64_t)h * head_dim;
        for (int i = 0; i < n_rot / 2; i++) {
            const float c = tab[2 * i], s = tab[2 * i + 1];
            const float x0 = v[2 * i], x1 = v[2 * i + 1];
            v[2 * i] = x0 * c - x1 * s;
            v[2 * i + 1] = x0 * s + x1 * c;
        }
    }
    free(tab);
}

/* out[i] = silu(g[i]) * u[i],  silu(x) = x / (1 + exp(-x)) */
void gref_swiglu(const float *g, const float *u, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = (g[i] / (1.0f + expf(-g[i]))) * u[i];
}

/* One query token against an f16 KV cache (GQA).  q: [n_head][hd] f32 (rounded to f16 first, as the CPU
 * path converts the K.Q operand to the cache type); kc/vc: [n_pos][n_kv][hd] f16 with element stride
 * kv_stride between positions; scores = q.k/sqrt(hd) ; softmax in f32 with a double sum; out = P.V with
 * f32 accumulation.  (Upstream's CPU flash-attention keeps f16 V accumulators; f32 here is the tighter
 * statement -- the difference is inside the 1e-2 logit tolerance. [UPSTREAM-MEM]) */
void gref_attn_decode(const float *q, const uint16_t *kc, const uint16_t *vc, float *out,
                      int n_head, int n_kv, int hd, int n_pos, int64_t kv_stride) {
    const int group = n_head / n_kv;
    const float scale = 1.0f / sqrtf((float)hd);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_head; h++) {
        const int kvh = h / group;
        float *sc = (float *)malloc(sizeof(float) * (size_t)n_pos);
        float *qh = (float *)malloc(sizeof(float) * (size_t)hd);
        for (int i = 0; i < hd; i++) qh[i] = gref_fp16_to_fp32(gref_fp32_to_fp16(q[(int64_t)h * hd + i]));
        float mx = -INFINITY;
        for (int p = 0; p < n_pos; p++) {
            const uint16_t *kr = kc + p * kv_stride + (int64_t)kvh * hd;
            double s = 0.0;
            for (int i = 0; i < hd; i++) s += (double)(gref_fp16_to_fp32(kr[i]) * qh[i]);
            sc[p] = (float)s * scale;
            if (sc[p] > mx) mx = sc[p];
        }
        double sum = 0.0;
        for (int p = 0; p < n_pos; p++) { sc[p] = expf(sc[p] - mx); sum += (double)sc[p]; }
        const float inv = (float)(1.0 / sum);
        float *o = out + (int64_t)h * hd;
        for (int i = 0; i < hd; i++) o[i] = 0.f;
        for (int p = 0; p < n_pos; p++) {
            const uint16_t *vr = vc + p * kv_stride + (int64_t)kvh * hd;
            const float pw = sc[p] * inv;
            for (int i = 0; i < hd; i++) o[i] += pw * gref_fp16_to_fp32(vr[i]);
        }
        free(sc); free(qh);
    }
}

/* first index of the maximum (greedy sampler) */
int64_t gref_argmax(const float *x, int64_t n) {
    int64_t best = 0;
    for (int64_t i = 1; i < n; i++) if (x[i] > x[best]) best = i;
    return best;
}

/* ==================================================================================================
 * ORDER-INDEPENDENT ("canon") VARIANTS
 *
 * ggml's f32 accumulation order is an implementation detail that differs between its own generic, AVX2,
 * AVX-512 and NEON kernels, so no single order is "the" reference.  The functions above restate the generic
 * order.  The functions below compute THE SAME integer dot products and the same per-unit f32 terms but add
 * them in f64, which makes the result independent of summation order (up to a 2^-53 effect that survives the
 * final rounding to f32 with probability ~1e-8 per output).  The CUDA kernels implement exactly this
 * definition, so GPU and oracle agree bit-for-bit and greedy-token parity is decided by logic, not by
 * rounding luck.  tests/ additionally bound canon against the generic order (<= 1e-5 relative per matvec).
 *
 * exp(): libm and CUDA expf differ by ulps, so canon uses gref_exp_ref(), a fixed sequence of IEEE
 * operations (rint, fma, multiply by a power of two) that both sides evaluate identically.
 * ================================================================================================== */

float gref_exp_ref(float x) {
    if (x < -103.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860682030941723212e-6f, r);
    float p = 1.0f / 5040.0f;
    p = fmaf(p, r, 1.0f / 720.0f);
    p = fmaf(p, r, 1.0f / 120.0f);
    p = fmaf(p, r, 1.0f / 24.0f);
    p = fmaf(p, r, 1.0f / 6.0f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ni = (int)n;
    if (ni < -126) { p *= 5.42101086242752217e-20f; /* 2^-64, exact */ ni += 64; }
    uint32_t bits = (uint32_t)(ni + 127) << 23;
    float s;
    memcpy(&s, &bits, 4);
    return p * s;
}

/* per "unit" (64 consecutive elements: a pair of 32-element sub-blocks) one f32 term, summed in f64 */
static double vec_dot_q4_K_q8_K_f64(int64_t k, const void *vw, const void *va) {
    const blk_q4_K *w = (const blk_q4_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d, dmin = gref_fp16_to_fp32(w[b].dmin) * a[b].d;
        for (int g = 0; g < 4; g++) {
            const uint8_t *q = w[b].qs + 32 * g;
            const int8_t *x = a[b].qs + 64 * g;
            int32_t dlo = 0, dhi = 0;
            for (int l = 0; l < 32; l++) { dlo += (q[l] & 0xF) * x[l]; dhi += (q[l] >> 4) * x[32 + l]; }
            uint8_t s0, m0, s1, m1;
            k4_scale_min(2 * g, w[b].scales, &s0, &m0);
            k4_scale_min(2 * g + 1, w[b].scales, &s1, &m1);
            const int32_t isum = s0 * dlo + s1 * dhi;
            const int32_t msum = m0 * (a[b].bsums[4 * g] + a[b].bsums[4 * g + 1]) + m1 * (a[b].bsums[4 * g + 2] + a[b].bsums[4 * g + 3]);
            const float term = d * (float)isum - dmin * (float)msum; /* -ffp-contract=off: mul, mul, sub */
            acc += (double)term;
        }
    }
    return acc;
}

static double vec_dot_q5_K_q8_K_f64(int64_t k, const void *vw, const void *va) {
    const blk_q5_K *w = (const blk_q5_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d, dmin = gref_fp16_to_fp32(w[b].dmin) * a[b].d;
        for (int g = 0; g < 4; g++) {
            const uint8_t *q = w[b].qs + 32 * g, *qh = w[b].qh;
            const int8_t *x = a[b].qs + 64 * g;
            int32_t dlo = 0, dhi = 0;
            for (int l = 0; l < 32; l++) {
                dlo += ((q[l] & 0xF) + (((qh[l] >> (2 * g)) & 1) << 4)) * x[l];
                dhi += ((q[l] >> 4) + (((qh[l] >> (2 * g + 1)) & 1) << 4)) * x[32 + l];
            }
            uint8_t s0, m0, s1, m1;
            k4_scale_min(2 * g, w[b].scales, &s0, &m0);
            k4_scale_min(2 * g + 1, w[b].scales, &s1, &m1);
            const int32_t isum = s0 * dlo + s1 * dhi;
            const int32_t msum = m0 * (a[b].bsums[4 * g] + a[b].bsums[4 * g + 1]) + m1 * (a[b].bsums[4 * g + 2] + a[b].bsums[4 * g + 3]);
            const float term = d * (float)isum - dmin * (float)msum;
            acc += (double)term;
        }
    }
    return acc;
}

/* unit = (half n, column t): elements 128n + 32r + 16t + (0..15), r = 0..3 */
static double vec_dot_q6_K_q8_K_f64(int64_t k, const void *vw, const void *va) {
    const blk_q6_K *w = (const blk_q6_K *)vw;
    const blk_q8_K *a = (const blk_q8_K *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK_K; b++) {
        const float d = gref_fp16_to_fp32(w[b].d) * a[b].d;
        for (int n = 0; n < 2; n++)
            for (int t = 0; t < 2; t++) {
                const uint8_t *ql = w[b].ql + 64 * n, *qh = w[b].qh + 32 * n;
                int32_t isum = 0;
                for (int r = 0; r < 4; r++) {
                    int32_t v = 0;
                    for (int i = 0; i < 16; i++) {
                        const int l = 16 * t + i;
                        const uint8_t qb = ql[32 * (r & 1) + l];
                        const int lo = (r & 2) ? (qb >> 4) : (qb & 0xF);
                        const int q = (lo | (((qh[l] >> (2 * r)) & 3) << 4)) - 32;
                        v += q * a[b].qs[128 * n + 32 * r + l];
                    }
                    isum += (int32_t)w[b].scales[8 * n + 2 * r + t] * v;
                }
                acc += (double)(d * (float)isum);
            }
    }
    return acc;
}

static double vec_dot_q8_0_q8_0_f64(int64_t k, const void *vw, const void *va) {
    const blk_q8_0 *w = (const blk_q8_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += w[b].qs[j] * a[b].qs[j];
        acc += (double)((float)s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d)));
    }
    return acc;
}

static double vec_dot_q4_0_q8_0_f64(int64_t k, const void *vw, const void *va) {
    const blk_q4_0 *w = (const blk_q4_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += q4_0_code(&w[b], j) * a[b].qs[j];
        acc += (double)((float)s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d)));
    }
    return acc;
}
static double vec_dot_q5_0_q8_0_f64(int64_t k, const void *vw, const void *va) {
    const blk_q5_0 *w = (const blk_q5_0 *)vw;
    const blk_q8_0 *a = (const blk_q8_0 *)va;
    double acc = 0.0;
    for (int64_t b = 0; b < k / QK8_0; b++) {
        int32_t s = 0;
        for (int j = 0; j < QK8_0; j++) s += q5_0_code(&w[b], j) * a[b].qs[j];
        acc += (double)((float)s * (gref_fp16_to_fp32(w[b].d) * gref_fp16_to_fp32(a[b].d)));
    }
    return acc;
}

float gref_vec_dot_q4_K_q8_K_canon(int64_t k, const void *w, const void *a) { return (float)vec_dot_q4_K_q8_K_f64(k, w, a); }
float gref_vec_dot_q5_K_q8_K_canon(int64_t k, const void *w, const void *a) { return (float)vec_dot_q5_K_q8_K_f64(k, w, a); }
float gref_vec_dot_q6_K_q8_K_canon(int64_t k, const void *w, const void *a) { return (float)vec_dot_q6_K_q8_K_f64(k, w, a); }
float gref_vec_dot_q8_0_q8_0_canon(int64_t k, const void *w, const void *a) { return (float)vec_dot_q8_0_q8_0_f64(k, w, a); }

/* unrounded canon row sums (what tensor-parallel ranks exchange): Y[rows] f64 for one activation vector */
int gref_matvec_f64(int type, const void *W, int64_t rows, int64_t k, const float *x, double *Y, int nthreads) {
    const int64_t wrow = gref_row_bytes(type, k);
    const int64_t arow = gref_act_row_bytes(type, k);
    if (wrow <= 0 || type == GREF_F32 || type == GREF_F16) return -1;
    uint8_t *act = (uint8_t *)malloc((size_t)arow);
    if (!act) return -2;
    gref_quantize_act(type, x, act, k);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const uint8_t *wr = (const uint8_t *)W + r * wrow;
        switch (type) {
            case GREF_Q4_K: Y[r] = vec_dot_q4_K_q8_K_f64(k, wr, act); break;
            case GREF_Q5_K: Y[r] = vec_dot_q5_K_q8_K_f64(k, wr, act); break;
            case GREF_Q6_K: Y[r] = vec_dot_q6_K_q8_K_f64(k, wr, act); break;
            case GREF_Q4_0: Y[r] = vec_dot_q4_0_q8_0_f64(k, wr, act); break;
            case GREF_Q5_0: Y[r] = vec_dot_q5_0_q8_0_f64(k, wr, act); break;
            default: Y[r] = vec_dot_q8_0_q8_0_f64(k, wr, act); break;
        }
    }
    free(act);
    return 0;
}

/* mode 0 = generic ggml order, 1 = canon */
int gref_matmul_mode(int type, const void *W, int64_t rows, int64_t k, const float *X, int64_t m, float *Y, int nthreads, int mode) {
    if (!mode) return gref_matmul(type, W, rows, k, X, m, Y, nthreads);
    const int64_t wrow = gref_row_bytes(type, k);
    const int64_t arow = gref_act_row_bytes(type, k);
    if (wrow <= 0 || type == GREF_F32 || type == GREF_F16) return -1;
    uint8_t *act = (uint8_t *)malloc((size_t)(arow * m));
    if (!act) return -2;
    for (int64_t j = 0; j < m; j++) gref_quantize_act(type, X + j * k, act + j * arow, k);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        const uint8_t *wr = (const uint8_t *)W + r * wrow;
        for (int64_t j = 0; j < m; j++) {
            const void *ar = act + j * arow;
            float v;
            switch (type) {
                case GREF_Q4_K: v = gref_vec_dot_q4_K_q8_K_canon(k, wr, ar); break;
                case GREF_Q5_K: v = gref_vec_dot_q5_K_q8_K_canon(k, wr, ar); break;
                case GREF_Q6_K: v = gref_vec_dot_q6_K_q8_K_canon(k, wr, ar); break;
                case GREF_Q4_0: v = (float)vec_dot_q4_0_q8_0_f64(k, wr, ar); break;
                case GREF_Q5_0: v = (float)vec_dot_q5_0_q8_0_f64(k, wr, ar); break;
                default: v = gref_vec_dot_q8_0_q8_0_canon(k, wr, ar); break;
            }
            Y[j * rows + r] = v;
        }
    }
    free(act);
    return 0;
}

void gref_swiglu_canon(const float *g, const float *u, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = (g[i] / (1.0f + gref_exp_ref(-g[i]))) * u[i];
}

/* rope table through f64 cos/sin rounded once to f32 (what the engine's host code builds with numpy) */
void gref_rope_table_canon(int32_t pos, int n_rot, float freq_base, const float *freq_factors, float *out) {
    const float theta_scale = (float)pow((double)freq_base, (double)(-2.0f / (float)n_rot));
    float theta = (float)pos;
    for (int i = 0; i < n_rot / 2; i++) {
        const float t = freq_factors ? theta / freq_factors[i] : theta;
        out[2 * i] = (float)cos((double)t);
        out[2 * i + 1] = (float)sin((double)t);
        theta *= theta_scale;
    }
}

void gref_rope_apply(float *x, int n_heads, int head_dim, int n_rot, const float *tab) {
    for (int h = 0; h < n_heads; h++) {
        float *v = x + (int64_t)h * head_dim;
        for (int i = 0; i < n_rot / 2; i++) {
            const float c = tab[2 * i], s = tab[2 * i + 1];
            const float x0 = v[2 * i], x1 = v[2 * i + 1];
            v[2 * i] = x0 * c - x1 * s;
            v[2 * i + 1] = x0 * s + x1 * c;
        }
    }
}

/* attention, canon: s_p = (float)(sum_f64 k.q) * scale ; M = max ; e_p = exp_ref(s_p - M) ;
 * out_d = (float)( sum_f64 fl(e_p * v_pd) / sum_f64 e_p )   (normalised at the end, like flash attention) */
void gref_attn_decode_canon(const float *q, const uint16_t *kc, const uint16_t *vc, float *out,
                            int n_head, int n_kv, int hd, int n_pos, int64_t kv_stride) {
    const int group = n_head / n_kv;
    const float scale = 1.0f / sqrtf((float)hd);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_head; h++) {
        const int kvh = h / group;
        float *sc = (float *)malloc(sizeof(float) * (size_t)n_pos);
        float *qh = (float *)malloc(sizeof(float) * (size_t)hd);
        double *o = (double *)calloc((size_t)hd, sizeof(double));
        for (int i = 0; i < hd; i++) qh[i] = gref_fp16_to_fp32(gref_fp32_to_fp16(q[(int64_t)h * hd + i]));
        float mx = -INFINITY;
        for (int p = 0; p < n_pos; p++) {
            const uint16_t *kr = kc + p * kv_stride + (int64_t)kvh * hd;
            double s = 0.0;
            for (int i = 0; i < hd; i++) s += (double)(gref_fp16_to_fp32(kr[i]) * qh[i]);
            sc[p] = (float)s * scale;
            if (sc[p] > mx) mx = sc[p];
        }
        double sum = 0.0;
        for (int p = 0; p < n_pos; p++) {
            const float e = gref_exp_ref(sc[p] - mx);
            sum += (double)e;
            const uint16_t *vr = vc + p * kv_stride + (int64_t)kvh * hd;
            for (int i = 0; i < hd; i++) o[i] += (double)e * (double)gref_fp16_to_fp32(vr[i]);   /* exact product (24 x 11 bits), like K.Q */
        }
        for (int i = 0; i < hd; i++) out[(int64_t)h * hd + i] = (float)(o[i] / sum);
        free(sc); free(qh); free(o);
    }
}
