// layout.cuh -- the HBM weight layout ("tile-SoA") that every GEMV/GEMM kernel streams.
//
// GGUF stores a weight row as consecutive AoS blocks (Q4_K 144 B, Q6_K 210 B, Q8_0 34 B per block; the last
// two are only 2-byte aligned -- gguf/quants.py:396-401, 475-522, 552-572).  At load time each row is
// re-ordered, WITHOUT changing its byte count, into K-tiles of 2048 elements (8 super-blocks; the last tile
// of a row may hold fewer) whose fields are grouped into 16-byte-aligned sections, so that a warp reads a tile
// with perfectly coalesced 128-bit loads and lane u owns "unit" u = 64 consecutive-ish elements:
//
//   Q4_K tile (144*nsb B):  Q0[u] 16 B | Q1[u] 16 B | HDR[sb] 16 B            u = 4*sb + g, U = 4*nsb
//       canonical 32-byte group g of block sb = {sub-block 2g in low nibbles, 2g+1 in high nibbles};
//       Q0[u] = its bytes 0..15, Q1[u] = bytes 16..31; HDR[sb] = {f16 d, f16 dmin, 4 x 24-bit fields}:
//       field g = sc[2g] | sc[2g+1] << 6 | min[2g] << 12 | min[2g+1] << 18 -- the same sixteen 6-bit values as the
//       canonical 12 scale bytes (gguf/quants.py:478-502), regrouped so that lane u = 4*sb + g extracts the four
//       it needs with one byte-permute and three shifts instead of ~20 bit operations.
//       unit u covers elements 64u .. 64u+63 of the tile.
//   Q6_K tile (210*nsb B):  QLA[u] | QLB[u] | QH[u] (16 B each) | SC[sb] 16 B | D[sb] 2 B
//       u = 4*sb + 2*n + t (n = 128-element half, t = 16-element column); with l = 16t..16t+15:
//       QLA[u] = ql[64n + l], QLB[u] = ql[64n + 32 + l], QH[u] = qh[32n + l];
//       elements 128n + 32r + l (r = 0..3) use scale SC[sb][8n + 2r + t].
//   Q8_0 tile (34*nb B, nb = 8*nsb):  QS0[u] | QS1[u] | QS2[u] | QS3[u] (16 B each) | D[blk] 2 B
//       u = pair of 32-element blocks (2u, 2u+1); QSi[u] = qs[block 2u + (i>>1)][16*(i&1) .. +16].
//   Q5_K tile (176*nsb B):  Q0[u] | Q1[u] (as Q4_K) | QHU[u] 8 B | HDR[sb] 16 B
//       QHU[u] = {u32 fifth bits of sub-block 2g (bit l = element l), u32 fifth bits of sub-block 2g+1}.
//
// Row stride = 16-byte round-up of the canonical row bytes (identical for every K in BASELINE.json's configs;
// e.g. TinyLlama's K=5632 Q6_K rows grow by 4 bytes).  Rooflines are always computed on canonical GGUF bytes.
#pragma once
#include <stdint.h>

#define GGB_TILE_ELEMS 2048
#define GGB_TILE_SB 8 /* super-blocks per full tile */
/* tiled activation images (ggb_act_prep_tiled): per (K-tile, token) codes 2048 | per-16 sums 256 | block scales 32, padded to a
 * token stride of 4 words mod 32 banks */
#define GGB_ACT_TILE_BS_OFF 2048
#define GGB_ACT_TILE_DSC_OFF 2304
#define GGB_ACT_TILE_STRIDE 2448

#ifdef __CUDACC__
#define GGB_HD __host__ __device__ __forceinline__
#else
#define GGB_HD static inline
#endif

GGB_HD int ggb_sb_bytes(int type) { /* canonical bytes per 256 elements */
    switch (type) {
        case GGB_TYPE_Q4_K: return 144;
        case GGB_TYPE_Q5_K: return 176;
        case GGB_TYPE_Q6_K: return 210;
        case GGB_TYPE_Q8_0: return 272;
        default: return 0;
    }
}
GGB_HD int64_t ggb_canon_row_bytes(int type, int64_t k) { return (k / 256) * ggb_sb_bytes(type); }
GGB_HD int64_t ggb_row_stride(int type, int64_t k) { return (ggb_canon_row_bytes(type, k) + 15) & ~(int64_t)15; }
GGB_HD int ggb_tiles_per_row(int64_t k) { return (int)((k + GGB_TILE_ELEMS - 1) / GGB_TILE_ELEMS); }
GGB_HD int ggb_tile_nsb(int64_t k, int t) { /* super-blocks in tile t */
    int64_t nsb = k / 256 - (int64_t)t * GGB_TILE_SB;
    return nsb > GGB_TILE_SB ? GGB_TILE_SB : (int)nsb;
}

// canonical 6-bit (scale, min) of sub-block j from the 12 packed scale bytes (gguf/quants.py:478-502)
GGB_HD void ggb_k4_scale_min(int j, const uint8_t* s, int* sc, int* mn) {
    if (j < 4) { *sc = s[j] & 63; *mn = s[j + 4] & 63; }
    else { *sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); *mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}
// byte i (0..11) of the re-encoded scale area of a tile-SoA Q4_K/Q5_K header
GGB_HD uint8_t ggb_hdr2_byte(const uint8_t* canon_scales, int i) {
    const int g = i / 3;
    int s0, m0, s1, m1;
    ggb_k4_scale_min(2 * g, canon_scales, &s0, &m0);
    ggb_k4_scale_min(2 * g + 1, canon_scales, &s1, &m1);
    const uint32_t f = (uint32_t)s0 | ((uint32_t)s1 << 6) | ((uint32_t)m0 << 12) | ((uint32_t)m1 << 18);
    return (uint8_t)(f >> (8 * (i - 3 * g)));
}
// (scale, min) of sub-block j from a tile-SoA header
GGB_HD void ggb_hdr2_scale_min(const uint8_t* hdr, int j, int* sc, int* mn) {
    const uint8_t* p = hdr + 4 + 3 * (j >> 1);
    const uint32_t f = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    *sc = (f >> (6 * (j & 1))) & 63;
    *mn = (f >> (12 + 6 * (j & 1))) & 63;
}

// byte offset inside the CANONICAL row for the byte at offset `o` inside the REPACKED row (o < canonical row
// bytes).  Used by the repack kernel (one thread per 2 bytes: every field boundary in both layouts is even).
GGB_HD int64_t ggb_repacked_to_canon(int type, int64_t k, int64_t o) {
    const int sbb = ggb_sb_bytes(type);
    const int64_t tile_full = (int64_t)sbb * GGB_TILE_SB;
    const int t = (int)(o / tile_full);
    const int ot = (int)(o - (int64_t)t * tile_full);
    const int nsb = ggb_tile_nsb(k, t);
    const int U = 4 * nsb;
    const int64_t sb0 = (int64_t)t * GGB_TILE_SB; /* first super-block of the tile */
    if (type == GGB_TYPE_Q4_K || type == GGB_TYPE_Q5_K) {
        const int qoff = (type == GGB_TYPE_Q4_K) ? 16 : 48; /* qs offset in the canonical block */
        if (ot < 32 * U) { /* Q0 | Q1 */
            const int half = ot >= 16 * U;
            const int o2 = ot - half * 16 * U;
            const int u = o2 >> 4, b = o2 & 15, sb = u >> 2, g = u & 3;
            return (sb0 + sb) * sbb + qoff + 32 * g + 16 * half + b;
        }
        int o3 = ot - 32 * U;
        if (type == GGB_TYPE_Q5_K) {
            if (o3 < 8 * U) return -1; /* QHU is bit-gathered, not a byte permutation: handled by the kernel */
            o3 -= 8 * U;
        }
        if ((o3 & 15) >= 4) return -2; /* re-encoded 6-bit scales/mins: computed by the repack kernel */
        return (sb0 + (o3 >> 4)) * sbb + (o3 & 15);
    }
    if (type == GGB_TYPE_Q6_K) {
        if (ot < 48 * U) {
            const int sec = ot / (16 * U); /* 0 QLA, 1 QLB, 2 QH */
            const int o2 = ot - sec * 16 * U;
            const int u = o2 >> 4, b = o2 & 15, sb = u >> 2, n = (u >> 1) & 1, tt = u & 1;
            const int inblk = (sec == 2) ? (128 + 32 * n + 16 * tt + b) : (64 * n + 32 * sec + 16 * tt + b);
            return (sb0 + sb) * sbb + inblk;
        }
        const int o3 = ot - 48 * U;
        if (o3 < 16 * nsb) return (sb0 + (o3 >> 4)) * sbb + 192 + (o3 & 15);
        const int o4 = o3 - 16 * nsb;
        return (sb0 + (o4 >> 1)) * sbb + 208 + (o4 & 1);
    }
    if (type == GGB_TYPE_Q8_0) {
        const int64_t blk0 = sb0 * 8; /* first 32-element block of the tile */
        if (ot < 64 * U) {
            const int i = ot / (16 * U);
            const int o2 = ot - i * 16 * U;
            const int u = o2 >> 4, b = o2 & 15;
            return (blk0 + 2 * u + (i >> 1)) * 34 + 2 + 16 * (i & 1) + b;
        }
        const int o3 = ot - 64 * U;
        return (blk0 + (o3 >> 1)) * 34 + (o3 & 1);
    }
    return -1;
}
