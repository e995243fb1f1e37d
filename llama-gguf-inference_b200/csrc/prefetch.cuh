// prefetch.cuh -- L2 prefetch of a LATER launch's weights (include/ggufb200.h: ggb_prefetch).
// One lane per participating warp issues cp.async.bulk.prefetch.L2 (SASS UBLKPF.L2) for its share of the window
// [skip, skip + bytes) of the byte stream consumer CTA `c` will read: its row slice of segment 0, then 1, then 2 --
// the same even row split the GEMV uses (rows = q * grid + r; CTA c starts at c*q + min(c, r)).
#pragma once
#include <stdint.h>

#include "../../include/ggufb200.h"

struct PfK {
    const uint8_t* w[GGB_MAX_SEG];
    int64_t row_bytes[GGB_MAX_SEG];
    int rq[GGB_MAX_SEG], rr[GGB_MAX_SEG];
    int n_seg, grid, when;
    int64_t skip, bytes;
};
struct PfSet { PfK f[GGB_PF_MAX]; };

static inline PfK make_pfk(const ggb_prefetch& p) {
    PfK k = {};
    if (p.n_seg <= 0 || p.n_seg > GGB_MAX_SEG || p.grid <= 0 || p.bytes <= 0) return k;
    k.n_seg = p.n_seg; k.grid = p.grid; k.when = p.when;
    k.skip = p.skip & ~(int64_t)15; k.bytes = p.bytes & ~(int64_t)15;
    for (int s = 0; s < p.n_seg; s++) {
        k.w[s] = (const uint8_t*)p.w[s];
        k.row_bytes[s] = p.row_bytes[s];
        k.rq[s] = p.rows[s] / p.grid; k.rr[s] = p.rows[s] % p.grid;
        if (!k.w[s] || (p.row_bytes[s] & 15)) { k.n_seg = 0; break; }
    }
    return k;
}
static inline PfSet make_pfset(const ggb_prefetch* p) {
    PfSet s = {};
    if (p) for (int i = 0; i < GGB_PF_MAX; i++) s.f[i] = make_pfk(p[i]);
    return s;
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// called by ONE lane of each of `nparts` warps (part = 0..nparts-1) of CTA c
__device__ __forceinline__ void l2_prefetch_window(const PfK& F, int c, int part, int nparts) {
    if (F.n_seg == 0 || c >= F.grid) return;
    constexpr int64_t CH = 8192;
    const int64_t lo = F.skip, hi = F.skip + F.bytes;
    int64_t base = 0;
    for (int s = 0; s < F.n_seg; s++) {
        const int a = c * F.rq[s] + min(c, F.rr[s]), b = (c + 1) * F.rq[s] + min(c + 1, F.rr[s]);
        const int64_t len = (int64_t)(b - a) * F.row_bytes[s];
        const int64_t s0 = max(lo, base) - base, s1 = min(hi, base + len) - base;
        if (s1 > s0) {
            const uint8_t* p = F.w[s] + (int64_t)a * F.row_bytes[s] + s0;
            const int64_t n = s1 - s0;
            const int64_t per = (((n + nparts - 1) / nparts) + 15) & ~(int64_t)15;
            int64_t o = (int64_t)part * per;
            const int64_t e = min(n, o + per);
            for (; o < e; o += CH) l2_prefetch_bulk(p + o, (uint32_t)min(CH, e - o));
        }
        base += len;
        if (base >= hi) break;
    }
}
__device__ __forceinline__ void l2_prefetch_set(const PfSet& S, int when, int c, int part, int nparts) {
#pragma unroll
    for (int i = 0; i < GGB_PF_MAX; i++)
        if (S.f[i].when & when) l2_prefetch_window(S.f[i], c, part, nparts);
}
