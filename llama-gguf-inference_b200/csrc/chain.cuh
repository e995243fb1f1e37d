// chain.cuh -- flag-chained activation vectors between the launches of one decoded token.
//
// A chained vector holds {f32 value, u32 tag} words instead of plain floats.  The producer's epilogue writes each word
// with ONE 8-byte store, so the arrival of the value IS its flag (the "LL" idea of collective libraries, as in peer.cuh,
// here between two launches on the same GPU); the consumer polls the words where it needs them.  No grid dependency is
// needed between the two launches: the consumer -- already resident through programmatic dependent launch, its weight
// ring full -- starts as soon as the last word has landed in L2, instead of waiting for the producer grid to exit, for
// the dependency release behind it (~0.7 us) and then for a first trip to L2.
//
//   tag    = epoch << 10 | id.  The epoch is a device counter advanced once per token by ggb_chain_tick (the first
//            launch of every token, a normal stream-ordered launch), so every launch of the token reads the same value;
//            id (1..1023) names the write within the token (the residual stream is overwritten 2L+1 times per token).
//            A word of an earlier token or an earlier write never matches.
//   order  word e of a vector sits at chain_slot(e): inside a block of 256 the 16-byte chunk (two words) index is
//            (i >> 1) * 32 + lane for element 8 * lane + i, so that a warp that owns the block -- lane l quantises
//            elements 8l..8l+7 -- reads it with four fully coalesced 512-byte requests.
//   safety a consumer launch can only begin after EVERY CTA of its producer launch has executed
//            griddepcontrol.launch_dependents (or exited), i.e. is resident and running: a polling CTA never occupies
//            resources its producer still waits for.  Every poll is bounded (trap after ~2 s), so a logic error
//            surfaces as a CUDA error, not as a hung GPU.
#pragma once
#include <stdint.h>

#include "common.cuh"

#define GGB_CHAIN_TAG_BITS 10
#define GGB_CHAIN_TAG_MAX ((1 << GGB_CHAIN_TAG_BITS) - 1)

__device__ __forceinline__ uint32_t chain_epoch_bits(const uint32_t* epoch) {
    uint32_t e;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(epoch) : "memory");
    return e << GGB_CHAIN_TAG_BITS;
}
__device__ __forceinline__ int chain_slot(int e) {
    const int j = e & 255;
    return (e & ~255) + (((((j & 7) >> 1) << 5) + (j >> 3)) << 1) + (j & 1);
}
__device__ __forceinline__ void chain_store(uint2* vec, int e, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(vec + chain_slot(e)), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 chain_ld16(const void* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint2 chain_ld8(const void* p) {
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
// bounded back-off: sleeps a little, and every 4096 spins checks the clock; after ~2 s the context is killed
struct ChainSpin {
    unsigned n = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void pause(unsigned ns) {
        __nanosleep(ns);
        if (++n == 4096) {
            n = 0;
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            else if (t - t0 > 2000000000ull) {
                printf("ggufb200: chained vector never arrived (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
};
// lane 0 waits (sleeping) until the FIRST word of a block carries the tag: while the producer is still far from done the
// poll traffic is one sector per warp, not the whole vector
__device__ __forceinline__ void chain_wait_first(const uint2* block, uint32_t want, int lane) {
    if (lane == 0) {
        ChainSpin sp;
        while (chain_ld8(block).y != want) sp.pause(64);
    }
    __syncwarp();
}
// the lane's eight elements of block b (elements 8*lane .. 8*lane+7), polled until all carry the tag
__device__ __forceinline__ void chain_load_block(const uint2* vec, int b, int lane, uint32_t want, float (&v)[8]) {
    const uint4* p = reinterpret_cast<const uint4*>(vec + b * 256) + lane;
    ChainSpin sp;
    uint4 a0, a1, a2, a3;
    for (;;) {
        a0 = chain_ld16(p); a1 = chain_ld16(p + 32); a2 = chain_ld16(p + 64); a3 = chain_ld16(p + 96);
        const bool ok = a0.y == want && a0.w == want && a1.y == want && a1.w == want && a2.y == want && a2.w == want && a3.y == want && a3.w == want;
        if (ok) break;
        sp.pause(32);
    }
    v[0] = __uint_as_float(a0.x); v[1] = __uint_as_float(a0.z); v[2] = __uint_as_float(a1.x); v[3] = __uint_as_float(a1.z);
    v[4] = __uint_as_float(a2.x); v[5] = __uint_as_float(a2.z); v[6] = __uint_as_float(a3.x); v[7] = __uint_as_float(a3.z);
}
__device__ __forceinline__ float chain_load_one(const uint2* vec, int e, uint32_t want) {
    const uint2* p = vec + chain_slot(e);
    ChainSpin sp;
    uint2 w = chain_ld8(p);
    while (w.y != want) { sp.pause(32); w = chain_ld8(p); }
    return __uint_as_float(w.x);
}
