// peer.cuh -- layout of a rank's "exchange region" for the tensor-parallel all-reduce over NVLink peer memory, and
// the system-scope load/store primitives both sides use.  One region per rank, mapped into every peer (CUDA IPC):
//   recv   [2 parities][n ranks][d_cap] x 16 bytes   partial sums written BY the peers (slot = sender's rank).  An f64
//          travels as two 8-byte words {low half | epoch << 32} {high half | epoch << 32}: each word is written with
//          one atomic 8-byte store, so the arrival of the data IS the flag -- no fence, no separate signal
//          (the "LL" idea of collective libraries, here inside the GEMV epilogue);
//   state  [2] int32 (own 128-byte line)              local only: reduce CTAs finished | epoch of the last exchange
#pragma once
#include <stdint.h>

#define GGB_PEER_MAX 8

__host__ __device__ inline size_t ggb_peer_state_off(int n, int64_t d_cap) { return (size_t)2 * n * d_cap * 16; }
__host__ __device__ inline size_t ggb_peer_region_size(int n, int64_t d_cap) { return ggb_peer_state_off(n, d_cap) + 128; }

#ifdef __CUDACC__
__device__ __forceinline__ void st_ll_f64(void* dst16, double v, uint32_t epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)epoch << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst16), "l"((b & 0xFFFFFFFFull) | e), "l"((b >> 32) | e) : "memory");
}
// returns true once both halves carry `epoch`
__device__ __forceinline__ bool ld_ll_f64(const void* src16, uint32_t epoch, double& v) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src16) : "memory");
    if ((uint32_t)(w0 >> 32) != epoch || (uint32_t)(w1 >> 32) != epoch) return false;
    v = __longlong_as_double((long long)((w0 & 0xFFFFFFFFull) | (w1 << 32)));
    return true;
}
#endif
