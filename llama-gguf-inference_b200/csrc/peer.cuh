// peer.cuh -- layout of a rank's "exchange region" for the tensor-parallel all-reduce over NVLink peer memory, and
// the system-scope load/store primitives both sides use.  One region per rank, mapped into every peer (CUDA IPC):
//   recv   [2 parities][n ranks][d_cap] f64   partial sums written BY the peers (slot = sender's rank)
//   flags  [n] int32 (own 128-byte line)       flags[r] = epoch of the newest complete partial from rank r
//   state  [2] int32 (own 128-byte line)       local only: CTAs of the producing launch that have finished | epoch
#pragma once
#include <stdint.h>

#define GGB_PEER_MAX 8

__host__ __device__ inline size_t ggb_peer_flags_off(int n, int64_t d_cap) { return (size_t)2 * n * d_cap * sizeof(double); }
__host__ __device__ inline size_t ggb_peer_state_off(int n, int64_t d_cap) { return ggb_peer_flags_off(n, d_cap) + 128; }
__host__ __device__ inline size_t ggb_peer_region_size(int n, int64_t d_cap) { return ggb_peer_state_off(n, d_cap) + 128; }

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {   /* L2: written by a peer GPU, never cached in L1 */
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
#endif
