// Micro-benchmark: how fast can all 148 SMs read the SAME small vector (the GEMV prologue's access pattern)?
//   mode 0  LDG.128, every CTA reads the same N bytes (8 warps x 4 float4 pairs in flight, like the prologue)
//   mode 1  LDG.128, every CTA reads a PRIVATE copy (no hot lines)
//   mode 2  one cp.async.bulk (TMA) of the same N bytes per CTA into shared memory
//   mode 3  TMA, cluster of CL CTAs: each CTA fetches 1/CL of the bytes and multicasts it to all CTAs of the cluster
// Reports the in-kernel time (globaltimer, max over CTAs of end - min of start) per repetition.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ src, int n16, int priv, unsigned long long* t, float* sink) {
    const unsigned long long t0 = gtime();
    const float4* p = src + (priv ? (size_t)blockIdx.x * n16 : 0);
    float acc = 0.f;
    for (int i = threadIdx.x; i < n16; i += 256 * 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const int e = i + j * 256; v[j] = p[e < n16 ? e : i]; }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    if (acc == 123.456f) sink[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = gtime(); }
}

template <int CL>
__global__ void __launch_bounds__(256) k_tma(const uint8_t* __restrict__ src, int nbytes, unsigned long long* t, float* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const unsigned long long t0 = gtime();
    const uint32_t b = smem_u32(&bar);
    uint32_t crank = 0;
    if (CL > 1) crank = cg::this_cluster().block_rank();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL > 1) cg::this_cluster().sync();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes) : "memory");
        if (CL == 1) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm)), "l"(src), "r"(nbytes), "r"(b) : "memory");
        } else {
            const int part = nbytes / CL;
            const uint16_t mask = (uint16_t)((1u << CL) - 1);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(sm) + crank * part), "l"(src + (size_t)crank * part), "r"(part), "r"(b), "h"(mask) : "memory");
        }
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
    if (sm[threadIdx.x] == 77 && sm[nbytes - 1] == 78) sink[0] = 1.f;
    __syncthreads();
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = gtime(); }
    if (CL > 1) cg::this_cluster().sync();
}

static double span_us(const unsigned long long* h, int n) {
    unsigned long long a = ~0ull, b = 0;
    for (int i = 0; i < n; i++) { if (h[2 * i] < a) a = h[2 * i]; if (h[2 * i + 1] > b) b = h[2 * i + 1]; }
    return (b - a) / 1e3;
}
static double med_us(const unsigned long long* h, int n) {
    double s = 0; for (int i = 0; i < n; i++) s += (double)(h[2 * i + 1] - h[2 * i]); return s / n / 1e3;
}

template <int CL> static void run_tma(const uint8_t* src, int nb, unsigned long long* t, float* sink, int G) {
    cudaFuncSetAttribute(k_tma<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G / CL * CL); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = nb;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_tma<CL>, src, nb, t, sink);
}

int main() {
    const int G = 148;
    uint8_t* buf; cudaMalloc(&buf, (size_t)G * 128 * 1024); cudaMemset(buf, 1, (size_t)G * 128 * 1024);
    uint8_t* flush; cudaMalloc(&flush, 512u << 20);
    unsigned long long* t; cudaMalloc(&t, 2 * G * 8); float* sink; cudaMalloc(&sink, 4);
    unsigned long long h[2 * G];
    const int sizes[] = {5120, 16384, 18432, 57344};
    for (int nb : sizes) {
        for (int mode = 0; mode < 6; mode++) {
            double best = 1e9, bmed = 0;
            for (int rep = 0; rep < 5; rep++) {
                cudaMemset(buf, rep + 1, (size_t)G * 128 * 1024);   /* the vector was just WRITTEN (dirty in L2), like a phase output */
                cudaDeviceSynchronize();
                if (mode == 0) k_ldg<<<G, 256>>>((const float4*)buf, nb / 16, 0, t, sink);
                else if (mode == 1) k_ldg<<<G, 256>>>((const float4*)buf, nb / 16, 1, t, sink);
                else if (mode == 2) run_tma<1>(buf, nb, t, sink, G);
                else if (mode == 3) run_tma<2>(buf, nb, t, sink, G);
                else if (mode == 4) run_tma<4>(buf, nb, t, sink, G);
                else run_tma<8>(buf, nb, t, sink, G);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
                const int n = mode >= 3 ? (G / (1 << (mode - 2)) * (1 << (mode - 2))) : G;
                const double s = span_us(h, n);
                if (s < best) { best = s; bmed = med_us(h, n); }
            }
            const char* names[] = {"LDG same", "LDG private", "TMA same", "TMA mcast cl2", "TMA mcast cl4", "TMA mcast cl8"};
            printf("%6d bytes  %-14s span %6.2f us   mean per-CTA %6.2f us\n", nb, names[mode], best, bmed);
        }
    }
    return 0;
}
