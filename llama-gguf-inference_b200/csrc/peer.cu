// peer.cu -- the exchange step of tensor parallelism (the all-reduce after attn_output and ffn_down,
// BASELINE.json north_star) done over NVLink peer memory instead of a library collective:
//   producer  the row-split GEMV's epilogue (gemv.cu, GGB_EPI_PEER_F64) stores its f64 partial sums straight into
//             EVERY rank's exchange region (slot = own rank) as {data, epoch} words -- the transfer rides on the
//             GEMV, CTA by CTA, with no fence, no signal and no separate send;
//   consumer  peer_reduce_residual_kernel polls the n slots of element i until they carry the epoch, then
//             x[i] += (float)(sum over ranks, in rank order, of partial[r][i]) -- the same f64 sum on every rank, so
//             the replicas stay bit-identical.  Its last CTA advances the epoch.
// Two parities of the receive buffers are enough: a rank finishes the reduce of epoch e only after every peer has
// written e, and a peer writes e+1 only after it finished its own reduce of e.
// The processes exchange cudaIpc handles of their regions once at start-up (host side: model.py).
#include "common.cuh"
#include "peer.cuh"

extern "C" int64_t ggb_peer_region_bytes(int n, int64_t d_cap) {
    if (n < 1 || n > GGB_PEER_MAX || d_cap <= 0) return -1;
    return (int64_t)ggb_peer_region_size(n, d_cap);
}

extern "C" int ggb_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ptr || !handle64 || bytes == 0) GGB_FAIL(GGB_ERR_ARG, "ggb_peer_alloc: bad argument");
    void* p = nullptr;
    GGB_CUDA(cudaMalloc(&p, bytes));
    GGB_CUDA(cudaMemset(p, 0, bytes));
    GGB_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); GGB_FAIL(GGB_ERR_CUDA, "ggb_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return GGB_OK;
}
extern "C" int ggb_peer_open(const unsigned char* handle64, void** ptr) {
    if (!ptr || !handle64) GGB_FAIL(GGB_ERR_ARG, "ggb_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GGB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GGB_OK;
}
extern "C" int ggb_peer_close(void* ptr) {
    if (ptr) GGB_CUDA(cudaIpcCloseMemHandle(ptr));
    return GGB_OK;
}
extern "C" int ggb_peer_free(void* ptr) {
    if (ptr) GGB_CUDA(cudaFree(ptr));
    return GGB_OK;
}

__global__ void __launch_bounds__(256) peer_reduce_residual_kernel(float* __restrict__ x, uint64_t own_base, int n, int64_t d, int64_t d_cap) {
    pdl_wait();
    pdl_launch_dependents();   /* after the wait, like the GEMV: at most the next launch is resident meanwhile (70B tp2: 139.5 tok/s; without: 134.6) */
    uint8_t* base = reinterpret_cast<uint8_t*>(own_base);
    int* state = reinterpret_cast<int*>(base + ggb_peer_state_off(n, d_cap));
    const uint32_t e = (uint32_t)(*reinterpret_cast<volatile int*>(state + 1) + 1);   /* the exchange this rank just fed */
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) {
        const uint8_t* recv = base + (size_t)(e & 1) * n * d_cap * 16;
        double s = 0.0;
        for (int r = 0; r < n; r++) {
            const void* src = recv + ((size_t)r * d_cap + i) * 16;
            double v;
            unsigned long long t0 = 0;
            int spins = 0;
            while (!ld_ll_f64(src, e, v)) {
                if (++spins == 4096) {   /* a dead peer must not hang the GPU: give up after ~20 s and kill the context */
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                    if (!t0) t0 = t;
                    else if (t - t0 > 20000000000ull) { printf("ggufb200: peer exchange timed out waiting for rank %d (epoch %u)\n", r, e); __trap(); }
                    spins = 0;
                }
            }
            s += v;
        }
        x[i] = __fadd_rn(x[i], (float)s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(state, 1);
        if (prev == (int)gridDim.x - 1) {          /* every CTA has consumed epoch e: the next exchange may reuse e's parity after one more */
            *reinterpret_cast<volatile int*>(state) = 0;
            *reinterpret_cast<volatile int*>(state + 1) = (int)e;
        }
    }
}

extern "C" int ggb_peer_reduce_residual(float* x, const void* own_region, int n, int64_t d, int64_t d_cap, int use_pdl, void* stream) {
    if (!x || !own_region || n < 1 || n > GGB_PEER_MAX || d <= 0 || d > d_cap) GGB_FAIL(GGB_ERR_ARG, "ggb_peer_reduce_residual: bad argument");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((d + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
    GGB_CUDA(cudaLaunchKernelEx(&cfg, peer_reduce_residual_kernel, x, (uint64_t)(uintptr_t)own_region, n, d, d_cap));
    return GGB_OK;
}
