// dequant.cu -- K0: bit-exact dequantisation of canonical GGUF blocks, the load-time repack into the
// tile-SoA layout (layout.cuh), its inverse check, and the embedding-row gather.
//
// Stands in for ggml's dequantize_row_q8_0 / q4_K / q5_K / q6_K and get_rows [UPSTREAM-MEM]; the arithmetic
// order is the one gguf-py documents (gguf/quants.py:396-401, 475-522, 525-549, 552-572): scales are widened
// f16->f32, multiplied in f32, and for Q4_K/Q5_K the result is (d*sc)*q - (dmin*m) as two roundings and a
// subtraction -- written with __fmul_rn/__fsub_rn so nvcc cannot contract it into an FMA.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "layout.cuh"

// ------------------------------------------------------------------ error plumbing (shared by all .cu files)
static thread_local char g_err[512] = "";
void ggb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* ggb_last_error(void) { return g_err; }
extern "C" int ggb_abi_version(void) { return GGB_ABI_VERSION; }
extern "C" int ggb_device_info(char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    int dev = 0;
    GGB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GGB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return GGB_OK;
}

// ------------------------------------------------------------------ per-element decode of canonical blocks
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t* s, int& sc, int& mn) {
    if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
    else { sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}
__device__ __forceinline__ uint16_t ld_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// element e (0..255) of one canonical block
__device__ __forceinline__ float deq_q4k(const uint8_t* b, int e) {
    const float d = h2f(ld_u16(b)), dmin = h2f(ld_u16(b + 2));
    const int j = e >> 5, l = e & 31;
    int sc, mn;
    k4_scale_min(j, b + 4, sc, mn);
    const uint8_t byte = b[16 + 32 * (j >> 1) + l];
    const int q = (j & 1) ? (byte >> 4) : (byte & 0xF);
    return __fsub_rn(__fmul_rn(__fmul_rn(d, (float)sc), (float)q), __fmul_rn(dmin, (float)mn));
}
__device__ __forceinline__ float deq_q5k(const uint8_t* b, int e) {
    const float d = h2f(ld_u16(b)), dmin = h2f(ld_u16(b + 2));
    const int j = e >> 5, l = e & 31;
    int sc, mn;
    k4_scale_min(j, b + 4, sc, mn);
    const uint8_t byte = b[48 + 32 * (j >> 1) + l];
    const int q = ((j & 1) ? (byte >> 4) : (byte & 0xF)) + (((b[16 + l] >> j) & 1) << 4);
    return __fsub_rn(__fmul_rn(__fmul_rn(d, (float)sc), (float)q), __fmul_rn(dmin, (float)mn));
}
__device__ __forceinline__ float deq_q6k(const uint8_t* b, int e) {
    const int n = e >> 7, r = (e >> 5) & 3, l = e & 31;
    const uint8_t qlb = b[64 * n + 32 * (r & 1) + l];
    const int lo = (r & 2) ? (qlb >> 4) : (qlb & 0xF);
    const int hi = (b[128 + 32 * n + l] >> (2 * r)) & 3;
    const int q = (int)(int8_t)(lo | (hi << 4)) - 32;
    const int sc = (int8_t)b[192 + 8 * n + 2 * r + (l >> 4)];
    const float d = h2f(ld_u16(b + 208));
    return __fmul_rn(__fmul_rn(d, (float)sc), (float)q);
}
__device__ __forceinline__ float deq_q8_0(const uint8_t* b, int e) { /* b -> 34-byte block */
    return __fmul_rn((float)(int8_t)b[2 + e], h2f(ld_u16(b)));
}

__global__ void dequant_canon_kernel(int type, const uint8_t* __restrict__ w, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v;
    switch (type) {
        case GGB_TYPE_Q4_K: v = deq_q4k(w + (i >> 8) * 144, (int)(i & 255)); break;
        case GGB_TYPE_Q5_K: v = deq_q5k(w + (i >> 8) * 176, (int)(i & 255)); break;
        case GGB_TYPE_Q6_K: v = deq_q6k(w + (i >> 8) * 210, (int)(i & 255)); break;
        case GGB_TYPE_Q8_0: v = deq_q8_0(w + (i >> 5) * 34, (int)(i & 31)); break;
        case GGB_TYPE_F16: v = h2f(reinterpret_cast<const uint16_t*>(w)[i]); break;
        default: v = reinterpret_cast<const float*>(w)[i]; break;
    }
    out[i] = v;
}

static int check_type(int type, bool allow_float) {
    switch (type) {
        case GGB_TYPE_Q4_K: case GGB_TYPE_Q5_K: case GGB_TYPE_Q6_K: case GGB_TYPE_Q8_0: return 1;
        case GGB_TYPE_F32: case GGB_TYPE_F16: return allow_float;
        default: return 0;
    }
}

extern "C" int ggb_dequant(int type, const void* w, float* out, int64_t n, void* stream) {
    if (!check_type(type, true)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_dequant: unsupported tensor type %d", type);
    if (n < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_dequant: negative element count");
    if (n == 0) return GGB_OK;
    if (!w || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_dequant: null pointer");
    const int blk = (type == GGB_TYPE_Q8_0) ? 32 : ((type == GGB_TYPE_F32 || type == GGB_TYPE_F16) ? 1 : 256);
    if (n % blk) GGB_FAIL(GGB_ERR_ARG, "ggb_dequant: %lld elements is not a multiple of the block size %d", (long long)n, blk);
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    dequant_canon_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(type, (const uint8_t*)w, out, n);
    GGB_CHECK_LAUNCH("ggb_dequant");
    return GGB_OK;
}

// ------------------------------------------------------------------ repack
__global__ void repack_kernel(int type, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t rows, int64_t k,
                              int64_t canon_row, int64_t stride) {
    const int64_t pairs_per_row = stride / 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * pairs_per_row) return;
    const int64_t row = idx / pairs_per_row;
    const int64_t o = (idx - row * pairs_per_row) * 2;
    uint16_t v = 0;
    if (o < canon_row) {
        const uint8_t* srow = src + row * canon_row;
        const int64_t so = ggb_repacked_to_canon(type, k, o);
        if (so >= 0) {
            v = (uint16_t)(srow[so] | (srow[so + 1] << 8));
        } else if (so == -2) { /* Q4_K/Q5_K header: regrouped 6-bit scales/mins (layout.cuh) */
            const int sbb = ggb_sb_bytes(type);
            const int64_t tile_full = (int64_t)sbb * GGB_TILE_SB;
            const int t = (int)(o / tile_full);
            const int nsb = ggb_tile_nsb(k, t);
            const int o3 = (int)(o - t * tile_full) - (type == GGB_TYPE_Q5_K ? 40 : 32) * 4 * nsb; /* offset inside HDR */
            const int sb = o3 >> 4, i = (o3 & 15) - 4;
            const uint8_t* sc = srow + ((int64_t)t * GGB_TILE_SB + sb) * sbb + 4;
            v = (uint16_t)(ggb_hdr2_byte(sc, i) | (ggb_hdr2_byte(sc, i + 1) << 8));
        } else { /* Q5_K QHU: gather the fifth bits of 16 elements of one sub-block into two bytes */
            const int64_t tile_full = (int64_t)176 * GGB_TILE_SB;
            const int t = (int)(o / tile_full);
            const int nsb = ggb_tile_nsb(k, t);
            const int o3 = (int)(o - t * tile_full) - 32 * 4 * nsb; /* offset inside QHU */
            const int u = o3 >> 3, byte = o3 & 7;                   /* byte 0..3 -> sub-block 2g, 4..7 -> 2g+1 */
            const int sb = u >> 2, g = u & 3;
            const int bit = 2 * g + (byte >> 2);
            const uint8_t* qh = srow + ((int64_t)t * GGB_TILE_SB + sb) * 176 + 16;
            const int l0 = (byte & 3) * 8;
            uint32_t acc = 0;
            for (int l = 0; l < 16; l++) acc |= (uint32_t)((qh[l0 + l] >> bit) & 1) << l;
            v = (uint16_t)acc;
        }
    }
    *reinterpret_cast<uint16_t*>(dst + row * stride + o) = v;
}

extern "C" int64_t ggb_repacked_row_stride(int type, int64_t k) {
    if (!check_type(type, false) || k <= 0 || (k % 256)) return -1;
    return ggb_row_stride(type, k);
}

extern "C" int ggb_repack(int type, const void* canon, void* dst, int64_t rows, int64_t k, void* stream) {
    if (!check_type(type, false)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_repack: unsupported tensor type %d", type);
    if (k <= 0 || (k % 256)) GGB_FAIL(GGB_ERR_ARG, "ggb_repack: k=%lld must be a positive multiple of 256", (long long)k);
    if (rows < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_repack: negative row count");
    if (rows == 0) return GGB_OK;
    if (!canon || !dst) GGB_FAIL(GGB_ERR_ARG, "ggb_repack: null pointer");
    if (((uintptr_t)dst & 15)) GGB_FAIL(GGB_ERR_ARG, "ggb_repack: dst must be 16-byte aligned");
    const int64_t stride = ggb_row_stride(type, k), canon_row = ggb_canon_row_bytes(type, k);
    const int64_t total = rows * (stride / 2);
    const int threads = 256;
    repack_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        type, (const uint8_t*)canon, (uint8_t*)dst, rows, k, canon_row, stride);
    GGB_CHECK_LAUNCH("ggb_repack");
    return GGB_OK;
}

// ------------------------------------------------------------------ dequant of tile-SoA rows (inverse check)
__device__ float deq_repacked(int type, const uint8_t* row, int64_t k, int64_t e) {
    const int t = (int)(e / GGB_TILE_ELEMS);
    const int et = (int)(e - (int64_t)t * GGB_TILE_ELEMS);
    const int nsb = ggb_tile_nsb(k, t), U = 4 * nsb;
    const uint8_t* tile = row + (int64_t)t * ggb_sb_bytes(type) * GGB_TILE_SB;
    const int sb = et >> 8, eb = et & 255;
    if (type == GGB_TYPE_Q4_K || type == GGB_TYPE_Q5_K) {
        const int j = eb >> 5, l = eb & 31, g = j >> 1, u = 4 * sb + g;
        const uint8_t byte = tile[(l >= 16 ? 16 * U : 0) + 16 * u + (l & 15)];
        int q = (j & 1) ? (byte >> 4) : (byte & 0xF);
        const uint8_t* hdr = tile + 32 * U + (type == GGB_TYPE_Q5_K ? 8 * U : 0) + 16 * sb;
        if (type == GGB_TYPE_Q5_K) {
            const uint32_t bits = *reinterpret_cast<const uint32_t*>(tile + 32 * U + 8 * u + 4 * (j & 1));
            q += ((bits >> l) & 1) << 4;
        }
        int sc, mn;
        ggb_hdr2_scale_min(hdr, j, &sc, &mn);
        const float d = h2f(ld_u16(hdr)), dmin = h2f(ld_u16(hdr + 2));
        return __fsub_rn(__fmul_rn(__fmul_rn(d, (float)sc), (float)q), __fmul_rn(dmin, (float)mn));
    }
    if (type == GGB_TYPE_Q6_K) {
        const int n = eb >> 7, r = (eb >> 5) & 3, l = eb & 31, tt = l >> 4, u = 4 * sb + 2 * n + tt;
        const uint8_t qlb = tile[((r & 1) ? 16 * U : 0) + 16 * u + (l & 15)];
        const int lo = (r & 2) ? (qlb >> 4) : (qlb & 0xF);
        const int hi = (tile[32 * U + 16 * u + (l & 15)] >> (2 * r)) & 3;
        const int q = (int)(int8_t)(lo | (hi << 4)) - 32;
        const int sc = (int8_t)tile[48 * U + 16 * sb + 8 * n + 2 * r + tt];
        const float d = h2f(ld_u16(tile + 48 * U + 16 * nsb + 2 * sb));
        return __fmul_rn(__fmul_rn(d, (float)sc), (float)q);
    }
    /* Q8_0 */
    const int blk = et >> 5, l = et & 31, u = blk >> 1, i = ((blk & 1) << 1) | (l >> 4);
    const int q = (int8_t)tile[i * 16 * U + 16 * u + (l & 15)];
    const float d = h2f(ld_u16(tile + 64 * U + 2 * blk));
    return __fmul_rn((float)q, d);
}

__global__ void dequant_repacked_kernel(int type, const uint8_t* __restrict__ w, float* __restrict__ out, int64_t rows, int64_t k, int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * k) return;
    const int64_t row = i / k;
    out[i] = deq_repacked(type, w + row * stride, k, i - row * k);
}

extern "C" int ggb_dequant_repacked(int type, const void* w, float* out, int64_t rows, int64_t k, void* stream) {
    if (!check_type(type, false)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_dequant_repacked: unsupported tensor type %d", type);
    if (k <= 0 || (k % 256) || rows < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_dequant_repacked: bad shape rows=%lld k=%lld", (long long)rows, (long long)k);
    if (rows == 0) return GGB_OK;
    if (!w || !out) GGB_FAIL(GGB_ERR_ARG, "ggb_dequant_repacked: null pointer");
    const int64_t total = rows * k;
    dequant_repacked_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        type, (const uint8_t*)w, out, rows, k, ggb_row_stride(type, k));
    GGB_CHECK_LAUNCH("ggb_dequant_repacked");
    return GGB_OK;
}

// ------------------------------------------------------------------ embedding gather (get_rows), canonical layout
__global__ void embed_row_kernel(int type, const uint8_t* __restrict__ emb, int64_t k, int64_t row_bytes,
                                 const int32_t* __restrict__ tok, float* __restrict__ x) {
    pdl_wait();
    tok += blockIdx.y;                       /* blockIdx.y = which token of a batch (ggb_embed_rows); 0 for the single-row form */
    x += (int64_t)blockIdx.y * k;
    const uint8_t* row = emb + (int64_t)(*tok) * row_bytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
        float v;
        switch (type) {
            case GGB_TYPE_Q4_K: v = deq_q4k(row + (i >> 8) * 144, (int)(i & 255)); break;
            case GGB_TYPE_Q5_K: v = deq_q5k(row + (i >> 8) * 176, (int)(i & 255)); break;
            case GGB_TYPE_Q6_K: v = deq_q6k(row + (i >> 8) * 210, (int)(i & 255)); break;
            case GGB_TYPE_Q8_0: v = deq_q8_0(row + (i >> 5) * 34, (int)(i & 31)); break;
            case GGB_TYPE_F16: v = h2f(reinterpret_cast<const uint16_t*>(row)[i]); break;
            default: v = reinterpret_cast<const float*>(row)[i]; break;
        }
        x[i] = v;
    }
}

int64_t ggb_canon_any_row_bytes(int type, int64_t k) {
    switch (type) {
        case GGB_TYPE_F32: return k * 4;
        case GGB_TYPE_F16: return k * 2;
        case GGB_TYPE_Q8_0: return k / 32 * 34;
        default: return ggb_canon_row_bytes(type, k);
    }
}

extern "C" int ggb_embed_row(int type, const void* token_embd, int64_t k, const int32_t* tok_dev, float* x, void* stream) {
    if (!check_type(type, true)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_embed_row: unsupported tensor type %d", type);
    if (!token_embd || !tok_dev || !x || k <= 0) GGB_FAIL(GGB_ERR_ARG, "ggb_embed_row: bad argument");
    embed_row_kernel<<<(unsigned)((k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        type, (const uint8_t*)token_embd, k, ggb_canon_any_row_bytes(type, k), tok_dev, x);
    GGB_CHECK_LAUNCH("ggb_embed_row");
    return GGB_OK;
}

// the batched form (ggml get_rows with many ids): ONE launch, blockIdx.y = token (it used to be one launch per token --
// 2048 launches, 7.6 ms of a 67 ms prefill)
extern "C" int ggb_embed_rows(int type, const void* token_embd, int64_t k, const int32_t* ids_dev, int tokens, float* out, void* stream) {
    if (tokens < 0) GGB_FAIL(GGB_ERR_ARG, "ggb_embed_rows: negative token count");
    if (tokens == 0) return GGB_OK;
    if (!check_type(type, true)) GGB_FAIL(GGB_ERR_UNSUPPORTED, "ggb_embed_rows: unsupported tensor type %d", type);
    if (!token_embd || !ids_dev || !out || k <= 0) GGB_FAIL(GGB_ERR_ARG, "ggb_embed_rows: bad argument");
    for (int t0 = 0; t0 < tokens; t0 += 32768) {
        const int n = tokens - t0 < 32768 ? tokens - t0 : 32768;
        embed_row_kernel<<<dim3((unsigned)((k + 255) / 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
            type, (const uint8_t*)token_embd, k, ggb_canon_any_row_bytes(type, k), ids_dev + t0, out + (int64_t)t0 * k);
        GGB_CHECK_LAUNCH("ggb_embed_rows");
    }
    return GGB_OK;
}
