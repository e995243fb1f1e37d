/*
 * ggufb200.h -- C-ABI of libggufb200.so, the B200 (sm_100a) compute library of the GGUF decode engine.
 *
 * What this replaces.  The reference (zepfu/llama-gguf-inference) contains no native code: its compute is
 * the third-party `/app/llama-server` process that scripts/start.sh:473-480,516 spawns and that
 * scripts/gateway.py:699-804 proxies to.  The process/HTTP contract of that binary is re-implemented by the
 * python host in this repository (llama-gguf-inference_b200/server.py, cli.py); this header is the boundary
 * between that host and the hand-written CUDA kernels -- the place where, in the reference's backend,
 * llama.cpp calls into ggml / ggml-cuda [UPSTREAM-MEM].  Each entry point names the ggml operation it
 * stands in for.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types, no exceptions across the boundary;
 *   - every function returns 0 on success or a negative GGB_ERR_* code; ggb_last_error() returns the
 *     thread-local message of the last failure on the calling thread;
 *   - device pointers are caller-owned (the host allocates with torch); the library never frees them and
 *     keeps no hidden device state besides cached device attributes;
 *   - every launch takes the caller's CUDA stream (cudaStream_t passed as void*), performs no host
 *     synchronisation and no allocation, so all entry points are CUDA-graph capturable;
 *   - there is NO CPU fallback anywhere in this library.
 */
#ifndef GGUFB200_H
#define GGUFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGB_ABI_VERSION 3

/* ggml tensor type ids (gguf/constants.py:4059-4093) */
#define GGB_TYPE_F32 0
#define GGB_TYPE_F16 1
#define GGB_TYPE_Q8_0 8
#define GGB_TYPE_Q4_K 12
#define GGB_TYPE_Q5_K 13
#define GGB_TYPE_Q6_K 14

#define GGB_OK 0
#define GGB_ERR_ARG (-1)     /* bad argument (null pointer, unsupported type, misaligned or ragged size) */
#define GGB_ERR_CUDA (-2)    /* a CUDA runtime call or launch failed */
#define GGB_ERR_UNSUPPORTED (-3)

int ggb_abi_version(void);
const char* ggb_last_error(void);
/* name / SM count / compute capability / total memory of the current device */
int ggb_device_info(char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- K0: dequantisation (ggml dequantize_row_q8_0 / q4_K / q5_K / q6_K; bit-exact)
 * `w` holds n elements in CANONICAL GGUF block layout; out = n floats. */
int ggb_dequant(int type, const void* w, float* out, int64_t n, void* stream);

/* ---- load-time repack: canonical GGUF rows -> tile-SoA rows (csrc/layout.cuh).  Same byte count per row
 * up to a 16-byte round-up of the row stride.  k must be a multiple of 256. */
int64_t ggb_repacked_row_stride(int type, int64_t k);
int ggb_repack(int type, const void* canon, void* dst, int64_t rows, int64_t k, void* stream);
/* inverse check / get_rows on repacked weights: out[rows][k] floats, bit-exact with ggb_dequant */
int ggb_dequant_repacked(int type, const void* w, float* out, int64_t rows, int64_t k, void* stream);

/* ---- activation quantisation (ggml quantize_row_q8_K / quantize_row_q8_0), m rows of k floats.
 * Q8_K: qs[m][k] int8, d[m][k/256] f32, bsums[m][k/16] int16.   Q8_0: qs[m][k] int8, d[m][k/32] f16 bits. */
int ggb_quantize_q8_K(const float* x, int8_t* qs, float* d, int16_t* bsums, int64_t k, int m, void* stream);
int ggb_quantize_q8_0(const float* x, int8_t* qs, uint16_t* d, int64_t k, int m, void* stream);

/* ---- K1: fused dequant-GEMV  y = W.x  (ggml mul_mat with one activation column: quantize_row_q8_K +
 * ggml_vec_dot_q*_K_q8_K).  One launch covers up to GGB_MAX_SEG weight matrices that share the input vector.
 *
 *   prologue (how the K-vector the CTAs quantise is produced)
 *     GGB_PRO_PLAIN     x
 *     GGB_PRO_RMSNORM   rms_norm(x, eps) * norm_w                     (ggml_rms_norm + ggml_mul)
 *   epilogue
 *     GGB_EPI_STORE     y_s[r] = dot
 *     GGB_EPI_RESIDUAL  y_0[r] = residual[r] + dot                    (ggml_add; y_0 may alias residual)
 *     GGB_EPI_SWIGLU    y_0[r] = silu(dot_0[r]) * dot_1[r]            (segments 0 = gate, 1 = up)
 *     GGB_EPI_ROPE_KV   segments 0,1,2 = q,k,v: rope(q) -> y_0 (f32); rope(k) -> f16 kcache[pos]; v -> f16
 *                       vcache[pos]  (ggml_rope NORM mode + ggml_cpy into the KV cache); pos is read from
 *                       device memory so a captured graph can be replayed for every position
 *     GGB_EPI_ARGMAX    y_0[r] = dot and per-CTA (max, first index) partials for the greedy sampler
 *     GGB_EPI_PEER_F64  tensor parallel, fused with the exchange: the f64 row sums go straight into every rank's
 *                       exchange region over NVLink peer memory as {data, epoch} words (the data is its own arrival flag);
 *                       ggb_peer_reduce_residual on each rank then adds the cross-rank sum to x
 *     GGB_EPI_STORE_F64 y_0 is a double*: the unrounded f64 row sums.  Tensor-parallel ranks that hold a K-slice of
 *                       W exchange these (all-reduce in f64) so the single rounding to f32 happens after the
 *                       cross-rank sum and the result is bit-identical with the unsharded GEMV.
 */
#define GGB_MAX_SEG 3
#define GGB_PRO_PLAIN 0
#define GGB_PRO_RMSNORM 1
#define GGB_EPI_STORE 0
#define GGB_EPI_RESIDUAL 1
#define GGB_EPI_SWIGLU 2
#define GGB_EPI_ROPE_KV 3
#define GGB_EPI_ARGMAX 4
#define GGB_EPI_STORE_F64 5
#define GGB_EPI_PEER_F64 6
#define GGB_PEER_MAX 8

typedef struct ggb_gemv_seg {
    const void* w;   /* tile-SoA weights [rows][row_stride] */
    int32_t type;    /* GGB_TYPE_Q4_K / Q5_K / Q6_K / Q8_0 */
    int32_t rows;
    float* y;        /* output vector of this segment (f32) or NULL when the epilogue consumes it */
} ggb_gemv_seg;

typedef struct ggb_gemv_args {
    int32_t n_seg;
    int32_t k;                 /* shared inner dimension, multiple of 256 */
    ggb_gemv_seg seg[GGB_MAX_SEG];
    int32_t prologue;
    int32_t epilogue;
    const float* x;            /* [k] input */
    const float* norm_w;       /* [k] RMSNorm gain (GGB_PRO_RMSNORM) */
    float eps;
    int32_t use_pdl;           /* launch with programmatic stream serialisation */
    const float* residual;     /* GGB_EPI_RESIDUAL */
    /* GGB_EPI_ROPE_KV */
    const int32_t* pos_dev;    /* device scalar: position of this token */
    const float* rope_tab;     /* [n_ctx][n_rot/2][2] cos,sin */
    int32_t n_rot, head_dim;
    uint16_t* kcache;          /* [n_ctx][rows_k] f16 bits */
    uint16_t* vcache;          /* [n_ctx][rows_v] f16 bits */
    /* GGB_EPI_ARGMAX */
    float* part_val;           /* [grid] */
    int32_t* part_idx;         /* [grid] */
    int32_t grid;              /* 0 = library default (a multiple of the SM count) */
    /* GGB_EPI_PEER_F64: exchange regions of all ranks as mapped in THIS process (ggb_peer_alloc / ggb_peer_open) */
    int32_t peer_n, peer_rank;
    int64_t peer_d_cap;
    uint64_t peer_base[GGB_PEER_MAX];
    /* ask for at least this much dynamic shared memory (0 = what the launch needs): with more than half of an SM's shared
     * memory two CTAs of the launch can never share an SM, which keeps the placement even when it becomes resident while
     * another kernel's small CTAs are still running (the output projection behind the attention) */
    int32_t min_smem;
    /* 1 = every ggb_gemv launch of this model has k % 2048 == 0: Q4_K / Q6_K launches then use the kernel instance that
     * carries no partial-tile code (the kernel is instruction-fetch bound: leaner code is faster code).  0 = generic. */
    int32_t full_k_model;
} ggb_gemv_args;

int ggb_gemv(const ggb_gemv_args* args, void* stream);
/* dynamic shared memory ggb_gemv would request for these args (the host uses it to decide which adjacent launches can
 * be co-resident); negative = error code */
int64_t ggb_gemv_smem_bytes(const ggb_gemv_args* args);
/* number of CTAs ggb_gemv will launch for these args (size of part_val/part_idx) */
int ggb_gemv_grid(const ggb_gemv_args* args);

/* ---- K2: dequant-GEMM on tcgen05 / TMEM (ggml mul_mat with many activation columns: prefill, large batches)
 *   Y[tokens][y_stride >= rows] (f32) = X[tokens][k] (f16) . W[rows][k]^T,  W in tile-SoA layout, k % 128 == 0.
 * Weights are dequantised exactly (f32) and rounded to f16 inside the kernel (kind::f16 with fp16 operands: 11 significand bits); accumulation is f32 in TMEM. */
int ggb_gemm(int type, const void* w, int rows, int k, const void* x_f16, int tokens, float* y, int64_t y_stride, void* stream);
/* two weight matrices of the same format and shape (ffn_gate, ffn_up) against the same activations in one launch */
int ggb_gemm2(int type, const void* w0, const void* w1, int rows, int k, const void* x_f16, int tokens, float* y0, float* y1,
              int64_t y_stride, void* stream);
int ggb_f32_to_f16(const float* x, void* y_f16, int64_t n, void* stream);
/* y = f16(d * q): x [m][k] quantised as the CPU path quantises the activation operand (Q8_K per 256 elements; Q8_0 per 32
 * when q8_0 != 0) and dequantised again -- the activation operand of ggb_gemm that keeps it within f16 rounding of ggml's
 * integer dot (quantize_row_q8_K / q8_0 + vec_dot) */
int ggb_act_fakequant_f16(const float* x, void* y_f16, int64_t k, int m, int q8_0, void* stream);
/* the same on silu(gate) * up (ggml silu + mul + the quantisation of ffn_down's activation operand in one pass) */
int ggb_swiglu_fakequant_f16(const float* gate, const float* up, void* y_f16, int64_t k, int m, int q8_0, void* stream);
/* x [m][k] += add (optional, may be NULL), then y = the quantised-dequantised f16 of rms_norm(x) * w: ggml add + rms_norm + mul +
 * the activation quantisation in front of the next ggb_gemm, one pass per token row */
int ggb_add_rmsnorm_fakequant_f16(float* x, const float* add, const float* w, void* y_f16, int64_t k, int m, float eps, int q8_0, void* stream);

/* ---- batched glue of the prefill path (ggml get_rows / rope / cpy / flash_attn_ext / add on T tokens) */
int ggb_embed_rows(int type, const void* token_embd, int64_t k, const int32_t* ids_dev, int tokens, float* out, void* stream);
/* q [T][n_head*hd] rotated in place; k rotated, v copied, both stored as f16 at cache rows pos0..pos0+T-1 */
int ggb_rope_kv_prefill(float* q, const float* k, const float* v, int tokens, int pos0, int n_head, int n_kv, int head_dim,
                        int n_rot, const float* rope_tab, uint16_t* kcache, uint16_t* vcache, void* stream);
/* causal GQA attention of T query tokens (positions pos0..) over the f16 cache; out [T][n_head*hd] f32 */
int ggb_attn_prefill(const float* q, const uint16_t* kcache, const uint16_t* vcache, int tokens, int pos0, int n_head, int n_kv,
                     int head_dim, float* out, void* stream);
int ggb_add_f32(float* x, const float* y, int64_t n, void* stream);

/* ---- small fused glue */
/* embedding gather (ggml get_rows on canonical token_embd): x[k] = dequant(row tok) ; tok read from device */
int ggb_embed_row(int type, const void* token_embd, int64_t k, const int32_t* tok_dev, float* x, void* stream);
/* reduce the GEMV's argmax partials to the token id (first index of the maximum), append it to
 * out_tokens[*step_dev], advance *pos_dev and *step_dev, and gather the next embedding row into x */
int ggb_argmax_next(const float* part_val, const int32_t* part_idx, int n_part, int32_t* tok_dev,
                    int32_t* pos_dev, int32_t* step_dev, int32_t* out_tokens, int32_t out_cap,
                    int emb_type, const void* token_embd, int64_t k, float* x, void* stream);
/* plain ops, exported for tests and for the prefill path */
int ggb_rms_norm(const float* x, const float* w, float* y, int64_t k, int m, float eps, void* stream);
int ggb_swiglu(const float* g, const float* u, float* out, int64_t n, void* stream);
int ggb_argmax(const float* x, int64_t n, int32_t* out_idx, void* stream);

/* ---- tensor-parallel glue (one process per GPU; the collectives themselves are NCCL calls made by the host)
 * x[i] += (float)y64[i]  -- applied after the f64 all-reduce of GGB_EPI_STORE_F64 partials (ggml_add) */
int ggb_residual_add_f64(float* x, const double* y64, int64_t n, int use_pdl, void* stream);
/* vocabulary-sharded greedy sampling: reduce this rank's GEMV arg-max partials (row indices local to the shard,
 * row_offset = first vocabulary row of the shard) to ONE sortable 64-bit key (larger logit wins, then smaller index);
 * the host all-reduces the keys with MAX; unpack turns the winner into the token id and does the rest of
 * ggb_argmax_next (append to out_tokens, advance pos/step, gather the next embedding row). */
/* Exchange over peer memory (csrc/peer.cu): each rank allocates a region of ggb_peer_region_bytes(n, d_cap) bytes
 * with ggb_peer_alloc (cudaMalloc, zeroed; handle64 = its 64-byte cudaIpc handle), sends the handle to its peers,
 * and maps theirs with ggb_peer_open.  ggb_peer_reduce_residual polls until all n ranks' partials of the current
 * exchange have arrived, then x[i] += (float)(sum over ranks, in rank order, of their f64 partial i). */
int64_t ggb_peer_region_bytes(int n, int64_t d_cap);
int ggb_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int ggb_peer_open(const unsigned char* handle64, void** ptr);
int ggb_peer_close(void* ptr);
int ggb_peer_free(void* ptr);
int ggb_peer_reduce_residual(float* x, const void* own_region, int n, int64_t d, int64_t d_cap, int use_pdl, void* stream);
int ggb_argmax_pack(const float* part_val, const int32_t* part_idx, int n_part, int32_t row_offset, int64_t* key, void* stream);
int ggb_argmax_unpack_next(const int64_t* key, int32_t* tok_dev, int32_t* pos_dev, int32_t* step_dev, int32_t* out_tokens,
                           int32_t out_cap, int emb_type, const void* token_embd, int64_t k, float* x, void* stream);

/* ---- KV-cache attention, one query token (ggml flash_attn_ext / soft_max path, GQA)
 * q [n_head*hd] f32 (already rotated); caches [n_ctx][n_kv*hd] f16; attends positions 0..*pos_dev inclusive.
 * use_pdl: bit 0 = launch with programmatic stream serialisation; bit 1 = also release the NEXT launch once this one
 * has passed its dependency wait (only useful when that launch cannot land twice on an SM: ggb_gemv_args.min_smem).
 * bit 2 = the sequence is long (a few thousand positions): run the two softmax passes as launches over the whole GPU (position
 * slices x groups of four query heads; head_dim 128, 4 or 8 query heads per KV head, n_ctx > 2048) that exchange scores and
 * partial sums through ws -- same result bit for bit, faster from ~2 000 positions on, slower below.
 * ws: workspace of ggb_attn_decode_ws_bytes_ctx() bytes, 256-byte aligned (16 bytes when the long-sequence path does not apply).
 * out [n_head*hd] f32. */
size_t ggb_attn_decode_ws_bytes(int n_head, int head_dim);
size_t ggb_attn_decode_ws_bytes_ctx(int n_head, int n_kv, int head_dim, int n_ctx);
int ggb_attn_decode(const float* q, const uint16_t* kcache, const uint16_t* vcache, const int32_t* pos_dev,
                    int n_head, int n_kv, int head_dim, int n_ctx, void* ws, float* out, int use_pdl, void* stream);

/* ---- small-batch decode: several sequences ("slots"), ONE new token each, in one pass over the weights.
 * Stands in for llama-server's continuous batching (ggml mul_mat with a few activation columns); the arithmetic
 * per token is the batch-1 GEMV's, so a sequence decodes to bit-identical logits alone or inside a batch.
 *
 * ggb_act_prep: what the GEMV prologue does, once per token instead of once per CTA: optional rms_norm * norm_w,
 *   then quantisation (Q8_K, or Q8_0 when q8_0 != 0) into an "activation image" of ggb_act_image_bytes(k) bytes
 *   per token (int8 codes in the kernel's bank-swizzled order | per-16 sums | block scales).  x is [nb][k].
 * ggb_gemv_batch: y_s[b][r] = dot(W_s[r], x[b]) for nb tokens; epilogues STORE, RESIDUAL (y_0[b][r] =
 *   residual[b][r] + dot, y_0 may alias residual), SWIGLU.  Outputs are row-major [nb][rows_s].  Launches as many
 *   passes of <= 8 tokens as shared memory allows. */
int64_t ggb_act_image_bytes(int64_t k);
int ggb_act_prep(const float* x, const float* norm_w, float eps, int64_t k, int nb, int q8_0, void* act, int use_pdl, void* stream);

typedef struct ggb_gemv_batch_args {
    int32_t n_seg;
    int32_t k;
    ggb_gemv_seg seg[GGB_MAX_SEG];
    int32_t epilogue;          /* GGB_EPI_STORE / GGB_EPI_RESIDUAL / GGB_EPI_SWIGLU */
    int32_t nb;                /* tokens */
    const void* act;           /* [nb][ggb_act_image_bytes(k)] from ggb_act_prep */
    const float* residual;     /* [nb][rows_0] (GGB_EPI_RESIDUAL) */
    int32_t use_pdl;
    int32_t grid;              /* 0 = one CTA per SM */
    int32_t act_tiled;         /* 1: act is the tiled layout of ggb_act_prep_tiled (see ggb_gemv_batch_prefers_tiled) */
    int32_t reserved;
} ggb_gemv_batch_args;
int ggb_gemv_batch(const ggb_gemv_batch_args* args, void* stream);
/* Long vectors, more than 8 tokens: sixteen whole-vector images (ffn_down, K = 14336) do not fit shared memory and the launch
 * would stream the weights in two 8-token passes.  The tiled layout [K-tile of 2048][token][2448 B] lets the kernel walk its
 * rows tile by tile and stream the image slices: ONE pass, same integers, same f32 terms, same f64 sums -- bit-identical.
 * ggb_gemv_batch_prefers_tiled(args) = 1 when a launch of that shape (act ignored) should be fed this way: then fill act with
 * ggb_act_prep_tiled (K-quant activations, ggb_act_tiled_bytes(k, nb) bytes) and set act_tiled = 1. */
int ggb_gemv_batch_prefers_tiled(const ggb_gemv_batch_args* args);
int64_t ggb_act_tiled_bytes(int64_t k, int nb);
int ggb_act_prep_tiled(const float* x, const float* norm_w, float eps, int64_t k, int nb, void* act, int use_pdl, void* stream);

/* Per-token cache addressing for the batch: token b belongs to slot slot_dev[b] at position pos_dev[b]; the caches
 * of all slots live in one allocation, slot s starting slot_stride elements after slot s-1
 * (cache row = [n_kv*hd] f16).  pos_dev[b] < 0 marks an idle entry (skipped).
 * ggb_rope_kv_batch: rotate q [nb][n_head*hd] in place, rotate k and store k, v as f16 at (slot, pos)
 *   (ggml_rope NORM mode + ggml_cpy).  ggb_attn_decode_batch: the one-token attention of ggb_attn_decode for every
 *   batch entry, out [nb][n_head*hd].  ggb_argmax_rows: first index of the maximum of each row of x [nb][n]. */
int ggb_rope_kv_batch(float* q, const float* k, const float* v, int nb, const int32_t* pos_dev, const int32_t* slot_dev,
                      int64_t slot_stride, int n_head, int n_kv, int head_dim, int n_rot, const float* rope_tab,
                      uint16_t* kcache, uint16_t* vcache, void* stream);
int ggb_attn_decode_batch(const float* q, const uint16_t* kcache, const uint16_t* vcache, const int32_t* pos_dev,
                          const int32_t* slot_dev, int64_t slot_stride, int nb, int n_head, int n_kv, int head_dim,
                          int n_ctx, float* out, int use_pdl, void* stream);
int ggb_argmax_rows(const float* x, int64_t n, int nb, int32_t* out_idx, void* stream);
/* Sampler candidates (sampled requests: temperature > 0 with top-k): for every row of x [nb][n] the elements >= the k-th largest
 * value -- k of them, more on ties -- unordered, as (value, index) pairs in out_val / out_idx [nb][cap]; out_cnt[b] = how many
 * row b has (> cap: ties overflowed, read the row instead).  Exact; the host sorts them and runs top-p / min-p / temperature on
 * k numbers instead of reading back and partitioning n (llama.cpp's top-k sampler, [UPSTREAM-MEM: llama-sampling.cpp]). */
int ggb_topk_rows(const float* x, int64_t n, int nb, int k, int cap, float* out_val, int32_t* out_idx, int32_t* out_cnt, void* stream);
/* out[b][j] = x[b][idx[b][j]] (0 for an index outside 0..n-1): the raw logits of the tokens in a request's penalty window, which
 * together with the top-(k + window) candidates determine the penalised top-k exactly (penalties only touch window tokens). */
int ggb_gather_rows(const float* x, int64_t n, int nb, const int32_t* idx, int m, float* out, void* stream);
/* Tensor-parallel batch: x [nb][n] is this rank's vocabulary shard starting at global row row_offset.  ggb_argmax_rows_key
 * packs (maximum, global index) of every row into one sortable signed 64-bit key (larger value, then smaller index);
 * after a MAX all-reduce of the keys ggb_argmax_keys_unpack yields the global first-maximum index of every row. */
int ggb_argmax_rows_key(const float* x, int64_t n, int nb, int32_t row_offset, int64_t* keys, void* stream);
int ggb_argmax_keys_unpack(const int64_t* keys, int nb, int32_t* out_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GGUFB200_H */
