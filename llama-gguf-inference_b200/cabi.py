"""ctypes binding of libggufb200.so (include/ggufb200.h).  The only way the python host reaches the GPU.

There is no fallback: if the shared library is missing or fails to load, `lib()` raises -- compute entry
points never route anywhere else.  Device pointers are plain integers (torch `Tensor.data_ptr()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GGB_LIB_PATH") or os.path.join(HERE, "libggufb200.so")   # override: experiment builds (tools/)

F32, F16, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 8, 12, 13, 14
MAX_SEG = 3
PRO_PLAIN, PRO_RMSNORM = 0, 1
EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU, EPI_ROPE_KV, EPI_ARGMAX, EPI_STORE_F64, EPI_PEER_F64 = 0, 1, 2, 3, 4, 5, 6
PEER_MAX = 8

# every symbol include/ggufb200.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "ggb_abi_version", "ggb_last_error", "ggb_device_info",
    "ggb_dequant", "ggb_repacked_row_stride", "ggb_repack", "ggb_dequant_repacked",
    "ggb_quantize_q8_K", "ggb_quantize_q8_0",
    "ggb_gemv", "ggb_gemv_grid", "ggb_gemv_smem_bytes", "ggb_gemm", "ggb_gemm2", "ggb_f32_to_f16", "ggb_act_fakequant_f16", "ggb_swiglu_fakequant_f16", "ggb_add_rmsnorm_fakequant_f16",
    "ggb_embed_row", "ggb_argmax_next", "ggb_rms_norm", "ggb_swiglu", "ggb_argmax",
    "ggb_attn_decode_ws_bytes", "ggb_attn_decode_ws_bytes_ctx", "ggb_attn_decode",
    "ggb_residual_add_f64", "ggb_argmax_pack", "ggb_argmax_unpack_next",
    "ggb_embed_rows", "ggb_rope_kv_prefill", "ggb_attn_prefill", "ggb_add_f32",
    "ggb_peer_region_bytes", "ggb_peer_alloc", "ggb_peer_open", "ggb_peer_close", "ggb_peer_free", "ggb_peer_reduce_residual",
    "ggb_act_image_bytes", "ggb_act_prep", "ggb_act_prep_tiled", "ggb_act_tiled_bytes", "ggb_gemv_batch_prefers_tiled", "ggb_gemv_batch", "ggb_rope_kv_batch", "ggb_attn_decode_batch", "ggb_argmax_rows", "ggb_topk_rows", "ggb_gather_rows", "ggb_argmax_rows_key", "ggb_argmax_keys_unpack",
]


class GGBError(RuntimeError):
    pass


class GemvSeg(C.Structure):
    _fields_ = [("w", C.c_void_p), ("type", C.c_int32), ("rows", C.c_int32), ("y", C.c_void_p)]


class GemvArgs(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int32), ("k", C.c_int32),
        ("seg", GemvSeg * MAX_SEG),
        ("prologue", C.c_int32), ("epilogue", C.c_int32),
        ("x", C.c_void_p), ("norm_w", C.c_void_p),
        ("eps", C.c_float), ("use_pdl", C.c_int32),
        ("residual", C.c_void_p),
        ("pos_dev", C.c_void_p), ("rope_tab", C.c_void_p),
        ("n_rot", C.c_int32), ("head_dim", C.c_int32),
        ("kcache", C.c_void_p), ("vcache", C.c_void_p),
        ("part_val", C.c_void_p), ("part_idx", C.c_void_p),
        ("grid", C.c_int32),
        ("peer_n", C.c_int32), ("peer_rank", C.c_int32), ("peer_d_cap", C.c_int64), ("peer_base", C.c_uint64 * PEER_MAX),
        ("min_smem", C.c_int32), ("full_k_model", C.c_int32),
    ]


class GemvBatchArgs(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int32), ("k", C.c_int32),
        ("seg", GemvSeg * MAX_SEG),
        ("epilogue", C.c_int32), ("nb", C.c_int32),
        ("act", C.c_void_p), ("residual", C.c_void_p),
        ("use_pdl", C.c_int32), ("grid", C.c_int32),
        ("act_tiled", C.c_int32), ("reserved", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GGBError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). This engine has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
    sig = {
        "ggb_abi_version": ([], i32),
        "ggb_last_error": ([], C.c_char_p),
        "ggb_device_info": ([C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)], i32),
        "ggb_dequant": ([i32, vp, vp, i64, vp], i32),
        "ggb_repacked_row_stride": ([i32, i64], i64),
        "ggb_repack": ([i32, vp, vp, i64, i64, vp], i32),
        "ggb_dequant_repacked": ([i32, vp, vp, i64, i64, vp], i32),
        "ggb_quantize_q8_K": ([vp, vp, vp, vp, i64, i32, vp], i32),
        "ggb_quantize_q8_0": ([vp, vp, vp, i64, i32, vp], i32),
        "ggb_gemv": ([C.POINTER(GemvArgs), vp], i32),
        "ggb_gemv_grid": ([C.POINTER(GemvArgs)], i32),
        "ggb_gemv_smem_bytes": ([C.POINTER(GemvArgs)], i64),
        "ggb_gemm": ([i32, vp, i32, i32, vp, i32, vp, i64, vp], i32),
        "ggb_gemm2": ([i32, vp, vp, i32, i32, vp, i32, vp, vp, i64, vp], i32),
        "ggb_f32_to_f16": ([vp, vp, i64, vp], i32),
        "ggb_act_fakequant_f16": ([vp, vp, i64, i32, i32, vp], i32),
        "ggb_swiglu_fakequant_f16": ([vp, vp, vp, i64, i32, i32, vp], i32),
        "ggb_add_rmsnorm_fakequant_f16": ([vp, vp, vp, vp, i64, i32, f32, i32, vp], i32),
        "ggb_embed_row": ([i32, vp, i64, vp, vp, vp], i32),
        "ggb_argmax_next": ([vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, i64, vp, vp], i32),
        "ggb_rms_norm": ([vp, vp, vp, i64, i32, f32, vp], i32),
        "ggb_swiglu": ([vp, vp, vp, i64, vp], i32),
        "ggb_argmax": ([vp, i64, vp, vp], i32),
        "ggb_attn_decode_ws_bytes": ([i32, i32], sz),
        "ggb_attn_decode_ws_bytes_ctx": ([i32, i32, i32, i32], sz),
        "ggb_attn_decode": ([vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp], i32),
        "ggb_residual_add_f64": ([vp, vp, i64, i32, vp], i32),
        "ggb_embed_rows": ([i32, vp, i64, vp, i32, vp, vp], i32),
        "ggb_rope_kv_prefill": ([vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
        "ggb_attn_prefill": ([vp, vp, vp, i32, i32, i32, i32, i32, vp, vp], i32),
        "ggb_add_f32": ([vp, vp, i64, vp], i32),
        "ggb_peer_region_bytes": ([i32, i64], i64),
        "ggb_peer_alloc": ([sz, C.POINTER(vp), C.c_char_p], i32),
        "ggb_peer_open": ([C.c_char_p, C.POINTER(vp)], i32),
        "ggb_peer_close": ([vp], i32),
        "ggb_peer_free": ([vp], i32),
        "ggb_peer_reduce_residual": ([vp, vp, i32, i64, i64, i32, vp], i32),
        "ggb_act_image_bytes": ([i64], i64),
        "ggb_act_prep": ([vp, vp, f32, i64, i32, i32, vp, i32, vp], i32),
        "ggb_act_prep_tiled": ([vp, vp, f32, i64, i32, vp, i32, vp], i32),
        "ggb_act_tiled_bytes": ([i64, i32], i64),
        "ggb_gemv_batch_prefers_tiled": ([C.POINTER(GemvBatchArgs)], i32),
        "ggb_gemv_batch": ([C.POINTER(GemvBatchArgs), vp], i32),
        "ggb_rope_kv_batch": ([vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp], i32),
        "ggb_attn_decode_batch": ([vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, i32, vp], i32),
        "ggb_argmax_rows": ([vp, i64, i32, vp, vp], i32),
        "ggb_topk_rows": ([vp, i64, i32, i32, i32, vp, vp, vp, vp], i32),
        "ggb_gather_rows": ([vp, i64, i32, vp, i32, vp, vp], i32),
        "ggb_argmax_rows_key": ([vp, i64, i32, i32, vp, vp], i32),
        "ggb_argmax_keys_unpack": ([vp, i32, vp, vp], i32),
        "ggb_argmax_pack": ([vp, vp, i32, i32, vp, vp], i32),
        "ggb_argmax_unpack_next": ([vp, vp, vp, vp, vp, i32, i32, vp, i64, vp, vp], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ggb_last_error()
        raise GGBError(f"{what or 'ggufb200'} failed ({rc}): {msg.decode() if msg else '?'}")


def device_info() -> dict:
    name = C.create_string_buffer(256)
    sm, maj, mnr, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t()
    check(lib().ggb_device_info(name, 256, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)), "ggb_device_info")
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}


def make_gemv_args(segs, k, x, *, prologue=PRO_PLAIN, epilogue=EPI_STORE, norm_w=0, eps=0.0, use_pdl=0, residual=0,
                   pos_dev=0, rope_tab=0, n_rot=0, head_dim=0, kcache=0, vcache=0, part_val=0, part_idx=0, grid=0,
                   peer=None, min_smem=0, full_k_model=0) -> GemvArgs:
    """segs: list of (w_ptr, type, rows, y_ptr)."""
    a = GemvArgs()
    a.n_seg = len(segs)
    a.k = k
    for i, (w, t, rows, y) in enumerate(segs):
        a.seg[i].w = w
        a.seg[i].type = t
        a.seg[i].rows = rows
        a.seg[i].y = y
    a.prologue, a.epilogue = prologue, epilogue
    a.x, a.norm_w, a.eps, a.use_pdl, a.residual = x, norm_w, eps, use_pdl, residual
    a.pos_dev, a.rope_tab, a.n_rot, a.head_dim = pos_dev, rope_tab, n_rot, head_dim
    a.kcache, a.vcache, a.part_val, a.part_idx, a.grid = kcache, vcache, part_val, part_idx, grid
    a.min_smem = min_smem
    a.full_k_model = full_k_model
    if peer is not None:    # (bases of every rank's exchange region in this process, own rank, capacity in rows)
        bases, rank, d_cap = peer
        a.peer_n, a.peer_rank, a.peer_d_cap = len(bases), rank, d_cap
        for i, b in enumerate(bases):
            a.peer_base[i] = b
    return a


def make_gemv_batch_args(segs, k, act, nb, *, epilogue=EPI_STORE, residual=0, use_pdl=0, grid=0, act_tiled=0) -> GemvBatchArgs:
    """segs: list of (w_ptr, type, rows, y_ptr); outputs are [nb][rows]."""
    a = GemvBatchArgs()
    a.n_seg, a.k = len(segs), k
    for i, (w, t, rows, y) in enumerate(segs):
        a.seg[i].w, a.seg[i].type, a.seg[i].rows, a.seg[i].y = w, t, rows, y
    a.epilogue, a.nb, a.act, a.residual, a.use_pdl, a.grid, a.act_tiled = epilogue, nb, act, residual, use_pdl, grid, act_tiled
    return a


def copy_args(a: GemvArgs, **override) -> GemvArgs:
    """a field-by-field copy of a launch description with some fields replaced"""
    b = GemvArgs()
    C.memmove(C.byref(b), C.byref(a), C.sizeof(GemvArgs))
    for k, v in override.items():
        setattr(b, k, v)
    return b
