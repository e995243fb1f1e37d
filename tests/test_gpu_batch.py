"""-m gpu: the small-batch decode kernels (csrc/gemv_batch.cu, batched rope / attention / arg-max) through the C-ABI.

Bar: every token of a batch gets bit-identical results to the oracle's order-independent "canon" restatement --
i.e. exactly what the batch-1 kernels produce for that token alone (tests/test_gpu_kernels.py).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import rand_blocks

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _unswizzle(codes: np.ndarray) -> np.ndarray:
    """image order -> element order: 16-byte chunk c of the vector is stored at chunk c ^ ((c >> 2) & 7)"""
    ch = codes.reshape(-1, 16)
    c = np.arange(ch.shape[0])
    return ch[c ^ ((c >> 2) & 7)].reshape(-1)


def _prep(x: np.ndarray, k: int, norm_w=None, eps=0.0, q8_0=False):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    nb = x.shape[0]
    img = L.ggb_act_image_bytes(k)
    assert img == k + k // 4
    xd = U.to_dev(x.astype(np.float32))
    act = torch.zeros(max(nb, 1) * img, dtype=torch.uint8, device=U.DEV)
    gd = U.to_dev(norm_w) if norm_w is not None else None
    cabi.check(L.ggb_act_prep(xd.data_ptr(), gd.data_ptr() if gd is not None else 0, eps, k, nb, int(q8_0), act.data_ptr(), 0, U.stream_ptr()))
    U.sync()
    return act


@pytest.mark.parametrize("k", [256, 4096, 14336])
def test_act_prep_q8_K_bit_exact(oracle, k):
    rng = np.random.default_rng(k + 5)
    nb = 5
    x = (rng.standard_normal((nb, k)) * np.exp(rng.uniform(-3, 3, (nb, 1)))).astype(np.float32)
    x[0, :256] = 0.0
    img = _prep(x, k).cpu().numpy().reshape(nb, -1)
    for b in range(nb):
        rd, rq, rb = oracle.q8_K_fields(oracle.quantize_q8_K(x[b]))
        assert np.array_equal(_unswizzle(img[b, :k]).view(np.int8), rq.reshape(-1))
        assert np.array_equal(img[b, k:k + k // 8].copy().view(np.int16), rb.reshape(-1))
        assert np.array_equal(img[b, k + k // 8:k + k // 8 + 4 * (k // 256)].copy().view(np.uint32), _bits(rd))


def test_act_prep_rmsnorm_and_q8_0(oracle):
    k, nb, eps = 2048, 3, 1e-5
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((nb, k)) * 2).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    img = _prep(x, k, norm_w=g, eps=eps).cpu().numpy().reshape(nb, -1)
    for b in range(nb):
        h = oracle.rms_norm(x[b], g, eps)
        rd, rq, rb = oracle.q8_K_fields(oracle.quantize_q8_K(h))
        assert np.array_equal(_unswizzle(img[b, :k]).view(np.int8), rq.reshape(-1))
        assert np.array_equal(img[b, k + k // 8:k + k // 8 + 4 * (k // 256)].copy().view(np.uint32), _bits(rd))
    img = _prep(x, k, q8_0=True).cpu().numpy().reshape(nb, -1)
    for b in range(nb):
        ref = oracle.quantize_q8_0(x[b]).reshape(-1, 34)
        assert np.array_equal(_unswizzle(img[b, :k]).view(np.int8), ref[:, 2:].copy().view(np.int8).reshape(-1))
        d = ref[:, :2].copy().view(np.float16).reshape(-1).astype(np.float32)
        assert np.array_equal(img[b, k + k // 8:k + k // 4].copy().view(np.uint32), _bits(d))


def _batch_gemv(segs, k, act, nb, **kw):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    ys = [torch.zeros(max(nb * r, 1), dtype=torch.float32, device=U.DEV) for _, _, r in segs]
    a = cabi.make_gemv_batch_args([(w.data_ptr(), t, r, y.data_ptr()) for (w, t, r), y in zip(segs, ys)], k, act.data_ptr(), nb, **kw)
    cabi.check(L.ggb_gemv_batch(C.byref(a), U.stream_ptr()), "ggb_gemv_batch")
    U.sync()
    return [y.cpu().numpy()[:nb * r].reshape(nb, r) for y, (_, _, r) in zip(ys, segs)]


@pytest.mark.parametrize("name,qt", [("q8_0", 8), ("q4_k", 12), ("q5_k", 13), ("q6_k", 14)])
@pytest.mark.parametrize("rows,k,nb", [(1, 256, 1), (7, 512, 2), (300, 2048, 3), (64, 5632, 5), (1000, 4096, 8), (300, 14336, 11),
                                       (2048, 4096, 16)])
def test_gemv_batch_store_matches_oracle(oracle, name, qt, rows, k, nb):
    import gpu_util as U
    rng = np.random.default_rng(rows + k + nb + qt)
    be, _ = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    x = (rng.standard_normal((nb, k)) * np.exp(rng.uniform(-1, 1, (nb, 1)))).astype(np.float32)
    w = U.gpu_repack(qt, raw, rows, k)
    act = _prep(x, k, q8_0=(qt == 8))
    (got,) = _batch_gemv([(w, qt, rows)], k, act, nb)
    for b in range(nb):
        canon = oracle.matmul(qt, raw, rows, k, x[b], mode="canon")
        assert np.array_equal(_bits(got[b]), _bits(canon)), f"token {b}"
    # and identical to the batch-1 kernel
    (one,) = U.gpu_gemv([(w, qt, rows)], k, x[nb - 1])
    assert np.array_equal(_bits(got[nb - 1]), _bits(one))


def test_gemv_batch_mixed_segments_rmsnorm(oracle):
    """three formats in one launch (Q4_K | Q5_K | Q6_K: the generic K-quant kernel) sharing rms-normed inputs, 6 tokens"""
    import gpu_util as U
    k, nb, eps = 4096, 6, 1e-5
    rng = np.random.default_rng(71)
    x = (rng.standard_normal((nb, k)) * 3).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    segs, raws = [], []
    for qt, rows in ((12, 1024), (13, 256), (14, 256)):
        raw = rand_blocks(qt, rows * k // 256, rng)
        raws.append((qt, raw, rows))
        segs.append((U.gpu_repack(qt, raw, rows, k), qt, rows))
    act = _prep(x, k, norm_w=g, eps=eps)
    outs = _batch_gemv(segs, k, act, nb)
    for b in range(nb):
        h = oracle.rms_norm(x[b], g, eps)
        for got, (qt, raw, rows) in zip(outs, raws):
            assert np.array_equal(_bits(got[b]), _bits(oracle.matmul(qt, raw, rows, k, h, mode="canon")))


def test_gemv_batch_residual_and_swiglu(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    k, rows, nb = 2048, 1536, 7
    rng = np.random.default_rng(111)
    x = rng.standard_normal((nb, k)).astype(np.float32)
    rg = rand_blocks(12, rows * k // 256, rng)
    ru = rand_blocks(12, rows * k // 256, rng)
    wg, wu = U.gpu_repack(12, rg, rows, k), U.gpu_repack(12, ru, rows, k)
    act = _prep(x, k)
    out = torch.zeros(nb * rows, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_batch_args([(wg.data_ptr(), 12, rows, out.data_ptr()), (wu.data_ptr(), 12, rows, 0)], k, act.data_ptr(), nb,
                                  epilogue=cabi.EPI_SWIGLU)
    cabi.check(L.ggb_gemv_batch(C.byref(a), U.stream_ptr()))
    U.sync()
    got = out.cpu().numpy().reshape(nb, rows)
    res = rng.standard_normal((nb, rows)).astype(np.float32)
    rd = U.to_dev(res)
    a = cabi.make_gemv_batch_args([(wg.data_ptr(), 12, rows, rd.data_ptr())], k, act.data_ptr(), nb, epilogue=cabi.EPI_RESIDUAL,
                                  residual=rd.data_ptr())
    cabi.check(L.ggb_gemv_batch(C.byref(a), U.stream_ptr()))
    U.sync()
    got_res = rd.cpu().numpy().reshape(nb, rows)
    for b in range(nb):
        gg, uu = oracle.matmul(12, rg, rows, k, x[b], mode="canon"), oracle.matmul(12, ru, rows, k, x[b], mode="canon")
        assert np.array_equal(_bits(got[b]), _bits(oracle.swiglu(gg, uu, mode="canon")))
        assert np.array_equal(_bits(got_res[b]), _bits(res[b] + gg))


def test_gemv_batch_bad_args():
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    x = torch.zeros(1024, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_batch_args([(x.data_ptr(), 12, 4, x.data_ptr())], 100, x.data_ptr(), 2)
    assert L.ggb_gemv_batch(C.byref(a), 0) == -1
    a = cabi.make_gemv_batch_args([(x.data_ptr(), 2, 4, x.data_ptr())], 256, x.data_ptr(), 2)
    assert L.ggb_gemv_batch(C.byref(a), 0) == -3
    a = cabi.make_gemv_batch_args([(x.data_ptr(), 12, 4, x.data_ptr())], 256, x.data_ptr(), 2, epilogue=cabi.EPI_ARGMAX)
    assert L.ggb_gemv_batch(C.byref(a), 0) == -1
    a = cabi.make_gemv_batch_args([(x.data_ptr(), 12, 4, x.data_ptr())], 256, x.data_ptr(), 0)
    assert L.ggb_gemv_batch(C.byref(a), 0) == 0   # no tokens: nothing to do
    assert L.ggb_act_image_bytes(100) == -1


@pytest.mark.parametrize("gqa", ["2", "0"])
@pytest.mark.parametrize("n_head,n_kv,hd,n_ctx", [(8, 2, 64, 96), (8, 2, 128, 96), (16, 2, 128, 2048), (16, 2, 128, 2304), (4, 4, 128, 96)])
def test_rope_kv_attn_argmax_batch(oracle, monkeypatch, gqa, n_head, n_kv, hd, n_ctx):
    """per-token (slot, position) addressing: every entry equals the one-token kernels' / the oracle's result.  head_dim 128
    with 4 or 8 query heads per KV head takes the grouped-query kernel when GGB_ATTN_GQA=2 forces it (by default only batches with
    enough clusters to fill the GPU do): one cluster of four CTAs per four query heads, 4 or 8 positions per lane group in flight by
    context size; the other shapes, and GGB_ATTN_GQA=0, the per-head kernel."""
    monkeypatch.setenv("GGB_ATTN_GQA", gqa)
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    from ggufb200.model import rope_table
    L = cabi.lib()
    n_slots = 4
    qd, kvd = n_head * hd, n_kv * hd
    rng = np.random.default_rng(5)
    slots = np.array([2, 0, 3, 1, 0], dtype=np.int32)       # entry 4 is idle (pos -1)
    pos = np.array([17, 0, n_ctx - 1, n_ctx // 2 + 3, -1], dtype=np.int32)
    nb = len(slots)
    q = rng.standard_normal((nb, qd)).astype(np.float32)
    k = rng.standard_normal((nb, kvd)).astype(np.float32)
    v = rng.standard_normal((nb, kvd)).astype(np.float32)
    kc0 = (rng.standard_normal((n_slots, n_ctx, kvd)) * 0.5).astype(np.float16)
    vc0 = rng.standard_normal((n_slots, n_ctx, kvd)).astype(np.float16)
    tabh = rope_table(n_ctx, hd, 10000.0)
    qdv, kdv, vdv = U.to_dev(q), U.to_dev(k), U.to_dev(v)
    kcd, vcd = U.to_dev(kc0.view(np.int16)), U.to_dev(vc0.view(np.int16))
    posd, slotd, tab = U.to_dev(pos), U.to_dev(slots), U.to_dev(tabh)
    cabi.check(L.ggb_rope_kv_batch(qdv.data_ptr(), kdv.data_ptr(), vdv.data_ptr(), nb, posd.data_ptr(), slotd.data_ptr(), n_ctx * kvd,
                                   n_head, n_kv, hd, hd, tab.data_ptr(), kcd.data_ptr(), vcd.data_ptr(), U.stream_ptr()))
    out = torch.zeros((nb, qd), dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_attn_decode_batch(qdv.data_ptr(), kcd.data_ptr(), vcd.data_ptr(), posd.data_ptr(), slotd.data_ptr(), n_ctx * kvd,
                                       nb, n_head, n_kv, hd, n_ctx, out.data_ptr(), 0, U.stream_ptr()))
    U.sync()
    kc1, vc1 = kcd.cpu().numpy().view(np.uint16), vcd.cpu().numpy().view(np.uint16)
    exp_k, exp_v = kc0.view(np.uint16).copy(), vc0.view(np.uint16).copy()
    for b in range(nb - 1):
        ctab = oracle.rope_table_canon(int(pos[b]), hd, 10000.0)
        q_ref = oracle.rope_apply(q[b], n_head, hd, hd, ctab)
        exp_k[slots[b], pos[b]] = oracle.fp32_to_fp16(oracle.rope_apply(k[b], n_kv, hd, hd, ctab))
        exp_v[slots[b], pos[b]] = oracle.fp32_to_fp16(v[b])
        assert np.array_equal(_bits(qdv[b].cpu().numpy()), _bits(q_ref))
    assert np.array_equal(kc1, exp_k) and np.array_equal(vc1, exp_v)      # nothing but the addressed rows written
    assert np.array_equal(_bits(qdv[nb - 1].cpu().numpy()), _bits(q[nb - 1]))  # idle entry untouched
    for b in range(nb - 1):
        canon = oracle.attn_decode(qdv[b].cpu().numpy(), exp_k[slots[b]], exp_v[slots[b]], n_head, n_kv, hd, int(pos[b]) + 1, mode="canon")
        assert np.array_equal(_bits(out[b].cpu().numpy()), _bits(canon)), f"entry {b}"
    assert not out[nb - 1].cpu().numpy().any()
    # arg-max per row, first index wins ties
    x = rng.standard_normal((3, 5000)).astype(np.float32)
    x[1, 77] = x[1, 4000] = 9.0
    idx = torch.zeros(3, dtype=torch.int32, device=U.DEV)
    cabi.check(L.ggb_argmax_rows(U.to_dev(x).data_ptr(), 5000, 3, idx.data_ptr(), U.stream_ptr()))
    U.sync()
    assert idx.cpu().tolist() == [int(np.argmax(x[0])), 77, int(np.argmax(x[2]))]


# ------------------------------------------------------------------ tiled activation images (one pass over the weights for 9..16 tokens of a long vector)
TILE_STRIDE, TILE_BS, TILE_DSC = 2448, 2048, 2304


def _prep_tiled(x: np.ndarray, k: int, norm_w=None, eps=0.0):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    nb = x.shape[0]
    size = L.ggb_act_tiled_bytes(k, nb)
    assert size == -(-k // 2048) * nb * TILE_STRIDE
    xd = U.to_dev(x.astype(np.float32))
    act = torch.zeros(size, dtype=torch.uint8, device=U.DEV)
    gd = U.to_dev(norm_w) if norm_w is not None else None
    cabi.check(L.ggb_act_prep_tiled(xd.data_ptr(), gd.data_ptr() if gd is not None else 0, eps, k, nb, act.data_ptr(), 0, U.stream_ptr()))
    U.sync()
    return act


@pytest.mark.parametrize("k,norm", [(2048, False), (5632, True), (14336, False), (4096, True)])
def test_act_prep_tiled_bit_exact(oracle, k, norm):
    rng = np.random.default_rng(k + 11)
    nb, eps = 11, 1e-5
    x = (rng.standard_normal((nb, k)) * np.exp(rng.uniform(-3, 3, (nb, 1)))).astype(np.float32)
    x[0, :256] = 0.0
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32) if norm else None
    T = -(-k // 2048)
    img = _prep_tiled(x, k, norm_w=g, eps=eps).cpu().numpy().reshape(T, nb, TILE_STRIDE)
    for b in range(nb):
        h = oracle.rms_norm(x[b], g, eps) if norm else x[b]
        rd, rq, rb = oracle.q8_K_fields(oracle.quantize_q8_K(h))
        for t in range(T):
            nsb = min(8, k // 256 - 8 * t)
            sl = img[t, b]
            assert np.array_equal(_unswizzle(sl[:2048].copy())[:256 * nsb].view(np.int8), rq[8 * t:8 * t + nsb].reshape(-1)), (b, t)
            assert np.array_equal(sl[TILE_BS:TILE_BS + 32 * nsb].copy().view(np.int16), rb[8 * t:8 * t + nsb].reshape(-1))
            assert np.array_equal(sl[TILE_DSC:TILE_DSC + 4 * nsb].copy().view(np.uint32), _bits(rd[8 * t:8 * t + nsb]))


@pytest.mark.parametrize("name,qt", [("q4_k", 12), ("q6_k", 14)])
@pytest.mark.parametrize("rows,k,nb", [(300, 14336, 11), (4096, 14336, 16), (600, 5632, 9), (100, 4096, 16), (2, 2048, 5), (8192, 8192, 13)])
def test_gemv_batch_tiled_images_match_oracle_and_whole_images(oracle, name, qt, rows, k, nb):
    """the tile-major walk over streamed image slices gives the bits of the whole-image kernel and of the oracle, for every epilogue"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    rng = np.random.default_rng(rows + k + nb + qt + 1)
    be, _ = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    x = (rng.standard_normal((nb, k)) * np.exp(rng.uniform(-1, 1, (nb, 1)))).astype(np.float32)
    w = U.gpu_repack(qt, raw, rows, k)
    (whole,) = _batch_gemv([(w, qt, rows)], k, _prep(x, k), nb)
    act = _prep_tiled(x, k)
    (got,) = _batch_gemv([(w, qt, rows)], k, act, nb, act_tiled=1)
    assert np.array_equal(_bits(got), _bits(whole))
    for b in (0, nb - 1):
        assert np.array_equal(_bits(got[b]), _bits(oracle.matmul(qt, raw, rows, k, x[b], mode="canon"))), f"token {b}"
    # residual and unrounded-f64 epilogues
    res = rng.standard_normal((nb, rows)).astype(np.float32)
    y = U.to_dev(res.copy())
    a = cabi.make_gemv_batch_args([(w.data_ptr(), qt, rows, y.data_ptr())], k, act.data_ptr(), nb, epilogue=cabi.EPI_RESIDUAL, residual=y.data_ptr(), act_tiled=1)
    cabi.check(L.ggb_gemv_batch(C.byref(a), U.stream_ptr()), "residual")
    U.sync()
    assert np.array_equal(_bits(y.cpu().numpy()), _bits(res + whole))
    y64 = torch.zeros(nb * rows, dtype=torch.float64, device=U.DEV)
    a = cabi.make_gemv_batch_args([(w.data_ptr(), qt, rows, y64.data_ptr())], k, act.data_ptr(), nb, epilogue=cabi.EPI_STORE_F64, act_tiled=1)
    cabi.check(L.ggb_gemv_batch(C.byref(a), U.stream_ptr()), "store_f64")
    U.sync()
    assert np.array_equal(_bits(y64.cpu().numpy().astype(np.float32).reshape(nb, rows)), _bits(whole))


def test_gemv_batch_prefers_tiled_only_where_two_passes_would_run():
    from ggufb200 import cabi
    L = cabi.lib()

    def pref(rows, k, nb, qt=12, nseg=1, epi=None):
        a = cabi.make_gemv_batch_args([(16, qt, rows, 16)] * nseg, k, 16, nb, epilogue=cabi.EPI_STORE if epi is None else epi)
        return L.ggb_gemv_batch_prefers_tiled(C.byref(a))

    assert pref(4096, 14336, 16) == 1 and pref(4096, 14336, 16, qt=14) == 1 and pref(4096, 14336, 9) == 1      # ffn_down of Llama-3-8B
    assert pref(4096, 14336, 8) == 0                 # one pass anyway
    assert pref(4096, 4096, 16) == 0                 # sixteen whole images fit
    assert pref(14336, 4096, 16, nseg=2, epi=cabi.EPI_SWIGLU) == 0
    assert pref(4096, 14336, 16, qt=8) == 0          # Q8_0 stays on the dp4a kernel
    assert pref(128256, 14336, 16) == 0              # too many row groups per CTA for the partial sums
    # errors: tiled images with a shape / format the tiled kernel does not take
    a = cabi.make_gemv_batch_args([(16, 8, 64, 16)], 4096, 16, 4, act_tiled=1)
    assert L.ggb_gemv_batch(C.byref(a), 0) == -3    # GGB_ERR_UNSUPPORTED
