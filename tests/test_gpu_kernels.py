"""-m gpu: parity of every kernel of libggufb200 against the CPU oracle, called through the C-ABI.

Bars (BASELINE.json north_star): dequantisation bit-exact; activation quantisation bit-exact (int8 codes,
scales, sums); matvec within a relative tolerance of the reference's (ggml-order) result.  The kernels
accumulate their f32 terms in f64, so they are additionally required to equal the oracle's order-independent
"canon" variant BIT FOR BIT (oracle/ggml_ref.c), and to sit within 2e-5 of the ggml-order restatement --
far inside the stated 1e-2.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import rand_blocks

pytestmark = pytest.mark.gpu

TYPES = {"q8_0": 8, "q4_k": 12, "q5_k": 13, "q6_k": 14}
GEMV_TYPES = {"q8_0": 8, "q4_k": 12, "q5_k": 13, "q6_k": 14}


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _same_bits_or_nan(a, b):
    return np.all((_bits(a) == _bits(b)) | (np.isnan(a) & np.isnan(b)))


# ----------------------------------------------------------------------------- K0 dequant
@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("wild", [False, True])
def test_dequant_bit_exact(oracle, name, wild):
    import gpu_util as U
    qt = TYPES[name]
    rng = np.random.default_rng(100 + qt + wild)
    nblk = 777
    raw = rand_blocks(qt, nblk, rng, wild=wild)
    n = nblk * oracle.BLOCK[qt][0]
    ref = oracle.dequantize(raw, qt, n)
    got = U.gpu_dequant(qt, raw, n)
    assert _same_bits_or_nan(ref, got)


def test_dequant_empty_and_errors(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    assert U.gpu_dequant(12, np.zeros(0, np.uint8), 0).size == 0
    out = torch.empty(256, dtype=torch.float32, device=U.DEV)
    assert L.ggb_dequant(10, out.data_ptr(), out.data_ptr(), 256, 0) == -3  # Q2_K unsupported
    assert L.ggb_dequant(12, out.data_ptr(), out.data_ptr(), 100, 0) == -1  # ragged
    assert b"multiple" in L.ggb_last_error()


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("k", [256, 512, 2048, 2304, 4096, 5632, 14336])
def test_repack_roundtrip_bit_exact(oracle, name, k):
    """canonical -> tile-SoA -> dequant equals canonical dequant, for full, partial and multi-tile rows."""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    qt = TYPES[name]
    rows = 5
    rng = np.random.default_rng(k + qt)
    be, bb = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    ref = oracle.dequantize(raw, qt, rows * k)
    w = U.gpu_repack(qt, raw, rows, k)
    out = torch.empty(rows * k, dtype=torch.float32, device=U.DEV)
    cabi.check(cabi.lib().ggb_dequant_repacked(qt, w.data_ptr(), out.data_ptr(), rows, k, U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(ref), _bits(out.cpu().numpy()))
    # same byte count per row up to the 16-byte stride round-up
    stride = cabi.lib().ggb_repacked_row_stride(qt, k)
    assert 0 <= stride - k // be * bb < 16


# ----------------------------------------------------------------------------- activation quantisation
@pytest.mark.parametrize("k", [256, 4096, 14336])
def test_quantize_q8_K_bit_exact(oracle, k):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    rng = np.random.default_rng(k)
    m = 3
    x = (rng.standard_normal((m, k)) * np.exp(rng.uniform(-3, 3, (m, 1)))).astype(np.float32)
    x[0, :256] = 0.0                      # an all-zero block
    x[1, 5] = x[1, 9] = -np.abs(x[1]).max() * 2  # tie on |max|: the first one decides the sign
    x[1, 9] *= -1
    xd = U.to_dev(x)
    qs = torch.empty((m, k), dtype=torch.int8, device=U.DEV)
    d = torch.empty((m, k // 256), dtype=torch.float32, device=U.DEV)
    bs = torch.empty((m, k // 16), dtype=torch.int16, device=U.DEV)
    cabi.check(cabi.lib().ggb_quantize_q8_K(xd.data_ptr(), qs.data_ptr(), d.data_ptr(), bs.data_ptr(), k, m, U.stream_ptr()))
    U.sync()
    for r in range(m):
        rd, rq, rb = oracle.q8_K_fields(oracle.quantize_q8_K(x[r]))
        assert np.array_equal(rq.reshape(-1), qs[r].cpu().numpy())
        assert np.array_equal(_bits(rd), _bits(d[r].cpu().numpy()))
        assert np.array_equal(rb.reshape(-1), bs[r].cpu().numpy())


@pytest.mark.parametrize("k", [32, 4096])
def test_quantize_q8_0_bit_exact(oracle, k):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    rng = np.random.default_rng(k + 1)
    x = rng.standard_normal(k).astype(np.float32)
    x[:32] = 0
    xd = U.to_dev(x)
    qs = torch.empty(k, dtype=torch.int8, device=U.DEV)
    dd = torch.empty(k // 32, dtype=torch.int16, device=U.DEV)
    cabi.check(cabi.lib().ggb_quantize_q8_0(xd.data_ptr(), qs.data_ptr(), dd.data_ptr(), k, 1, U.stream_ptr()))
    U.sync()
    ref = oracle.quantize_q8_0(x).reshape(-1, 34)
    assert np.array_equal(ref[:, 2:].copy().view(np.int8).reshape(-1), qs.cpu().numpy())
    assert np.array_equal(ref[:, :2].copy().view(np.uint16).reshape(-1), dd.cpu().numpy().view(np.uint16))


# ----------------------------------------------------------------------------- K1 GEMV
def _gemv_case(oracle, qt, rows, k, seed, std=1.0):
    rng = np.random.default_rng(seed)
    be, _ = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    x = (rng.standard_normal(k) * std).astype(np.float32)
    return raw, x


@pytest.mark.parametrize("name", list(GEMV_TYPES))
@pytest.mark.parametrize("rows,k", [(1, 256), (7, 512), (300, 2048), (64, 5632), (1000, 4096), (500, 14336), (4096, 4096)])
def test_gemv_store_matches_oracle(oracle, name, rows, k):
    import gpu_util as U
    qt = GEMV_TYPES[name]
    raw, x = _gemv_case(oracle, qt, rows, k, rows * 31 + k + qt)
    ref = oracle.matmul(qt, raw, rows, k, x)
    canon = oracle.matmul(qt, raw, rows, k, x, mode="canon")
    w = U.gpu_repack(qt, raw, rows, k)
    (got,) = U.gpu_gemv([(w, qt, rows)], k, x)
    assert np.array_equal(_bits(got), _bits(canon)), f"{(got != canon).sum()} of {rows} outputs differ from the canon oracle"
    tol = 2e-5 * max(np.abs(ref).max(), 1e-30)
    assert np.abs(got - ref).max() <= tol, (np.abs(got - ref).max(), tol)


@pytest.mark.parametrize("full_k_model", [0, 1])
@pytest.mark.parametrize("name,rows,k", [("q4_k", 300, 4096), ("q6_k", 200, 14336), ("q4_k", 64, 8192), ("q5_k", 40, 4096)])
def test_gemv_activation_scales_outside_the_inline_division_range(oracle, name, rows, k, full_k_model):
    """Blocks whose largest magnitude is below 2^-90 or above 2^90 leave the prologue's inline-division quantiser and are
    redone in the reference form (gemv.cu: `redo`); every prologue shape (2, 4 and > 4 blocks per warp), both kernel
    instances.  ggml: quantize_row_q8_K [UPSTREAM-MEM ggml-quants.c]."""
    import gpu_util as U
    qt = GEMV_TYPES[name]
    raw, x = _gemv_case(oracle, qt, rows, k, rows + k)
    x[:256] *= np.float32(1e-33)
    x[256:512] *= np.float32(1e30)
    x[512:768] = 0
    x[k - 256:] *= np.float32(3e-30)
    canon = oracle.matmul(qt, raw, rows, k, x, mode="canon")
    w = U.gpu_repack(qt, raw, rows, k)
    (got,) = U.gpu_gemv([(w, qt, rows)], k, x, full_k_model=full_k_model)
    assert np.array_equal(_bits(got), _bits(canon)), f"{(got != canon).sum()} of {rows} outputs differ from the canon oracle"


def test_gemv_zero_rows_and_bad_args(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    x = torch.zeros(256, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_args([(0, 12, 0, x.data_ptr())], 256, x.data_ptr())
    assert L.ggb_gemv(C.byref(a), 0) == 0  # nothing to do
    a = cabi.make_gemv_args([(x.data_ptr(), 12, 4, x.data_ptr())], 100, x.data_ptr())
    assert L.ggb_gemv(C.byref(a), 0) == -1
    a = cabi.make_gemv_args([(x.data_ptr(), 2, 4, x.data_ptr())], 256, x.data_ptr())  # Q4_0 weights
    assert L.ggb_gemv(C.byref(a), 0) == -3


def test_gemv_mixed_segments_rmsnorm_prologue(oracle):
    """q/k in Q4_K and v in Q6_K sharing one rms-normed input (the QKV launch shape of a Q4_K_M file)."""
    import gpu_util as U
    from ggufb200 import cabi
    k = 4096
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(k) * 3).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    eps = 1e-5
    h = oracle.rms_norm(x, g, eps)
    segs, refs = [], []
    for qt, rows in ((12, 4096), (12, 1024), (14, 1024)):
        raw = rand_blocks(qt, rows * k // 256, rng)
        refs.append((oracle.matmul(qt, raw, rows, k, h), oracle.matmul(qt, raw, rows, k, h, mode="canon")))
        segs.append((U.gpu_repack(qt, raw, rows, k), qt, rows))
    gd = U.to_dev(g)
    outs = U.gpu_gemv(segs, k, x, prologue=cabi.PRO_RMSNORM, norm_w=gd.data_ptr(), eps=eps)
    for got, (ref, canon) in zip(outs, refs):
        assert np.array_equal(_bits(got), _bits(canon))
        assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("mix", [((12, 1024), (12, 256), (13, 256)), ((13, 512), (13, 128), (14, 128)), ((13, 300), (13, 300))])
def test_gemv_q5k_mixed_segments(oracle, mix):
    """Q5_K next to Q4_K (attn_v of a real 70B Q4_K_M file) and next to Q6_K (a Q5_K_M file's QKV): the generic K-quant kernel"""
    import gpu_util as U
    from ggufb200 import cabi
    k = 2304   # one full tile + a 1-super-block tail
    rng = np.random.default_rng(len(mix) * 31 + mix[0][0])
    x = (rng.standard_normal(k) * 2).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    h = oracle.rms_norm(x, g, 1e-5)
    segs, refs = [], []
    for qt, rows in mix:
        raw = rand_blocks(qt, rows * k // 256, rng)
        refs.append(oracle.matmul(qt, raw, rows, k, h, mode="canon"))
        segs.append((U.gpu_repack(qt, raw, rows, k), qt, rows))
    gd = U.to_dev(g)    # keep the gains alive across the launch
    outs = U.gpu_gemv(segs, k, x, prologue=cabi.PRO_RMSNORM, norm_w=gd.data_ptr(), eps=1e-5)
    for got, canon in zip(outs, refs):
        assert np.array_equal(_bits(got), _bits(canon))


def test_gemv_min_smem_and_smem_query(oracle):
    """min_smem only changes the shared-memory request (placement), never the result; ggb_gemv_smem_bytes reports it"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    k, rows = 4096, 1000
    rng = np.random.default_rng(23)
    raw = rand_blocks(12, rows * k // 256, rng)
    x = rng.standard_normal(k).astype(np.float32)
    w, xd = U.gpu_repack(12, raw, rows, k), U.to_dev(x)
    y = torch.zeros(rows, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_args([(w.data_ptr(), 12, rows, y.data_ptr())], k, xd.data_ptr())
    base = L.ggb_gemv_smem_bytes(C.byref(a))
    assert 60 * 1024 < base < 114 * 1024
    a.min_smem = 118 * 1024
    assert L.ggb_gemv_smem_bytes(C.byref(a)) == 118 * 1024
    cabi.check(L.ggb_gemv(C.byref(a), U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(y.cpu().numpy()), _bits(oracle.matmul(12, raw, rows, k, x, mode="canon")))
    a.min_smem = 16            # smaller than needed: ignored
    assert L.ggb_gemv_smem_bytes(C.byref(a)) == base
    a.k = 100                  # invalid args: the query reports the error code
    assert L.ggb_gemv_smem_bytes(C.byref(a)) == -1


def test_gemv_residual_and_swiglu_epilogues(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    k, rows = 2048, 1536
    rng = np.random.default_rng(11)
    x = rng.standard_normal(k).astype(np.float32)
    rg = rand_blocks(12, rows * k // 256, rng)
    ru = rand_blocks(12, rows * k // 256, rng)
    g, u = oracle.matmul(12, rg, rows, k, x, mode="canon"), oracle.matmul(12, ru, rows, k, x, mode="canon")
    wg, wu = U.gpu_repack(12, rg, rows, k), U.gpu_repack(12, ru, rows, k)
    xd = U.to_dev(x)
    out = torch.zeros(rows, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_args([(wg.data_ptr(), 12, rows, out.data_ptr()), (wu.data_ptr(), 12, rows, 0)], k, xd.data_ptr(),
                            epilogue=cabi.EPI_SWIGLU)
    cabi.check(L.ggb_gemv(C.byref(a), U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(oracle.swiglu(g, u, mode="canon")))
    ref = oracle.swiglu(g, u)  # libm expf
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    # residual, in place
    res = rng.standard_normal(rows).astype(np.float32)
    rd = U.to_dev(res)
    a = cabi.make_gemv_args([(wg.data_ptr(), 12, rows, rd.data_ptr())], k, xd.data_ptr(), epilogue=cabi.EPI_RESIDUAL,
                            residual=rd.data_ptr())
    cabi.check(L.ggb_gemv(C.byref(a), U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(rd.cpu().numpy()), _bits(res + g))


def test_gemv_rope_kv_epilogue(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    from ggufb200.model import rope_table
    L = cabi.lib()
    k, n_head, n_kv, hd, n_ctx, pos = 1024, 8, 2, 64, 32, 17
    rng = np.random.default_rng(13)
    x = rng.standard_normal(k).astype(np.float32)
    raws = [rand_blocks(qt, rows * k // 256, rng) for qt, rows in ((12, n_head * hd), (12, n_kv * hd), (14, n_kv * hd))]
    q = oracle.matmul(12, raws[0], n_head * hd, k, x, mode="canon")
    kk = oracle.matmul(12, raws[1], n_kv * hd, k, x, mode="canon")
    v = oracle.matmul(14, raws[2], n_kv * hd, k, x, mode="canon")
    ctab = oracle.rope_table_canon(pos, hd, 10000.0)
    q_ref = oracle.rope_apply(q, n_head, hd, hd, ctab)
    k_ref = oracle.fp32_to_fp16(oracle.rope_apply(kk, n_kv, hd, hd, ctab))
    v_ref = oracle.fp32_to_fp16(v)
    # libm-table rope (ggml order) stays within float noise of the canon table
    assert np.abs(oracle.rope_norm(q, n_head, hd, hd, pos, 10000.0) - q_ref).max() <= 1e-5 * np.abs(q_ref).max()
    ws = [U.gpu_repack(qt, r, rows, k) for r, (qt, rows) in zip(raws, ((12, n_head * hd), (12, n_kv * hd), (14, n_kv * hd)))]
    tab = U.to_dev(rope_table(n_ctx, hd, 10000.0))
    # the host table equals the oracle's canon table bit for bit, and the libm table to float noise
    assert np.array_equal(_bits(rope_table(n_ctx, hd, 10000.0)[pos]), _bits(ctab))
    assert np.abs(rope_table(n_ctx, hd, 10000.0)[pos] - oracle.rope_table(pos, hd, 10000.0)).max() < 1e-6
    xd = U.to_dev(x)
    qo = torch.zeros(n_head * hd, dtype=torch.float32, device=U.DEV)
    kc = torch.zeros((n_ctx, n_kv * hd), dtype=torch.int16, device=U.DEV)
    vc = torch.zeros((n_ctx, n_kv * hd), dtype=torch.int16, device=U.DEV)
    posd = torch.tensor([pos], dtype=torch.int32, device=U.DEV)
    a = cabi.make_gemv_args([(ws[0].data_ptr(), 12, n_head * hd, qo.data_ptr()), (ws[1].data_ptr(), 12, n_kv * hd, 0),
                             (ws[2].data_ptr(), 14, n_kv * hd, 0)], k, xd.data_ptr(), epilogue=cabi.EPI_ROPE_KV,
                            pos_dev=posd.data_ptr(), rope_tab=tab.data_ptr(), n_rot=hd, head_dim=hd,
                            kcache=kc.data_ptr(), vcache=vc.data_ptr())
    cabi.check(L.ggb_gemv(C.byref(a), U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(qo.cpu().numpy()), _bits(q_ref))
    assert np.array_equal(kc.cpu().numpy().view(np.uint16)[pos], k_ref)
    assert np.array_equal(vc.cpu().numpy().view(np.uint16)[pos], v_ref)
    assert not kc.cpu().numpy()[:pos].any() and not kc.cpu().numpy()[pos + 1:].any()  # only row `pos` written


def test_gemv_argmax_epilogue(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    k, rows = 512, 5000
    rng = np.random.default_rng(17)
    x = rng.standard_normal(k).astype(np.float32)
    raw = rand_blocks(14, rows * k // 256, rng)
    ref = oracle.matmul(14, raw, rows, k, x, mode="canon")
    w = U.gpu_repack(14, raw, rows, k)
    xd = U.to_dev(x)
    y = torch.zeros(rows, dtype=torch.float32, device=U.DEV)
    a = cabi.make_gemv_args([(w.data_ptr(), 14, rows, y.data_ptr())], k, xd.data_ptr(), epilogue=cabi.EPI_ARGMAX)
    n_part = L.ggb_gemv_grid(C.byref(a))
    pv = torch.zeros(n_part, dtype=torch.float32, device=U.DEV)
    pi = torch.zeros(n_part, dtype=torch.int32, device=U.DEV)
    a.part_val, a.part_idx = pv.data_ptr(), pi.data_ptr()
    cabi.check(L.ggb_gemv(C.byref(a), U.stream_ptr()))
    tok = torch.zeros(1, dtype=torch.int32, device=U.DEV)
    pos = torch.tensor([4], dtype=torch.int32, device=U.DEV)
    step = torch.zeros(1, dtype=torch.int32, device=U.DEV)
    outt = torch.full((8,), -1, dtype=torch.int32, device=U.DEV)
    cabi.check(L.ggb_argmax_next(pv.data_ptr(), pi.data_ptr(), n_part, tok.data_ptr(), pos.data_ptr(), step.data_ptr(),
                                 outt.data_ptr(), 8, 0, 0, 0, 0, U.stream_ptr()))
    U.sync()
    got = y.cpu().numpy()
    assert np.array_equal(_bits(got), _bits(ref))
    assert int(tok[0]) == int(np.argmax(got)) == oracle.argmax(got)
    assert int(pos[0]) == 5 and int(step[0]) == 1 and int(outt[0]) == int(tok[0])


# ----------------------------------------------------------------------------- attention / small ops
@pytest.mark.parametrize("hd,n_head,n_kv", [(64, 8, 2), (128, 32, 8), (128, 8, 1), (64, 4, 4)])
@pytest.mark.parametrize("pos", [0, 1, 37, 300, 1023])
def test_attn_decode_matches_oracle(oracle, hd, n_head, n_kv, pos):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    n_ctx = 1024
    rng = np.random.default_rng(hd + n_head + pos)
    q = rng.standard_normal(n_head * hd).astype(np.float32)
    kc = (rng.standard_normal((n_ctx, n_kv * hd)) * 0.5).astype(np.float16)
    vc = rng.standard_normal((n_ctx, n_kv * hd)).astype(np.float16)
    ref = oracle.attn_decode(q, kc.view(np.uint16), vc.view(np.uint16), n_head, n_kv, hd, pos + 1)
    canon = oracle.attn_decode(q, kc.view(np.uint16), vc.view(np.uint16), n_head, n_kv, hd, pos + 1, mode="canon")
    qd, kd, vd = U.to_dev(q), U.to_dev(kc.view(np.int16)), U.to_dev(vc.view(np.int16))
    ws = torch.zeros(L.ggb_attn_decode_ws_bytes(n_head, hd), dtype=torch.uint8, device=U.DEV)
    out = torch.zeros(n_head * hd, dtype=torch.float32, device=U.DEV)
    posd = torch.tensor([pos], dtype=torch.int32, device=U.DEV)
    cabi.check(L.ggb_attn_decode(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), posd.data_ptr(), n_head, n_kv, hd, n_ctx,
                                 ws.data_ptr(), out.data_ptr(), 0, U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(canon))
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-6)


@pytest.mark.parametrize("n_head,n_kv", [(32, 8), (8, 1)])
@pytest.mark.parametrize("pos,n_ctx", [(0, 6000), (5, 6000), (4999, 6000), (5999, 6000), (300, 2304), (2303, 2304)])
def test_attn_decode_long_context_split_matches_oracle(oracle, n_head, n_kv, pos, n_ctx):
    """use_pdl bit 2 (the caller says the sequence is long): the two softmax passes as launches over position slices x groups of
    four query heads with the exchange through the workspace -- bit-identical with the oracle's canon attention at any position"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    hd = 128
    rng = np.random.default_rng(n_head + pos)
    q = rng.standard_normal(n_head * hd).astype(np.float32)
    kc = (rng.standard_normal((n_ctx, n_kv * hd)) * 0.5).astype(np.float16)
    vc = rng.standard_normal((n_ctx, n_kv * hd)).astype(np.float16)
    canon = oracle.attn_decode(q, kc.view(np.uint16), vc.view(np.uint16), n_head, n_kv, hd, pos + 1, mode="canon")
    qd, kd, vd = U.to_dev(q), U.to_dev(kc.view(np.int16)), U.to_dev(vc.view(np.int16))
    nws = L.ggb_attn_decode_ws_bytes_ctx(n_head, n_kv, hd, n_ctx)
    assert nws > 4 * n_head * n_ctx                      # the split path's workspace, not the 16-byte stub
    ws = torch.zeros(nws, dtype=torch.uint8, device=U.DEV)
    out = torch.zeros(n_head * hd, dtype=torch.float32, device=U.DEV)
    posd = torch.tensor([pos], dtype=torch.int32, device=U.DEV)
    for use_pdl in (4, 5, 0):              # split path without / with PDL, and the cluster kernel on the same inputs
        out.zero_()
        cabi.check(L.ggb_attn_decode(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), posd.data_ptr(), n_head, n_kv, hd, n_ctx,
                                     ws.data_ptr(), out.data_ptr(), use_pdl, U.stream_ptr()))
        U.sync()
        assert np.array_equal(_bits(out.cpu().numpy()), _bits(canon)), f"use_pdl={use_pdl}"
    assert L.ggb_attn_decode_ws_bytes_ctx(n_head, n_kv, hd, 1024) == 16 and L.ggb_attn_decode_ws_bytes_ctx(6, 2, hd, n_ctx) == 16


def test_rms_norm_swiglu_argmax_embed(oracle):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    rng = np.random.default_rng(3)
    k = 4096
    x = (rng.standard_normal((2, k)) * 5).astype(np.float32)
    g = rng.standard_normal(k).astype(np.float32)
    xd, gd = U.to_dev(x), U.to_dev(g)
    y = torch.empty_like(xd)
    cabi.check(L.ggb_rms_norm(xd.data_ptr(), gd.data_ptr(), y.data_ptr(), k, 2, 1e-5, U.stream_ptr()))
    U.sync()
    for r in range(2):
        assert np.array_equal(_bits(oracle.rms_norm(x[r], g, 1e-5)), _bits(y[r].cpu().numpy()))
    a, b = x[0], x[1]
    out = torch.empty(k, dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_swiglu(xd[0].data_ptr(), xd[1].data_ptr(), out.data_ptr(), k, U.stream_ptr()))
    U.sync()
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(oracle.swiglu(a, b, mode="canon")))
    ref = oracle.swiglu(a, b)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    # argmax: first index on ties
    v = rng.standard_normal(100000).astype(np.float32)
    v[777] = v[90000] = 50.0
    idx = torch.zeros(1, dtype=torch.int32, device=U.DEV)
    cabi.check(L.ggb_argmax(U.to_dev(v).data_ptr(), v.size, idx.data_ptr(), U.stream_ptr()))
    U.sync()
    assert int(idx[0]) == 777 == oracle.argmax(v)
    # embedding gather = dequant of one canonical row
    for qt in (12, 14, 8):
        be, bb = oracle.BLOCK[qt]
        raw = rand_blocks(qt, 10 * k // be, rng)
        emb = U.to_dev(raw.reshape(-1))
        tok = torch.tensor([6], dtype=torch.int32, device=U.DEV)
        xo = torch.empty(k, dtype=torch.float32, device=U.DEV)
        cabi.check(L.ggb_embed_row(qt, emb.data_ptr(), k, tok.data_ptr(), xo.data_ptr(), U.stream_ptr()))
        U.sync()
        rb = k // be * bb
        assert np.array_equal(_bits(oracle.dequantize(raw.reshape(-1)[6 * rb:7 * rb], qt, k)), _bits(xo.cpu().numpy()))


@pytest.mark.parametrize("n,nb,k", [(128256, 3, 40), (32000, 16, 1), (5000, 2, 240), (1001, 1, 7), (256, 4, 256)])
def test_topk_rows_returns_exactly_the_logits_from_the_kth_largest_up(n, nb, k):
    """ggb_topk_rows: per row the set {x >= k-th largest x} (unordered), incl. ties, negatives, -0.0 / +0.0 and a row of equal values"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    cap = 256
    rng = np.random.default_rng(n + k)
    x = (rng.standard_normal((nb, n)) * 4).astype(np.float32)
    x[0, rng.integers(0, n, 30)] = x[0].max()                     # ties at the top of row 0
    if nb > 1:
        x[1, :] = np.where(rng.random(n) < 0.5, -0.0, 0.0).astype(np.float32)   # row 1: only zeros of both signs
        x[1, rng.integers(0, n, max(1, k // 2))] = 1.5
    xd = U.to_dev(x)
    out = torch.zeros(nb * (2 * cap + 1), dtype=torch.int32, device=U.DEV)
    cabi.check(L.ggb_topk_rows(xd.data_ptr(), n, nb, k, cap, out.data_ptr(), out.data_ptr() + 4 * nb * cap, out.data_ptr() + 8 * nb * cap,
                               U.stream_ptr()), "topk_rows")
    U.sync()
    h = out.cpu().numpy()
    for b in range(nb):
        c = int(h[2 * nb * cap + b])
        key = x[b].view(np.int32).astype(np.int64)
        key = np.where(key < 0, -(key & 0x7FFFFFFF) - 1, key)     # the kernel's order: -0.0 below +0.0
        kth = np.sort(key)[-k]
        want = np.flatnonzero(key >= kth)
        assert c == want.size, (b, c, want.size)
        if c <= cap:
            idx = h[nb * cap + b * cap:nb * cap + b * cap + c]
            val = h[b * cap:b * cap + c].view(np.float32)
            assert np.array_equal(np.sort(idx), want)
            assert np.array_equal(val.view(np.uint32), x[b][idx].view(np.uint32))
    assert L.ggb_topk_rows(xd.data_ptr(), n, nb, 0, cap, out.data_ptr(), out.data_ptr(), out.data_ptr(), 0) == -1
    assert L.ggb_topk_rows(xd.data_ptr(), n, nb, cap + 1, cap, out.data_ptr(), out.data_ptr(), out.data_ptr(), 0) == -1
