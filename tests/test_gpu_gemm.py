"""-m gpu: the tcgen05/TMEM dequant-GEMM (csrc/gemm.cu) against the oracle.

The GEMM is the tolerance-level path (fp16 operands, f32 accumulation in TMEM) -- the same trade upstream's CUDA
backend makes for batches.  Two checks: (1) against an exact emulation of its own arithmetic (oracle-dequantised
weights rounded to f16, f16 activations, float64 accumulation) the error must be f32-accumulation noise; (2)
against the reference integer path (oracle mul_mat, Q8_K activations) it must sit inside the stated 1e-2."""
import numpy as np
import pytest

from conftest import rand_blocks

pytestmark = pytest.mark.gpu
TYPES = {"q4_k": 12, "q5_k": 13, "q6_k": 14, "q8_0": 8}


def f16_round(a: np.ndarray) -> np.ndarray:
    return np.clip(np.ascontiguousarray(a, dtype=np.float32), -65504.0, 65504.0).astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("rows,k,tokens", [(128, 256, 256), (300, 2048, 17), (256, 4096, 300), (1000, 5632, 64), (4096, 4096, 512)])
def test_gemm_matches_oracle(oracle, name, rows, k, tokens):
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    qt = TYPES[name]
    rng = np.random.default_rng(rows + k + tokens + qt)
    be, _ = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    X = rng.standard_normal((tokens, k)).astype(np.float32)
    w = U.gpu_repack(qt, raw, rows, k)
    xd = U.to_dev(X)
    xb = torch.empty((tokens, k), dtype=torch.float16, device=U.DEV)
    cabi.check(L.ggb_f32_to_f16(xd.data_ptr(), xb.data_ptr(), tokens * k, U.stream_ptr()))
    y = torch.full((tokens, rows), float("nan"), dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_gemm(qt, w.data_ptr(), rows, k, xb.data_ptr(), tokens, y.data_ptr(), rows, U.stream_ptr()))
    U.sync()
    got = y.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.array_equal(xb.float().cpu().numpy(), f16_round(X))
    Wd = oracle.dequantize(raw, qt, rows * k).reshape(rows, k)
    exact = f16_round(X).astype(np.float64) @ f16_round(Wd).astype(np.float64).T
    scale = np.abs(exact).max()
    # Q8_0 operands are exactly "dequantise, round to f16" (d is an f16, d * q is rounded once): only f32-accumulation noise
    # is left.  K-quant operands round the sub-block scale d * sc (and dmin * m) to f16 before the packed-half multiply
    # (csrc/gemm.cu): up to ~3 f16 roundings per weight, coherent inside a sub-block.
    tol = 2e-5 * max(1.0, (k / 1024) ** 0.5) if name == "q8_0" else 1e-3
    assert np.abs(got - exact).max() <= tol * scale, np.abs(got - exact).max() / scale
    if tokens <= 64:   # the reference integer path, column by column (slow on the CPU: keep it small)
        ref = oracle.matmul(qt, raw, rows, k, X)
        assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max()


@pytest.mark.parametrize("name,rows,k,tokens", [("q4_k", 300, 2048, 700), ("q6_k", 128, 256, 17), ("q8_0", 1000, 1024, 513)])
def test_gemm2_equals_two_gemms(name, rows, k, tokens):
    """ggb_gemm2 (ffn_gate + ffn_up in one launch: the second matrix's row tiles ride behind the first's) must give exactly what
    two ggb_gemm launches give"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    qt = TYPES[name]
    rng = np.random.default_rng(rows + k + tokens)
    import oracle.oracle as O
    be, _ = O.BLOCK[qt]
    w0 = U.gpu_repack(qt, rand_blocks(qt, rows * k // be, rng), rows, k)
    w1 = U.gpu_repack(qt, rand_blocks(qt, rows * k // be, rng), rows, k)
    xb = torch.randn((tokens, k), device=U.DEV).to(torch.float16)
    ya, yb, y0, y1 = (torch.full((tokens, rows), float("nan"), dtype=torch.float32, device=U.DEV) for _ in range(4))
    s = U.stream_ptr()
    cabi.check(L.ggb_gemm(qt, w0.data_ptr(), rows, k, xb.data_ptr(), tokens, ya.data_ptr(), rows, s))
    cabi.check(L.ggb_gemm(qt, w1.data_ptr(), rows, k, xb.data_ptr(), tokens, yb.data_ptr(), rows, s))
    cabi.check(L.ggb_gemm2(qt, w0.data_ptr(), w1.data_ptr(), rows, k, xb.data_ptr(), tokens, y0.data_ptr(), y1.data_ptr(), rows, s))
    U.sync()
    assert torch.isfinite(y0).all() and torch.equal(y0, ya) and torch.equal(y1, yb)


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("rows,k,tokens", [(300, 2048, 17), (256, 4096, 33), (64, 14336, 8)])
def test_gemm_on_requantised_activations_tracks_the_integer_path(oracle, name, rows, k, tokens):
    """The prefill feeds the GEMM d * q -- the activations quantised exactly as the CPU path quantises them (Q8_K / Q8_0),
    dequantised to f16 (ggb_act_fakequant_f16).  Against ggml's integer dot (oracle mul_mat) only the f16 rounding of the
    two operands is left: <= 1e-3 of the largest output, ten times inside the north-star 1e-2."""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    qt = TYPES[name]
    rng = np.random.default_rng(rows * 7 + k + tokens + qt)
    be, _ = oracle.BLOCK[qt]
    raw = rand_blocks(qt, rows * k // be, rng)
    X = (rng.standard_normal((tokens, k)) * rng.uniform(0.05, 20.0, size=(tokens, 1))).astype(np.float32)
    w = U.gpu_repack(qt, raw, rows, k)
    xd = U.to_dev(X)
    xh = torch.empty((tokens, k), dtype=torch.float16, device=U.DEV)
    cabi.check(L.ggb_act_fakequant_f16(xd.data_ptr(), xh.data_ptr(), k, tokens, 1 if qt == 8 else 0, U.stream_ptr()))
    y = torch.full((tokens, rows), float("nan"), dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_gemm(qt, w.data_ptr(), rows, k, xh.data_ptr(), tokens, y.data_ptr(), rows, U.stream_ptr()))
    U.sync()
    got = y.cpu().numpy()
    ref = oracle.matmul(qt, raw, rows, k, X)
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() <= 1e-3, err.max()


def test_gemm_rejects_bad_arguments():
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    t = torch.zeros(1024, dtype=torch.float32, device=U.DEV)
    assert L.ggb_gemm(12, t.data_ptr(), 4, 100, t.data_ptr(), 4, t.data_ptr(), 4, 0) == -1
    assert L.ggb_gemm(2, t.data_ptr(), 4, 256, t.data_ptr(), 4, t.data_ptr(), 4, 0) == -3
    assert L.ggb_gemm(12, t.data_ptr(), 0, 256, t.data_ptr(), 4, t.data_ptr(), 4, 0) == 0


@pytest.mark.parametrize("kernel", ["tcgen05", "mma.sync"])
@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 8, 2), (64, 4, 4), (128, 4, 1)])
@pytest.mark.parametrize("T,pos0", [(1, 0), (150, 37), (64, 0), (200, 300), (300, 517), (128, 128)])
def test_attn_prefill_tensor_core_matches_oracle(oracle, monkeypatch, kernel, hd, n_head, n_kv, T, pos0):
    """causal prefill attention against the oracle's one-token attention, token by token: the tcgen05 / TMEM / TMA kernel
    (csrc/prefill_tc.cu, the default) and the mma.sync flash-attention it falls back to (csrc/prefill.cu)"""
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    monkeypatch.setenv("GGB_ATTN_PREFILL_TC", "1" if kernel == "tcgen05" else "0")
    L = cabi.lib()
    n_ctx = pos0 + T
    rng = np.random.default_rng(hd + T + pos0)
    q = rng.standard_normal((T, n_head * hd)).astype(np.float32)
    kc = (rng.standard_normal((n_ctx, n_kv * hd)) * 0.5).astype(np.float16)
    vc = rng.standard_normal((n_ctx, n_kv * hd)).astype(np.float16)
    qd, kd, vd = U.to_dev(q), U.to_dev(kc.view(np.int16)), U.to_dev(vc.view(np.int16))
    out = torch.zeros((T, n_head * hd), dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_attn_prefill(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), T, pos0, n_head, n_kv, hd, out.data_ptr(), U.stream_ptr()))
    U.sync()
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    for t in sorted({min(x, T - 1) for x in (0, 1, T // 2, max(T - 2, 0), T - 1, 63, 64, 127, 128, 255, 256)}):
        ref = oracle.attn_decode(q[t], kc.view(np.uint16), vc.view(np.uint16), n_head, n_kv, hd, pos0 + t + 1)
        assert np.abs(got[t] - ref).max() <= 3e-3 * max(np.abs(ref).max(), 1e-3), f"token {t}"
