"""-m gpu: the tcgen05/TMEM dequant-GEMM (csrc/gemm.cu) against the oracle.

The GEMM is the tolerance-level path (bf16 operands, f32 accumulation in TMEM) -- the same trade upstream's CUDA
backend makes for batches.  Two checks: (1) against an exact emulation of its own arithmetic (oracle-dequantised
weights rounded to bf16, bf16 activations, float64 accumulation) the error must be f32-accumulation noise; (2)
against the reference integer path (oracle mul_mat, Q8_K activations) it must sit inside the stated 1e-2."""
import numpy as np
import pytest

from conftest import rand_blocks

pytestmark = pytest.mark.gpu
TYPES = {"q4_k": 12, "q6_k": 14, "q8_0": 8}


def bf16_round(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


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
    xb = torch.empty((tokens, k), dtype=torch.bfloat16, device=U.DEV)
    cabi.check(L.ggb_f32_to_bf16(xd.data_ptr(), xb.data_ptr(), tokens * k, U.stream_ptr()))
    y = torch.full((tokens, rows), float("nan"), dtype=torch.float32, device=U.DEV)
    cabi.check(L.ggb_gemm(qt, w.data_ptr(), rows, k, xb.data_ptr(), tokens, y.data_ptr(), rows, U.stream_ptr()))
    U.sync()
    got = y.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.array_equal(xb.float().cpu().numpy(), bf16_round(X))
    Wd = oracle.dequantize(raw, qt, rows * k).reshape(rows, k)
    exact = bf16_round(X).astype(np.float64) @ bf16_round(Wd).astype(np.float64).T
    scale = np.abs(exact).max()
    assert np.abs(got - exact).max() <= 2e-5 * scale * max(1.0, (k / 1024) ** 0.5), np.abs(got - exact).max() / scale
    if tokens <= 64:   # the reference integer path, column by column (slow on the CPU: keep it small)
        ref = oracle.matmul(qt, raw, rows, k, X)
        assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max()


def test_gemm_rejects_bad_arguments():
    import torch
    import gpu_util as U
    from ggufb200 import cabi
    L = cabi.lib()
    t = torch.zeros(1024, dtype=torch.float32, device=U.DEV)
    assert L.ggb_gemm(12, t.data_ptr(), 4, 100, t.data_ptr(), 4, t.data_ptr(), 4, 0) == -1
    assert L.ggb_gemm(2, t.data_ptr(), 4, 256, t.data_ptr(), 4, t.data_ptr(), 4, 0) == -3
    assert L.ggb_gemm(12, t.data_ptr(), 0, 256, t.data_ptr(), 4, t.data_ptr(), 4, 0) == 0
