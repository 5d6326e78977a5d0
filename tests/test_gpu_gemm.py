"""tcgen05 GEMM engine (minppo_b200/csrc/umma_gemm.cuh) against torch fp32 matmul on the same
bf16-rounded operands.  Tolerance: fp32 accumulation-order noise only (1e-4 relative to the
largest |C|), since both sides see identical bf16 inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(mode, M, N, K, splits, device, seed=0):
    import torch

    from minppo_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    stream = torch.cuda.current_stream(device).cuda_stream
    c = torch.full((splits, M, N), float("nan"), dtype=torch.float32, device=device)
    rowidx = None
    if mode == 0:
        a = torch.randn(M, K, generator=g).to(device).bfloat16()
        b = torch.randn(N, K, generator=g).to(device).bfloat16()
        ref = a.float() @ b.float().T
        lda = K
    elif mode == 1:
        a = torch.randn(K, M, generator=g).to(device).bfloat16()       # At
        b = torch.randn(K, N, generator=g).to(device).bfloat16()       # Bt
        ref = a.float().T @ b.float()
        lda = M
    elif mode == 2:
        R = 3 * M + 7
        lda = K + 64
        a = torch.randn(R, lda, generator=g).to(device).bfloat16()
        rowidx = torch.randint(0, R, (M,), generator=g, dtype=torch.int32).to(device)
        b = torch.randn(N, K, generator=g).to(device).bfloat16()
        ref = a[rowidx.long(), :K].float() @ b.float().T
    else:
        R = 2 * K + 5
        lda = ((M + 63) // 64) * 64
        a = torch.randn(R, lda, generator=g).to(device).bfloat16()
        rowidx = torch.randint(0, R, (K,), generator=g, dtype=torch.int32).to(device)
        b = torch.randn(K, N, generator=g).to(device).bfloat16()
        ref = a[rowidx.long(), :M].float().T @ b.float()
    rc = lib.minppo_debug_gemm(mode, a.data_ptr(), b.data_ptr(), None if rowidx is None else rowidx.data_ptr(),
                               c.data_ptr(), M, N, K, lda, splits, stream)
    _lib.check(rc)
    torch.cuda.synchronize(device)
    got = c.sum(0)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    return err


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(128, 256, 256, 1), (256, 256, 512, 2), (128, 64, 64, 1), (384, 128, 1024, 4),
                                   (128, 192, 192, 3)])
def test_umma_gemm_matches_fp32(mode, shape, cuda_device):
    M, N, K, splits = shape
    err = _run(mode, M, N, K, splits, cuda_device, seed=mode * 100 + M + N + K)
    assert err < 1e-4, f"mode {mode} shape {shape}: relative error {err:.3e}"
