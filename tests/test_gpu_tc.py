"""GPU: pins the tcgen05 / TMEM conventions (descriptors, TMEM layouts) the tensor-core kernel relies on."""
import numpy as np
import pytest
import torch

from bnn_chaos_model_b200 import _lib

pytestmark = pytest.mark.gpu


def tf32_exact(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def run_probe(K, N, variant, seed=0):
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    A = tf32_exact(torch.randn(128, K, generator=g))
    B = tf32_exact(torch.randn(N, K, generator=g))
    D = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.to(dev).contiguous(), B.to(dev).contiguous()  # keep alive: ptr() of a temporary dangles
    _lib.check(lib.bnn_tc_probe(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), K, N, variant,
                                _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T).float()
    return float((D.cpu() - ref).abs().max()), float(ref.abs().max())


@pytest.mark.parametrize("K,N", [(8, 16), (32, 48), (48, 48), (48, 32), (96, 48)])
def test_tcgen05_tf32_gemm_matches(K, N):
    # variant 0 = LBO is the byte distance between the two 16-byte K chunks, SBO between 8-row groups
    # (variant 1, the swapped reading, faults with an illegal address and is not exercised here)
    err0, scale = run_probe(K, N, 0)
    print(f"K={K} N={N}: err={err0:.3e} scale={scale:.2f}")
    assert err0 < 1e-5 * scale, err0
