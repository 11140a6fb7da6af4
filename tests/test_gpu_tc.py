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


def rn_tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("R,MJ,NK,N,bias_round,two", [
    (104, 40, 41, 48, 1, 0),    # dW1-shaped: g_a2^T [h1 | 1], round-to-nearest by the +0x1000 bias
    (104, 20, 41, 48, 1, 1),    # dW2-shaped, accumulated across two commits
    (104, 40, 73, 80, 1, 0),    # dW0-shaped: g_a1^T [x' | n(live) | 1]
    (104, 40, 83, 96, 1, 1),
    (104, 40, 41, 48, 0, 0),    # raw fp32 operands: the tensor core truncates the 13 low mantissa bits
    (8, 5, 3, 16, 1, 0),
])
def test_tcgen05_ss_row_contraction(R, MJ, NK, N, bias_round, two):
    """The training kernel's weight-gradient GEMM: both operands in shared memory, contraction over the rows, M = 128
    with surplus feature rows reading neighbouring data (their lanes are ignored)."""
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator().manual_seed(R * 1000 + MJ)
    G = torch.randn(R, MJ, generator=g)
    Hm = torch.randn(R, NK, generator=g).abs()
    D = torch.full((128, N), float("nan"), device=dev)
    Gd, Hd = G.to(dev).contiguous(), Hm.to(dev).contiguous()
    _lib.check(lib.bnn_tc_probe_ss(_lib.ptr(Gd), _lib.ptr(Hd), _lib.ptr(D), R, MJ, NK, N, bias_round, two,
                                   _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    rnd = rn_tf32 if bias_round else tf32_exact
    ref = (rnd(G).double().T @ rnd(Hm).double()).float()
    got = D.cpu()[:MJ, :NK]
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    print(f"R={R} MJ={MJ} NK={NK} N={N} bias={bias_round} two={two}: err={err:.3e} scale={scale:.2f}")
    assert err < 2e-6 * scale * max(1.0, (R / 8) ** 0.5), err
    # and the single-pass tf32 product is within the stated 1-term tolerance of the exact fp32 product
    exact = (G.double().T @ Hm.double()).float()
    if bias_round:
        assert float((got - exact).abs().max()) < 2e-3 * scale
