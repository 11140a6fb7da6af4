"""GPU parity: truncated-normal NLL fwd/bwd (K3) and SWAG moment collection (K5)."""
import json

import numpy as np
import pytest
import torch

from conftest import make_swag_model
from bnn_chaos_model_b200 import _lib
from bnn_chaos_model_b200 import spock_reg_model as S

pytestmark = pytest.mark.gpu


def test_nll_forward_backward_vs_reference(gold_loss):
    dev = torch.device("cuda:0")
    lib = _lib.load()
    testy = torch.from_numpy(gold_loss["testy"]).to(dev).contiguous()
    y = torch.from_numpy(gold_loss["y"]).to(dev).contiguous()
    B = testy.shape[0]
    loss = torch.empty(B, device=dev); tot = torch.empty(1, device=dev); grad = torch.empty((B, 2), device=dev)
    _lib.check(lib.bnn_nll_fwd_bwd(_lib.ptr(testy), _lib.ptr(y), B, _lib.ptr(loss), _lib.ptr(tot), _lib.ptr(grad),
                                   _lib.current_stream_ptr()))
    ref, gref = gold_loss["loss_ref"], gold_loss["grad_ref"]
    np.testing.assert_allclose(loss.cpu().numpy(), ref, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(grad.cpu().numpy(), gref, rtol=1e-4, atol=1e-5)
    assert float(tot) == pytest.approx(float(ref.astype(np.float64).sum()), rel=1e-5)
    # both branches and the x < -1 polynomial branch of safe_log_erf are exercised
    yy = gold_loss["y"]
    assert (yy >= 9).any() and (yy < 9).any()
    # mirror method
    m = make_swag_model(0, dev)
    np.testing.assert_allclose(m._lossfnc(testy, y).cpu().numpy(), ref, rtol=1e-5, atol=2e-6)
    # non-finite terms are replaced by 100 per label, not raised (:563-570)
    bad = torch.tensor([[float("nan"), 1.0]], device=dev)
    assert float(m._lossfnc(bad, torch.tensor([[5.0, 9.5]], device=dev))) == 200.0


def test_aggregate_model_trajectory(gold_aggregate):
    dev = torch.device("cuda:0")
    g = gold_aggregate
    hp = json.loads(str(g["hparams"]))
    m = S.SWAGModel(hp).init_params({"K": int(g["K"]), "c": int(g["c"]), "swa_lr": 1e-4, "swa_start": 0}).to(dev)
    ws = torch.from_numpy(g["ws"]).to(dev)
    for epoch in range(int(g["n_epochs"])):
        m.load(ws[epoch])
        m.current_epoch = epoch
        m.aggregate_model()
        np.testing.assert_allclose(m.w_avg.cpu().numpy(), g[f"w_avg_{epoch}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m.w2_avg.cpu().numpy(), g[f"w2_avg_{epoch}"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(m.pre_D.cpu().numpy(), g[f"pre_D_{epoch}"])  # pure copies: bit-exact
    assert m.n_models == int(g["n_epochs"]) and m.pre_D.shape[1] == int(g["K"])
