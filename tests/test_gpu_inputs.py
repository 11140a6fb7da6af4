"""GPU parity of the input packing kernel (K6) against the reference's data_setup_kernel golden and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from bnn_chaos_model_b200 import synth
from bnn_chaos_model_b200.inputs import data_setup_kernel, pack_trios
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    """Units in the last place between two fp32 arrays (same sign assumed where it matters)."""
    ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


def test_pack_inputs_vs_reference_golden():
    dev = torch.device("cuda:0")
    z = load_golden("pack.npz")
    x = data_setup_kernel(torch.from_numpy(z["masses"]).to(dev), torch.from_numpy(z["tseries"]).to(dev)).cpu().numpy()
    ref = z["x_ref"]
    assert x.shape == ref.shape == (6, 100, 41) and np.isfinite(x).all()
    # float64 arithmetic on both sides; CUDA's and numpy's double cos/sin may differ in the last double ulp, which can
    # flip the final fp32 rounding: at most 1 fp32 ulp, and only in the 18 cos/sin columns
    d = ulp_diff(x, ref)
    assert d.max() <= 1
    non_angle = [c for c in range(41) if not (11 <= c <= 16 or 20 <= c <= 25 or 29 <= c <= 34)]
    assert d[..., non_angle].max() == 0
    assert (d > 0).mean() < 1e-3


def test_pack_inputs_large_vs_oracle_and_feeds_model():
    dev = torch.device("cuda:0")
    raw = synth.raw_systems(3000, seed=77)
    ts, ms = raw[:, :, :26].copy(), raw[:, 0, 26:29].copy()
    rng = np.random.default_rng(0)
    bad = rng.integers(0, ts.size, 500)
    ts.reshape(-1)[bad] = rng.choice([np.nan, np.inf, -np.inf], 500)
    x = data_setup_kernel(torch.from_numpy(ms).to(dev), torch.from_numpy(ts).to(dev))
    ref = R.pack_inputs(ms, ts, synth.SSX_MEAN, synth.SSX_SCALE)
    assert ulp_diff(x.cpu().numpy(), ref).max() <= 1
    # trio flattening (multiswag_5_planet.py:287): [N, 3, T, 26] -> [3N, T, 41]
    xt = pack_trios(torch.from_numpy(ts).to(dev).reshape(1000, 3, 100, 26), torch.from_numpy(ms).to(dev).reshape(1000, 3, 3))
    assert torch.equal(xt, x)
    with pytest.raises(NotImplementedError):
        data_setup_kernel(torch.zeros(2, 3, device=dev), torch.zeros(2, 100, 25, device=dev))
