"""GPU parity of the posterior post-processing kernels (K7).  The order statistics are deterministic and compared
with numpy on identical samples (1e-6); the sampling stage uses counter-based Philox instead of numpy's global
RNG, so it is compared in distribution (Kolmogorov-Smirnov against the oracle's fast_truncnorm / prior table)."""
import numpy as np
import pytest
import torch
from scipy import stats as sps

from bnn_chaos_model_b200.posterior import STAT_NAMES, posterior_summary, sample_instability, summarize_instability
from oracle import restatement as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("N,Rt,U", [(7, 1, 2000), (5, 3, 100), (3, 3, 1), (2, 2, 4097)])
def test_summary_statistics_vs_numpy(dev, N, Rt, U):
    rng = np.random.default_rng(N * 100 + U)
    t = rng.uniform(4.0, 12.0, (N * Rt, U)).astype(np.float32)
    pred = np.stack([rng.uniform(4.0, 12.0, (N * Rt, U)), rng.uniform(0.5, 6.0, (N * Rt, U))], -1).astype(np.float32)
    got = summarize_instability(torch.from_numpy(t).to(dev), torch.from_numpy(pred).to(dev), Rt).cpu().numpy()
    samps = t.reshape(N, Rt, U).transpose(2, 0, 1)            # [U, N, R] like the reference's arrays
    p = pred.reshape(N, Rt, U, 2).transpose(2, 0, 1, 3)       # [U, N, R, 2]
    ref = R.posterior_stats(samps, p)
    for i, name in enumerate(STAT_NAMES):
        np.testing.assert_allclose(got[:, i], ref[name], rtol=2e-6, err_msg=name)


def test_radix_select_summary_for_many_weight_samples(dev):
    """U = 60000 (30 models x 2000 samples, BASELINE config 3) exceeds the shared-memory sort: exact radix select.
    Also cross-checked against the sort path at a size both support, with ties, negatives and a NaN."""
    N, Rt, U = 2, 2, 60000
    rng = np.random.default_rng(7)
    t = rng.uniform(4.0, 30.0, (N * Rt, U)).astype(np.float32)
    pred = np.stack([rng.uniform(4.0, 12.0, (N * Rt, U)), rng.uniform(0.5, 6.0, (N * Rt, U))], -1).astype(np.float32)
    got = summarize_instability(torch.from_numpy(t).to(dev), torch.from_numpy(pred).to(dev), Rt).cpu().numpy()
    ref = R.posterior_stats(t.reshape(N, Rt, U).transpose(2, 0, 1), pred.reshape(N, Rt, U, 2).transpose(2, 0, 1, 3))
    for i, name in enumerate(STAT_NAMES):
        np.testing.assert_allclose(got[:, i], ref[name], rtol=2e-6, err_msg=name)
    U = 3001
    t = np.round(rng.uniform(-3.0, 12.0, (3, U)), 1).astype(np.float32)  # many ties, negative values
    pred = np.stack([np.round(rng.uniform(4, 12, (3, U)), 1), rng.uniform(0.5, 6, (3, U))], -1).astype(np.float32)
    pred[1, 5, 0] = np.nan
    a = summarize_instability(torch.from_numpy(t).to(dev), torch.from_numpy(pred).to(dev), 1)
    from bnn_chaos_model_b200 import _lib

    _lib.check(_lib.load().bnn_set_summary_variant(1))   # force the radix select (diagnostic switch)
    try:
        b = summarize_instability(torch.from_numpy(t).to(dev), torch.from_numpy(pred).to(dev), 1)
    finally:
        _lib.check(_lib.load().bnn_set_summary_variant(0))
    assert torch.equal(a[:, 1:], b[:, 1:]) and torch.allclose(a[:, 0], b[:, 0], rtol=1e-6)


def test_sampling_matches_reference_distribution(dev):
    # (a) one (mu, std) repeated: KS against the oracle's fast_truncnorm + prior resampling
    U = 200_000
    mu, sd = 7.5, 2.0
    pred = torch.tensor([mu, sd], device=dev).repeat(1, U, 1)
    t = sample_instability(pred, seed=11).cpu().numpy()[0]
    assert t.min() > 4.0 and t.max() <= 100.0
    rng = np.random.default_rng(5)
    ref = R.fast_truncnorm(np.full(U, mu, np.float32), np.full(U, sd, np.float32), left=4, nsamp=40, d=50000, rng=rng)
    past9 = ref >= 9
    ref[past9] = R.prior_samples_table(rng.random(int(past9.sum())))
    assert sps.ks_2samp(t, ref).pvalue > 1e-3
    # the two branches separately: truncated normal below 9, prior above
    below = t[t < 9]
    a, b = (4 - mu) / sd, (9 - mu) / sd
    assert sps.kstest(below, sps.truncnorm(a, b, loc=mu, scale=sd).cdf).pvalue > 1e-3
    frac9 = (1 - sps.norm.cdf(b)) / (1 - sps.norm.cdf(a))
    assert abs((t >= 9).mean() - frac9) < 5 * np.sqrt(frac9 * (1 - frac9) / U)
    assert sps.kstest(t[t >= 9], R.prior_cdf).pvalue > 1e-3
    # (b) no draw can pass: the first draw is returned (mask.argmax == 0), exactly z0*std + mu
    lowpred = torch.tensor([-200.0, 1.0], device=dev).repeat(1, 1000, 1)
    low = sample_instability(lowpred, seed=3).cpu().numpy()[0]
    assert (low < 4).all() and abs(low.mean() + 200) < 0.2 and abs(low.std() - 1) < 0.1


def test_sampling_is_keyed_on_global_indices(dev):
    """Shards (row_offset) reproduce the single-call draws bit for bit; different seeds differ."""
    g = torch.Generator().manual_seed(0)
    pred = torch.stack([torch.rand((12, 300), generator=g) * 8 + 4, torch.rand((12, 300), generator=g) * 5 + 0.5], -1).to(dev)
    full = sample_instability(pred, seed=9)
    parts = torch.cat([sample_instability(pred[:5].contiguous(), seed=9), sample_instability(pred[5:].contiguous(), seed=9, row_offset=5)])
    assert torch.equal(full, parts)
    assert not torch.equal(full, sample_instability(pred, seed=10))
    s = posterior_summary(pred, n_trios=3, seed=9)
    assert s.shape == (4, 8) and bool(torch.isfinite(s).all())
    assert bool((s[:, 5] <= s[:, 3]).all() and (s[:, 3] <= s[:, 1]).all() and (s[:, 1] <= s[:, 2]).all() and (s[:, 2] <= s[:, 4]).all())
    with pytest.raises(ValueError):
        summarize_instability(full[:5], pred[:5], n_trios=3)


def test_ensemble_posterior_summary_end_to_end(dev):
    """raw 5-planet style input -> pack (K6) -> predict (K1+K2) -> sample + summarise (K7): sharded == single."""
    from conftest import make_swag_model
    from bnn_chaos_model_b200 import synth
    from bnn_chaos_model_b200.inputs import pack_trios
    from bnn_chaos_model_b200.multiswag import MultiSWAG, shard_range

    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
    N, Rt, S_ = 20, 3, 50
    raw = synth.raw_systems(N * Rt, seed=4)
    x = pack_trios(torch.from_numpy(raw[:, :, :26]).to(dev).reshape(N, Rt, 100, 26),
                   torch.from_numpy(raw[:, 0, 26:29]).to(dev).reshape(N, Rt, 3))
    full = ens.posterior_summary(x, S_, n_trios=Rt, seed=2)
    assert full.shape == (N, 8) and bool(torch.isfinite(full).all())
    assert bool((full[:, 1] >= 4).all()) and bool((full[:, 6] >= 4).all() and (full[:, 7] >= 0.5).all())
    g = ens.system_granule()
    parts = []
    for r in range(2):
        lo, hi = shard_range(N, r, 2, g)  # system boundaries at multiples of g keep rows aligned for any n_trios
        parts.append(ens.posterior_summary(x[lo * Rt:hi * Rt].contiguous(), S_, n_trios=Rt, seed=2, system_offset=lo))
    assert torch.equal(torch.cat(parts), full)
    # the deterministic columns agree with a host-side evaluation of the same predictions
    pred = ens.predict(x, S_, seed=2, system_major=True).cpu().numpy().reshape(N, Rt, 2 * S_, 2).transpose(2, 0, 1, 3)
    ref = R.posterior_stats(pred[..., 0], pred)
    np.testing.assert_allclose(full[:, 6].cpu().numpy(), ref["median_mu"], rtol=2e-6)
    np.testing.assert_allclose(full[:, 7].cpu().numpy(), ref["median_std"], rtol=2e-6)


def test_posterior_summary_chunking_is_invisible(dev):
    """posterior_summary walks the systems in chunks so that [rows, U, 2] never exists as a whole (BASELINE configs[2]:
    9 GB per GPU otherwise); the chunk size must not show in the result, bit for bit, for 1 and 3 trios."""
    from conftest import make_swag_model
    from bnn_chaos_model_b200 import synth
    from bnn_chaos_model_b200.multiswag import MultiSWAG

    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(17, dev)], device=dev)
    x = torch.from_numpy(synth.make_systems(3 * 47, seed=33)).to(dev)
    for n_trios, N in ((1, 141), (3, 47)):
        S_ = 20
        full = ens.posterior_summary(x, S_, n_trios=n_trios, seed=5, system_offset=10)
        assert full.shape == (N, 8)
        for budget in (12 * 40 * n_trios * 5, 12 * 40 * n_trios * 17, 12 * 40 * n_trios * 46):   # 5, 15, 45 systems per chunk
            got = ens.posterior_summary(x, S_, n_trios=n_trios, seed=5, system_offset=10, max_block_bytes=budget)
            assert torch.equal(got, full), (n_trios, budget)
        # units walked in chunks too, their weights sampled per (system chunk, unit chunk)
        for unit_chunk in (7, 16):
            got = ens.posterior_summary(x, S_, n_trios=n_trios, seed=5, system_offset=10, unit_chunk=unit_chunk,
                                        max_block_bytes=12 * 40 * n_trios * 17)
            assert torch.equal(got, full), (n_trios, unit_chunk)
