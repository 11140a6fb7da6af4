"""GPU parity tests of the MultiSWAG posterior-predictive path (K1 sampler + K2 fused predict),
all through the C ABI.  Tolerances: per-system mu / std within 1e-5 relative (fp32), theta within
1e-6 (north_star)."""
import numpy as np
import pytest
import torch

from conftest import make_swag_model, rel_err, swag_stats
from bnn_chaos_model_b200 import _lib, synth
from bnn_chaos_model_b200.multiswag import MultiSWAG, shard_range
from oracle import restatement as R

pytestmark = pytest.mark.gpu
SEEDS = (0, 3, 17)
TOL = 1e-5
PV = {"auto": 0, "tc": 1, "v2": 2, "v1": 3}   # bnn_set_predict_variant (include/bnnchaos_diag.h)


@pytest.fixture
def predict_variant():
    """Process-wide diagnostic override of the predictive kernel, reset to automatic after the test."""
    lib = _lib.load()

    def set_(name):
        _lib.check(lib.bnn_set_predict_variant(PV[name]))

    yield set_
    _lib.check(lib.bnn_set_predict_variant(0))
    _lib.check(lib.bnn_set_predict_unit_chunk(0))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def cpu_stats(seed):
    st = swag_stats(seed)
    return tuple(torch.from_numpy(st[k]) for k in ("w_avg", "w2_avg", "pre_D"))


@pytest.mark.parametrize("seed", SEEDS)
def test_sampler_explicit_draws_vs_reference_theta(gold_predict, dev, seed):
    ens = MultiSWAG([make_swag_model(seed, dev)], device=dev)
    z1 = torch.from_numpy(gold_predict[f"z1_s{seed}"]).to(dev)
    z2 = torch.from_numpy(gold_predict[f"z2_s{seed}"]).to(dev)
    theta, thp = ens.sample_thetas(z1.shape[0], seed=0, scale=0.5, z1=z1, z2=z2)
    ref = torch.from_numpy(gold_predict[f"theta_ref_s{seed}"])
    # element-wise terms are computed with the reference's roundings; only the K=30 dot
    # product of D z2 may differ in summation order
    np.testing.assert_allclose(theta.cpu().numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)
    assert float((theta.cpu() == ref).float().mean()) > 0.5


def test_pack_theta_layout(gold_predict, dev):
    m = make_swag_model(0, dev)
    cfg = m.config()
    theta = torch.from_numpy(gold_predict["theta_ref_s0"]).to(dev)
    thp = m._packed(cfg, theta).cpu()
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    live = [c for c in range(41) if c not in spec.zero_cols]
    for u in range(theta.shape[0]):
        p = R.unflatten(spec, theta[u].cpu())
        W0 = p["feature_nn.0.weight"]
        o = 0
        W0p = thp[u, o:o + len(live) * 48].reshape(len(live), 4, 12); o += len(live) * 48
        assert torch.equal(W0p[:, :, :10].reshape(len(live), 40), W0[:, live].T)
        assert float(W0p[:, :, 10:].abs().max()) == 0.0
        b0p = thp[u, o:o + 48].reshape(4, 12); o += 48
        assert torch.equal(b0p[:, :10].reshape(40), p["feature_nn.0.bias"])
        W1p = thp[u, o:o + 40 * 48].reshape(40, 4, 12); o += 40 * 48 + 48
        assert torch.equal(W1p[:, :, :10].reshape(40, 40), p["feature_nn.2.weight"].T)
        W2p = thp[u, o:o + 40 * 48].reshape(40, 4, 12); o += 40 * 48
        W2 = p["feature_nn.4.weight"]  # [20,40]
        assert torch.equal(W2p[:, :, 0:10:2].reshape(40, 20), W2.T)
        assert torch.equal(W2p[:, :, 1:10:2].reshape(40, 20), W2.T)
    assert thp.shape[1] == _lib.load().bnn_packed_param_count(cfg)


def _sample_both(ens, n_units, unit_offset, S_, seed, unit_model=None, z=None, want_flat=True):
    """(theta, packed) of bnn_swag_sample (fused kernel) and of the unfused two-launch diagnostic entry."""
    lib = _lib.load()
    cfg = ens.config()
    M, d = ens.w_avg.shape
    P = lib.bnn_packed_param_count(cfg)
    dev = ens.device
    um = None if unit_model is None else torch.as_tensor(unit_model, dtype=torch.int32, device=dev)
    z1, z2 = (None, None) if z is None else z
    outs = []
    for fn, flat in ((lib.bnn_swag_sample, want_flat), (lib.bnn_swag_sample_unfused, True)):
        theta = torch.full((n_units, d), float("nan"), device=dev) if flat else None
        thp = torch.full((n_units, P), float("nan"), device=dev)
        _lib.check(fn(cfg, _lib.ptr(ens.w_avg), _lib.ptr(ens.w2_avg), _lib.ptr(ens.pre_D), M, ens.K, _lib.ptr(um), n_units,
                      unit_offset, S_, 0.5, seed, _lib.ptr(z1), _lib.ptr(z2), _lib.ptr(theta), _lib.ptr(thp),
                      _lib.current_stream_ptr()))
        outs.append((theta, thp))
    torch.cuda.synchronize()
    return outs


@pytest.mark.parametrize("n_units,unit_offset,S_", [(1, 0, 1), (7, 3, 2), (601, 5, 3), (1200, 0, 400), (333, 1000, 7)])
def test_fused_sampler_equals_unfused_bitwise(dev, n_units, unit_offset, S_):
    """bnn_swag_sample = ONE fused launch (pre_D tiles shared by the units of a CTA, packed layout written from shared
    memory); it must reproduce the two-launch sampler + pack bit for bit: 1 / 2 / 4 units per CTA, ragged last CTA, CTAs
    whose units straddle two models (S_ = 3, 7), unit chunks (unit_offset), Philox draws."""
    ens = MultiSWAG([make_swag_model(s, dev) for s in SEEDS], device=dev)
    (th_f, thp_f), (th_u, thp_u) = _sample_both(ens, n_units, unit_offset, S_, seed=11)
    assert torch.equal(th_f, th_u)
    assert torch.equal(thp_f, thp_u)          # also: no NaN fill left, every packed float is written
    (none, thp_only), _ = _sample_both(ens, n_units, unit_offset, S_, seed=11, want_flat=False)
    assert none is None and torch.equal(thp_only, thp_u)


@pytest.mark.parametrize("K", [20, 7, 2])
def test_fused_sampler_other_deviation_ranks(dev, K):
    """The fused sampler moves the pre_D tiles in 16-, 8- or 4-byte words depending on K (K = 30 -> 8 bytes: every other
    test); K = 20 takes the 16-byte path, odd K the 4-byte path, K = 2 is the smallest rank the reference's
    sqrt(2 (K - 1)) allows.  Three models (an odd model index shifts the tile base by d K floats), bit-equal to the
    unfused sampler + pack."""
    models = [make_swag_model(s, dev) for s in SEEDS]
    for m in models:
        m.pre_D = m.pre_D[:, :K].contiguous()
        m.K = K
        m.swa_params["K"] = K
    ens = MultiSWAG(models, device=dev)
    assert ens.K == K and ens.pre_D.shape[2] == K
    for n_units, S_ in ((601, 3), (9, 5)):
        (th_f, thp_f), (th_u, thp_u) = _sample_both(ens, n_units, 2, S_, seed=23)
        assert bool(torch.isfinite(th_u).all()) and torch.equal(th_f, th_u) and torch.equal(thp_f, thp_u), (K, n_units)


def test_fused_sampler_explicit_draws_and_unit_model(dev):
    ens = MultiSWAG([make_swag_model(s, dev) for s in SEEDS], device=dev)
    g = torch.Generator(device="cpu").manual_seed(4)
    U, d = 37, ens.w_avg.shape[1]
    z = (torch.randn((U, d), generator=g).to(dev), torch.randn((U, ens.K), generator=g).to(dev))
    um = torch.randint(0, 3, (U,), generator=g).tolist()
    for unit_model, zz in ((um, None), (None, z), (um, z)):
        (th_f, thp_f), (th_u, thp_u) = _sample_both(ens, U, 0, 13, seed=2, unit_model=unit_model, z=zz)
        assert torch.equal(th_f, th_u) and torch.equal(thp_f, thp_u)


def test_predict_strided_unit_chunks_fill_a_block(dev):
    """bnn_predict_strided: units sampled and evaluated chunk by chunk into column blocks of one [N, U, 2] array equal
    the one-shot system-major prediction bit for bit (Philox keyed on global unit / system indices)."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
    S_, N = 45, 35
    x = torch.from_numpy(synth.make_systems(N, seed=8)).to(dev)
    full = ens.predict(x, S_, seed=6, system_offset=20, system_major=True)
    U = 2 * S_
    block = torch.full((N, U, 2), float("nan"), device=dev)
    for u0 in range(0, U, 32):
        u1 = min(u0 + 32, U)
        _, thp = ens.sample_thetas(S_, 6, unit_offset=u0, n_units=u1 - u0, want_flat=False)
        ens.predict_into(x, thp, block, u0, seed=6, system_offset=20)
    assert torch.equal(block, full)


@pytest.mark.parametrize("seed", SEEDS)
def test_predict_vs_reference_golden(gold_predict, dev, seed):
    """theta, eps and the 256 in-distribution systems of the reference-generated golden file."""
    m = make_swag_model(seed, dev)
    cfg = m.config()
    x = torch.from_numpy(synth.make_systems(256, seed=123)).to(dev)
    theta = torch.from_numpy(gold_predict[f"theta_ref_s{seed}"]).to(dev)
    eps = torch.from_numpy(gold_predict[f"eps_s{seed}"]).to(dev).contiguous()
    out, _ = m._predict(x, m._packed(cfg, theta), eps, cfg=cfg)
    ref = torch.from_numpy(gold_predict[f"out_ref_s{seed}"])
    assert rel_err(out.cpu(), ref) < TOL


@pytest.mark.parametrize("seed", SEEDS)
def test_forward_swag_fast_dropin_same_torch_draws(dev, seed):
    """Same torch CUDA generator state -> same draws as the reference's call order (randn(1,d),
    randn(K,1), randn_like x2) -> same prediction as the oracle fed those draws."""
    m = make_swag_model(seed, dev)
    spec = R.ModelSpec.from_hparams(swag_stats(seed)["hparams"])
    B = 77  # ragged: not a multiple of the 8-system tile
    x = torch.from_numpy(synth.make_systems(B, seed=5))
    torch.manual_seed(100 + seed)
    out = m.forward_swag_fast(x.to(dev), scale=0.5).cpu()
    torch.manual_seed(100 + seed)
    z1 = torch.randn((1, spec.d), device=dev).cpu()
    z2 = torch.randn((30, 1), device=dev).cpu()
    e1 = torch.randn((B, 20), device=dev).cpu()
    e2 = torch.randn((B, 20), device=dev).cpu()
    theta = R.sample_weights(*cpu_stats(seed), 30, 0.5, z1, z2)
    np.testing.assert_allclose(m.flatten().cpu().numpy(), theta.numpy(), rtol=1e-6, atol=1e-7)  # weights stay loaded
    ref = R.forward_swag_fast(spec, theta, x, e1, e2)
    assert out.shape == (B, 2)
    assert rel_err(out, ref) < TOL


def test_forward_swag_records_summary_kl(dev):
    m = make_swag_model(0, dev)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    x = torch.from_numpy(synth.make_systems(40, seed=6))
    torch.manual_seed(1)
    out = m.forward_swag(x.to(dev), scale=0.5).cpu()
    theta = m.flatten().cpu()
    torch.manual_seed(1)
    torch.randn((1, spec.d), device=dev); torch.randn((30, 1), device=dev)
    e1 = torch.randn((40, 20), device=dev).cpu(); e2 = torch.randn((40, 20), device=dev).cpu()
    ref, skl = R.forward(spec, theta, x, False, None, e1, e2, None)
    assert rel_err(out, ref) < TOL
    assert float(m.summary_kl()) == pytest.approx(float(skl.sum()), rel=1e-5)


def test_forward_noisy_and_split_api(dev):
    """VarModel.forward(noisy_val=True) draw order: eps_in, eps1, eps2, eps_sum; and the split
    compute_summary_stats / predict_instability pair."""
    m = make_swag_model(3, dev)
    m.load(m.w_avg.clone())
    spec = R.ModelSpec.from_hparams(swag_stats(3)["hparams"])
    B = 24
    x = torch.from_numpy(synth.make_systems(B, seed=8))
    theta = m.flatten().cpu()
    torch.manual_seed(9)
    out = m.forward(x.to(dev), noisy_val=True).cpu()
    skl = float(m.summary_kl())
    torch.manual_seed(9)
    eps_in = torch.randn_like(x.to(dev)).cpu()
    e1 = torch.randn((B, 20), device=dev).cpu(); e2 = torch.randn((B, 20), device=dev).cpu()
    es = torch.randn((B, 40), device=dev).cpu()
    ref, skl_ref = R.forward(spec, theta, x, True, eps_in, e1, e2, es)
    assert rel_err(out, ref) < TOL
    assert skl == pytest.approx(float(skl_ref.sum()), rel=1e-5)
    # split API: summary stats of the un-masked input, then the head
    torch.manual_seed(10)
    s = m.compute_summary_stats(x.to(dev))
    torch.manual_seed(10)
    e1 = torch.randn((B, 20), device=dev).cpu(); e2 = torch.randn((B, 20), device=dev).cpu()
    p = R.unflatten(spec, theta)
    s_ref = R.compute_summary_stats(spec, p, x, e1, e2)
    # 1e-5 (north_star), relative to the magnitude of what is pooled: the sampled std half element-wise; the sampled mean half
    # -- averages of latent series that may cancel to ~0 -- against |mean| + std of the same latent column
    got, want = s.cpu().numpy(), s_ref.numpy()
    np.testing.assert_allclose(got[:, 20:], want[:, 20:], rtol=1e-5, atol=0)
    assert (np.abs(got[:, :20] - want[:, :20]) <= 1e-5 * (np.abs(want[:, :20]) + want[:, 20:])).all()
    mu, sd = m.predict_instability(s)
    mu_ref, sd_ref = R.predict_instability(spec, p, s.cpu())
    assert mu.shape == (B, 1) and rel_err(mu.cpu(), mu_ref) < TOL and rel_err(sd.cpu(), sd_ref) < TOL


def test_batched_philox_vs_oracle(dev):
    """In-kernel Philox draws (sampler z1/z2, predict eps) against the oracle's restatement."""
    models = [make_swag_model(s, dev) for s in SEEDS]
    ens = MultiSWAG(models, device=dev)
    S_, N, seed = 3, 41, 2024
    x = torch.from_numpy(synth.make_systems(N, seed=12))
    theta, thp = ens.sample_thetas(S_, seed)
    got = ens.predict(x.to(dev), S_, seed=seed, thp=thp).cpu()
    assert got.shape == (len(SEEDS) * S_, N, 2)
    units = np.arange(len(SEEDS) * S_)
    z1 = torch.from_numpy(R.draw_z1(seed, units, 7583)); z2 = torch.from_numpy(R.draw_z2(seed, units, 30))
    eps = torch.from_numpy(R.draw_eps(seed, units, np.arange(N), 40))
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    for u in units:
        th = R.sample_weights(*cpu_stats(SEEDS[u // S_]), 30, 0.5, z1[u], z2[u])
        np.testing.assert_allclose(theta[u].cpu().numpy(), th.numpy(), rtol=1e-5, atol=1e-6)
        ref = R.forward_swag_fast(spec, theta[u].cpu(), x, eps[u, :, :20], eps[u, :, 20:])
        assert rel_err(got[u], ref) < TOL


def test_sharded_equals_single_bitwise(dev):
    """Counter-based draws keyed on global indices: evaluating shards with system_offset gives the
    single-launch result bit for bit; system-major output is the transpose."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(17, dev)], device=dev)
    N, S_ = 61, 5
    x = torch.from_numpy(synth.make_systems(N, seed=13)).to(dev)
    full = ens.predict(x, S_, seed=4)
    g = ens.system_granule()
    assert g in (1, 5)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_range(N, r, world, g)
            assert lo % g == 0
            parts.append(ens.predict(x[lo:hi].contiguous(), S_, seed=4, system_offset=lo, system_major=True))
        got = torch.cat(parts, 0)
        assert torch.equal(got, full.permute(1, 0, 2))


def test_unit_chunked_launches_equal_single_launch(dev, predict_variant):
    """bnn_predict walks many units in L2-sized chunks (60,000 units at BASELINE configs[2]); the chunking must not show
    in the results: both output layouts, the explicit-eps path and the summary output, bit for bit."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev), make_swag_model(17, dev)], device=dev)
    N, S_ = 43, 9   # 27 units
    x = torch.from_numpy(synth.make_systems(N, seed=19)).to(dev)
    _, thp = ens.sample_thetas(S_, 6)
    ref_um = ens.predict(x, S_, seed=6, thp=thp)
    ref_sm = ens.predict(x, S_, seed=6, thp=thp, system_major=True)
    lib = _lib.load()
    cfg = ens.config(100)
    U = thp.shape[0]
    eps = torch.randn((U, N, 40), device=dev)
    def explicit():
        out = torch.empty((U, N, 2), device=dev); summ = torch.empty((U, N, 40), device=dev)
        _lib.check(lib.bnn_predict(cfg, _lib.ptr(x), N, _lib.ptr(thp), U, _lib.ptr(eps), None, 0, 0, 0, 0, _lib.ptr(out),
                                   _lib.ptr(summ), None, _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        return out, summ
    ref_e, ref_s = explicit()
    for chunk in (4, 5, 13):
        _lib.check(lib.bnn_set_predict_unit_chunk(chunk))
        assert torch.equal(ens.predict(x, S_, seed=6, thp=thp), ref_um), chunk
        assert torch.equal(ens.predict(x, S_, seed=6, thp=thp, system_major=True), ref_sm), chunk
        o, sm_ = explicit()
        assert torch.equal(o, ref_e) and torch.equal(sm_, ref_s), chunk
    assert torch.equal(ref_sm, ref_um.permute(1, 0, 2))


def test_repeated_predictions_are_bit_identical(dev):
    """The tensor-core kernel hands tiles between epilogue, MMA-issuing and tail warps through mbarriers, named barriers
    and a weight ring; compute-sanitizer is not available on the pool, so a race would have to show here: 30 repetitions
    of a multi-tile, multi-unit launch (ragged last tile) agree bit for bit, for both output layouts."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
    N, S_ = 1523, 37
    x = torch.from_numpy(synth.make_systems(N, seed=29)).to(dev)
    _, thp = ens.sample_thetas(S_, 12)
    ref = ens.predict(x, S_, seed=12, thp=thp)
    ref_sm = ens.predict(x, S_, seed=12, thp=thp, system_major=True)
    assert bool(torch.isfinite(ref).all()) and torch.equal(ref_sm, ref.permute(1, 0, 2))
    for rep in range(30):
        assert torch.equal(ens.predict(x, S_, seed=12, thp=thp), ref), rep
    assert torch.equal(ens.predict(x, S_, seed=12, thp=thp, system_major=True), ref_sm)


def test_edge_cases(dev):
    m = make_swag_model(0, dev)
    m.load(m.w_avg.clone())
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    theta = m.flatten().cpu()
    # (a) one system, (b) NaN / Inf in a zeroed column poisons that system only, (c) T != 100
    for T in (100, 40, 8, 128):
        B = 5
        x = torch.from_numpy(synth.make_systems(B, seed=20 + T, t=T))
        torch.manual_seed(T)
        out = m.forward(x.to(dev), noisy_val=False).cpu()
        torch.manual_seed(T)
        e1 = torch.randn((B, 20), device=dev).cpu(); e2 = torch.randn((B, 20), device=dev).cpu()
        ref, _ = R.forward(spec, theta, x, False, None, e1, e2, None)
        assert rel_err(out, ref) < TOL, T
    x = torch.from_numpy(synth.make_systems(3, seed=30))
    x[1, 17, 3] = float("nan")
    x[2, 5, 38] = float("inf")
    out = m.forward(x.to(dev), noisy_val=False).cpu()
    assert torch.isfinite(out[0]).all() and torch.isnan(out[1]).all() and torch.isnan(out[2]).all()
    out1 = m.forward(x[:1].to(dev), noisy_val=False)
    assert out1.shape == (1, 2)
    with pytest.raises(ValueError):
        m.forward(torch.zeros(2, 100, 40, device=dev), noisy_val=False)
    # fewer than K recorded deviations: same failure mode as the reference's D @ z_2 (:835)
    m.pre_D = m.pre_D[:, :10]
    with pytest.raises(RuntimeError):
        m.sample_weights(0.5)


def test_fp16_layer1_input_range(dev, predict_variant):
    """The tensor-core kernel stages x as fp16 hi / lo: a system with |x| >= 2^15 in a live column is scaled down by a
    power of two (and its layer-1 accumulator scaled back), tiny inputs lose nothing that matters.  Both against the fp32
    FFMA kernel (same Philox draws) at 1e-5; the out-of-range system must not cost its tile neighbours any precision."""
    ens = MultiSWAG([make_swag_model(0, dev)], device=dev)
    S_, N = 6, 20
    base = synth.make_systems(N, seed=61)
    live = [c for c in range(41) if c not in R.ModelSpec.from_hparams(swag_stats(0)["hparams"]).zero_cols]
    big = base.copy()
    big[7, :, live[3]] *= 3.0e5          # system 7 (tile 1) far out of distribution: |x| ~ 1e6
    big[12, 40:60, live[10]] = -7.0e7    # system 12 (tile 2): a burst of huge values
    tiny = base * 1.0e-6
    outs = {}
    for name in ("tc", "v2"):
        predict_variant(name)
        outs[name] = [ens.predict(torch.from_numpy(a).to(dev), S_, seed=4) for a in (base, big, tiny)]
    for k, label in enumerate(("base", "big", "tiny")):
        assert bool(torch.isfinite(outs["tc"][k]).all()), label
        assert rel_err(outs["tc"][k].cpu(), outs["v2"][k].cpu()) < TOL, label
    # other tiles (5 systems each) are untouched bit for bit; the tile neighbours of an out-of-range system keep their
    # precision (a neighbour that shares a 32-row block with it gets its layer-1 bias from the epilogue instead of the MMA:
    # last-bit differences only)
    other_tiles = [n for n in range(N) if n // 5 not in (1, 2)]
    assert torch.equal(outs["tc"][1][:, other_tiles], outs["tc"][0][:, other_tiles])
    neighbours = [n for n in range(5, 15) if n not in (7, 12)]
    assert rel_err(outs["tc"][1][:, neighbours].cpu(), outs["tc"][0][:, neighbours].cpu()) < 2e-6
    assert not torch.equal(outs["tc"][1][:, 7], outs["tc"][0][:, 7])


def test_full_size_properties(dev):
    """BASELINE config-2-sized run (10k systems x 1000 samples would take the oracle hours):
    check size-independent properties at a large size + spot checks against the oracle."""
    ens = MultiSWAG([make_swag_model(0, dev)], device=dev)
    N, S_ = 10000, 64
    xh = synth.make_systems(N, seed=0)
    x = torch.from_numpy(xh).to(dev)
    theta, thp = ens.sample_thetas(S_, seed=77)
    out = ens.predict(x, S_, seed=77, thp=thp)
    assert out.shape == (S_, N, 2) and bool(torch.isfinite(out).all())
    mu, sd = out[..., 0], out[..., 1]
    assert float(mu.min()) >= 4.0 and float(mu.max()) <= 12.0 and float(sd.min()) >= 0.5 and float(sd.max()) <= 6.0
    # permutation equivariance over systems with explicit eps (no index-keyed draws)
    eps = torch.randn((S_, N, 40), device=dev)
    m = ens.models[0]
    cfg = m.config()
    a, _ = m._predict(x, thp, eps, cfg=cfg)
    # (bit-exact when whole position-independent groups are permuted; within tolerance for any permutation,
    # because the tensor-core kernel's pooling blocks depend on a system's slot in its 5-system tile)
    g = ens.system_granule()
    perm = (torch.randperm(N // g, device=dev)[:, None] * g + torch.arange(g, device=dev)[None, :]).reshape(-1)
    b, _ = m._predict(x[perm].contiguous(), thp, eps[:, perm].contiguous(), cfg=cfg)
    assert torch.equal(a[:, perm], b)
    perm = torch.randperm(N, device=dev)
    b, _ = m._predict(x[perm].contiguous(), thp, eps[:, perm].contiguous(), cfg=cfg)
    assert rel_err(b, a[:, perm]) < TOL
    # spot check 3 units x 64 systems against the oracle
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    idx = torch.arange(0, N, N // 64)[:64]
    for u in (0, 31, 63):
        ref = R.forward_swag_fast(spec, theta[u].cpu(), torch.from_numpy(xh)[idx], eps[u, idx, :20].cpu(),
                                  eps[u, idx, 20:].cpu())
        assert rel_err(a[u, idx].cpu(), ref) < TOL


def test_kernel_variants_agree_bitwise(dev, predict_variant):
    """The synchronous kernel (v1) and the warp-specialised TMA/mbarrier kernel (v2) share the arithmetic order:
    outputs must be identical, also on ragged sizes and when an item's unit range is split into chunks."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
    for N, S_ in ((13, 3), (203, 40), (1200, 70)):
        x = torch.from_numpy(synth.make_systems(N, seed=N)).to(dev)
        _, thp = ens.sample_thetas(S_, seed=N)
        outs = {}
        for v in ("v1", "v2"):
            predict_variant(v)
            outs[v] = ens.predict(x, S_, seed=N, thp=thp)
        torch.cuda.synchronize()
        assert torch.equal(outs["v2"], outs["v1"]), (N, S_)


TC_VARIANTS = ("tc",)


@pytest.mark.parametrize("variant", TC_VARIANTS + ("v2", "v1"))
def test_every_kernel_vs_reference_golden(gold_predict, dev, predict_variant, variant):
    """The tcgen05 3xTF32 kernel and both FFMA kernels: same golden theta / eps / systems, same 1e-5 tolerance."""
    predict_variant(variant)
    for seed in SEEDS:
        m = make_swag_model(seed, dev)
        cfg = m.config()
        x = torch.from_numpy(synth.make_systems(256, seed=123)).to(dev)
        theta = torch.from_numpy(gold_predict[f"theta_ref_s{seed}"]).to(dev)
        eps = torch.from_numpy(gold_predict[f"eps_s{seed}"]).to(dev).contiguous()
        out, _ = m._predict(x, m._packed(cfg, theta), eps, cfg=cfg)
        ref = torch.from_numpy(gold_predict[f"out_ref_s{seed}"])
        assert rel_err(out.cpu(), ref) < TOL, (variant, seed)


def test_tensor_core_variants_vs_fp32_kernel(dev, predict_variant):
    """Ragged sizes, unit chunking, many units per CTA (exercises the record ring back-pressure and the weight
    ring) and a NaN-poisoned system: every tc variant against the FFMA kernel, which is pinned to the oracle."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
    for N, S_ in ((13, 3), (203, 40), (1200, 70), (745, 33)):
        xh = synth.make_systems(N, seed=N)
        xh[N // 2, 7, 3] = float("nan")  # zeroed column: poisons that system only
        x = torch.from_numpy(xh).to(dev)
        _, thp = ens.sample_thetas(S_, seed=N)
        predict_variant("v1")
        want = ens.predict(x, S_, seed=N, thp=thp)
        ok = torch.isfinite(want[:, :, 0])
        assert not bool(ok[:, N // 2].any()) and bool(ok[:, : N // 2].all())
        for v in TC_VARIANTS:
            predict_variant(v)
            for rep in range(2):  # twice: races show up as run-to-run differences
                got = ens.predict(x, S_, seed=N, thp=thp)
                assert bool(torch.isnan(got[:, N // 2]).all()), (v, N)
                assert rel_err(got[ok], want[ok]) < TOL, (v, N, S_, rep)


def test_wide_tensor_core_path_all_41_columns(dev, predict_variant):
    """Flag sets with more than 32 live input columns (here include_mmr / nan / eplusminus = True: 40 live columns; the
    noisy forward feeds all 41) run the WIDE tensor-core variant (layer 1 as K = 40 + a second 8-column
    pass): selected by default, equal to the FFMA kernel within the tolerance on ragged sizes with many units per CTA,
    a NaN system poisoned alone, and equal to the oracle on a small case."""
    from bnn_chaos_model_b200 import spock_reg_model as S

    models = []
    for seed in (0, 3):
        st = swag_stats(seed)
        hp = dict(st["hparams"], include_mmr=True, include_nan=True, include_eplusminus=True)
        m = S.SWAGModel(hp).init_params(st["swa_params"]).to(dev)
        m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(st[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
        models.append(m)
    assert 41 - len(models[0].zero_columns()) > 32 and 36 not in models[0].zero_columns()   # (fix_megno still zeroes column 7)
    ens = MultiSWAG(models, device=dev)
    assert ens.system_granule() == 5          # = the tensor-core kernel's tile: it is what bnn_predict selects
    for N, S_ in ((13, 3), (203, 40), (745, 33)):
        xh = synth.make_systems(N, seed=100 + N)
        xh[N // 2, 7, 36] = float("nan")      # a live column here: that system's outputs are NaN, nobody else's
        x = torch.from_numpy(xh).to(dev)
        _, thp = ens.sample_thetas(S_, seed=N)
        predict_variant("v1")
        want = ens.predict(x, S_, seed=N, thp=thp)
        ok = torch.isfinite(want[:, :, 0])
        assert not bool(ok[:, N // 2].any()) and bool(ok[:, : N // 2].all())
        predict_variant("auto")
        for rep in range(2):
            got = ens.predict(x, S_, seed=N, thp=thp)
            assert bool(torch.isnan(got[:, N // 2]).all()), N
            assert rel_err(got[ok], want[ok]) < TOL, (N, S_, rep)
    # against the oracle: the dense model's forward with explicit draws
    m = models[0]
    m.load(m.w_avg.clone())
    spec = R.ModelSpec.from_hparams(m.hparams)
    x = torch.from_numpy(synth.make_systems(23, seed=77))
    torch.manual_seed(4)
    out = m.forward(x.to(dev), noisy_val=False).cpu()
    torch.manual_seed(4)
    e1 = torch.randn((23, 20), device=dev).cpu(); e2 = torch.randn((23, 20), device=dev).cpu()
    ref, _ = R.forward(spec, m.flatten().cpu(), x, False, None, e1, e2, None)
    assert rel_err(out, ref) < TOL


def test_predict_host_pipelined_equals_single_launch(dev):
    """Host-buffer entry with chunked H2D / compute / D2H overlap: bit-identical to one device-resident launch."""
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(17, dev)], device=dev)
    N, S_ = 203, 6
    xh = torch.from_numpy(synth.make_systems(N, seed=31)).pin_memory()
    want = ens.predict(xh.to(dev), S_, seed=12, system_major=True).cpu()
    for n_chunks in (1, 3, 8, 64, (0.04, 0.47, 0.47, 0.02), (0.5, 0.5)):
        got = ens.predict_host(xh, S_, seed=12, n_chunks=n_chunks)
        assert got.shape == (N, 2 * S_, 2) and torch.equal(got, want), n_chunks
    out = torch.empty((N, 2 * S_, 2)).pin_memory()
    assert ens.predict_host(xh, S_, seed=12, out_host=out) is out and torch.equal(out, want)
