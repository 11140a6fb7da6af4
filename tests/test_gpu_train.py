"""GPU parity of the fused SWAG training step (K4), the validation loss and the multi-seed trainer, through the
C ABI.  Tolerances: logged losses 1e-5 relative; gradient 2e-4 of the gradient's max-norm (fp32 sums over
B*T rows in a different order than autograd); gradient norm 1e-5; theta after a step 1e-6 relative + 1e-7 absolute
on the FP32 kernel.  The tensor-core kernel's weight-gradient GEMMs are single-pass TF32 on round-to-nearest operands:
their unbiased rounding noise falls as 1 / sqrt(B T), so at the SMALL batches of these tests (B = 20..64, where
automatic selection runs the FP32 kernel) theta after a step is held to 1e-6 relative + 2e-6 absolute and the
input_noise_logvar block (a difference of large terms) to 2e-3 of its own max; at the reference's B = 2000
(tests/test_gpu_surface.py::test_full_size_gradient_vs_oracle_autograd) the tight tolerances hold for it too."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_swag_model, swag_stats
from bnn_chaos_model_b200 import _lib, synth
from bnn_chaos_model_b200._lib import TrainHParams
from bnn_chaos_model_b200.swag_train import MultiSeedSWAGTrainer, seeds_of_rank
from oracle import restatement as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _ws(lib, cfg, B, S, dev):
    return torch.empty((lib.bnn_train_workspace_bytes(cfg, B, S) + 3) // 4, device=dev)


def _step(lib, cfg, hp, S, theta, mom, x, y, idx, B, eps, seed, step, dev, want_grad=True):
    grad = torch.empty_like(theta) if want_grad else None
    metrics = torch.zeros((S, 8), device=dev)
    ws = _ws(lib, cfg, B, S, dev)
    e_in, e12, e_sum = eps if eps is not None else (None, None, None)
    _lib.check(lib.bnn_train_step(cfg, hp, S, _lib.ptr(theta), _lib.ptr(mom), _lib.ptr(x), _lib.ptr(y), _lib.ptr(idx), B,
                                  _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), seed, step, _lib.ptr(grad),
                                  _lib.ptr(metrics), _lib.ptr(ws), _lib.current_stream_ptr()), "bnn_train_step")
    torch.cuda.synchronize()
    return grad, metrics


VARIANTS = {"tc": 1, "v3": 2}


@pytest.fixture(params=["tc", "v3"], autouse=True)
def train_variant(request):
    """Every test of this file runs on both training kernels: tc -- the eight GEMMs of a system on tcgen05 (3xTF32 row
    GEMMs, single-pass round-to-nearest TF32 weight-gradient GEMMs), the default; v3 -- the FP32 FFMA kernel kept as the
    fallback.  The override is a process-wide diagnostic switch (include/bnnchaos_diag.h), reset to automatic afterwards."""
    lib = _lib.load()
    _lib.check(lib.bnn_set_train_variant(VARIANTS[request.param]))
    yield request.param
    _lib.check(lib.bnn_set_train_variant(0))


def test_train_steps_vs_reference_golden(gold_train, dev, train_variant):
    """Three SGD-momentum steps of the reference (autograd + clip_grad_norm_ + torch.optim.SGD) with all four
    noise tensors fixed: logged scalars, full gradient, gradient norm and theta after every step."""
    g = gold_train
    lib = _lib.load()
    m = make_swag_model(0, dev)
    cfg = m.config(100)
    B = int(g["B"])
    x = torch.from_numpy(synth.make_systems(B, seed=int(g["x_seed"]))).to(dev)
    y = torch.from_numpy(g["y"]).to(dev)
    theta = torch.from_numpy(g["theta0"]).to(dev)[None].contiguous()
    mom = torch.zeros_like(theta)
    for s in range(3):
        eps = (torch.from_numpy(g["eps_in"][s].astype(np.float32)).to(dev)[None].contiguous(),
               torch.from_numpy(g["eps12"][s]).to(dev)[None].contiguous(),
               torch.from_numpy(g["eps_sum"][s]).to(dev)[None].contiguous())
        hp = TrainHParams(lr=float(g["lr"]), momentum=float(g["momentum"]), weight_decay=float(g["weight_decay"]),
                          clip_norm=float(g["clip"]), beta_in=m.beta_in, beta_out=m.beta_out, first_step=int(s == 0),
                          apply_update=1)
        grad, met = _step(lib, cfg, hp, 1, theta, mom, x, y, None, B, eps, 0, s, dev)
        gref = g[f"grad_ref_{s}"]
        scale = np.abs(gref).max()
        err = np.abs(grad[0].cpu().numpy() - gref).max() / scale
        assert err < 2e-4, (s, err)
        np.testing.assert_allclose(met[0, :4].cpu().numpy(), g[f"logs_ref_{s}"], rtol=1e-5)
        assert float(met[0, 1]) * B == pytest.approx(float(g[f"loss_ref_{s}"]), rel=1e-5)
        assert float(met[0, 4]) == pytest.approx(float(g[f"gradnorm_ref_{s}"]), rel=1e-5)
        assert float(met[0, 6]) == 0.0
        np.testing.assert_allclose(theta[0].cpu().numpy(), g[f"theta_ref_{s}"], rtol=1e-6,
                                   atol=2e-6 if train_variant == "tc" else 1e-7)


def test_gradient_by_parameter_group(gold_train, dev, train_variant):
    """Every parameter tensor's gradient on its own scale (a wrong small block would hide under a global max-norm)."""
    g = gold_train
    lib = _lib.load()
    m = make_swag_model(0, dev)
    cfg = m.config(100)
    B = int(g["B"])
    x = torch.from_numpy(synth.make_systems(B, seed=int(g["x_seed"]))).to(dev)
    y = torch.from_numpy(g["y"]).to(dev)
    theta = torch.from_numpy(g["theta0"]).to(dev)[None].contiguous()
    eps = (torch.from_numpy(g["eps_in"][0].astype(np.float32)).to(dev)[None].contiguous(),
           torch.from_numpy(g["eps12"][0]).to(dev)[None].contiguous(),
           torch.from_numpy(g["eps_sum"][0]).to(dev)[None].contiguous())
    hp = TrainHParams(lr=0.0, momentum=0.0, weight_decay=0.0, clip_norm=1e30, beta_in=m.beta_in, beta_out=m.beta_out,
                      first_step=1, apply_update=0)
    grad, _ = _step(lib, cfg, hp, 1, theta, None, x, y, None, B, eps, 0, 0, dev)
    assert torch.equal(theta[0].cpu(), torch.from_numpy(g["theta0"]))  # apply_update=0 leaves the weights alone
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    got, ref = grad[0].cpu().numpy(), g["grad_ref_0"]
    for name, (off, shp) in spec.offsets().items():
        n = int(np.prod(shp))
        a, b = got[off:off + n], ref[off:off + n]
        tol = 2e-3 if (train_variant == "tc" and name == "input_noise_logvar") else 3e-4
        assert np.abs(a - b).max() <= tol * np.abs(b).max() + 1e-7, name


def test_training_step_dropin_backward_and_optimizer(dev, train_variant):
    """SWAGModel.training_step with torch draws in the reference's order; loss.backward() installs the fused
    gradient, clip_grad_norm_ + the SGD of configure_optimizers() then match the oracle's step."""
    m = make_swag_model(3, dev)
    m.load(m.w_avg.clone())
    spec = R.ModelSpec.from_hparams(swag_stats(3)["hparams"])
    B = 32
    x = torch.from_numpy(synth.make_systems(B, seed=41))
    y = torch.from_numpy(synth.make_labels(B, seed=41))
    theta0 = m.flatten().cpu()
    (opt,), _ = m.configure_optimizers()
    torch.manual_seed(5)
    res = m.training_step((x.to(dev), y.to(dev)), 0)
    res["loss"].backward()
    clip = 0.1 * 7583
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), clip)
    opt.step()
    # oracle with the same draws
    torch.manual_seed(5)
    eps_in = torch.randn_like(x.to(dev)).cpu()
    e1 = torch.randn((B, 20), device=dev).cpu(); e2 = torch.randn((B, 20), device=dev).cpu()
    es = torch.randn((B, 40), device=dev).cpu()
    th = theta0.clone().requires_grad_(True)
    total, logs = R.training_loss(spec, th, x, y, eps_in, e1, e2, es)
    (gref,) = torch.autograd.grad(total, th)
    assert float(res["loss"]) == pytest.approx(float(total), rel=1e-5)
    for k in ("train_loss_no_reg", "train_loss_with_reg", "input_kl", "summary_kl"):
        assert float(res["log"][k]) == pytest.approx(float(logs[k]), rel=1e-5), k
    assert float(gn) == pytest.approx(float(gref.norm()), rel=1e-4)
    th1, _, _ = R.clip_and_sgd_step(theta0, gref, None, m.swa_params["swa_lr"], m.hparams["momentum"],
                                    m.hparams["weight_decay"], clip, True)
    np.testing.assert_allclose(m.flatten().cpu().numpy(), th1.numpy(), rtol=1e-6, atol=2e-6 if train_variant == "tc" else 1e-7)


def test_philox_noise_and_multi_seed_batches(dev):
    """(a) NULL noise pointers == the draws bnn_train_noise writes out, bit for bit; (b) n_seeds models with
    per-seed batch indices == each seed alone on its gathered batch; (c) the step is bit-reproducible."""
    lib = _lib.load()
    S, B, N = 3, 23, 90  # odd batch: the last pair of v2 has an inactive slot
    models = [make_swag_model(s, dev) for s in (0, 3, 17)]
    cfg = models[0].config(100)
    x = torch.from_numpy(synth.make_systems(N, seed=51)).to(dev)
    y = torch.from_numpy(synth.make_labels(N, seed=51)).to(dev)
    theta0 = torch.stack([m.w_avg for m in models]).contiguous()
    idx = torch.stack([torch.randperm(N, generator=torch.Generator().manual_seed(s))[:B] for s in range(S)]).to(torch.int32).to(dev)
    hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                      first_step=1, apply_update=1)
    seed, step = 99, 7
    th_a, mom_a = theta0.clone(), torch.zeros_like(theta0)
    g_a, met_a = _step(lib, cfg, hp, S, th_a, mom_a, x, y, idx, B, None, seed, step, dev)
    e_in = torch.empty((S, B, 100, 41), device=dev); e12 = torch.empty((S, B, 40), device=dev); e_sum = torch.empty((S, B, 40), device=dev)
    _lib.check(lib.bnn_train_noise(cfg, S, B, seed, step, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
    assert abs(float(e_in.mean())) < 0.01 and abs(float(e_in.std()) - 1) < 0.01
    th_b, mom_b = theta0.clone(), torch.zeros_like(theta0)
    g_b, met_b = _step(lib, cfg, hp, S, th_b, mom_b, x, y, idx, B, (e_in, e12, e_sum), seed, step, dev)
    assert torch.equal(g_a, g_b) and torch.equal(th_a, th_b) and torch.equal(met_a, met_b)
    th_c, mom_c = theta0.clone(), torch.zeros_like(theta0)
    g_c, _ = _step(lib, cfg, hp, S, th_c, mom_c, x, y, idx, B, None, seed, step, dev)
    assert torch.equal(g_a, g_c) and torch.equal(th_a, th_c)
    assert not torch.equal(e_in[0], e_in[1])  # seeds draw different noise
    for s in range(S):  # each seed alone, batch gathered on the host side
        xs, ys = x[idx[s].long()].contiguous(), y[idx[s].long()].contiguous()
        th_s, mom_s = theta0[s:s + 1].clone(), torch.zeros((1, theta0.shape[1]), device=dev)
        g_s, met_s = _step(lib, cfg, hp, 1, th_s, mom_s, xs, ys, None, B,
                           (e_in[s:s + 1].contiguous(), e12[s:s + 1].contiguous(), e_sum[s:s + 1].contiguous()), 0, 0, dev)
        # a seed alone runs on more CTAs (different partial-sum grouping): equal to rounding, not bit for bit
        assert float((g_s[0] - g_a[s]).abs().max()) <= 2e-5 * float(g_a[s].abs().max())
        np.testing.assert_allclose(met_s[0, :5].cpu().numpy(), met_a[s, :5].cpu().numpy(), rtol=2e-5)
    # oracle cross-check of one seed with the Philox draws
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    th = theta0[0].cpu().clone().requires_grad_(True)
    total, _ = R.training_loss(spec, th, x[idx[0].long()].cpu(), y[idx[0].long()].cpu(), e_in[0].cpu(), e12[0, :, :20].cpu(),
                               e12[0, :, 20:].cpu(), e_sum[0].cpu())
    (gref,) = torch.autograd.grad(total, th)
    assert float((g_a[0].cpu() - gref).abs().max()) <= 3e-4 * float(gref.abs().max())
    assert float(met_a[0, 1]) * B == pytest.approx(float(total), rel=1e-5)


def test_eval_loss_vs_reference(gold_train, dev):
    """validation_step (:787-799): lossfnc(noisy_val=False) at two weight vectors in one call."""
    g = gold_train
    lib = _lib.load()
    m = make_swag_model(0, dev)
    cfg = m.config(100)
    B = int(g["B"])
    x = torch.from_numpy(synth.make_systems(B, seed=int(g["x_seed"]))).to(dev)
    y = torch.from_numpy(g["y"]).to(dev)
    thetas = torch.stack([torch.from_numpy(g["theta0"]), torch.from_numpy(g["theta_ref_2"])]).to(dev)
    thp = m._packed(cfg, thetas)
    eps = torch.from_numpy(g["eps12"][0]).to(dev)[None].repeat(2, 1, 1).contiguous()
    out = torch.empty((2, B, 2), device=dev); loss = torch.empty(2, device=dev)
    _lib.check(lib.bnn_eval_loss(cfg, _lib.ptr(x), _lib.ptr(y), B, _lib.ptr(thp), 2, _lib.ptr(eps), 0, _lib.ptr(out),
                                 _lib.ptr(loss), None, _lib.current_stream_ptr()))
    assert float(loss[0]) == pytest.approx(float(g["val_loss_ref"]), rel=1e-5)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    o1, _ = R.forward(spec, thetas[1].cpu(), x.cpu(), False, None, eps[1, :, :20].cpu(), eps[1, :, 20:].cpu(), None)
    assert float(loss[1]) == pytest.approx(float(R.lossfnc_per_system(o1, y.cpu()).sum()), rel=1e-5)
    # workspace variant (no prediction output)
    ws = torch.empty(2 * B * 2, device=dev); loss2 = torch.empty(2, device=dev)
    _lib.check(lib.bnn_eval_loss(cfg, _lib.ptr(x), _lib.ptr(y), B, _lib.ptr(thp), 2, _lib.ptr(eps), 0, None,
                                 _lib.ptr(loss2), _lib.ptr(ws), _lib.current_stream_ptr()))
    assert torch.equal(loss, loss2)


def test_multi_seed_trainer_collects_reference_moments(dev):
    """Trainer.fit replacement: per-epoch aggregate_model over the weights the fused steps produced; the
    collected moments must equal the oracle's aggregate_model replayed on the recorded weight trajectory."""
    torch.manual_seed(0)
    models = [make_swag_model(s, dev) for s in (0, 17)]
    for m in models:
        m.load(m.w_avg.clone())
        m.init_params({"K": 3, "c": 2, "swa_lr": 1e-4, "swa_start": 0})
        m.hparams["swa_start"] = 2
    N = 150
    X = torch.from_numpy(synth.make_systems(N, seed=61)); y = torch.from_numpy(synth.make_labels(N, seed=61))
    tr = MultiSeedSWAGTrainer(models, X[:120], y[:120], X[120:], y[120:], batch_size=50, device=dev, seed=4)
    assert len(tr.epoch_batches()) == 3 and tr.epoch_batches()[-1][1] == 20  # ragged last batch
    traj = []
    states = [R.SwagState(K=3, c=2) for _ in models]
    for epoch in range(6):
        logs = tr.fit(1)
        assert torch.isfinite(logs[0]["val_loss_no_reg"]).all()
        traj.append(tr.theta.cpu().clone())
        if tr.global_step > 2:
            for i in range(2):
                states[i] = R.aggregate_model(states[i], traj[-1][i], epoch)
    assert not torch.equal(traj[0], traj[-1])
    out = tr.export()
    for i, m in enumerate(out):
        st = states[i]
        assert m.n_models == st.n_models and m.pre_D.shape[1] == st.pre_D.shape[1] == 3
        np.testing.assert_allclose(m.w_avg.cpu().numpy(), st.w_avg.numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m.w2_avg.cpu().numpy(), st.w2_avg.numpy(), rtol=1e-6, atol=1e-7)
        assert torch.equal(m.pre_D.cpu(), st.pre_D)
    # the exported mirrors sample and predict
    o = out[0].forward_swag_fast(X[:8].to(dev), scale=0.5)
    assert o.shape == (8, 2) and bool(torch.isfinite(o).all())
    assert seeds_of_rank(30, 0, 8) == [0, 1, 2, 3] and seeds_of_rank(30, 7, 8) == [27, 28, 29]


def test_multi_seed_pretrainer_vs_oracle_schedule(dev, train_variant):
    """find_minima.py phase: fused steps under the custom one-cycle schedule (lr AND momentum per step) and the KL
    annealing, for two seeds at once; replayed step by step on the oracle (autograd + clip + SGD) with the Philox draws
    the kernel made; ends at the schedule's ValueError and restores the best-validation weights."""
    from bnn_chaos_model_b200.swag_train import MultiSeedPretrainer

    lib = _lib.load()
    models = [make_swag_model(s, dev) for s in (0, 3)]
    for m in models:
        m.load(m.w_avg.clone())
        m.steps = m.hparams["steps"] = 10      # schedule over int(0.9 * 10) = 9 optimizer steps
        m.lr = m.hparams["lr"] = 1e-3
    N, B = 60, 20
    X = torch.from_numpy(synth.make_systems(N, seed=71)); y = torch.from_numpy(synth.make_labels(N, seed=71))
    tr = MultiSeedPretrainer(models, X[:40], y[:40], X[40:], y[40:], batch_size=B, device=dev, seed=11)
    assert tr.total_sched == 9 and len(tr.epoch_batches()) == 2
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    cfg = models[0].config(100)
    th_o = [tr.theta[i].cpu().clone() for i in range(2)]
    buf_o = [torch.zeros_like(t) for t in th_o]
    d = th_o[0].numel()
    Xc, yc = X[:40], y[:40]
    n_checked = 0
    for epoch in range(3):
        for idx, Bb in tr.epoch_batches():
            g = tr.global_step
            lr, mom, b_in, b_out = tr.schedule()
            assert (lr, mom) == R.one_cycle(g, 1e-3, 9)
            assert (b_in, b_out) == pytest.approx(R.kl_annealing(g, 10, models[0].beta_in, models[0].beta_out))
            e_in = torch.empty((2, Bb, 100, 41), device=dev); e12 = torch.empty((2, Bb, 40), device=dev); e_sum = torch.empty((2, Bb, 40), device=dev)
            _lib.check(lib.bnn_train_noise(cfg, 2, Bb, tr.seed, g, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
            tr.train_step(idx, Bb)
            torch.cuda.synchronize()
            for i in range(2):
                sel = idx[i].long().cpu()
                th = th_o[i].clone().requires_grad_(True)
                total, _ = R.training_loss(spec, th, Xc[sel], yc[sel], e_in[i].cpu(), e12[i, :, :20].cpu(), e12[i, :, 20:].cpu(),
                                           e_sum[i].cpu(), beta_in=b_in, beta_out=b_out)
                (gr,) = torch.autograd.grad(total, th)
                th_o[i], buf_o[i], _ = R.clip_and_sgd_step(th_o[i], gr, buf_o[i], lr, mom, 1e-14, 0.1 * d, g == 0)
                np.testing.assert_allclose(tr.theta[i].cpu().numpy(), th_o[i].numpy(), rtol=1e-5,
                                           atol=2e-5 if train_variant == "tc" else 2e-6)   # B = 20, lr up to 1e-3
                th_o[i] = tr.theta[i].cpu().clone()   # no drift: every step is compared from the same start
                buf_o[i] = tr.momentum[i].cpu().clone()
            n_checked += 1
            if n_checked == 4:
                break
        if n_checked == 4:
            break
    assert n_checked == 4 and tr.global_step == 4
    logs = tr.fit()                       # runs to the end of the schedule
    assert tr.finished and tr.global_step == 10   # steps 0..9 ran, step 10 raised
    with pytest.raises(ValueError):
        tr.schedule(10)
    assert torch.isfinite(tr.best_val).all() and torch.equal(tr.theta, tr.best_theta)
    out = tr.export()
    assert torch.equal(out[1].flatten().to(dev), tr.theta[1])


def test_saliency_vs_oracle_autograd(dev):
    """feature_importance.py's gradforward: d mu / d x through compute_summary_stats (fixed eps1, eps2) and
    predict_instability at w_avg, against torch autograd on the oracle; then the ensemble entry (Philox eps)."""
    from bnn_chaos_model_b200.multiswag import feature_importance

    m = make_swag_model(3, dev)
    m.load(m.w_avg.clone())
    spec = R.ModelSpec.from_hparams(swag_stats(3)["hparams"])
    B = 37   # odd: the last pair has an inactive slot
    x = torch.from_numpy(synth.make_systems(B, seed=81)).to(dev)
    torch.manual_seed(5)
    g, mu, sumsq = m.gradforward(x)
    torch.manual_seed(5)
    eps1 = torch.randn((B, 20), device=dev).cpu(); eps2 = torch.randn((B, 20), device=dev).cpu()
    p = R.unflatten(spec, m.w_avg.cpu())
    xz = R.zero_columns(spec, x.cpu()).clone().requires_grad_(True)
    s = R.compute_summary_stats(spec, p, xz, eps1, eps2)
    mu_o, _ = R.predict_instability(spec, p, s)
    (g_o,) = torch.autograd.grad(mu_o.sum(), xz)
    np.testing.assert_allclose(mu.cpu().numpy(), mu_o.detach().reshape(-1).numpy(), rtol=1e-5)
    scale = float(g_o.abs().max())
    assert float((g.cpu() - g_o).abs().max()) <= 2e-5 * scale
    for c in range(41):   # every column on its own scale (the zeroed columns carry gradient too)
        assert float((g[:, :, c].cpu() - g_o[:, :, c]).abs().max()) <= 1e-4 * float(g_o[:, :, c].abs().max()) + 1e-9
    np.testing.assert_allclose(sumsq.cpu().numpy(), (g_o ** 2).sum((0, 1)).numpy(), rtol=1e-4)
    # ensemble: three models, Philox eps; importance is positive, finite, and reproducible
    models = [make_swag_model(sd, dev) for sd in (0, 3, 17)]
    imp, mus = feature_importance(models, x, seed=9)
    imp2, _ = feature_importance(models, x, seed=9)
    assert imp.shape == (3, 41) and mus.shape == (3, B) and torch.equal(imp, imp2)
    assert bool(torch.isfinite(imp).all()) and bool((imp > 0).all())
    assert 4.0 <= float(mus.min()) and float(mus.max()) <= 12.0


def test_full_size_step_properties(dev):
    """BASELINE configs[3] per GPU (4 seeds x batch 2000 of 8000 resident systems): (a) the step with in-kernel Philox
    noise is bit-reproducible and equals the step fed the same draws explicitly (bnn_train_noise); (b) the two
    independently written kernels agree on the full gradient and the logged scalars when fed the same explicit noise
    (the B = 2000 gradient is checked against the oracle's autograd in tests/test_gpu_surface.py); (c) seeds with
    equal weights but different batches differ."""
    lib = _lib.load()
    S, B, N = 4, 2000, 8000
    m = make_swag_model(0, dev)
    cfg = m.config(100)
    x = torch.from_numpy(synth.make_systems(N, seed=3)).to(dev)
    y = torch.from_numpy(synth.make_labels(N, seed=3)).to(dev)
    theta0 = m.w_avg[None].repeat(S, 1).contiguous()
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    idx = torch.stack([torch.randperm(N, device=dev, generator=gen)[:B] for _ in range(S)]).to(torch.int32).contiguous()
    hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                      first_step=1, apply_update=0)
    e_in = torch.empty((S, B, 100, 41), device=dev); e12 = torch.empty((S, B, 40), device=dev); e_sum = torch.empty((S, B, 40), device=dev)
    res = {}
    for v in ("tc", "v3"):
        _lib.check(lib.bnn_set_train_variant(VARIANTS[v]))
        g, met = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, idx, B, None, 5, 11, dev)
        g2, met2 = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, idx, B, None, 5, 11, dev)
        assert torch.equal(g, g2) and torch.equal(met, met2), v
        assert bool(torch.isfinite(g).all()) and bool((met[:, 6] == 0).all())
        if v == "tc":   # the tensor-core kernel's draws, written out: both kernels are then fed exactly these
            _lib.check(lib.bnn_train_noise(cfg, S, B, 5, 11, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
            assert not torch.equal(g[0], g[1])
        ge, mete = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, idx, B, (e_in, e12, e_sum), 5, 11, dev)
        if v == "tc":
            assert torch.equal(ge, g) and torch.equal(mete, met)
        res[v] = (ge, mete)
    for s in range(S):
        scale = float(res["v3"][0][s].abs().max())
        err = float((res["tc"][0][s] - res["v3"][0][s]).abs().max()) / scale
        assert err <= 5e-5, (s, err)
    np.testing.assert_allclose(res["tc"][1][:, :5].cpu().numpy(), res["v3"][1][:, :5].cpu().numpy(), rtol=2e-5)


def test_repeated_steps_are_bit_identical(dev, train_variant):
    """compute-sanitizer is not available on the pool: a race between the producer warps, the pipeline warps, the
    TMEM stash and the L2 scratch hand-off would show as run-to-run differences.  40 repetitions of the same step
    (odd batch, three seeds, several tiles per CTA) must agree bit for bit, gradients and metrics."""
    lib = _lib.load()
    S, B, N = 3, 613, 700
    m = make_swag_model(17, dev)
    cfg = m.config(100)
    x = torch.from_numpy(synth.make_systems(N, seed=91)).to(dev)
    y = torch.from_numpy(synth.make_labels(N, seed=91)).to(dev)
    theta0 = m.w_avg[None].repeat(S, 1).contiguous()
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    idx = torch.stack([torch.randperm(N, device=dev, generator=gen)[:B] for _ in range(S)]).to(torch.int32).contiguous()
    hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                      first_step=1, apply_update=0)
    g0, met0 = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, idx, B, None, 8, 2, dev)
    assert bool(torch.isfinite(g0).all())
    for rep in range(40):
        g, met = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, idx, B, None, 8, 2, dev)
        assert torch.equal(g, g0) and torch.equal(met, met0), (train_variant, rep)


def test_seed_groups_do_not_change_a_seed(dev, train_variant):
    """bnn_train_step runs many seeds in groups of launches when one launch would leave SMs idle (30 seeds: 4 CTAs each =
    120 of 148 SMs).  Every buffer and Philox key is indexed by the global seed, so (a) with the CTA count per seed held
    fixed (small batch: one CTA per system) any grouping, ragged ones included, gives bit-identical gradients, metrics,
    weights and momenta; (b) at 30 seeds x batch 2000 the cost model's plan (more CTAs per seed) agrees with the
    single-launch plan to rounding (a seed's gradient is summed over a different number of CTA partials)."""
    lib = _lib.load()
    m = make_swag_model(3, dev)
    cfg = m.config(100)
    try:
        S, B, N = 7, 12, 40
        x = torch.from_numpy(synth.make_systems(N, seed=77)).to(dev)
        y = torch.from_numpy(synth.make_labels(N, seed=77)).to(dev)
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        theta0 = (m.w_avg[None] + 1e-3 * torch.randn((S, m.w_avg.numel()), device=dev, generator=gen)).contiguous()
        idx = torch.stack([torch.randperm(N, device=dev, generator=gen)[:B] for _ in range(S)]).to(torch.int32).contiguous()
        hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                          first_step=1, apply_update=1)
        ref = None
        for groups in (1, 2, 3, 7):
            _lib.check(lib.bnn_set_train_seed_groups(groups))
            th, mom = theta0.clone(), torch.zeros_like(theta0)
            g, met = _step(lib, cfg, hp, S, th, mom, x, y, idx, B, None, 21, 4, dev)
            if ref is None:
                ref = (g, met, th, mom)
                assert bool(torch.isfinite(g).all()) and not torch.equal(g[0], g[1])
            else:
                for a, b in zip(ref, (g, met, th, mom)):
                    assert torch.equal(a, b), (train_variant, groups)
        S, B, N = 30, 2000, 2000
        x = torch.from_numpy(synth.make_systems(N, seed=78)).to(dev)
        y = torch.from_numpy(synth.make_labels(N, seed=78)).to(dev)
        theta0 = (m.w_avg[None] + 1e-3 * torch.randn((S, m.w_avg.numel()), device=dev, generator=gen)).contiguous()
        hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3,
                          first_step=1, apply_update=0)
        out = {}
        for groups in (1, 0):
            _lib.check(lib.bnn_set_train_seed_groups(groups))
            out[groups] = _step(lib, cfg, hp, S, theta0.clone(), None, x, y, None, B, None, 22, 1, dev)
        for s in range(S):
            err = float((out[0][0][s] - out[1][0][s]).abs().max() / out[1][0][s].abs().max())
            assert err <= 5e-5, (train_variant, s, err)   # the tolerance of the cross-kernel test above (measured: 2.2e-5 on tc)
        np.testing.assert_allclose(out[0][1][:, :5].cpu().numpy(), out[1][1][:, :5].cpu().numpy(), rtol=2e-5)
    finally:
        _lib.check(lib.bnn_set_train_seed_groups(0))


def test_trainer_noisy_validation_and_lr_milestone_vs_oracle(dev):
    """MultiSeedSWAGTrainer.validation_losses with the reference's default noisy_val=True (:787-799): loss at the current
    weights and at w_avg, / test_len, against the oracle's noisy forward fed the Philox draws of the launch; and the
    MultiStepLR milestone (:709-720) as a per-step learning rate."""
    lib = _lib.load()
    models = [make_swag_model(s, dev) for s in (0, 17)]
    for m in models:
        m.load(m.w_avg.clone())
        m.init_params({"K": 3, "c": 1, "swa_lr": 1e-4, "swa_start": 3, "swa_recording_lr_factor": 0.5})
        m.hparams["swa_start"] = 0
    N = 90
    X = torch.from_numpy(synth.make_systems(N, seed=63)); y = torch.from_numpy(synth.make_labels(N, seed=63))
    assert not MultiSeedSWAGTrainer(models, X[:60], y[:60], batch_size=30, device=dev).noisy_val   # v50 hparams: noisy_val False
    tr = MultiSeedSWAGTrainer(models, X[:60], y[:60], X[60:], y[60:], batch_size=30, device=dev, seed=6, noisy_val=True)
    assert tr.noisy_val and tr.test_len == 8740
    assert [tr.step_lr(g) for g in (0, 2, 3, 4)] == [1e-4, 1e-4, 5e-5, 5e-5]   # torch MultiStepLR([3], 0.5), stepped per step
    assert MultiSeedSWAGTrainer(models, X[:60], y[:60], batch_size=30, device=dev, lr_milestone=False).step_lr(100) == 1e-4
    tr.fit(2, validate=False)          # two epochs: moments exist, theta != w_avg
    assert int(tr.n_models[0]) == 2
    v, s = tr.validation_losses()
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    cfg = models[0].config(100)
    B, U = 30, 4
    e_in = torch.empty((U, B, 100, 41), device=dev); e12 = torch.empty((U, B, 40), device=dev); e_sum = torch.empty((U, B, 40), device=dev)
    _lib.check(lib.bnn_train_noise(cfg, U, B, tr.seed ^ 0x5EED, tr.current_epoch, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
    thetas = torch.cat([tr.theta, tr.w_avg]).cpu()
    got = torch.cat([v, s]).cpu()
    for u in range(U):
        out, _ = R.forward(spec, thetas[u], X[60:], True, e_in[u].cpu(), e12[u, :, :20].cpu(), e12[u, :, 20:].cpu(), e_sum[u].cpu())
        want = float(R.lossfnc_per_system(out, y[60:]).sum()) / 8740
        assert float(got[u]) == pytest.approx(want, rel=1e-5), u
    # noise-free variant: same normalisation
    tr.noisy_val = False
    v0, s0 = tr.validation_losses()
    assert bool(torch.isfinite(v0).all()) and not torch.equal(v0, v)
