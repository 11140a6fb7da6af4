"""GPU parity of the public surface that round 1 left untested ("What's weak" 1-3): the drop-in
``sample_full_swag`` / 5-planet trio loop, ``predict_trios``, ``SWAGModel.validation_step`` / ``validation_epoch_end``,
``load_ensemble``, the C host-buffer entry ``bnn_multiswag_predict_host`` called with plain host memory, a full-size
(B = 2000) gradient against the oracle's autograd, the 1e7-evaluation tensor-core-vs-FFMA accuracy check, and the
sharded entry points under a real 2-rank NCCL process group."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, make_swag_model, rel_err, swag_stats
from bnn_chaos_model_b200 import _lib, synth
from bnn_chaos_model_b200 import spock_reg_model as S
from bnn_chaos_model_b200._lib import TrainHParams
from bnn_chaos_model_b200.multiswag import MultiSWAG, load_ensemble
from oracle import restatement as R

pytestmark = pytest.mark.gpu
SEEDS = (0, 3, 17)
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def cpu_stats(seed):
    st = swag_stats(seed)
    return tuple(torch.from_numpy(st[k]) for k in ("w_avg", "w2_avg", "pre_D"))


def _replay_forward_swag_fast(spec, seed_model, x_cpu, dev):
    """The oracle fed the draws forward_swag_fast makes next on `dev` (randn(1,d), randn(K,1), randn(B,L) x 2)."""
    B = x_cpu.shape[0]
    z1 = torch.randn((1, spec.d), device=dev).cpu()
    z2 = torch.randn((30, 1), device=dev).cpu()
    e1 = torch.randn((B, 20), device=dev).cpu()
    e2 = torch.randn((B, 20), device=dev).cpu()
    theta = R.sample_weights(*cpu_stats(seed_model), 30, 0.5, z1, z2)
    return R.forward_swag_fast(spec, theta, x_cpu, e1, e2)


def test_sample_full_swag_vs_oracle(dev):
    """figures/spock/regression.py:74-92 / main_figures.py:127-139: a uniformly random ensemble member (numpy's global
    RNG), its weights sampled, forward_swag_fast -- replayed on the oracle with the same numpy and torch streams."""
    ens = MultiSWAG([make_swag_model(s, dev) for s in SEEDS], device=dev)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    x = torch.from_numpy(synth.make_systems(33, seed=101))
    np.random.seed(12)
    torch.manual_seed(34)
    got = [ens.sample_full_swag(x.to(dev)).detach().cpu() for _ in range(6)]
    np.random.seed(12)
    torch.manual_seed(34)
    picks = []
    for k in range(6):
        i = np.random.randint(0, len(SEEDS))
        picks.append(i)
        ref = _replay_forward_swag_fast(spec, SEEDS[i], x, dev)
        assert got[k].shape == (33, 2) and rel_err(got[k], ref) < TOL, k
    assert len(set(picks)) > 1  # the stream did visit different members


def test_five_planet_trio_loop_vs_oracle(dev):
    """figures/multiswag_5_planet.py:280-298: X [N, 3, T, F] -> ssX (done by the caller) -> [3N, T, F] -> per weight
    sample 10 chunks, each through sample_full_swag -> [S, N, 3, 2].  ``MultiSWAG.sample_trios`` is that loop; every
    (sample, chunk) block is checked against the oracle fed the same numpy / torch streams."""
    ens = MultiSWAG([make_swag_model(s, dev) for s in SEEDS], device=dev)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    N, Rt, n_samples = 14, 3, 2
    X = torch.from_numpy(synth.make_systems(N * Rt, seed=102)).reshape(N, Rt, 100, 41)
    np.random.seed(5)
    torch.manual_seed(6)
    got = ens.sample_trios(X.to(dev), n_samples).cpu()
    assert got.shape == (n_samples, N, Rt, 2)
    np.random.seed(5)
    torch.manual_seed(6)
    flat = X.reshape(-1, 100, 41)
    for s in range(n_samples):
        rows = []
        for part in torch.chunk(flat, chunks=10):
            i = np.random.randint(0, len(SEEDS))
            rows.append(_replay_forward_swag_fast(spec, SEEDS[i], part, dev))
        ref = torch.cat(rows).reshape(N, Rt, 2)
        assert rel_err(got[s], ref) < TOL, s


def test_predict_trios_batched_vs_oracle(dev):
    """The batched form of the same computation: every (model, weight sample) unit evaluates all 3N trio rows in one
    launch (Philox draws keyed on (unit, trio row)); each trio row against R.forward_swag_fast, and the posterior
    summary of the trio pipeline (min over trios) against the oracle's posterior_stats on ORACLE predictions."""
    models = [make_swag_model(s, dev) for s in SEEDS[:2]]
    ens = MultiSWAG(models, device=dev)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    N, Rt, S_, seed = 11, 3, 4, 77
    X = torch.from_numpy(synth.make_systems(N * Rt, seed=103)).reshape(N, Rt, 100, 41)
    got = ens.predict_trios(X.to(dev), S_, seed=seed).cpu()
    U = len(models) * S_
    assert got.shape == (U, N, Rt, 2)
    units = np.arange(U)
    z1 = torch.from_numpy(R.draw_z1(seed, units, 7583)); z2 = torch.from_numpy(R.draw_z2(seed, units, 30))
    eps = torch.from_numpy(R.draw_eps(seed, units, np.arange(N * Rt), 40))
    flat = X.reshape(-1, 100, 41)
    ref = torch.empty((U, N * Rt, 2))
    for u in units:
        th = R.sample_weights(*cpu_stats(SEEDS[u // S_]), 30, 0.5, z1[u], z2[u])
        ref[u] = R.forward_swag_fast(spec, th, flat, eps[u, :, :20], eps[u, :, 20:])
    assert rel_err(got.reshape(U, N * Rt, 2), ref) < TOL
    # trio pipeline: the deterministic columns of the [N, 8] summary against the oracle's statistics of its own predictions
    stats = ens.posterior_summary(flat.to(dev), S_, n_trios=Rt, seed=seed).cpu().numpy()
    pred_o = ref.reshape(U, N, Rt, 2).numpy()
    want = R.posterior_stats(pred_o[..., 0], pred_o)
    np.testing.assert_allclose(stats[:, 6], want["median_mu"], rtol=2e-5)
    np.testing.assert_allclose(stats[:, 7], want["median_std"], rtol=2e-5)


@pytest.mark.parametrize("noisy", [True, False])
def test_validation_step_vs_oracle(dev, noisy):
    """SWAGModel.validation_step (:787-799): lossfnc at the current weights and, after the w_avg swap, at w_avg (the
    current weights are restored), both / test_len, with hparams['noisy_val']; validation_epoch_end sums and gates
    aggregate_model on hparams['swa_start'] (:801-813)."""
    m = make_swag_model(3, dev)
    spec = R.ModelSpec.from_hparams(swag_stats(3)["hparams"])
    m.hparams["noisy_val"] = noisy
    g = torch.Generator().manual_seed(1)
    cur = (m.w_avg.cpu() + 0.01 * torch.randn(7583, generator=g)).to(dev)
    m.load(cur)
    B = 48
    x = torch.from_numpy(synth.make_systems(B, seed=104)); y = torch.from_numpy(synth.make_labels(B, seed=104))
    torch.manual_seed(8)
    res = m.validation_step((x.to(dev), y.to(dev)), 0)
    assert torch.equal(m.flatten(), cur)  # the swap is undone (:796-797)
    torch.manual_seed(8)
    want = []
    for theta in (cur.cpu(), m.w_avg.cpu()):
        e_in = torch.randn_like(x.to(dev)).cpu() if noisy else None
        e1 = torch.randn((B, 20), device=dev).cpu(); e2 = torch.randn((B, 20), device=dev).cpu()
        es = torch.randn((B, 40), device=dev).cpu() if noisy else None
        out, _ = R.forward(spec, theta, x, noisy, e_in, e1, e2, es)
        want.append(float(R.lossfnc_per_system(out, y).sum()) / 8740)
    assert float(res["val_loss"]) == pytest.approx(want[0], rel=1e-5)
    assert float(res["swa_loss"]) == pytest.approx(want[1], rel=1e-5)
    # epoch end: sums of the step outputs; collection only once global_step > hparams['swa_start']
    n0 = m.n_models
    m.hparams["swa_start"], m.global_step, m.current_epoch = 10, 10, 0
    out = m.validation_epoch_end([res, res])
    assert float(out["log"]["val_loss_no_reg"]) == pytest.approx(2 * want[0], rel=1e-5)
    assert float(out["log"]["swa_loss_no_reg"]) == pytest.approx(2 * want[1], rel=1e-5)
    assert m.n_models == n0
    m.global_step = 11
    m.validation_epoch_end([res])
    assert m.n_models == n0 + 1


def test_load_ensemble_round_trip(dev, tmp_path):
    """[load_swag(f) for f in glob(...)] (main_figures.py:39-42) through save_swag files -> MultiSWAG."""
    paths = []
    for s in SEEDS[:2]:
        p = str(tmp_path / f"x_v50_{s}_output.pkl")
        S.save_swag(make_swag_model(s, dev), p)
        paths.append(p)
    ens = load_ensemble(paths, device=dev)
    ref = MultiSWAG([make_swag_model(s, dev) for s in SEEDS[:2]], device=dev)
    assert ens.n_models == 2 and torch.equal(ens.w_avg, ref.w_avg) and torch.equal(ens.pre_D, ref.pre_D)
    x = torch.from_numpy(synth.make_systems(10, seed=105)).to(dev)
    assert torch.equal(ens.predict(x, 3, seed=1), ref.predict(x, 3, seed=1))
    assert abs(ens.ssX.mean_[0] - 4954.58585) < 1e-4


def test_c_host_entry_with_plain_host_buffers(dev):
    """bnn_multiswag_predict_host: the entry a non-torch caller binds -- numpy (pageable) host buffers in and out, device
    statistics, caller-provided scratch -- equals the device-resident MultiSWAG.predict bit for bit."""
    lib = _lib.load()
    ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(17, dev)], device=dev)
    cfg = ens.config(100)
    S_, seed = 5, 21
    for N in (57, 333, 1280):   # one chunk; four pipelined chunks with a ragged tail; four chunks, whole tiles
        xh = np.ascontiguousarray(synth.make_systems(N, seed=106))
        U = ens.n_models * S_
        out_h = np.full((U, N, 2), np.nan, np.float32)
        nbytes = lib.bnn_multiswag_host_scratch_bytes(cfg, N, U)
        assert nbytes > xh.nbytes
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.bnn_multiswag_predict_host(cfg, xh.ctypes.data_as(ctypes.c_void_p), N, _lib.ptr(ens.w_avg),
                                                _lib.ptr(ens.w2_avg), _lib.ptr(ens.pre_D), ens.n_models, ens.K, S_, 0.5, seed,
                                                out_h.ctypes.data_as(ctypes.c_void_p), scratch.data_ptr(),
                                                _lib.current_stream_ptr())
        _lib.check(rc, "bnn_multiswag_predict_host")
        want = ens.predict(torch.from_numpy(xh).to(dev), S_, seed=seed).cpu().numpy()
        assert np.array_equal(out_h, want), N
    # argument errors come back as negative codes with a message, not as a crash
    rc = lib.bnn_multiswag_predict_host(cfg, None, N, _lib.ptr(ens.w_avg), _lib.ptr(ens.w2_avg), _lib.ptr(ens.pre_D),
                                        ens.n_models, ens.K, S_, 0.5, seed, out_h.ctypes.data_as(ctypes.c_void_p),
                                        scratch.data_ptr(), None)
    assert rc == -1 and b"null" in lib.bnn_last_error_string()


def test_full_size_gradient_vs_oracle_autograd(dev):
    """BASELINE configs[3]'s batch (B = 2000) for one seed, on the automatically selected (tensor-core) kernel: loss,
    logged scalars, the FULL gradient and theta after the clipped SGD-momentum step against the oracle's autograd fed the
    kernel's own Philox draws (bnn_train_noise).  Tolerances: gradient 1e-4 of its max-norm (measured 3e-6), every parameter
    block 3e-4 of its own max (input_noise_logvar, a difference of large single-pass-TF32 terms: 2e-3), theta 1e-6 / 1e-7."""
    lib = _lib.load()
    m = make_swag_model(0, dev)
    cfg = m.config(100)
    spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
    B = 2000
    x = torch.from_numpy(synth.make_systems(B, seed=107)).to(dev)
    y = torch.from_numpy(synth.make_labels(B, seed=107)).to(dev)
    theta = m.w_avg[None].clone().contiguous()
    e_in = torch.empty((1, B, 100, 41), device=dev); e12 = torch.empty((1, B, 40), device=dev); e_sum = torch.empty((1, B, 40), device=dev)
    _lib.check(lib.bnn_train_noise(cfg, 1, B, 31, 4, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
    hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=0.1 * 7583, beta_in=m.beta_in, beta_out=m.beta_out,
                      first_step=1, apply_update=1)
    theta_before = theta.clone()
    mom = torch.zeros_like(theta)
    grad = torch.empty_like(theta); met = torch.zeros((1, 8), device=dev)
    ws = torch.empty((lib.bnn_train_workspace_bytes(cfg, B, 1) + 3) // 4, device=dev)
    _lib.check(lib.bnn_train_step(cfg, hp, 1, _lib.ptr(theta), _lib.ptr(mom), _lib.ptr(x), _lib.ptr(y), None, B, None, None, None,
                                  31, 4, _lib.ptr(grad), _lib.ptr(met), _lib.ptr(ws), _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    theta_after, theta = theta, theta_before
    torch.set_num_threads(os.cpu_count() or 1)
    th = theta[0].cpu().clone().requires_grad_(True)
    total, logs = R.training_loss(spec, th, x.cpu(), y.cpu(), e_in[0].cpu(), e12[0, :, :20].cpu(), e12[0, :, 20:].cpu(), e_sum[0].cpu())
    (gref,) = torch.autograd.grad(total, th)
    assert float(met[0, 1]) * B == pytest.approx(float(total), rel=1e-5)
    assert float(met[0, 0]) == pytest.approx(float(logs["train_loss_no_reg"]), rel=1e-5)
    assert float(met[0, 4]) == pytest.approx(float(gref.norm()), rel=1e-4)
    err = float((grad[0].cpu() - gref).abs().max() / gref.abs().max())
    print(f"B=2000 gradient vs oracle autograd: max err / max-norm = {err:.2e}")
    assert err < 1e-4, err
    for name, (off, shp) in spec.offsets().items():
        n = int(np.prod(shp))
        a, b = grad[0, off:off + n].cpu(), gref[off:off + n]
        tol = 2e-3 if name == "input_noise_logvar" else 3e-4
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-7, name
    # theta after the clipped SGD-momentum step (torch.optim.SGD semantics, first step): SWAG moments are averages of these
    th1, _, _ = R.clip_and_sgd_step(theta[0].cpu(), gref, None, 1e-4, 0.9, 1e-14, 0.1 * 7583, True)
    np.testing.assert_allclose(theta_after[0].cpu().numpy(), th1.numpy(), rtol=1e-6, atol=1e-7)
    assert float(met[0, 5]) < 1.0   # the clip was active (the reference's run clips almost every step at this batch size)


def test_tensor_core_accuracy_at_config2_size(dev):
    """1e7 evaluations (BASELINE configs[1]: 10,000 systems x 1,000 weight samples): the default (tensor-core, 3xTF32)
    kernel against the FP32 FFMA kernel, which is pinned to the oracle -- every (mu, std) within 1e-5 relative."""
    ens = MultiSWAG([make_swag_model(0, dev)], device=dev)
    N, S_ = 10000, 1000
    x = torch.from_numpy(synth.make_systems(N, seed=108)).to(dev)
    _, thp = ens.sample_thetas(S_, seed=5)
    lib = _lib.load()
    got = ens.predict(x, S_, seed=5, thp=thp)
    _lib.check(lib.bnn_set_predict_variant(2))   # FFMA2 v2
    try:
        want = ens.predict(x, S_, seed=5, thp=thp)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.bnn_set_predict_variant(0))
    rel = ((got - want).abs() / want.abs()).amax(dim=(1, 2))
    worst = float(rel.max())
    print(f"tensor-core vs FFMA over {N * S_:.0e} evals: max rel err {worst:.2e}, median of per-unit max {float(rel.median()):.2e}")
    assert worst < TOL, worst


_NCCL_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import make_swag_model
from bnn_chaos_model_b200 import synth
from bnn_chaos_model_b200.multiswag import MultiSWAG, shard_range
rank = int(sys.argv[1])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2, device_id=dev)
ens = MultiSWAG([make_swag_model(0, dev), make_swag_model(3, dev)], device=dev)
g = ens.system_granule()
for N, S_ in ((60, 4), (43, 3)):          # equal shards, then a ragged split (padded all_gather)
    x = torch.from_numpy(synth.make_systems(N, seed=N))
    lo, hi = shard_range(N, rank, 2, g)
    full = ens.predict_sharded(x[lo:hi].to(dev), N, S_, seed=9)
    single = ens.predict(x.to(dev), S_, seed=9, system_major=True)
    assert full.shape == (N, 2 * S_, 2) and torch.equal(full, single), (rank, N, "predict_sharded")
    if N == 60:   # equal shards: the gather in pieces under the next chunk's kernel, NCCL and peer-memory writes, twice
        for peer in (False, True, True):   # (the second peer-push call uses the other of its two buffers)
            import warnings
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                got = ens.predict_sharded(x[lo:hi].to(dev), N, S_, seed=9, overlap_chunks=3, peer_push=peer)
            if peer and w:
                print("peer-memory gather fell back to NCCL:", w[0].message)
            assert got.shape == (N, 2 * S_, 2) and torch.equal(got, single), (rank, N, "overlapped gather", peer)
        # deferred: batch k's predictions travel under batch k+1's kernel and are taken afterwards
        single10 = ens.predict(x.to(dev), S_, seed=10, system_major=True)
        for peer in (False, True):
            p1 = ens.predict_sharded(x[lo:hi].to(dev), N, S_, seed=9, peer_push=peer, defer=True)
            p2 = ens.predict_sharded(x[lo:hi].to(dev), N, S_, seed=10, peer_push=peer, defer=True)
            assert torch.equal(p1.result(), single) and torch.equal(p2.result(), single10), (rank, "deferred gather", peer)
    # 5-planet style rows (3 trios per system): only [N, 8] is gathered
    Ns = N // 3
    import math
    gs = g // math.gcd(g, 3)
    lo, hi = shard_range(Ns, rank, 2, gs)
    xt = x[: Ns * 3]
    st = ens.posterior_summary_sharded(xt[lo * 3:hi * 3].to(dev), Ns, S_, n_trios=3, seed=9)
    st1 = ens.posterior_summary(xt.to(dev), S_, n_trios=3, seed=9)
    assert st.shape == (Ns, 8) and torch.equal(st, st1), (rank, N, "posterior_summary_sharded")
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_entry_points_under_nccl_world2(tmp_path):
    """predict_sharded / posterior_summary_sharded under a REAL 2-rank NCCL group == the single-GPU result, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
