"""CPU: host-side logic, the C-ABI surface, Philox known answers, and the N>1 gather (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, swag_stats
from oracle import restatement as R

HEADER = os.path.join(ROOT, "include", "bnnchaos.h")
LIB = os.path.join(ROOT, "bnn_chaos_model_b200", "libbnnchaos.so")


DIAG_HEADER = os.path.join(ROOT, "include", "bnnchaos_diag.h")


def declared_symbols(header=HEADER):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bnnchaos.h but not exported"
    from bnn_chaos_model_b200 import _lib

    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table covers the whole header
    diag = declared_symbols(DIAG_HEADER)  # probes / timers live in their own header, outside the product ABI
    assert not set(diag) & set(syms)
    for s in diag:
        assert hasattr(lib, s), f"{s} declared in include/bnnchaos_diag.h but not exported"
    assert sorted(_lib.DIAG_SIGNATURES) == diag
    assert _lib.load().bnn_abi_version() == 1


def test_layout_queries_and_config_errors():
    from bnn_chaos_model_b200 import _lib

    lib = _lib.load()
    cfg = _lib.ModelConfig(41, 40, 20, 1, 1, 100, 0x1C0000000FE, 4.0, 12.0, 0.5, 6.0)
    assert lib.bnn_param_count(cfg) == 7583
    P = lib.bnn_packed_param_count(cfg)
    assert P > 0 and P % 4 == 0
    bad = _lib.ModelConfig(41, 64, 20, 1, 1, 100, 0, 4.0, 12.0, 0.5, 6.0)
    assert lib.bnn_param_count(bad) == -2  # BNN_E_CONFIG
    assert b"hidden" in lib.bnn_last_error_string()


def test_training_seed_plan_host_query():
    """A CTA of the training kernels belongs to one seed, so many seeds are run in groups of launches when one launch
    would leave SMs idle (host-only query of the plan bnn_train_step uses; 148 SMs assumed without a device): the 3-4 seeds
    per GPU of an 8-GPU run stay one launch, all 30 seeds on one GPU become two launches of 15 seeds x 9 CTAs."""
    import ctypes as C

    from bnn_chaos_model_b200 import _lib

    lib = _lib.load()
    cfg = _lib.ModelConfig(41, 40, 20, 1, 1, 100, 0x1C0000000FE, 4.0, 12.0, 0.5, 6.0)

    def plan(B, S):
        a, g, p = C.c_int32(), C.c_int32(), C.c_int32()
        assert lib.bnn_train_seed_plan(cfg, B, S, C.byref(a), C.byref(g), C.byref(p)) == 0
        return a.value, g.value, p.value

    if torch.cuda.is_available() and torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("the expected plans are a B200's (148 SMs)")
    assert plan(2000, 4) == (37, 1, 4) and plan(2000, 3) == (49, 1, 3) and plan(2000, 1) == (148, 1, 1)
    assert plan(2000, 30) == (9, 2, 15)
    assert plan(2000, 15) == (9, 1, 15) and plan(2000, 8) == (18, 1, 8)
    assert plan(64, 30) == (4, 1, 30)                                 # short batches: the fixed cost per launch outweighs idle SMs
    for S in range(1, 41):                                            # every seed is covered, never more CTAs than SMs
        a, g, p = plan(2000, S)
        assert (g - 1) * p < S <= g * p and a * min(p, S) <= 148 and a >= 1
    assert lib.bnn_train_seed_plan(cfg, 0, 4, None, None, None) != 0


def test_no_cpu_fallback():
    from bnn_chaos_model_b200 import spock_reg_model as S
    from bnn_chaos_model_b200._lib import BnnChaosError

    st = swag_stats(0)
    m = S.SWAGModel(st["hparams"]).init_params(st["swa_params"])
    m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(st[k]) for k in ("w_avg", "w2_avg", "pre_D"))
    with pytest.raises(BnnChaosError):
        m.forward_swag_fast(torch.zeros(2, 100, 41))
    with pytest.raises(BnnChaosError):
        m.forward(torch.zeros(2, 100, 41), noisy_val=False)


def test_mirror_matches_reference_surface(tmp_path):
    from bnn_chaos_model_b200 import spock_reg_model as S

    st = swag_stats(3)
    m = S.SWAGModel(st["hparams"]).init_params(st["swa_params"])
    spec = R.ModelSpec.from_hparams(st["hparams"])
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec.layout()
    assert m.flatten().shape == (7583,)
    assert tuple(m.zero_columns()) == spec.zero_cols
    assert (m.K, m.c, m.lowest) == (30, 5, 0.5)
    for name in ("forward", "compute_summary_stats", "predict_instability", "_lossfnc", "lossfnc", "input_kl",
                 "summary_kl", "init_params", "flatten", "load", "aggregate_model", "sample_weights", "forward_swag",
                 "forward_swag_fast"):
        assert callable(getattr(m, name)), name
    # load() is the inverse of flatten()
    v = torch.randn(7583)
    m.load(v)
    assert torch.equal(m.flatten(), v)
    # save_swag / load_swag round trip with the reference's dict keys
    m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(st[k]) for k in ("w_avg", "w2_avg", "pre_D"))
    path = str(tmp_path / "x_v50_0_output.pkl")
    S.save_swag(m, path)
    raw = torch.load(path, weights_only=False)
    assert sorted(raw) == ["hparams", "pre_D", "swa_params", "w2_avg", "w_avg"]
    m2 = S.load_swag(path)
    assert torch.equal(m2.pre_D, m.pre_D) and m2.K == 30 and m2.ssX is not None
    assert abs(m2.ssX.mean_[0] - 4954.58585) < 1e-4
    assert float(m.input_kl()) == pytest.approx(float(R.input_kl({"input_noise_logvar": m.input_noise_logvar})))


def test_unsafe_pickle_rejected(tmp_path):
    import pickle

    from bnn_chaos_model_b200 import spock_reg_model as S

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))

    path = str(tmp_path / "evil.pkl")
    torch.save({"hparams": Evil()}, path)
    with pytest.raises(pickle.UnpicklingError):
        S.load_swag(path)


def test_unsafe_legacy_pickle_and_scaler_rejected(tmp_path):
    """torch's legacy (non-zip) loader calls pickle_module.load() on the leading records before it builds
    an Unpickler, and *_ssX.pkl was read with plain pickle.load: both must go through an allow-list, and the payload
    must never run."""
    import pickle

    from bnn_chaos_model_b200 import spock_reg_model as S

    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return (os.system, (f"touch {marker}",))

    legacy = tmp_path / "legacy_v50_0_output.pkl"
    with open(legacy, "wb") as f:  # a legacy-format torch file starts with a pickled magic number
        pickle.dump(Evil(), f, protocol=2)
        pickle.dump(1001, f, protocol=2)
    with pytest.raises(pickle.UnpicklingError):
        S.load_swag(str(legacy))
    assert not marker.exists()
    # a genuine checkpoint whose scaler side-file is malicious (non-'v50' path -> the ssX file is read)
    st = swag_stats(0)
    m = S.SWAGModel(st["hparams"]).init_params(st["swa_params"])
    m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(st[k]) for k in ("w_avg", "w2_avg", "pre_D"))
    good = tmp_path / "x_output.pkl"
    S.save_swag(m, str(good))
    with open(str(good)[:-4] + "_ssX.pkl", "wb") as f:
        pickle.dump(Evil(), f)
    with pytest.raises(pickle.UnpicklingError):
        S.load_swag(str(good))
    assert not marker.exists()
    # and a real StandardScaler pickle still loads
    from sklearn.preprocessing import StandardScaler

    ss = StandardScaler().fit(np.arange(12.0).reshape(4, 3))
    with open(str(good)[:-4] + "_ssX.pkl", "wb") as f:
        pickle.dump(ss, f)
    m2 = S.load_swag(str(good))
    assert np.allclose(m2.ssX.mean_, ss.mean_) and np.allclose(m2.ssX.scale_, ss.scale_)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = R.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert tuple(int(v) for v in got) == want
    z = R.draw_z1(1234, np.arange(64), 7583)
    assert z.shape == (64, 7583) and abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1) < 0.01
    e = R.draw_eps(5, np.arange(3), np.arange(1000), 40)
    assert e.shape == (3, 1000, 40) and np.isfinite(e).all()


def test_shard_range_partitions():
    from bnn_chaos_model_b200.multiswag import shard_range

    for n in (0, 1, 7, 8, 100000, 12345):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            for g in (5, 8):  # granule-aligned boundaries (tensor-core tile), still a partition
                spans = [shard_range(n, r, w, g) for r in range(w)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
                assert all(a % g == 0 for a, _ in spans if a < n)
                assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 2 * g


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from bnn_chaos_model_b200.multiswag import ChunkedSystemGather, gather_system_shards, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
# the same gather in asynchronous pieces (what predict_sharded(overlap_chunks > 1) runs under the next chunk's kernel)
n_loc, tail = 7, (3, 2)
ref = torch.arange(2 * n_loc * 6, dtype=torch.float32).reshape(2 * n_loc, 3, 2)
cg = ChunkedSystemGather(n_loc, tail, 2, torch.device("cpu"))
for a, b in ((0, 3), (3, 6), (6, 7)):
    cg.add(ref[rank * n_loc + a: rank * n_loc + b].clone(), a, b)
assert torch.equal(cg.finish(), ref), rank
for n_total, g in ((10, 1), (11, 1), (23, 5), (4, 5)):
    lo, hi = shard_range(n_total, rank, 2, g)
    full_ref = torch.arange(n_total * 3 * 2, dtype=torch.float32).reshape(n_total, 3, 2)
    full = gather_system_shards(full_ref[lo:hi].clone(), n_total, granule=g)
    assert torch.equal(full, full_ref), (rank, n_total, g)
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_system_shards_gloo_world2(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_custom_one_cycle_lr_mirror_vs_reference_golden():
    """The mirrored scheduler on a real torch SGD, stepped like Lightning steps it: lr and momentum of every step
    equal the reference's (tests/golden/schedule.npz), and it raises at the same step."""
    from conftest import load_golden
    from bnn_chaos_model_b200 import spock_reg_model as S

    z = load_golden("schedule.npz")
    for t in "ab":
        mx, tot = float(z[f"{t}_max_lr"]), int(z[f"{t}_total"])
        w = torch.nn.Parameter(torch.zeros(3))
        opt = torch.optim.SGD([w], lr=mx, momentum=0.9, weight_decay=1e-14)
        sch = S.CustomOneCycleLR(opt, mx, tot, final_div_factor=1e4)
        raised_at = -1
        for i in range(tot + 3):
            assert opt.param_groups[0]["lr"] == pytest.approx(float(z[f"{t}_lr_ref"][i]), rel=1e-14)
            assert opt.param_groups[0]["momentum"] == pytest.approx(float(z[f"{t}_momentum_ref"][i]), rel=1e-14)
            opt.step()
            try:
                sch.step()
            except ValueError:
                raised_at = i + 1
                break
        assert raised_at == int(z[f"{t}_raised_at"])
    lr, mom = S.one_cycle_lr_momentum(1000, 5e-4, 270000)
    i = list(z["c_steps"]).index(1080)
    assert S.one_cycle_lr_momentum(1080, 5e-4, 270000)[0] == pytest.approx(float(z["c_lr_ref"][i]), rel=1e-14)
    with pytest.raises(ValueError):
        S.CustomOneCycleLR(torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0), 1.0, 10, pct_start=2.0)
