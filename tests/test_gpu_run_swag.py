"""GPU: the 30-seed distributed SWAG driver (bnn_chaos_model_b200/run_swag.py; BASELINE configs[3]): seeds sharded over
ranks, one all_gather of the statistics, reference-format output files (run_swag.py:95-97, spock_reg_model.py:911-930)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from bnn_chaos_model_b200 import run_swag, synth
from bnn_chaos_model_b200 import spock_reg_model as S

pytestmark = pytest.mark.gpu


def _args(extra=()):
    return run_swag.build_parser().parse_args(["--angles", "--no_mmr", "--no_nan", "--no_eplusminus", "--swa_steps", "6",
                                               "--batch_size", "50", "--epochs", "6", *extra])


def _data():
    X = torch.from_numpy(synth.make_systems(150, seed=61)); y = torch.from_numpy(synth.make_labels(150, seed=61))
    return X[:120], y[:120], X[120:], y[120:]


def test_run_swag_single_gpu_writes_reference_format(tmp_path):
    torch.cuda.set_device(0)
    args = _args(["--n_seeds", "3"])
    paths, tr, logs = run_swag.run(args, *_data(), out_dir=str(tmp_path))
    assert len(paths) == 3 and len(logs) == 6 and tr.global_step == 18   # 3 batches per epoch (ragged last one)
    assert os.path.basename(paths[2]) == ("steps=300000_megno=0_angles=1_power=0_hidden=40_latent=20_nommr=1_nonan=1_"
                                          "noeplusminus=1_v53_2_output.pkl")   # parse_swag_args.py:28-46
    assert int(tr.n_models[0]) == 5                       # collected after every epoch with global_step > swa_start = 3
    for i, p in enumerate(paths):
        raw = torch.load(p, weights_only=False)
        assert sorted(raw) == ["hparams", "pre_D", "swa_params", "w2_avg", "w_avg"]   # :911-920
        assert raw["swa_params"]["K"] == 30 and raw["swa_params"]["c"] == 5 and raw["swa_params"]["swa_start"] == 3
        assert raw["hparams"]["seed"] == i and raw["hparams"]["include_mmr"] is False
        assert os.path.exists(p[:-4] + "_ssX.pkl")
        m = S.load_swag(p)
        assert torch.equal(m.w_avg, tr.w_avg[i].cpu()) and torch.equal(m.w2_avg, tr.w2_avg[i].cpu())
        ncol = int(tr.n_cols[i])
        assert m.pre_D.shape == (7583, ncol) and torch.equal(m.pre_D, tr.pre_D[i, :, :ncol].cpu())
    # seeds start from different initialisations and see different batches / noise
    assert not torch.equal(tr.w_avg[0], tr.w_avg[1])
    # the unmodified reference's own load_swag reads the files (its module is staged under baseline/_ref by build())
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(ref_root, "spock_reg_model.py")):
        from oracle import ref_shim

        ref_shim.REFERENCE_ROOT = ref_root
        rm = ref_shim.load_reference_swag(paths[1])
        assert torch.equal(rm.w_avg, tr.w_avg[1].cpu()) and torch.equal(rm.pre_D, tr.pre_D[1, :, :int(tr.n_cols[1])].cpu())
        assert rm.K == 30 and rm.c == 5 and abs(rm.ssX.mean_[0] - 4954.58585) < 1e-4


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from bnn_chaos_model_b200 import run_swag, synth
from bnn_chaos_model_b200 import spock_reg_model as S
from bnn_chaos_model_b200.swag_train import seeds_of_rank
rank = int(sys.argv[1])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
args = run_swag.build_parser().parse_args(["--angles", "--no_mmr", "--no_nan", "--no_eplusminus", "--swa_steps", "6",
                                           "--batch_size", "50", "--epochs", "5", "--n_seeds", "5"])
X = torch.from_numpy(synth.make_systems(150, seed=61)); y = torch.from_numpy(synth.make_labels(150, seed=61))
paths, tr, logs = run_swag.run(args, X[:120], y[:120], X[120:], y[120:], out_dir={out!r})
mine = seeds_of_rank(5, rank, 2)
assert mine == ([0, 1, 2] if rank == 0 else [3, 4]) and tr.S == len(mine)
dist.barrier()
for i, s in enumerate(mine):          # every rank finds ITS seeds' statistics in the files rank 0 wrote
    name = run_swag.checkpoint_filename(args, s) + "_output.pkl"
    m = S.load_swag(os.path.join({out!r}, name))
    assert torch.equal(m.w_avg, tr.w_avg[i].cpu()) and torch.equal(m.w2_avg, tr.w2_avg[i].cpu()), (rank, s)
    assert torch.equal(m.pre_D, tr.pre_D[i, :, :int(tr.n_cols[i])].cpu()), (rank, s)
    assert m.flatten().shape == tr.theta[i].shape      # (the files hold the SWAG statistics, not the last iterate: :911-920)
assert (paths is not None) == (rank == 0)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_run_swag_two_ranks_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=port, out=str(tmp_path / "out")))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
