"""Error distribution of the predict kernels against an fp64 evaluation of the same network (GPU).

The reference is fp32 PyTorch; two fp32 evaluations in different summation orders already differ at the 1e-6 .. 1e-5
level on ill-conditioned systems (var_sample = eps2*std_in_var + var near 0 feeds a sqrt).  This tool quantifies that:
for N systems x S weight samples with explicit eps it prints, per implementation, quantiles and the maximum of the
relative error of (mu, std) against fp64, plus the same for a plain torch fp32 evaluation (the reference's arithmetic).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_stats  # noqa: E402
from bnn_chaos_model_b200 import spock_reg_model as S, synth  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG  # noqa: E402
from oracle import restatement as R  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
U = int(sys.argv[2]) if len(sys.argv) > 2 else 500
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["v2", "tc"]
dev = torch.device("cuda:0")
z, hp, sp = load_stats(0)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
spec = R.ModelSpec.from_hparams(hp)
x = torch.from_numpy(synth.make_systems(N, seed=5)).to(dev)
theta, thp = ens.sample_thetas(U, seed=5)
eps = torch.randn((U, N, 40), device=dev)
cfg = m.config()


def torch_eval(dtype):
    """The oracle's arithmetic (oracle/restatement.py) on the GPU in `dtype`, unit by unit."""
    outs = []
    xm = R.zero_columns(spec, x.to(dtype))
    for u in range(U):
        p = {k: v.to(dtype) for k, v in R.unflatten(spec, theta[u]).items()}
        s = R.compute_summary_stats(spec, p, xm, eps[u, :, :20].to(dtype), eps[u, :, 20:].to(dtype))
        mu, sd = R.predict_instability(spec, p, s)
        outs.append(torch.cat((mu, sd), 1))
    return torch.stack(outs)


truth = torch_eval(torch.float64)
rows = {}
torch.backends.cuda.matmul.allow_tf32 = False
rows["torch_fp32"] = torch_eval(torch.float32).double()
for v in variants:
    from bnn_chaos_model_b200 import _lib
    _lib.check(_lib.load().bnn_set_predict_variant({"auto": 0, "tc": 1, "v2": 2, "v1": 3}[v]))
    out, _ = m._predict(x, thp, eps, cfg=cfg)
    rows[v] = out.double()
qs = torch.tensor([0.5, 0.99, 0.9999, 0.999999], device=dev, dtype=torch.float64)
for k, o in rows.items():
    err = ((o - truth).abs() / truth.abs()).reshape(-1)
    idx = torch.randperm(err.numel(), device=dev)[: 4_000_000]
    qv = torch.quantile(err[idx], qs[:3]).tolist()
    print(json.dumps({"impl": k, "evals": N * U, "median": qv[0], "p99": qv[1], "p99.99": qv[2], "max": float(err.max()),
                      "frac_above_1e-5": float((err > 1e-5).double().mean())}), flush=True)
a, b = rows[variants[0]], rows[variants[-1]]
d = ((a - b).abs() / b.abs()).reshape(-1)
print(json.dumps({"pair": f"{variants[0]} vs {variants[-1]}", "max": float(d.max()), "frac_above_1e-5": float((d > 1e-5).double().mean())}))
d = ((rows["torch_fp32"] - rows[variants[-1]]).abs() / rows["torch_fp32"].abs()).reshape(-1)
print(json.dumps({"pair": f"torch_fp32 vs {variants[-1]}", "max": float(d.max()), "frac_above_1e-5": float((d > 1e-5).double().mean())}))
