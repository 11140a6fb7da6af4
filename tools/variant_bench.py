"""Time the predict-kernel variants on the BASELINE config-2 workload and cross-check them
bit for bit (they share the arithmetic order).  GPU only; used while tuning."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import FLOP_PER_EVAL_V50, load_stats  # noqa: E402
from bnn_chaos_model_b200 import spock_reg_model as S, synth  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG  # noqa: E402

variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["v2", "tc"]
n_sys = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n_samp = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda:0")
z, hp, sp = load_stats(0)
if len(sys.argv) > 4 and sys.argv[4] == "dense":   # all 41 input columns live (zero_mask == 0): the wide tensor-core variant
    hp = dict(hp, include_mmr=True, include_nan=True, include_eplusminus=True)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
x = torch.from_numpy(synth.make_systems(n_sys, seed=1)).to(dev)
_, thp = ens.sample_thetas(n_samp, seed=1)
ref = None
for v in variants:
    from bnn_chaos_model_b200 import _lib
    _lib.check(_lib.load().bnn_set_predict_variant({"auto": 0, "tc": 1, "v2": 2, "v1": 3}[v]))
    out = ens.predict(x, n_samp, seed=1, thp=thp)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ens.predict(x, n_samp, seed=1, thp=thp)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    same = None if ref is None else bool(torch.equal(out, ref))
    relerr = None if ref is None else float(((out - ref).abs() / ref.abs()).max())
    if ref is None:
        ref = out.clone()
    tf = FLOP_PER_EVAL_V50 * n_sys * n_samp / (ms * 1e-3) / 1e12
    print(json.dumps({"variant": v, "ms": round(ms, 3), "all_ms": [round(t, 2) for t in ts], "evals_per_s": n_sys * n_samp / (ms * 1e-3),
                      "tflops": round(tf, 2), "frac_fp32_peak": round(tf / 74.45, 4), "bitwise_equal_to_first": same, "max_rel_err_vs_first": relerr,
                      "finite": bool(torch.isfinite(out).all())}), flush=True)
