"""5-planet style pipeline on one GPU (BASELINE configs[4] shape, scaled): raw series -> K6 pack -> K1 sample -> K2 predict
-> K7 sample + summarise.  Reports per-stage device time and the HBM-roofline fraction of the two memory-bound stages."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_stats  # noqa: E402
from bnn_chaos_model_b200 import spock_reg_model as S, synth, posterior  # noqa: E402
from bnn_chaos_model_b200.inputs import pack_trios  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000      # 5-planet systems
S_ = int(sys.argv[2]) if len(sys.argv) > 2 else 100       # weight samples (multiswag_5_planet.py:55)
Rt = 3
HBM = 6556.5e9
dev = torch.device("cuda:0")
z, hp, sp = load_stats(0)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
base = synth.raw_systems(3000, seed=8)
reps = (N * Rt + 2999) // 3000
raw = torch.from_numpy(base).to(dev).repeat(reps, 1, 1)[: N * Rt]
ts = raw[:, :, :26].contiguous().reshape(N, Rt, 100, 26)
ms = raw[:, 0, 26:29].contiguous().reshape(N, Rt, 3)
ev = lambda: torch.cuda.Event(enable_timing=True)

def run(seed):
    e = [ev() for _ in range(5)]
    e[0].record(); x = pack_trios(ts, ms)
    e[1].record(); _, thp = ens.sample_thetas(S_, seed)
    e[2].record(); pred = ens.predict(x, S_, seed, thp=thp, system_major=True)
    e[3].record(); st = posterior.posterior_summary(pred, Rt, seed)
    e[4].record(); torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(4)], st

for i in range(3):
    run(i)
best = None
for i in range(5):
    t, st = run(10 + i)
    best = t if best is None else [min(a, b) for a, b in zip(best, t)]
rows = N * Rt * 100
pack_bytes = rows * (26 * 8 + 41 * 4) + N * Rt * 24
post_bytes = N * Rt * S_ * (8 + 4 + 4 + 8 + 8) + N * 32   # pred read (sample), t write, t read, pred read x2 passes (summarise)
evals = N * Rt * S_
print(json.dumps({"systems": N, "trios": Rt, "samples": S_, "evals": evals,
                  "ms": {"pack": best[0], "swag_sample": best[1], "predict": best[2], "posterior": best[3], "total": sum(best)},
                  "evals_per_s_total": evals / (sum(best) * 1e-3), "systems_per_s_total": N / (sum(best) * 1e-3),
                  "pack_GBps": pack_bytes / (best[0] * 1e-3) / 1e9, "pack_frac_hbm": pack_bytes / (best[0] * 1e-3) / HBM,
                  "posterior_GBps": post_bytes / (best[3] * 1e-3) / 1e9, "posterior_frac_hbm": post_bytes / (best[3] * 1e-3) / HBM,
                  "summary_finite": bool(torch.isfinite(st).all())}))
