"""One GPU's share of BASELINE configs[2]: MultiSWAG 30 seed models x 2000 weight samples x 100,000 systems over 8 GPUs
= 12,500 systems per GPU x 60,000 units = 7.5e8 evals, predictions [12500, 60000, 2] (6 GB, system-major: the block a
rank all-gathers) and the per-system posterior summary [12500, 8] (K7, radix-select path for 60,000 samples).
The three shipped v50 statistics (tests/golden) are cycled over the 30 model slots; Philox units are distinct."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_stats  # noqa: E402
from bnn_chaos_model_b200 import spock_reg_model as S, synth, posterior  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 30
S_ = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 12500
dev = torch.device("cuda:0")
models = []
for i in range(M):
    z, hp, sp = load_stats((0, 3, 17)[i % 3])
    m = S.SWAGModel(hp).init_params(sp).to(dev)
    m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
    models.append(m)
ens = MultiSWAG(models, device=dev)
x = torch.from_numpy(synth.make_systems(N, seed=2)).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {}
for rep in range(2):
    e = [ev() for _ in range(4)]
    e[0].record(); _, thp = ens.sample_thetas(S_, 7 + rep)
    e[1].record(); pred = ens.predict(x, S_, 7 + rep, thp=thp, system_major=True)
    e[2].record(); st = posterior.posterior_summary(pred, 1, 7 + rep)
    e[3].record(); torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
    out = {"models": M, "samples": S_, "systems": N, "evals": M * S_ * N, "pred_shape": list(pred.shape),
           "ms": {"swag_sample+pack": t[0], "predict": t[1], "posterior_summary": t[2], "total": sum(t)},
           "evals_per_s_predict": M * S_ * N / (t[1] * 1e-3), "evals_per_s_total": M * S_ * N / (sum(t) * 1e-3),
           "frac_fp32_peak_predict": M * S_ * N * 734560 / (t[1] * 1e-3) / 74.45e12,
           "finite": bool(torch.isfinite(st).all()), "mu_range": [float(pred[..., 0].min()), float(pred[..., 0].max())],
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    del pred, st, thp
print(json.dumps(out))
