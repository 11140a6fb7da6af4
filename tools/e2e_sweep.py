import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_stats
from bnn_chaos_model_b200 import spock_reg_model as S, synth
from bnn_chaos_model_b200.multiswag import MultiSWAG
dev = torch.device("cuda:0")
z, hp, sp = load_stats(0)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
xh = torch.from_numpy(synth.make_systems(10000, seed=1)).pin_memory()
out = torch.empty((10000, 1000, 2)).pin_memory()
for cfg in (1, 2, 3, 4, 8, (0.04, 0.96), (0.04, 0.48, 0.48), (0.03, 0.17, 0.4, 0.4), (0.02, 0.08, 0.3, 0.3, 0.3)):
    for i in range(2): ens.predict_host(xh, 1000, seed=i, out_host=out, n_chunks=cfg)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(3): ens.predict_host(xh, 1000, seed=5 + i, out_host=out, n_chunks=cfg)
    b.record(); torch.cuda.synchronize()
    print(cfg, round(a.elapsed_time(b) / 3, 2), "ms")
