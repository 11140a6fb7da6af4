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
cfgs = ((0.04, 0.48, 0.48), (0.04, 0.94, 0.02), (0.04, 0.47, 0.47, 0.02), (0.04, 0.46, 0.46, 0.04), (0.04, 0.47, 0.47, 0.01, 0.01),
        (0.04, 0.31, 0.31, 0.31, 0.02, 0.01))
# interleaved rounds (box clocks drift under the power cap): median of 6 rounds x 2 steps per configuration
ts = {c: [] for c in cfgs}
for c in cfgs:
    ens.predict_host(xh, 1000, seed=0, out_host=out, n_chunks=c)
for rnd in range(6):
    for c in cfgs:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(2): ens.predict_host(xh, 1000, seed=5 + i, out_host=out, n_chunks=c)
        b.record(); torch.cuda.synchronize()
        ts[c].append(a.elapsed_time(b) / 2)
for c in cfgs:
    v = sorted(ts[c])
    print(c, "median", round(v[len(v) // 2], 3), "min", round(v[0], 3), "ms")
