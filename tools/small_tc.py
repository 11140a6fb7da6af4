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
x = torch.from_numpy(synth.make_systems(23, seed=1)).to(dev)
_, thp = ens.sample_thetas(6, seed=1)
os.environ["BNN_PREDICT_VARIANT"] = "v1"
ref = ens.predict(x, 6, seed=1, thp=thp); torch.cuda.synchronize()
os.environ["BNN_PREDICT_VARIANT"] = sys.argv[1] if len(sys.argv) > 1 else "tc4n4"
out = ens.predict(x, 6, seed=1, thp=thp); torch.cuda.synchronize()
print("max rel err", float(((out - ref).abs() / ref.abs()).max()))
