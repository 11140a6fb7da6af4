"""Clock64 timeline of CTA 0 of the tensor-core predictive kernel (debugging aid, GPU).
Needs the diagnostic library:  make -C bnn_chaos_model_b200/csrc predict_timeline ;
BNN_CHAOS_LIB=$PWD/bnn_chaos_model_b200/libbnnchaos_tctl.so python tools/tc_timeline.py
Roles: 0/1 = warp 0 of epilogue team 0/1, 2..5 = issuer of TMEM slot 0..3.  Codes: 100*slot + 10*phase + {1: phase start (before the wait
for D), 2: D ready, 3: A published, 4: pooled}; issuer: 100*slot + 10*layer + {1: A ready seen, 2: layer issued + committed}."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
dbg = torch.zeros(8 * 512, dtype=torch.int64, device=dev)
os.environ["BNN_TC_TIMELINE_PTR"] = str(dbg.data_ptr())
from bench import load_stats
from bnn_chaos_model_b200 import spock_reg_model as S, synth
from bnn_chaos_model_b200.multiswag import MultiSWAG
z, hp, sp = load_stats(0)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
x = torch.from_numpy(synth.make_systems(740, seed=1)).to(dev)
_, thp = ens.sample_thetas(24, seed=1)
out = ens.predict(x, 24, seed=1, thp=thp); torch.cuda.synchronize()
dbg.zero_()
out = ens.predict(x, 24, seed=1, thp=thp); torch.cuda.synchronize()
d = dbg.cpu().view(8, 512)
ev = []
for role in range(8):
    n = int(d[role, 0])
    for k in range(n):
        v = int(d[role, 1 + k]); ev.append((v & 0xFFFFFFFFFFFF, role, v >> 48))
ev.sort()
t0 = ev[0][0]
for role in range(6):
    seq = [(t - t0, code) for t, r, code in ev if r == role]
    print(f"role {role}: {len(seq)} stamps")
    prev = None
    for t, code in seq[:int(os.environ.get("TL_N", "120"))]:
        print(f"   {t:8d} (+{0 if prev is None else t - prev:5d})  slot {code // 100} phase/layer {(code % 100) // 10} ev {code % 10}")
        prev = t
