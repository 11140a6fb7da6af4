"""Dump a clock64 timeline of CTA 0 of the tensor-core kernel (debugging aid, GPU).
Needs a library built with the stamps compiled in:  make -C bnn_chaos_model_b200/csrc clean && make -C bnn_chaos_model_b200/csrc TIMELINE=1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
dbg = torch.zeros(8 * 512, dtype=torch.int64, device=dev)
os.environ["BNN_TC_TIMELINE_PTR"] = str(dbg.data_ptr())
os.environ["BNN_PREDICT_VARIANT"] = sys.argv[1] if len(sys.argv) > 1 else "tc4n4"
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
names = {1: "x staged", 2: "pooled", 3: "d1 wake", 4: "e1 done", 5: "d2 wake", 6: "e2 done", 7: "d3 wake", 8: "e3 done", 31: "  ld done", 32: "  st issued", 33: "  st done", 34: "  x st issued", 21: "  rec slot free", 22: "  pool sums done"}
tnames = {1: "wait unit", 2: "unit ready", 3: "tail done"}
xs = [t - t0 for t, role, code in ev if role == 0 and code == 1]
print("slot0 x-staged times:", xs)
print("slot0 job periods:", [b - a for a, b in zip(xs, xs[1:])])
for r in (4, 5):
    print(f"tail{r-4}:", [(code, t - t0) for t, role, code in ev if role == r][:40])
for t, role, code in ev[:int(os.environ.get("TL_N", "0"))]:
    if role == 7:
        kind = {0: "issue-start", 1: "issue-end  "}[code // 100]
        print(f"{t - t0:8d}  MMA0     {kind} L{code % 100 + 1}")
    else:
        who = f"EPI{role}   " if role < 3 else f"TAIL{role - 4}  "
        print(f"{t - t0:8d}  {who}  {(names if role < 3 else tnames).get(code, code)}")
