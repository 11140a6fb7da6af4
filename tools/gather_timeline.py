"""Where do the peer pushes of PeerPushGather run relative to the persistent predictive kernel?  (diagnostic, N >= 2)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gather_timeline.py

Per step, on every rank: K2's duration, when the step's pushes END relative to the end of the K2 that produced them, and
how long the NEXT K2 takes.  Pushes that run on the copy engines end ~0.1-0.8 ms after their K2; pushes executed by SM copy
kernels cannot start beside the next K2's one-CTA-per-SM grid and end ~a whole kernel later (or delay that kernel).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_stats  # noqa: E402
from bnn_chaos_model_b200 import spock_reg_model as S, synth  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG, PeerPushGather  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
z, hp, sp = load_stats(0)
m = S.SWAGModel(hp).init_params(sp).to(dev)
m.w_avg, m.w2_avg, m.pre_D = (torch.from_numpy(z[k]).to(dev) for k in ("w_avg", "w2_avg", "pre_D"))
ens = MultiSWAG([m], device=dev)
n_sys, n_samp, steps = 10000, 1000, 8
x = torch.from_numpy(synth.make_systems(n_sys, seed=1000 + rank)).to(dev)
gather = PeerPushGather.get(n_sys, (n_samp, 2), world, rank, dev, None)
main = torch.cuda.current_stream()
ev = lambda: torch.cuda.Event(enable_timing=True)
rows, pend = [], None
for i in range(steps):
    _, thp = ens.sample_thetas(n_samp, seed=i, want_flat=False)
    a, b, c = ev(), ev(), ev()
    a.record(main)
    part = ens.predict(x, n_samp, seed=i, system_offset=rank * n_sys, system_major=True, thp=thp)
    b.record(main)
    turn = gather.begin()
    gather.add(part, 0, n_sys, turn)
    c.record(gather.stream)            # behind this step's pushes
    gather.seal(turn)
    if pend is not None:
        gather.finish(pend)
    pend = turn
    rows.append((a, b, c))
gather.finish(pend)
dist.barrier()
torch.cuda.synchronize()
out = []
for i, (a, b, c) in enumerate(rows):
    nxt = rows[i + 1] if i + 1 < len(rows) else None
    out.append({"step": i, "k2_ms": round(a.elapsed_time(b), 3), "pushes_end_after_k2_ms": round(b.elapsed_time(c), 3),
                "next_k2_starts_after_ms": round(b.elapsed_time(nxt[0]), 3) if nxt else None})
print(json.dumps({"rank": rank, "world": world, "mb_pushed_per_step": world * part.numel() * 4 / 1e6, "steps": out}), flush=True)
dist.destroy_process_group()
