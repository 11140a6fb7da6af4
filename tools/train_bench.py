"""Time the fused training step (K4): BASELINE config 4 shape -- B=2000 systems x 100 x 41, n_seeds models per GPU."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_stats  # noqa: E402
from bnn_chaos_model_b200 import _lib, synth, spock_reg_model as S  # noqa: E402
from bnn_chaos_model_b200._lib import TrainHParams  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
n_data = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda:0")
lib = _lib.load()
z, hp_, sp = load_stats(0)
m = S.SWAGModel(hp_).init_params(sp).to(dev)
cfg = m.config(100)
x = torch.from_numpy(synth.make_systems(n_data, seed=3)).to(dev)
y = torch.from_numpy(synth.make_labels(n_data, seed=3)).to(dev)
theta = torch.from_numpy(z["w_avg"]).to(dev)[None].repeat(n_seeds, 1).contiguous()
mom = torch.zeros_like(theta)
met = torch.zeros((n_seeds, 8), device=dev)
ws = torch.empty((lib.bnn_train_workspace_bytes(cfg, B, n_seeds) + 3) // 4, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
idx = torch.stack([torch.randperm(n_data, device=dev, generator=g)[:B] for _ in range(n_seeds)]).to(torch.int32).contiguous()
hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3, first_step=1, apply_update=1)
def step(i):
    hp.first_step = int(i == 0)
    _lib.check(lib.bnn_train_step(cfg, hp, n_seeds, _lib.ptr(theta), _lib.ptr(mom), _lib.ptr(x), _lib.ptr(y), _lib.ptr(idx), B,
                                  None, None, None, 1, i, None, _lib.ptr(met), _lib.ptr(ws), _lib.current_stream_ptr()))
for i in range(3): step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(iters): step(3 + i)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
flop = 3 * 814_560 * B * n_seeds
print(json.dumps({"n_seeds": n_seeds, "B": B, "ms_per_step": round(ms, 4), "seed_steps_per_s": n_seeds / (ms * 1e-3),
                  "tflops": flop / (ms * 1e-3) / 1e12, "frac_fp32_peak": flop / (ms * 1e-3) / 1e12 / 74.45,
                  "loss": met[:, 0].tolist()[:2], "nonfinite": met[:, 6].tolist()[:2]}))
tl = (ctypes.c_ulonglong * 24)()
_lib.check(lib.bnn_train_timeline(tl, 24))
if any(tl):
    if lib.bnn_set_train_variant and os.environ.get("BNN_TRAIN_VARIANT", "tc") != "v3":
        names = ["wait x image", "stage x -> A", "wait D1", "epilogue h1", "wait D2", "epilogue h2", "wait D3", "head: V0^T + record",
                 "g_f", "wait D(g_a2)", "epilogue g_a2", "wait D(g_a1)", "epilogue g_a1", "head: f store", "head: pooling", "head: V0",
                 "head: V1", "head: output + NLL", "head: V1^T", "-", "issuer 0: loop", "issuer 0: inside issue",
                 "producer warp 0: work", "producer warp 0: wait for a free ring stage"]
        tot = float(sum(tl[:19]))
    else:
        names = ["stage", "S0 load", "L1 fwd", "L2 fwd", "L3 fwd", "head tail (g_f)", "B1 g_a2", "B2 g_a1", "outer", "g_x+sums", "epilogue",
                 "pool", "head V0", "head V1", "head V2+nll", "bwd V1", "bwd V0+rec", "gm/gv", "producer: work", "producer: wait for free scratch", "p20", "p21", "p22", "p23"]
        tot = float(sum(tl[:18]))
    print(json.dumps({"timeline_cycles_cta0_last_step": {n: int(v) for n, v in zip(names, tl) if n != "-"},
                      "share_of_row_thread_0": {n: round(v / tot, 3) for n, v in zip(names, tl) if n != "-"}}))
