"""Debug / bring-up harness of the tensor-core training kernel: one step on both kernels fed the same explicit noise,
per-parameter-group differences, then timing.  Run under `timeout` on the GPU box (a deadlocked kernel would hang)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import make_swag_model, swag_stats
from bnn_chaos_model_b200 import _lib, synth
from bnn_chaos_model_b200._lib import TrainHParams
from oracle import restatement as R

dev = torch.device("cuda:0")
lib = _lib.load()
S, B, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = make_swag_model(0, dev)
cfg = m.config(100)
spec = R.ModelSpec.from_hparams(swag_stats(0)["hparams"])
x = torch.from_numpy(synth.make_systems(N, seed=3)).to(dev)
y = torch.from_numpy(synth.make_labels(N, seed=3)).to(dev)
theta0 = m.w_avg[None].repeat(S, 1).contiguous()
gen = torch.Generator(device=dev); gen.manual_seed(0)
idx = torch.stack([torch.randperm(N, device=dev, generator=gen)[:B] for _ in range(S)]).to(torch.int32).contiguous()
hp = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3, first_step=1, apply_update=0)
ws = torch.empty((lib.bnn_train_workspace_bytes(cfg, B, S) + 3) // 4, device=dev)

def step(variant, eps):
    _lib.check(lib.bnn_set_train_variant(variant))
    g = torch.zeros_like(theta0); met = torch.zeros((S, 8), device=dev)
    e = eps if eps is not None else (None, None, None)
    _lib.check(lib.bnn_train_step(cfg, hp, S, _lib.ptr(theta0.clone()), None, _lib.ptr(x), _lib.ptr(y), _lib.ptr(idx), B,
                                  _lib.ptr(e[0]), _lib.ptr(e[1]), _lib.ptr(e[2]), 5, 11, _lib.ptr(g), _lib.ptr(met), _lib.ptr(ws), None))
    torch.cuda.synchronize()
    return g, met

_lib.check(lib.bnn_set_train_variant(1))
e_in = torch.empty((S, B, 100, 41), device=dev); e12 = torch.empty((S, B, 40), device=dev); e_sum = torch.empty((S, B, 40), device=dev)
_lib.check(lib.bnn_train_noise(cfg, S, B, 5, 11, _lib.ptr(e_in), _lib.ptr(e12), _lib.ptr(e_sum), None))
torch.cuda.synchronize()
print("noise ok", float(e_in.std()), flush=True)
g3, m3 = step(2, (e_in, e12, e_sum))
print("v3 ok", m3[0].tolist(), flush=True)
gt, mt = step(1, (e_in, e12, e_sum))
print("tc explicit ok", mt[0].tolist(), flush=True)
for s in range(S):
    print(f"seed {s}: metrics rel diff {((mt[s,:5]-m3[s,:5]).abs()/m3[s,:5].abs().clamp_min(1e-20)).tolist()}")
    for name, (off, shp) in spec.offsets().items():
        n = int(np.prod(shp))
        a, b = gt[s, off:off + n], g3[s, off:off + n]
        print(f"   {name:28s} max|ref| {float(b.abs().max()):.3e}  err/max {float((a-b).abs().max()/b.abs().max().clamp_min(1e-30)):.2e}")
    print(f"   total err / max-norm: {float((gt[s]-g3[s]).abs().max()/g3[s].abs().max()):.2e}")
gp, mp = step(1, None)
print("tc philox == explicit:", bool(torch.equal(gp, gt)), bool(torch.equal(mp, mt)), flush=True)
gp2, _ = step(1, None)
print("tc repeat bit-identical:", bool(torch.equal(gp, gp2)))
if iters:
    for variant in (1, 2):
        _lib.check(lib.bnn_set_train_variant(variant))
        th = theta0.clone(); mom = torch.zeros_like(th); met = torch.zeros((S, 8), device=dev)
        hp2 = TrainHParams(lr=1e-4, momentum=0.9, weight_decay=1e-14, clip_norm=758.3, beta_in=1e-5, beta_out=1e-3, first_step=1, apply_update=1)
        def run(i):
            hp2.first_step = int(i == 0)
            _lib.check(lib.bnn_train_step(cfg, hp2, S, _lib.ptr(th), _lib.ptr(mom), _lib.ptr(x), _lib.ptr(y), _lib.ptr(idx), B,
                                          None, None, None, 1, i, None, _lib.ptr(met), _lib.ptr(ws), None))
        for i in range(3): run(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters): run(3 + i)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        tf = 3 * 814560 * B * S / (ms * 1e-3) / 1e12
        print(json.dumps({"variant": "tc" if variant == 1 else "v3", "S": S, "B": B, "ms_per_step": ms, "seed_steps_per_s": S / (ms * 1e-3),
                          "tflops_algorithmic": tf, "frac_fp32_peak": tf / 74.45, "finite": bool(torch.isfinite(met).all())}), flush=True)
