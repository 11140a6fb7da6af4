"""Time K1 (bnn_swag_sample: fused sample + pack) against the unfused two-launch form.  python tools/k1_time.py [units]"""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import make_swag_model  # noqa: E402
from bnn_chaos_model_b200 import _lib  # noqa: E402
from bnn_chaos_model_b200.multiswag import MultiSWAG  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda:0")
lib = _lib.load()
ens = MultiSWAG([make_swag_model(0, dev)], device=dev)
cfg = ens.config()
M, d = ens.w_avg.shape
P = lib.bnn_packed_param_count(cfg)
theta = torch.empty((U, d), device=dev)
thp = torch.empty((U, P), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"units": U, "K": ens.K, "d": d, "P": P}
for name, fn, th in (("fused_packed_only", lib.bnn_swag_sample, None), ("fused_both", lib.bnn_swag_sample, theta),
                     ("unfused", lib.bnn_swag_sample_unfused, theta)):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(cfg, _lib.ptr(ens.w_avg), _lib.ptr(ens.w2_avg), _lib.ptr(ens.pre_D), M, ens.K, None, U, 0, U, 0.5, it,
                      None, None, _lib.ptr(th), _lib.ptr(thp), _lib.current_stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name + "_ms"] = sorted(ts[2:])[len(ts[2:]) // 2]
print(json.dumps(res))
