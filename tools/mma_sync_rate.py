"""Issue rate of mma.sync.m16n8k8 tf32 (register operands) on every SM at once: cycles per MMA per scheduler (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_chaos_model_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
sink = torch.zeros(148 * 32 * 32, device="cuda:0")
for warps in (4, 8, 12, 16):
    for nacc in (1, 4, 8, 15):
        _lib.check(lib.bnn_mma_sync_rate(warps, nacc, 2000, _lib.ptr(out), _lib.ptr(sink), None))
        torch.cuda.synchronize()
        cyc, n = out.tolist()
        per_sched = cyc / (n * warps / 4)   # MMAs issued by one scheduler = n * warps / 4
        print(f"warps/CTA={warps:2d} independent accumulators={nacc:2d}: {cyc / n:6.2f} cycles per MMA per warp, "
              f"{per_sched:5.2f} per scheduler -> {1024 / per_sched * 4:7.1f} MAC/cycle/SM "
              f"({1024 / per_sched * 4 / 128:4.1f} x the FP32 FMA rate)", flush=True)
