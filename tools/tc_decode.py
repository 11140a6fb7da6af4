"""Decode the tcgen05 operand layouts empirically with indicator inputs (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_chaos_model_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")

def run(A, B, variant=0):
    K, N = A.shape[1], B.shape[0]
    D = torch.full((128, N), float("nan"), device=dev)
    Ad, Bd = A.to(dev).contiguous(), B.to(dev).contiguous()  # keep alive: ptr() of a temporary dangles
    _lib.check(lib.bnn_tc_probe(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), K, N, variant, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    return D.cpu()

torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
K, N = 16, 16
A = (torch.arange(128)[:, None] * 16 + torch.arange(K)[None, :]).float() + 0.5
B = (torch.arange(N)[:, None] * 100 + torch.arange(K)[None, :]).float()
for v, name in ((6, "A region, no MMA"), (2, "A region after MMA"), (0, "D")):
    D = run(A, B, v)
    if v:
        bad = (D != A).any(1).nonzero().flatten().tolist()
        print(name, ": rows where readback != A:", bad[:40], "n_bad", len(bad))
        if bad:
            print(" row", bad[0], D[bad[0]].tolist())
    else:
        ref = (A.double() @ B.double().T).float()
        bad = ((D - ref).abs() > 1e-3 * ref.abs().max()).any(1).nonzero().flatten().tolist()
        print(name, ": bad rows:", bad[:64], "n_bad", len(bad))
