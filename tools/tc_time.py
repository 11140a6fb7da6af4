"""Cycles per tcgen05.mma (M=128, kind::tf32) as a function of N, K chain length and A source (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnn_chaos_model_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
modes = [(int(a), None) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [(0, None), (1, None)]
for from_smem, _ in modes:
    for N in ((16, 32, 48, 64, 80, 96) if from_smem > 1 else (16, 32, 48, 80, 96, 144, 192, 240, 256)):
        for K, reps in ((48, 1), (48, 20)):
            _lib.check(lib.bnn_tc_time(K, N, reps, from_smem, _lib.ptr(out), None))
            torch.cuda.synchronize()
            n = reps * K // 8
            tot, iss = out.tolist()
            print(f"A_from_smem={from_smem & 1} issuers={2 if from_smem & 256 else 1} N={N:3d} mmas={n:4d} total={tot:7d} cyc ({tot/n:7.1f}/mma) issue={iss:6d} ({iss/n:6.1f}/mma)", flush=True)
