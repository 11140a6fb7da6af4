import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe
K, N, v = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
try:
    print("K", K, "N", N, "variant", v, "->", run_probe(K, N, v), flush=True)
except Exception as e:
    print("K", K, "N", N, "variant", v, "EXC", type(e).__name__, str(e)[:200], flush=True)
