import sys, os, torch
sys.path.insert(0, '/root/repo')
from bnn_chaos_model_b200 import posterior
dev = torch.device('cuda:0')
N, U = 12500, 60000
g = torch.Generator(device=dev); g.manual_seed(0)
pred = torch.empty((N, U, 2), device=dev)
pred[..., 0] = 4 + 6 * torch.rand((N, U), device=dev, generator=g)
pred[..., 1] = 0.3 + torch.rand((N, U), device=dev, generator=g)
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e = [ev() for _ in range(3)]
    e[0].record(); t = posterior.sample_instability(pred, 5)
    e[1].record(); st = posterior.summarize_instability(t, pred, 1)
    e[2].record(); torch.cuda.synchronize()
    print('sample ms', e[0].elapsed_time(e[1]), 'summarize ms', e[1].elapsed_time(e[2]), flush=True)
