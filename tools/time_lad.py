import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geo4d_b200 import ops
dev = "cuda"
for G, n in [(1, 16 * 163840), (2, 16 * 163840), (1, 163840)]:
    x = torch.rand(G, n, device=dev) + 0.1
    y = 2.5 * x + 0.3 + 0.05 * torch.randn(G, n, device=dev)
    for iters in (1000, 5000):
        state = torch.zeros(G, 9, device=dev); state[:, 0] = 2.0
        acc = torch.zeros(G * 4, device=dev, dtype=torch.float64)
        torch.cuda.synchronize(); t0 = time.time()
        ops.lad_fit(x, y, n, G, state, acc, 1e-2, iters)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"G={G} n={n} iters={iters}: {dt*1e3:.1f} ms, steps run {state[:, 7].tolist()} done {state[:, 8].tolist()} s,t {state[0, :2].tolist()}", flush=True)
