"""Loss chain (K1c, K3, K4 fwd+bwd, both views) on the generalised grid, for ncu launch lists.
usage: python tools/prof_loss.py N grid D [iters]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from wseg_b200 import loss as L
from wseg_b200 import ops, synth

n, hw, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
mk = lambda: (F.relu(0.02 * torch.randn(n, d, hw, hw, generator=g, device=dev) + 0.01).requires_grad_(True),
              (0.5 * torch.randn(n, 21, hw, hw, generator=g, device=dev)).abs())
(f1, c1), (f2, c2) = mk(), mk()
label = synth.label_with_bg(synth.labels(n, seed=n)).to(dev)
P = n * hw * hw
k1, k2 = ops.hash_keys(P, 1, dev), ops.hash_keys(P, 2, dev)
for _ in range(iters):
    t = L.contrast_terms(f1, c1, f2, c2, label, 0.2, grid=(hw, hw), keys=(k1, k2))
    g1, g2 = torch.autograd.grad(t["loss_nce"], (f1, f2))
torch.cuda.synchronize()
print("ok", n, hw, d, float(t["loss_nce"]), float(g1.abs().sum()), float(g2.abs().sum()))
