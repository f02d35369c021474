"""One eager pass of the lazy loss body for ncu launch lists.  usage: python tools/prof_lazy.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import loss as L, ops, synth
dev = torch.device("cuda:0")
n = 8
v1 = {k: v.to(dev) for k, v in synth.head_outputs(n, (448, 448), seed=1).items()}
v2 = {k: v.to(dev) for k, v in synth.head_outputs(n, (128, 128), seed=2).items()}
lab = synth.labels(n, 3).to(dev)
small = {"cam1": (0.05 * torch.randn(n, 21, 56, 56, device=dev)).requires_grad_(True),
         "cam2": (0.05 * torch.randn(n, 21, 16, 16, device=dev)).requires_grad_(True),
         "rv1": v1["cam_rv_down"].clone().requires_grad_(True), "rv2": v2["cam_rv_down"].clone().requires_grad_(True),
         "f1": v1["f_proj"].clone().requires_grad_(True), "f2": v2["f_proj"].clone().requires_grad_(True)}
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = L.contrast_step_loss_lazy(small["cam1"], small["rv1"], small["f1"], small["cam2"], small["rv2"], small["f2"], lab,
                                    (448, 448))
    grads = torch.autograd.grad(out["loss"], list(small.values()))
torch.cuda.synchronize()
print("ok", float(out["loss"]))
