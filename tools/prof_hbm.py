"""The HBM-bound kernels at the bench sizes, for ncu --set full captures: K1a (cam_norm) and K1b fwd+bwd at N=8, 448^2.
usage: python tools/prof_hbm.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
dev = torch.device("cuda:0")
n = 8
v1 = synth.head_outputs(n, (448, 448), seed=1)
label = synth.label_with_bg(synth.labels(n, 3)).to(dev)
cam_small = (0.05 * torch.randn(n, 21, 56, 56)).to(dev)
rv = v1["cam_rv"].to(dev)
go = torch.randn(n, 21, 128, 128).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    flush.zero_()
    camn = ops.cam_norm(cam_small)
    y, saved = ops.fullres_norm_fwd(rv, label, (128, 128))
    g = ops.fullres_norm_bwd(rv, y, saved, go)
torch.cuda.synchronize()
print("ok", float(camn.sum()), float(y.sum()), float(g.abs().sum()))
