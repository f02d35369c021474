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
f1 = v1["f_proj"].to(dev)
f2 = synth.head_outputs(n, (128, 128), seed=2)["f_proj"].to(dev)
rv_small = v1["cam_rv_down"].to(dev)
rv2_small = synth.head_outputs(n, (128, 128), seed=2)["cam_rv_down"].to(dev)
k1, k2 = ops.hash_keys(n * 256, 1, dev), ops.hash_keys(n * 256, 2, dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    flush.zero_()
    camn = ops.cam_norm(cam_small)
    up = ops.resize_fwd(cam_small, (448, 448))                  # the x8 enlargement of Net.forward's tail
    y, saved = ops.fullres_norm_fwd(rv, label, (128, 128))
    g = ops.fullres_norm_bwd(rv, y, saved, go)
    # the loss chain at the literal 16x16 grid: glue, K1c, K3, K4
    fs, cs = ops.resize_fwd(f1, (16, 16)), ops.resize_fwd(rv_small, (16, 16))
    n1, p1 = ops.seed_labels(cs, label, 0.2)
    n2, p2 = ops.seed_labels(rv2_small, label, 0.2)
    pr1, _ = ops.prototypes(n1, fs)
    pr2, _ = ops.prototypes(n2, f2)
    ops.contrast_loss_raw(fs, f2, pr1, pr2, p1, p2, k1, k2)
torch.cuda.synchronize()
print("ok", float(camn.sum()), float(y.sum()), float(g.abs().sum()))
