"""Forward / backward main-kernel durations at the headline shape (N=8, 56x56) on the current pair settings and on 72 pairs.
usage: python tools/pcm_quick.py   (or through tools/ab_run.py <lib>)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops

dev = torch.device("cuda:0")
ops.pcm_profile(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator().manual_seed(3)
n, h = 8, 56
f = torch.randn(n, 192, h, h, generator=gen).to(dev)
cam = (0.05 * torch.randn(n, 21, h, h, generator=gen)).to(dev)
g = (1e-3 * torch.randn(n, 21, h, h, generator=gen)).to(dev)
camn = ops.cam_norm(cam)
for pairs in (None, (72, 72), (52, 52)):
    if pairs:
        ops.set_pcm_pairs(*pairs)
    tf, tb = [], []
    for _ in range(7):
        flush.zero_()
        out, ws = ops.pcm_forward(camn, f)
        gf = ops.pcm_backward(f, camn, out, g, ws)
        torch.cuda.synchronize()
        a, b = ops.pcm_main_ms()
        tf.append(a); tb.append(b)
    tf.sort(); tb.sort()
    print(f"pairs {pairs or 'default'}: fwd {tf[3]*1e3:.1f} us  bwd {tb[3]*1e3:.1f} us")
