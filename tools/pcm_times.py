"""PCM main-kernel durations (CUDA events around the tcgen05 kernels) over a few shapes.  usage: python tools/pcm_times.py [large]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth

dev = torch.device("cuda:0")
ops.pcm_profile(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = ((8, 56), (16, 56), (8, 64), (8, 32), (8, 96), (2, 56))
if len(sys.argv) > 1 and sys.argv[1] == "large":
    shapes = ((128, 56), (64, 80), (32, 96), (128, 32), (64, 64))
for n, h in shapes:
    gen = torch.Generator().manual_seed(3)
    f = torch.randn(n, 192, h, h, generator=gen).to(dev)
    cam = (0.05 * torch.randn(n, 21, h, h, generator=gen)).to(dev)
    g = (1e-3 * torch.randn(n, 21, h, h, generator=gen)).to(dev)
    camn = ops.cam_norm(cam)
    tf, tb = [], []
    for _ in range(5):
        flush.zero_()
        out, ws = ops.pcm_forward(camn, f)
        gf = ops.pcm_backward(f, camn, out, g, ws)
        torch.cuda.synchronize()
        a, b = ops.pcm_main_ms()
        tf.append(a); tb.append(b)
    HW = h * h
    ff, fb = 2.0 * n * HW * HW * 213, 2.0 * n * HW * HW * 405
    tf.sort(); tb.sort()
    print(f"N={n} {h}x{h}: fwd {tf[2]*1e3:.1f} us ({ff/tf[2]/1e9:.0f} TF/s)  bwd {tb[2]*1e3:.1f} us ({fb/tb[2]/1e9:.0f} TF/s)")
