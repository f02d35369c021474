"""PCM main-kernel durations (CUDA events around the tcgen05 kernels) over a few shapes.  usage: python tools/pcm_times.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth

dev = torch.device("cuda:0")
ops.pcm_profile(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, h in ((8, 56), (16, 56), (8, 64), (8, 32), (8, 96), (2, 56)):
    d = synth.pcm_inputs(n, h, h, seed=3)
    cam, f, g = d["cam"].to(dev), d["f"].to(dev), d["grad_out"].to(dev)
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
