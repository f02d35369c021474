"""x8 enlargement of Net.forward's tail (network/resnet38_contrast.py:57-59): forward / backward durations, L2 flushed.
usage: python tools/time_resize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from wseg_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, c, h, up in ((8, 21, 56, 448), (8, 21, 16, 128), (16, 21, 64, 512)):
    x = torch.randn(n, c, h, h, device=dev)
    ref = F.interpolate(x, (up, up), mode="bilinear", align_corners=True)
    y = ops.resize_fwd(x, (up, up))
    g = torch.randn_like(y)
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y = ops.resize_fwd(x, (up, up)); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    nbytes = y.numel() * 4
    print(f"[{n},{c},{h},{h}] -> {up}x{up}: forward {ts[3]:.1f} us = {nbytes / ts[3] / 1e6:.2f} TB/s of stores; "
          f"max |y - F.interpolate| {float((y - ref).abs().max()):.2e}")
