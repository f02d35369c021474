"""Minimal PCM fwd+bwd loop for ncu captures.  usage: python tools/prof_pcm.py [fast|precise] [iters]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "precise"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ops.set_pcm_mode(mode)
dev = torch.device("cuda:0")
d = synth.pcm_inputs(8, 56, 56, seed=3)
cam, f, g = d["cam"].to(dev), d["f"].to(dev), d["grad_out"].to(dev)
camn = ops.cam_norm(cam)
for _ in range(iters):
    out, ws = ops.pcm_forward(camn, f)
    gf = ops.pcm_backward(f, camn, out, g, ws)
torch.cuda.synchronize()
print("ok", mode, float(out.abs().max()), float(gf.abs().max()))
