"""Dump the in-kernel timeline of CTA 0 of the PCM backward kernel.  usage: python tools/trace_pcm_bwd.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import _lib, ops, synth

dev = torch.device("cuda:0")
d = synth.pcm_inputs(8, 56, 56, seed=3)
cam, f, g = d["cam"].to(dev), d["f"].to(dev), d["grad_out"].to(dev)
camn = ops.cam_norm(cam)
for _ in range(3):
    out, ws = ops.pcm_forward(camn, f)
    ops.pcm_backward(f, camn, out, g, ws)
buf = torch.zeros(768, dtype=torch.int64, device=dev)
_lib.lib().wsegk_debug_pcm_trace_bwd(buf.data_ptr())
ops.pcm_profile(True)
ops.pcm_backward(f, camn, out, g, ws)
_lib.lib().wsegk_debug_pcm_trace_bwd(None)
torch.cuda.synchronize()
print(f"kernel by events {ops.pcm_main_ms()[1] * 1e3:.1f} us")
t = buf.cpu().reshape(3, 64, 4)
t0 = int(t[t > 0].min())
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
print(" t | producers: V slot free, X slot free | issuer T(t): start, V landed | G(t): start, E ready, E+F ready | tmem warp 0: fetch, loaded, arrived")
for i in range(44):
    print(f"{i:2d} | {rel(t[0, i, 0]):7d} {rel(t[0, i, 1]):7d} | {rel(t[1, i, 0]):7d} {rel(t[1, i, 1]):7d} | {rel(t[1, i, 2]):7d} {rel(t[0, i, 2]):7d} {rel(t[1, i, 3]):7d} | "
          f"{rel(t[2, i, 0]):7d} {rel(t[2, i, 1]):7d} {rel(t[2, i, 2]):7d}")
