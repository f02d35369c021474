"""Dump the in-kernel timeline of CTA (0,0,0) of the PCM forward kernel.  usage: python tools/trace_pcm.py [fast|precise]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import _lib, ops, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
ops.set_pcm_mode(mode)
dev = torch.device("cuda:0")
d = synth.pcm_inputs(8, 56, 56, seed=3)
cam, f = d["cam"].to(dev), d["f"].to(dev)
camn = ops.cam_norm(cam)
for _ in range(3):
    ops.pcm_forward(camn, f)
buf = torch.zeros(768, dtype=torch.int64, device=dev)
_lib.lib().wsegk_debug_pcm_trace(buf.data_ptr())
ops.pcm_forward(camn, f)
_lib.lib().wsegk_debug_pcm_trace(None)
torch.cuda.synchronize()
t = buf.cpu().reshape(3, 64, 4)
t0 = int(t[t > 0].min())
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
print("mode", mode, "(clock64 ticks relative to first stamp)")
print(" t | prod:empty_f ok | mma: full_f, peer_full_f (for S(t)) | loop: top, p_full ok | tmem warp: wait s_full, got, arrived")
for i in range(44):
    print(f"{i:2d} | {rel(t[0, i, 0]):7d} | {rel(t[1, i, 0]):7d} {rel(t[1, i, 1]):7d} | {rel(t[1, i, 2]):7d} {rel(t[1, i, 3]):7d} | "
          f"{rel(t[2, i, 0]):7d} {rel(t[2, i, 1]):7d} {rel(t[2, i, 2]):7d}")
