"""Forward main-kernel duration only (ablation builds produce garbage values; only the time is read).  usage: python tools/fwd_time.py [N h]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
dev = torch.device("cuda:0")
n, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 56)
d = synth.pcm_inputs(n, h, h, seed=3)
cam, f = d["cam"].to(dev), d["f"].to(dev)
camn = ops.cam_norm(cam)
ops.pcm_profile(True)
ts = []
for i in range(10):
    ops.pcm_forward(camn, f)
    torch.cuda.synchronize()
    ts.append(ops.pcm_main_ms()[0])
ts = sorted(ts[3:])
print(f"N={n} {h}x{h} fwd main {ts[len(ts)//2]*1e3:.1f} us (min {ts[0]*1e3:.1f})  fixup {ops.pcm_fixup_ms()*1e3:.1f} us")
