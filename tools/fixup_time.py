"""Duration of the sign fix-up kernel at the headline shape.  usage: [WSEGK_FIX_CHUNK=n] python tools/fixup_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
dev = torch.device("cuda:0")
ops.pcm_profile(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d = synth.pcm_inputs(8, 56, 56, seed=3)
cam, f = d["cam"].to(dev), d["f"].to(dev)
camn = ops.cam_norm(cam)
ts = []
for _ in range(9):
    flush.zero_()
    ops.pcm_forward(camn, f)
    torch.cuda.synchronize()
    ts.append(ops.pcm_fixup_ms() * 1e3)
ts.sort()
print(f"chunk {os.environ.get('WSEGK_FIX_CHUNK', 'default')}: fix-up {ts[4]:.1f} us")
