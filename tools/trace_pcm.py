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
buf = torch.zeros(1024 + 2 * 148 + 2, dtype=torch.int64, device=dev)
ops.pcm_profile(True)
_lib.lib().wsegk_debug_pcm_trace(buf.data_ptr())
ops.pcm_forward(camn, f)
_lib.lib().wsegk_debug_pcm_trace(None)
torch.cuda.synchronize()
per = buf[1024:1024 + 296].cpu().reshape(148, 2)
per = per[per[:, 0] > (1 << 50)]          # globaltimer stamps (ns since the epoch); CTA 0's clock64 pair sits behind them
ncta = per.shape[0]
g0 = int(per[:, 0].min())
print(f"kernel by events {ops.pcm_main_ms()[0] * 1e3:.1f} us; {per.shape[0]} CTAs: start {int(per[:, 0].max()) - g0} ns after the first at the latest; "
      f"end min {int(per[:, 1].min()) - g0} max {int(per[:, 1].max()) - g0} ns; longest CTA {int((per[:, 1] - per[:, 0]).max())} ns, shortest {int((per[:, 1] - per[:, 0]).min())} ns")
t = buf[:1024].cpu().reshape(4, 64, 4).clone()
gt0, gt1 = int(buf[1024]), int(buf[1025])
ck0, ck1 = int(buf[1024 + 2 * ncta]), int(buf[1024 + 2 * ncta + 1])
print(f"CTA 0: {ck1 - ck0} clk in {gt1 - gt0} ns -> SM clock {(ck1 - ck0) / max(1, gt1 - gt0) * 1e3:.0f} MHz")
t0 = int(t[:3][t[:3] > 0].min())
tp0 = int(t[3, :, :2][t[3, :, :2] > 0].min()) - (int(t[2, :, 1][t[2, :, 1] > 0].min()) - t0) if (t[3, :, :2] > 0).any() else 0
relp = lambda x: int(x) - tp0 if int(x) > 0 else -1
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
print("mode", mode, "(clock64 ticks relative to first stamp)")
print(" t | prod:empty_f ok | S issuer: start, waits done, MMAs issued, commits done | PV issuer: start, waits done, MMAs issued | tmem warp 0: fetch, loaded, arrived | warp 7 arrived | peer CTA warp 0: loaded, arrived (aligned on the first load) | PV: p_full ok")
for i in range(44):
    print(f"{i:2d} | {rel(t[0, i, 0]):7d} | {rel(t[1, i, 0]):7d} {rel(t[1, i, 1]):7d} {rel(t[0, i, 1]):7d} {rel(t[0, i, 2]):7d} | {rel(t[1, i, 2]):7d} {rel(t[1, i, 3]):7d} {rel(t[0, i, 3]):7d} | "
          f"{rel(t[2, i, 0]):7d} {rel(t[2, i, 1]):7d} {rel(t[2, i, 2]):7d} | {rel(t[3, i, 2]):7d} | {relp(t[3, i, 0]):7d} {relp(t[3, i, 1]):7d} | {rel(t[3, i, 3]):7d}")
