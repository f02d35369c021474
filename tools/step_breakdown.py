"""Warm, in-situ duration of every operator of the hot-path step (CUDA events around each op, one stream, eager)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from wseg_b200 import ops

dev = torch.device("cuda:0")
hp = bench.HotPath(bench.N_PER_GPU, dev, seed=1234)
d = hp.d
acc = collections.OrderedDict()
ITERS = 20


def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record()
    acc.setdefault(name, []).append((a, b))
    return out


def one():
    camn1 = timed("K1a v1", lambda: ops.cam_norm(d["cam1"]))
    rv1, ws1 = timed("K2 fwd v1 (4 kernels)", lambda: ops.pcm_forward(camn1, d["f1"]))
    camn2 = timed("K1a v2", lambda: ops.cam_norm(d["cam2"]))
    rv2, ws2 = timed("K2 fwd v2", lambda: ops.pcm_forward(camn2, d["f2"]))
    f1 = timed("glue resize x2", lambda: (ops.resize_fwd(d["fproj1"], hp.grid), ops.resize_fwd(rv1, hp.grid)))
    f1, c1 = f1
    norm1, pl1 = timed("K1c v1", lambda: ops.seed_labels(c1, d["label"], 0.2))
    norm2, pl2 = timed("K1c v2", lambda: ops.seed_labels(rv2, d["label"], 0.2))
    proto1, _ = timed("K3 v1", lambda: ops.prototypes(norm1, f1))
    proto2, _ = timed("K3 v2", lambda: ops.prototypes(norm2, d["fproj2"]))
    timed("K4", lambda: ops.contrast_loss_raw(f1, d["fproj2"], proto1, proto2, pl1, pl2, hp.key1, hp.key2))
    for name, size, gname in (("cam1_full", bench.VIEW2, "g_n1"), ("rv1_full", bench.VIEW2, "g_rv1"),
                              ("cam2_full", None, "g_n2"), ("rv2_full", None, "g_rv2")):
        y, saved = timed("K1b fwd " + name, lambda: ops.fullres_norm_fwd(d[name], d["label"], size))
        timed("K1b bwd " + name, lambda: ops.fullres_norm_bwd(d[name], y, saved, d[gname]))
    timed("K2 bwd v1 (3 kernels)", lambda: ops.pcm_backward(d["f1"], camn1, rv1, d["g_rv_down1"], ws1))
    timed("K2 bwd v2", lambda: ops.pcm_backward(d["f2"], camn2, rv2, d["g_rv_down2"], ws2))


for _ in range(5):
    one()
acc.clear()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(ITERS):
    one()
t1.record()
torch.cuda.synchronize()
tot = 0.0
for k, v in acc.items():
    ms = sum(a.elapsed_time(b) for a, b in v) / len(v)
    tot += ms
    print(f"{ms * 1e3:8.1f} us  {k}")
print(f"{tot * 1e3:8.1f} us  sum of ops;  {t0.elapsed_time(t1) / ITERS * 1e3:.1f} us per step wall (eager, single stream, includes launch gaps)")
