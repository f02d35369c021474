"""Duration of each dependency chain of the hot-path step, each captured alone as a CUDA graph (warm, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wseg_b200 import ops

dev = torch.device("cuda:0")
hp = bench.HotPath(bench.N_PER_GPU, dev, seed=1234)
d = hp.d
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
state = {}


def k1b(name, size, gname):
    y, saved = ops.fullres_norm_fwd(d[name], d["label"], size)
    return y, saved, ops.fullres_norm_bwd(d[name], y, saved, d[gname])


def chain_a():
    state["rv1"], state["ws1"], state["camn1"] = ops.pcm_forward_logits(d["cam1"], d["f1"])


def chain_a2():
    state["rv2"], state["ws2"], state["camn2"] = ops.pcm_forward_logits(d["cam2"], d["f2"])


def chain_b():
    return k1b("rv1_full", bench.VIEW2, "g_rv1")


def chain_c():
    return ops.pcm_backward(d["f1"], state["camn1"], state["rv1"], d["g_rv_down1"], state["ws1"])


def chain_c2():
    return k1b("rv2_full", None, "g_rv2"), ops.pcm_backward(d["f2"], state["camn2"], state["rv2"], d["g_rv_down2"], state["ws2"])


def chain_e():
    f1 = ops.resize_fwd(d["fproj1"], hp.grid)
    c1 = ops.resize_fwd(state["rv1"], hp.grid)
    n1, pl1 = ops.seed_labels(c1, d["label"], 0.2)
    n2, pl2 = ops.seed_labels(state["rv2"], d["label"], 0.2)
    p1, _ = ops.prototypes(n1, f1)
    p2, _ = ops.prototypes(n2, d["fproj2"])
    return ops.contrast_loss_raw(f1, d["fproj2"], p1, p2, pl1, pl2, hp.key1, hp.key2)


def chain_e_parts():
    f1 = ops.resize_fwd(d["fproj1"], hp.grid)
    c1 = ops.resize_fwd(state["rv1"], hp.grid)
    n1, pl1 = ops.seed_labels(c1, d["label"], 0.2)
    n2, pl2 = ops.seed_labels(state["rv2"], d["label"], 0.2)
    state.update(f1=f1, n1=n1, pl1=pl1, n2=n2, pl2=pl2)


def chain_k3():
    state["p1"], _ = ops.prototypes(state["n1"], state["f1"])
    state["p2"], _ = ops.prototypes(state["n2"], d["fproj2"])


def chain_k4():
    return ops.contrast_loss_raw(state["f1"], d["fproj2"], state["p1"], state["p2"], state["pl1"], state["pl2"], hp.key1, hp.key2)


def chain_f():
    return k1b("cam1_full", bench.VIEW2, "g_n1"), k1b("cam2_full", None, "g_n2")


def measure(name, fn, iters=20):
    for _ in range(3):
        keep = fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    print(f"{tot / iters * 1e3:8.1f} us  {name}")
    return keep


measure("A  view 1: K1a + K2 forward (4 kernels)", chain_a)
measure("A2 view 2: K1a + K2 forward", chain_a2)
measure("B  K1b fwd+bwd on cam_rv1 (448^2 -> 128^2)", chain_b)
measure("C  K2 backward view 1 (3 kernels + memset)", chain_c)
measure("C2 view 2: K1b cam_rv2 + K2 backward", chain_c2)
measure("E  loss chain: resize x2, K1c x2, K3 x2, K4", chain_e)
measure("E1   resize x2 + K1c x2", chain_e_parts)
measure("E2   K3 x2", chain_k3)
measure("E3   K4", chain_k4)
measure("F  K1b fwd+bwd on cam1 and cam2", chain_f)
measure("whole step (forked streams)", hp.step)
