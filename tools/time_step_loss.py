"""Time loss.contrast_step_loss (hot path in CUDA kernels + auxiliary SEAM losses in PyTorch) fwd+bwd at N=8, 448^2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import loss as L, ops, synth
dev = torch.device("cuda:0")
n = 8
v1 = {k: v.to(dev).requires_grad_(True) for k, v in synth.head_outputs(n, (448, 448), seed=1).items()}
v2 = {k: v.to(dev).requires_grad_(True) for k, v in synth.head_outputs(n, (128, 128), seed=2).items()}
lab = synth.labels(n, 3).to(dev)
def run(aux=True):
    if aux:
        out = L.contrast_step_loss(v1["cam"], v1["cam_rv"], v1["f_proj"], v1["cam_rv_down"], v2["cam"], v2["cam_rv"],
                                   v2["f_proj"], v2["cam_rv_down"], lab, 0.2, (128, 128))
        out["loss"].backward()
    else:
        label = synth.label_with_bg(lab)
        t = L.contrast_terms(v1["f_proj"], v1["cam_rv_down"], v2["f_proj"], v2["cam_rv_down"], label, 0.2, grid=(16, 16))
        tot = t["loss_nce"]
        for x, size in ((v1["cam"], (128, 128)), (v1["cam_rv"], (128, 128)), (v2["cam"], None), (v2["cam_rv"], None)):
            tot = tot + ops.fullres_norm(x, label, size).sum()
        tot.backward()
for aux in (True, False):
    for _ in range(3): run(aux)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): run(aux)
    torch.cuda.synchronize()
    print(f"aux losses {'in' if aux else 'ex'}cluded: {(time.time() - t0) / 10 * 1e3:.2f} ms per fwd+bwd (eager)")

# the same, captured as a CUDA graph (GPU time without the eager launch overhead)
from wseg_b200 import ops as _ops
label = synth.label_with_bg(lab)
keys = (_ops.hash_keys(n * 256, 1, dev), _ops.hash_keys(n * 256, 2, dev))
leaves = [v1["cam"], v1["cam_rv"], v1["f_proj"], v2["cam"], v2["cam_rv"], v2["f_proj"]]
def body():
    n_ = n
    cls1 = _ops.cls_loss(v1["cam"], label); rvmin1 = _ops.min_pool_loss(v1["cam_rv"], label)
    n1 = _ops.fullres_norm(v1["cam"], label, (128, 128)); rv1 = _ops.fullres_norm(v1["cam_rv"], label, (128, 128))
    cls2 = _ops.cls_loss(v2["cam"], label); rvmin2 = _ops.min_pool_loss(v2["cam_rv"], label)
    n2 = _ops.fullres_norm(v2["cam"], label, None); rv2 = _ops.fullres_norm(v2["cam_rv"], label, None)
    er, ecr = _ops.er_ecr_loss(n1, n2, rv1, rv2)
    t = L.contrast_terms(v1["f_proj"], v1["cam_rv_down"], v2["f_proj"], v2["cam_rv_down"], label, 0.2, grid=(16, 16), keys=keys)
    total = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2 + er + ecr + t["loss_nce"]
    return total, torch.autograd.grad(total, leaves)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): body()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    res = body()
for _ in range(3): g.replay()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(20): g.replay()
torch.cuda.synchronize()
print(f"whole loss body (aux + contrast) as a CUDA graph: {(time.time() - t0) / 20 * 1e3:.3f} ms per fwd+bwd; loss {float(res[0]):.5f}")

# ---- Net.forward's tail (two x8 enlargements per view) + the loss body, materialised vs lazy -------------
small = {"cam1": (0.05 * torch.randn(n, 21, 56, 56, device=dev)).requires_grad_(True),
         "cam2": (0.05 * torch.randn(n, 21, 16, 16, device=dev)).requires_grad_(True),
         "rv1": v1["cam_rv_down"].detach().clone().requires_grad_(True), "rv2": v2["cam_rv_down"].detach().clone().requires_grad_(True),
         "f1": v1["f_proj"].detach().clone().requires_grad_(True), "f2": v2["f_proj"].detach().clone().requires_grad_(True)}
sleaves = list(small.values())
def graph_ms(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = fn()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): gr.replay()
    torch.cuda.synchronize()
    return (time.time() - t0) / reps * 1e3, out
def materialised():
    cam1, rv1 = _ops.bilinear_resize(small["cam1"], (448, 448)), _ops.bilinear_resize(small["rv1"], (448, 448))
    cam2, rv2 = _ops.bilinear_resize(small["cam2"], (128, 128)), _ops.bilinear_resize(small["rv2"], (128, 128))
    cls1 = _ops.cls_loss(cam1, label); rvmin1 = _ops.min_pool_loss(rv1, label)
    n1 = _ops.fullres_norm(cam1, label, (128, 128)); r1 = _ops.fullres_norm(rv1, label, (128, 128))
    cls2 = _ops.cls_loss(cam2, label); rvmin2 = _ops.min_pool_loss(rv2, label)
    n2 = _ops.fullres_norm(cam2, label, None); r2 = _ops.fullres_norm(rv2, label, None)
    er, ecr = _ops.er_ecr_loss(n1, n2, r1, r2)
    t = L.contrast_terms(small["f1"], small["rv1"], small["f2"], small["rv2"], label, 0.2, grid=(16, 16), keys=keys)
    total = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2 + er + ecr + t["loss_nce"]
    return total, torch.autograd.grad(total, sleaves)
def lazy():
    cls1 = _ops.cls_loss_up(small["cam1"], label, (448, 448)); rvmin1 = _ops.min_pool_loss_up(small["rv1"], label, (448, 448))
    n1 = _ops.fullres_norm_up(small["cam1"], label, (448, 448), (128, 128)); r1 = _ops.fullres_norm_up(small["rv1"], label, (448, 448), (128, 128))
    cls2 = _ops.cls_loss_up(small["cam2"], label, (128, 128)); rvmin2 = _ops.min_pool_loss_up(small["rv2"], label, (128, 128))
    n2 = _ops.fullres_norm_up(small["cam2"], label, (128, 128), None); r2 = _ops.fullres_norm_up(small["rv2"], label, (128, 128), None)
    er, ecr = _ops.er_ecr_loss(n1, n2, r1, r2)
    t = L.contrast_terms(small["f1"], small["rv1"], small["f2"], small["rv2"], label, 0.2, grid=(16, 16), keys=keys)
    total = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2 + er + ecr + t["loss_nce"]
    return total, torch.autograd.grad(total, sleaves)
ms_m, out_m = graph_ms(materialised)
ms_l, out_l = graph_ms(lazy)
print(f"enlargements + whole loss body as a CUDA graph: materialised {ms_m:.3f} ms (loss {float(out_m[0]):.5f}), "
      f"lazy {ms_l:.3f} ms (loss {float(out_l[0]):.5f})")
