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
