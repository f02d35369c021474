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
