"""Loss-path diagnostics on a B200: K1b/K1c/K3/K4 vs the oracle (ATen fp32 on the same GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth, loss as L
from oracle import reference_fn as R

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run(n, H, W, v2=(128, 128), seed=0, d=128):
    print(f"=== N={n} view1={H}x{W} view2={v2} D={d}")
    a = {k: v.to(dev) for k, v in synth.head_outputs(n, (H, W), seed=seed * 10 + 1, d=d).items()}
    b = {k: v.to(dev) for k, v in synth.head_outputs(n, v2, seed=seed * 10 + 2, d=d).items()}
    lab20 = synth.labels(n, seed).to(dev)
    label = synth.label_with_bg(lab20)
    # ---- K1b
    for name, t, size in (("cam1", a["cam"], v2), ("cam_rv2", b["cam_rv"], None)):
        x = t.clone().requires_grad_(True)
        y = ops.fullres_norm(x, label, size)
        g = torch.randn_like(y)
        y.backward(g)
        xr = t.clone().requires_grad_(True)
        yr = R.fullres_norm(xr, label.reshape(n, -1, 1, 1), size)
        yr.backward(g)
        print(f"  K1b {name}: fwd rel {rel(y, yr):.2e}  bwd rel {rel(x.grad, xr.grad):.2e}")
    # ---- contrast terms
    f1 = a["f_proj"].clone().requires_grad_(True); f2 = b["f_proj"].clone().requires_grad_(True)
    mine = L.contrast_terms(f1, a["cam_rv_down"], f2, b["cam_rv_down"], label, 0.2, grid=(v2[0] // 8, v2[1] // 8), rng=7)
    mine["loss_nce"].backward()
    r1 = a["f_proj"].clone().requires_grad_(True); r2 = b["f_proj"].clone().requires_grad_(True)
    ref = R.contrast_terms(r1, a["cam_rv_down"], r2, b["cam_rv_down"], label.reshape(n, -1, 1, 1), 0.2,
                           sampler=R.ReferenceSampler(7), grid=(v2[0] // 8, v2[1] // 8), argmax_of="softmax")
    ref["loss_nce"].backward()
    for v in "12":
        print(f"  view{v}: norm_cam max abs {(mine['norm_cam' + v] - ref['norm_cam' + v]).abs().max().item():.2e}  "
              f"pseudo equal {bool((mine['pseudo_label' + v].long() == ref['pseudo_label' + v]).all())}  "
              f"topk idx equal {bool((mine['top_indices' + v].long() == ref['top_indices' + v]).all())}  "
              f"proto max abs {(mine['prototypes' + v] - ref['prototypes' + v]).abs().max().item():.2e}")
    for k in ("loss_nce", "loss_intra_nce", "loss_cross_nce", "loss_cross_nce2"):
        print(f"  {k}: mine {float(mine[k]):.7f} ref {float(ref[k]):.7f} rel {abs(float(mine[k]) - float(ref[k])) / abs(float(ref[k])):.2e}")
    print(f"  grad f_proj1 rel {rel(f1.grad, r1.grad):.2e}  grad f_proj2 rel {rel(f2.grad, r2.grad):.2e}")
    # mined sets
    m = L.contrast_terms(a["f_proj"], a["cam_rv_down"], b["f_proj"], b["cam_rv_down"], label, 0.2,
                         grid=(v2[0] // 8, v2[1] // 8), rng=7, want_mined=True)
    for v in "12":
        cnt = torch.zeros_like(m["mined" + v])
        for _, rnd, hard in ref["sets" + v][1]:
            cnt.index_add_(0, rnd, torch.ones_like(rnd, dtype=torch.int32))
            cnt.index_add_(0, hard, torch.ones_like(hard, dtype=torch.int32))
        print(f"  view{v}: mined multiplicities equal {bool((cnt == m['mined' + v]).all())} (mismatch {(cnt != m['mined' + v]).sum().item()})")
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(20):
        L.contrast_terms(a["f_proj"], a["cam_rv_down"], b["f_proj"], b["cam_rv_down"], label, 0.2, grid=(v2[0] // 8, v2[1] // 8))
    torch.cuda.synchronize()
    print(f"  contrast_terms wall {1e3 * (time.time() - t0) / 20:.3f} ms/iter (eager, incl. python)")


if __name__ == "__main__":
    run(2, 192, 192, seed=1)
    run(8, 448, 448, seed=2)
    run(4, 256, 320, v2=(160, 96), seed=3)
