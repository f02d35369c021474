"""PCM backward diagnostics on a B200: closed-form tcgen05 backward vs autograd of the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
from oracle import reference_fn as R

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def run(n, h, w, k=192, c=21, seed=0, positive=False, gscale=1.0):
    d = synth.pcm_inputs(n, h, w, seed, k=k, c=c, positive=positive)
    cam, f, g = d["cam"].to(dev), d["f"].to(dev), d["grad_out"].to(dev) * gscale
    camn = ops.cam_norm(cam)
    f_ref = f.clone().requires_grad_(True)
    ref = R.pcm(camn, f_ref)
    ref.backward(g)
    f_my = f.clone().requires_grad_(True)
    out = ops.pcm(camn, f_my)
    out.backward(g)
    torch.cuda.synchronize()
    gr, gm = f_ref.grad, f_my.grad
    err = (gr - gm).abs().max().item()
    print(f"[{n}x{h}x{w} pos={positive}] fwd rel(max) {(out - ref).abs().max().item() / ref.abs().max().item():.3e}  "
          f"bwd max abs err {err:.3e}, max |ref grad| {gr.abs().max().item():.3e}, rel(max) {err / gr.abs().max().item():.3e}, "
          f"rel(fro) {((gr - gm).norm() / gr.norm()).item():.3e}")
    # timing of bwd alone
    _, ws = ops.pcm_forward(camn, f)
    for _ in range(3):
        ops.pcm_backward(f, camn, out.detach(), g, ws)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        ops.pcm_backward(f, camn, out.detach(), g, ws)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    fl = 2.0 * n * (h * w) ** 2 * (2 * k + c)
    print(f"  bwd {ms * 1e3:.1f} us -> {fl / ms / 1e9:.1f} TFLOP/s algorithmic")


if __name__ == "__main__":
    run(1, 16, 16, seed=1)
    run(2, 20, 13, seed=2)
    run(8, 56, 56, seed=3)
    run(8, 56, 56, seed=4, positive=True)
    run(1, 94, 125, seed=5)
