"""First-light diagnostics for the PCM forward on a B200: checks each stage against torch on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
from oracle import reference_fn as R

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def decode_tiles(raw_u8, n, T, rows, cols):
    """[n][T] canonical K-major tiles of rows x cols bf16 -> [n, T*rows(if rows is pixel) ...]"""
    x = raw_u8.view(torch.bfloat16).reshape(n, T, cols // 8, rows, 8)      # [n,T,g,r,e]
    return x.permute(0, 1, 3, 2, 4).reshape(n, T, rows, cols).float()


def run(n, h, w, k=192, c=21, seed=0, positive=False):
    d = synth.pcm_inputs(n, h, w, seed, k=k, c=c, positive=positive)
    cam, f = d["cam"].to(dev), d["f"].to(dev)
    camn = ops.cam_norm(cam)
    ref_camn = R.cam_d_norm(cam)
    print(f"[{n}x{h}x{w}] cam_norm max abs err {(camn - ref_camn).abs().max().item():.3e}")
    out, ws = ops.pcm_forward(camn, f)
    torch.cuda.synchronize()
    lay = ops.pcm_layout(n, c, k, h, w)
    T, HWp, ns = lay["T"], lay["HWp"], lay["nsplit"]
    HW = h * w
    fh = f.reshape(n, k, HW)
    fh = fh / (fh.norm(dim=1, keepdim=True) + 1e-5)
    fh_p = torch.zeros(n, HWp, k, device=dev)
    fh_p[:, :HW] = fh.transpose(1, 2)
    fh_bf = fh_p.bfloat16().float()
    got_f = decode_tiles(ws[lay["fhat"]:lay["fhat"] + n * T * 128 * k * 2], n, T, 128, k).reshape(n, HWp, k)
    print(f"  prep f: max abs diff vs bf16(fhat) {(got_f - fh_bf).abs().max().item():.3e}")
    camb_raw = ws[lay["camb"]:lay["camb"] + n * T * 48 * 128 * 2].view(torch.bfloat16).reshape(n, T, 16, 48, 8)
    camb = camb_raw.permute(0, 3, 1, 2, 4).reshape(n, 48, HWp).float()       # [n,row,pixel]
    cam_p = torch.zeros(n, c, HWp, device=dev)
    cam_p[:, :, :HW] = ref_camn.reshape(n, c, HW)
    hi = cam_p.bfloat16().float()
    lo = (cam_p - hi).bfloat16().float()
    print(f"  prep cam: hi {(camb[:, :c] - hi).abs().max().item():.3e} lo {(camb[:, 24:24 + c] - lo).abs().max().item():.3e}"
          f" ones {(camb[:, c, :HW] - 1).abs().max().item():.3e}")
    S = torch.matmul(fh_bf, fh_bf.transpose(1, 2))                              # [n, j, i]
    P = torch.relu(S).bfloat16().float()
    num = torch.matmul(P, camb.transpose(1, 2))                                 # [n, j, 48]
    part = ws[lay["part"]:lay["part"] + n * ns * HWp * 48 * 4].view(torch.float32).reshape(n, ns, HWp, 48).sum(1)
    err = (part - num).abs()
    print(f"  main: partial sums max abs err {err.max().item():.3e} (scale {num.abs().max().item():.3e}); "
          f"colsum err {(part[..., c] - num[..., c]).abs().max().item():.3e}")
    if err.max().item() > 1e-2 * num.abs().max().item():
        bad = (err > 1e-2 * num.abs().max()).nonzero()
        print("  first bad entries (n, j, col):", bad[:8].tolist(), " count", bad.shape[0])
        print("  got ", part[0, :4, :6].tolist())
        print("  want", num[0, :4, :6].tolist())
    ref = R.pcm(ref_camn, f)
    e = (out - ref).abs().max().item()
    print(f"  cam_rv: max abs err {e:.3e}, max |ref| {ref.abs().max().item():.3e}, rel(max) {e / ref.abs().max().item():.3e}")
    # timing
    for _ in range(3):
        ops.pcm_forward(camn, f)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        ops.pcm_forward(camn, f)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    fl = 2.0 * n * HW * HW * (k + c)
    print(f"  fwd {ms * 1e3:.1f} us  -> {fl / ms / 1e9:.1f} TFLOP/s algorithmic")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(1, 16, 16, seed=1)
    run(2, 20, 13, seed=2)
    run(8, 56, 56, seed=3)
    run(8, 56, 56, seed=4, positive=True)
    run(1, 94, 125, seed=5)
