"""Bring-up diagnostics for the CTA-pair PCM kernels on a B200: every stage against torch on the GPU.
usage: python tools/diag_pcm.py [fast|precise] [case ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth
from oracle import reference_fn as R

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def section(ws, lay, name, nbytes):
    return ws[lay[name]:lay[name] + nbytes]


def decode_tiles(raw_u8, n, tiles, rows, cols, dtype=torch.float16):
    x = raw_u8.view(dtype).reshape(n, tiles, cols // 8, rows, 8)          # [n,T,g,r,e]
    return x.permute(0, 1, 3, 2, 4).reshape(n, tiles * rows, cols).float()


def run(n, h, w, k=192, c=21, seed=0, positive=False, time_it=True):
    mode = ops.get_pcm_mode()
    d = synth.pcm_inputs(n, h, w, seed, k=k, c=c, positive=positive)
    cam, f, g = d["cam"].to(dev), d["f"].to(dev), d["grad_out"].to(dev)
    camn = ops.cam_norm(cam)
    lay = ops.pcm_layout(n, c, k, h, w)
    PJ, HWpj, wb, NS, HWpi, ns = lay["PJ"], lay["HWpj"], lay["wb"], lay["NS"], lay["HWpi"], lay["nsplit"]
    HW = h * w
    print(f"[{mode} {n}x{h}x{w} pos={positive}] PJ={PJ} wb={wb} NS={NS} nsplit={ns} ws={lay['bytes'] / 1e6:.1f} MB")
    out, ws = ops.pcm_forward(camn, f)
    torch.cuda.synchronize()
    # ---- prep
    fh = f.reshape(n, k, HW).double()
    fh = fh / (fh.norm(dim=1, keepdim=True) + 1e-5)
    fhp = torch.zeros(n, max(HWpj, HWpi), k, device=dev, dtype=torch.float64)
    fhp[:, :HW] = fh.transpose(1, 2)
    hi = fhp.float().half().double()
    a_hi = decode_tiles(section(ws, lay, "fa", n * 2 * PJ * 128 * k * 2), n, 2 * PJ, 128, k).double()
    b_hi = decode_tiles(section(ws, lay, "fb", n * 2 * NS * wb * k * 2), n, 2 * NS, wb, k).double()
    print(f"  prep: A err {(a_hi - hi[:, :HWpj]).abs().max().item():.2e}  B err {(b_hi - hi[:, :HWpi]).abs().max().item():.2e}")
    if mode == "precise":
        f32 = section(ws, lay, "fh32", n * HWpj * k * 4).view(torch.float32).reshape(n, HWpj, k).double()
        print(f"  prep: fh32 err {(f32 - fhp[:, :HWpj]).abs().max().item():.2e} (fp16 alone {(a_hi - fhp[:, :HWpj]).abs().max().item():.2e})")
        nz = section(ws, lay, "nzmask", n * NS * HWpj * 16).view(torch.int32).reshape(n, NS, HWpj, 4)
        order0 = torch.stack([torch.arange(16), torch.arange(16) + 16], 1).reshape(-1).to(dev).int()
        nzb = ((nz.unsqueeze(-1) >> order0) & 1).reshape(n, NS, HWpj, 128)[..., :2 * wb].permute(0, 2, 1, 3).reshape(n, HWpj, HWpi).bool()
        Sh = torch.matmul(hi[:, :HWpj], hi[:, :HWpi].transpose(1, 2))
        S0 = torch.matmul(fhp[:, :HWpj], fhp[:, :HWpi].transpose(1, 2))
        inside = torch.zeros_like(nzb); inside[:, :HW, :HW] = True
        print(f"  near-zero mask: {int(nzb.sum())} entries marked ({int(nzb.sum()) / (n * HW * HW):.2e}); outside the image {int((nzb & ~inside).sum())}; "
              f"max |S_fp16| marked {Sh[nzb].abs().max().item() if int(nzb.sum()) else 0:.2e}; "
              f"min |S_fp16| unmarked inside {Sh[~nzb & inside].abs().min().item():.2e}; max |S_fp16 - S| {(Sh - S0)[inside].abs().max().item():.2e}")
        del Sh, S0, nzb
    # ---- S sign mask
    S = torch.matmul(fhp[:, :HWpj], fhp[:, :HWpi].transpose(1, 2))              # [n, j, i] fp64
    mask_raw = section(ws, lay, "mask", n * NS * HWpj * 16).view(torch.int32).reshape(n, NS, HWpj, 4)
    # word c, bit m = column 32c + 2m, bit 16+m = column 32c + 2m + 1
    order = torch.stack([torch.arange(16), torch.arange(16) + 16], 1).reshape(-1).to(dev).int()
    bits = ((mask_raw.unsqueeze(-1) >> order) & 1).reshape(n, NS, HWpj, 128)
    bits = bits[..., :2 * wb].permute(0, 2, 1, 3).reshape(n, HWpj, HWpi).bool()
    want = S > 0
    live = torch.zeros_like(want)
    live[:, :HW, :HW] = True
    bad = (bits != want) & live
    nbad = int(bad.sum())
    print(f"  mask: {nbad} of {int(live.sum())} sign bits differ from fp64 ({nbad / max(1, int(live.sum())):.2e})"
          + (f"; max |S| among them {S[bad].abs().max().item():.2e}" if nbad else ""))
    # ---- partial sums
    part = section(ws, lay, "part", n * ns * HWpj * 64 * 4).view(torch.float32).reshape(n, ns, HWpj, 64).sum(1).double()
    camp = torch.zeros(n, c, HWpi, device=dev, dtype=torch.float64)
    camp[:, :, :HW] = camn.reshape(n, c, HW).double()
    P = torch.relu(S)
    num = torch.matmul(P, camp.transpose(1, 2))                                   # [n, j, c]
    colsum = P[:, :, :HW].sum(-1)
    got_num = part[..., :c] + part[..., 32:32 + c]
    e_num = (got_num[:, :HW] - num[:, :HW]).abs().max().item()
    e_col = (part[:, :HW, c] - colsum[:, :HW]).abs().max().item()
    print(f"  main: numerator err {e_num:.2e} (scale {num.abs().max().item():.2e}); colsum err {e_col:.2e} (scale {colsum.max().item():.2e})")
    ref = R.pcm(camn, f)
    e = (out - ref).abs().max().item()
    print(f"  cam_rv: max abs err {e:.3e}, rel(max) {e / ref.abs().max().item():.3e}")
    # ---- backward
    f_ref = f.clone().requires_grad_(True)
    R.pcm(camn, f_ref).backward(g)
    gm = ops.pcm_backward(f, camn, out, g, ws)
    torch.cuda.synchronize()
    gr = f_ref.grad
    print(f"  grad_f: rel(fro) {((gr - gm).norm() / gr.norm()).item():.3e}  rel(max) {((gr - gm).abs().max() / gr.abs().max()).item():.3e}")
    if time_it:
        ops.pcm_profile(True)
        fw, bw = [], []
        for i in range(8):
            o2, ws2 = ops.pcm_forward(camn, f)
            ops.pcm_backward(f, camn, o2, g, ws2)
            a, b = ops.pcm_main_ms()
            if i >= 3:
                fw.append(a); bw.append(b)
        ops.pcm_profile(False)
        fl_f = 2.0 * n * HW * HW * (k + c); fl_b = 2.0 * n * HW * HW * (2 * k + c)
        a, b = sum(fw) / len(fw), sum(bw) / len(bw)
        print(f"  main kernels: fwd {a * 1e3:.1f} us ({fl_f / a / 1e9:.0f} TF/s alg)  bwd {b * 1e3:.1f} us ({fl_b / b / 1e9:.0f} TF/s alg)"
              f"  sign fix-up {ops.pcm_fixup_ms() * 1e3:.1f} us")


CASES = {
    "a": dict(n=1, h=16, w=16, seed=1),
    "b": dict(n=2, h=20, w=13, seed=2),
    "c": dict(n=1, h=64, w=64, seed=6, positive=True),
    "d": dict(n=8, h=56, w=56, seed=3),
    "e": dict(n=8, h=56, w=56, seed=4, positive=True),
    "f": dict(n=1, h=94, w=125, seed=5),
}

if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] in ("fast", "precise"):
        ops.set_pcm_mode(args[0])
        args = args[1:]
    print(torch.cuda.get_device_name(0), "mode", ops.get_pcm_mode())
    for name in (args or list(CASES)):
        run(**CASES[name])
