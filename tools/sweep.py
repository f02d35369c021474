"""per-kernel sweep over grid size, batch and projection width, plus the stock-ATen fp32
PCM (the reference's own GPU formulation, network/resnet38_contrast.py:70-73 through autograd) timed on the same GPU.

usage: python tools/sweep.py [--out gpurun_out/sweep.json] [--quick]
Each number: median of `iters` launches, CUDA events on the launching stream, 256 MiB L2 flush before every launch.
Nothing here is a bench value; bench.py owns the headline metric."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from wseg_b200 import loss as L
from wseg_b200 import ops, synth

DEV = torch.device("cuda:0")
HBM_GBS, BF16_TFS = 6445.0, 1700.0
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fh:
        _p = json.load(fh)
    HBM_GBS = float(_p.get("hbm_gbs", HBM_GBS))
    BF16_TFS = float(_p.get("bf16_tflops", BF16_TFS))
except Exception:
    pass


def timed(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


def aten_pcm(cam, f):
    """Plain torch fp32 statement of the module (what the reference executes on a GPU)."""
    n, c, h, w = cam.shape
    cam = cam.view(n, -1, h * w)
    f = f.view(n, -1, h * w)
    f = f / (torch.norm(f, dim=1, keepdim=True) + 1e-5)
    aff = F.relu(torch.matmul(f.transpose(1, 2), f))
    aff = aff / (torch.sum(aff, dim=1, keepdim=True) + 1e-5)
    return torch.matmul(cam, aff).view(n, -1, h, w)


def sweep_pcm(rows, sizes, batches, iters, flush):
    ops.pcm_profile(True)
    for hw in sizes:
        for n in batches:
            lay = ops.pcm_layout(n, 21, 192, hw, hw)
            if lay["bytes"] > 60e9:
                continue
            g = torch.Generator(device="cpu").manual_seed(1000 + hw + n)
            f = torch.randn(n, 192, hw, hw, generator=g).to(DEV)
            cam = (0.05 * torch.randn(n, 21, hw, hw, generator=g)).to(DEV)
            gout = (1e-3 * torch.randn(n, 21, hw, hw, generator=g)).to(DEV)
            camn = ops.cam_norm(cam)
            state = {}

            def fwd():
                state["out"], state["ws"] = ops.pcm_forward(camn, f)

            def bwd():
                state["gf"] = ops.pcm_backward(f, camn, state["out"], gout, state["ws"])

            t_f = timed(fwd, iters, flush)
            main_f = ops.pcm_main_ms()[0]
            t_b = timed(bwd, iters, flush)
            main_b = ops.pcm_main_ms()[1]
            with torch.no_grad():
                t_i = timed(lambda: ops.pcm(camn, f), iters, flush)
            HW = hw * hw
            ff, fb = 2.0 * n * HW * HW * (192 + 21), 2.0 * n * HW * HW * (2 * 192 + 21)
            row = {"kernel": "pcm", "N": n, "h": hw, "fwd_ms": t_f, "fwd_main_ms": main_f, "bwd_ms": t_b,
                   "bwd_main_ms": main_b, "infer_ms": t_i,
                   "fwd_main_tflops": ff / main_f / 1e9, "bwd_main_tflops": fb / main_b / 1e9,
                   "fwdbwd_tflops_incl_prep": (ff + fb) / (t_f + t_b) / 1e9,
                   "frac_of_bf16_peak_incl_prep": (ff + fb) / (t_f + t_b) / 1e9 / BF16_TFS,
                   "infer_tflops": ff / t_i / 1e9, "wb": lay["wb"], "nsplit": lay["nsplit"],
                   "workspace_MB": lay["bytes"] / 1e6}
            if n == batches[0]:
                # stock ATen fp32 (allow_tf32 False): forward, and forward + backward w.r.t. f
                torch.backends.cuda.matmul.allow_tf32 = False
                fr = f.clone().requires_grad_(True)
                with torch.no_grad():
                    row["aten_fwd_ms"] = timed(lambda: aten_pcm(camn, f), max(3, iters // 2), flush)

                def both():
                    fr.grad = None
                    aten_pcm(camn, fr).backward(gout)

                row["aten_fwdbwd_ms"] = timed(both, max(3, iters // 2), flush)
                row["speedup_fwdbwd_vs_aten"] = row["aten_fwdbwd_ms"] / (t_f + t_b)
                del fr
            rows.append(row)
            print(json.dumps(row), flush=True)
            del f, cam, gout, camn, state
            torch.cuda.empty_cache()
    ops.pcm_profile(False)


def sweep_loss(rows, sizes, batches, dims, iters, flush):
    """K1c + K3 + K4 on the GENERALISED grid (loss grid = feature grid), forward + backward in one call chain."""
    for hw in sizes:
        for n in batches:
            for d in dims:
                if n * hw * hw > (1 << 17):          # keeps the sweep short (no size limit in the kernels)
                    continue
                g = torch.Generator(device=DEV).manual_seed(5 + hw + n + d)
                a = {"f_proj": F.relu(0.02 * torch.randn(n, d, hw, hw, generator=g, device=DEV) + 0.01),
                     "cam_rv_down": (0.5 * torch.randn(n, 21, hw, hw, generator=g, device=DEV)).abs()}
                b = {"f_proj": F.relu(0.02 * torch.randn(n, d, hw, hw, generator=g, device=DEV) + 0.01),
                     "cam_rv_down": (0.5 * torch.randn(n, 21, hw, hw, generator=g, device=DEV)).abs()}
                label = synth.label_with_bg(synth.labels(n, seed=n)).to(DEV)
                f1 = a["f_proj"].clone().requires_grad_(True)
                f2 = b["f_proj"].clone().requires_grad_(True)
                P = n * hw * hw
                k1, k2 = ops.hash_keys(P, 1, DEV), ops.hash_keys(P, 2, DEV)

                def run():
                    t = L.contrast_terms(f1, a["cam_rv_down"], f2, b["cam_rv_down"], label, 0.2, grid=(hw, hw),
                                         keys=(k1, k2))
                    torch.autograd.grad(t["loss_nce"], (f1, f2))

                ms = timed(run, iters, flush)
                # the same as ONE CUDA graph (what a training step replays; the eager number includes the host's launch
                # overhead of the chain's ~45 small kernels)
                ms_graph = None
                try:
                    side = torch.cuda.Stream(device=DEV)
                    side.wait_stream(torch.cuda.current_stream(DEV))
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            run()
                    torch.cuda.current_stream(DEV).wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        run()
                    ms_graph = timed(graph.replay, iters, flush)
                    del graph
                except Exception as exc:                      # keep the sweep going; the eager number stands
                    print(f"# graph capture failed at N={n} grid={hw} D={d}: {type(exc).__name__}: {exc}", flush=True)
                # algorithmic bytes: K1c + K3 + K4 of both views
                c = 21
                g2 = hw * hw
                k = g2 // 8
                by = 2 * (n * c * g2 * 4 * 2 + n * g2 * 4) + 2 * (n * c * g2 * 4 + c * k * d * 4 + c * d * 4) \
                    + (2 * P * d * 4 * 2 + 2 * c * d * 4 + P * 4 * 2)
                row = {"kernel": "seed+proto+nce (K1c,K3,K4 fwd+bwd, both views)", "N": n, "grid": hw, "D": d,
                       "ms": ms, "algorithmic_MB": by / 1e6, "GBps": by / ms / 1e6, "frac_of_hbm": by / ms / 1e6 / HBM_GBS}
                if ms_graph:
                    row.update({"ms_graph": ms_graph, "GBps_graph": by / ms_graph / 1e6,
                                "frac_of_hbm_graph": by / ms_graph / 1e6 / HBM_GBS})
                rows.append(row)
                print(json.dumps(row), flush=True)
                del a, b, f1, f2
                torch.cuda.empty_cache()


def sweep_fullres(rows, cases, iters, flush):
    for n, hw, size in cases:
        g = torch.Generator(device="cpu").manual_seed(77)
        p = torch.randn(n, 21, hw, hw, generator=g).to(DEV).requires_grad_(True)
        label = synth.label_with_bg(synth.labels(n, seed=3)).to(DEV)
        oh, ow = size if size is not None else (hw, hw)
        go = torch.randn(n, 21, oh, ow, generator=g).to(DEV)
        st = {}

        def fwd():
            st["y"] = ops.fullres_norm(p, label, size)

        def bwd():
            st["g"] = torch.autograd.grad(st["y"], p, go, retain_graph=True)

        tf = timed(fwd, iters, flush)
        tb = timed(bwd, iters, flush)
        by = n * 21 * (hw * hw + oh * ow) * 4
        rows.append({"kernel": "fullres_norm (K1b)", "N": n, "H": hw, "out": [oh, ow], "fwd_ms": tf, "bwd_ms": tb,
                     "algorithmic_MB_each_way": by / 1e6, "fwd_GBps": by / tf / 1e6, "bwd_GBps": by / tb / 1e6,
                     "fwd_frac_of_hbm": by / tf / 1e6 / HBM_GBS, "bwd_frac_of_hbm": by / tb / 1e6 / HBM_GBS})
        print(json.dumps(rows[-1]), flush=True)
        del p, go, st
        torch.cuda.empty_cache()
        # the same operator on the LAZY map: input = the stride-8 map, the x8 enlargement evaluated on the fly;
        # compared with enlargement + K1b on the materialised tensor (forward and backward each)
        small = torch.randn(n, 21, hw // 8, hw // 8, generator=g).to(DEV).requires_grad_(True)
        go = torch.randn(n, 21, oh, ow, generator=g).to(DEV)
        st = {}

        def lazy_fwd():
            st["y"] = ops.fullres_norm_up(small, label, (hw, hw), size)

        def lazy_bwd():
            st["g"] = torch.autograd.grad(st["y"], small, go, retain_graph=True)

        def eager_fwd():
            st["ye"] = ops.fullres_norm(ops.bilinear_resize(small, (hw, hw)), label, size)

        def eager_bwd():
            st["ge"] = torch.autograd.grad(st["ye"], small, go, retain_graph=True)

        lf, lb = timed(lazy_fwd, iters, flush), timed(lazy_bwd, iters, flush)
        ef, eb = timed(eager_fwd, iters, flush), timed(eager_bwd, iters, flush)
        rows.append({"kernel": "x8 enlargement + fullres_norm: materialised vs lazy", "N": n, "H": hw, "out": [oh, ow],
                     "materialised_fwd_ms": ef, "materialised_bwd_ms": eb, "lazy_fwd_ms": lf, "lazy_bwd_ms": lb,
                     "speedup_fwd_bwd": (ef + eb) / (lf + lb)})
        print(json.dumps(rows[-1]), flush=True)
        del small, go, st
        torch.cuda.empty_cache()


def main():
    ops.set_pcm_pairs(72, 72)      # kernels alone on the GPU: 72 of the 74 SM pairs (the training step runs them on fewer, bench.py)
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    iters = 5 if args.quick else 9
    rows = []
    t0 = time.time()
    if args.quick:
        sweep_pcm(rows, [32, 56], [8], iters, flush)
        sweep_loss(rows, [32], [8], [128], iters, flush)
        sweep_fullres(rows, [(2, 448, (128, 128))], iters, flush)
    else:
        sweep_pcm(rows, [32, 48, 56, 64, 80, 96], [8, 16, 32, 64, 128], iters, flush)
        sweep_loss(rows, [32, 56, 96], [8, 32, 128], [128, 256], iters, flush)
        sweep_fullres(rows, [(8, 448, (128, 128)), (8, 128, None), (16, 512, (256, 256)), (16, 256, None)], iters, flush)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump({"peaks": {"hbm_gbs": HBM_GBS, "bf16_tflops": BF16_TFS}, "rows": rows,
                   "seconds": time.time() - t0}, fh, indent=1)
    print("sweep done in %.1f s, %d rows" % (time.time() - t0, len(rows)))


if __name__ == "__main__":
    main()
