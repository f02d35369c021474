"""ORACLE — test infrastructure.  Generates tests/golden/*.npz by running the REFERENCE'S OWN code
(oracle/verbatim.py: exec of contrast_train.py:129-395, network/resnet38_contrast.py Net.forward / Net.PCM) on CPU
in the build container, on inputs from wseg_b200.synth (regenerated from seeds at test time; checksums stored).

    python -m oracle.make_golden          # needs /root/reference (or $WSEG_REFERENCE)

torch.topk ties are pinned to the canonical order (value, then lowest index) while the reference code runs —
see oracle/verbatim.py.  Everything else is the reference verbatim.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import verbatim                      # noqa: E402
from wseg_b200 import synth                      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LOSS_CASE = dict(n=2, view1=(192, 192), view2=(128, 128), seed1=11, seed2=12, label_seed=5, py_seed=7,
                 bg_threshold=0.2)
PCM_CASE = dict(n=2, h=20, w=13, seed=21)
NET_CASE = dict(n=1, size=(72, 56), seed=31)
# tool/torchutils.py PolyOptimizer as contrast_train.py:91-96 builds it (four groups: lr, 2lr, 10lr, 20lr; weight decay on
# the weight groups only), stepped past max_step
OPT_CASE = dict(seed=41, shapes=((5, 3), (7,), (4, 2, 3), (6,)), lr=0.01, wt_dec=5e-4, max_step=5, steps=7)


def optim_inputs():
    """Initial parameters and per-step gradients of OPT_CASE (regenerated from the seed at test time)."""
    c = OPT_CASE
    g = synth.gen(c["seed"])
    params = [torch.randn(*s, generator=g) for s in c["shapes"]]
    grads = [[0.1 * torch.randn(*s, generator=g) for s in c["shapes"]] for _ in range(c["steps"])]
    return params, grads


def optim_groups(params, lr, wt_dec):
    return [{'params': [params[0]], 'lr': lr, 'weight_decay': wt_dec}, {'params': [params[1]], 'lr': 2 * lr, 'weight_decay': 0},
            {'params': [params[2]], 'lr': 10 * lr, 'weight_decay': wt_dec}, {'params': [params[3]], 'lr': 20 * lr, 'weight_decay': 0}]


def checksum(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * torch.arange(1, t.numel() + 1) % 7).sum().item()])


def sample(t: torch.Tensor, stride: int = 97) -> np.ndarray:
    return t.detach().reshape(-1)[::stride].numpy().copy()


def loss_case():
    c = LOSS_CASE
    v1 = synth.head_outputs(c["n"], c["view1"], seed=c["seed1"])
    v2 = synth.head_outputs(c["n"], c["view2"], seed=c["seed2"])
    lab = synth.labels(c["n"], c["label_seed"])
    a1 = {k: v.clone().requires_grad_(True) for k, v in v1.items()}
    a2 = {k: v.clone().requires_grad_(True) for k, v in v2.items()}
    stub = verbatim.HeadStub()
    stub.add(c["view1"], a1["cam"], a1["cam_rv"], a1["f_proj"], a1["cam_rv_down"])
    stub.add(c["view2"], a2["cam"], a2["cam_rv"], a2["f_proj"], a2["cam_rv_down"])
    img = torch.zeros(c["n"], 3, *c["view1"])
    ns = verbatim.run_loop_body(stub, img, lab, c["bg_threshold"], py_seed=c["py_seed"], canonical_ties=True)
    ns["loss"].backward()
    out = {k: np.float64(float(ns[k])) for k in ("loss", "loss_cls", "loss_er", "loss_ecr", "loss_nce",
                                                  "loss_intra_nce", "loss_cross_nce", "loss_cross_nce2")}
    out.update({
        "pseudo_label1": ns["pseudo_label1"].numpy().astype(np.int32),
        "pseudo_label2": ns["pseudo_label2"].numpy().astype(np.int32),
        "top_indices1": ns["top_indices"].numpy().astype(np.int32),
        "top_indices2": ns["top_indices2"].numpy().astype(np.int32),
        "prototypes1": ns["prototypes1"].numpy(), "prototypes2": ns["prototypes2"].numpy(),
        "norm_cam1": ns["norm_cam1"].numpy(), "norm_cam2": ns["norm_cam2"].numpy(),
        "grad_f_proj1": a1["f_proj"].grad.numpy(), "grad_f_proj2": a2["f_proj"].grad.numpy(),
        "grad_cam1_sample": sample(a1["cam"].grad), "grad_cam_rv1_sample": sample(a1["cam_rv"].grad),
        "grad_cam2_sample": sample(a2["cam"].grad), "grad_cam_rv2_sample": sample(a2["cam_rv"].grad),
        "grad_cam1_sum": checksum(a1["cam"].grad), "grad_cam_rv1_sum": checksum(a1["cam_rv"].grad),
        "in_sum_v1": np.stack([checksum(v1[k]) for k in ("cam", "cam_rv", "f_proj", "cam_rv_down")]),
        "in_sum_v2": np.stack([checksum(v2[k]) for k in ("cam", "cam_rv", "f_proj", "cam_rv_down")]),
        "label": lab.numpy(),
    })
    np.savez_compressed(os.path.join(OUT, "loss_n2_192_128.npz"), **out)
    print("loss case:", {k: float(out[k]) for k in ("loss", "loss_nce", "loss_intra_nce")})


def pcm_case():
    c = PCM_CASE
    mod = verbatim.import_reference_net()
    d = synth.pcm_inputs(c["n"], c["h"], c["w"], c["seed"])
    # Net.PCM verbatim (resnet38_contrast.py:63-75) with the f9 conv replaced by the identity
    holder = types.SimpleNamespace(f9=torch.nn.Identity())
    # cam_d_norm exactly as Net.forward computes it (:41-48), via a captured call below in net_case; here restate
    # is NOT used: run the verbatim lines on the synthetic logits by calling forward's block through exec
    src = verbatim._source_lines("network/resnet38_contrast.py", 42, 48)
    env = {"torch": torch, "F": torch.nn.functional, "cam": d["cam"], "n": c["n"], "c": 21}
    with torch.no_grad():
        exec(compile(src, "resnet38_contrast.py[42:48]", "exec"), env)
    cam_d_norm = env["cam_d_norm"]
    f = d["f"].clone().requires_grad_(True)
    out = mod.Net.PCM(holder, cam_d_norm, f)
    out.backward(d["grad_out"])
    np.savez_compressed(os.path.join(OUT, "pcm_n2_20x13.npz"), cam_d_norm=cam_d_norm.numpy(),
                        cam_rv=out.detach().numpy(), grad_f=f.grad.numpy(),
                        in_sum=np.stack([checksum(d[k]) for k in ("cam", "f", "grad_out")]))
    print("pcm case: cam_rv", tuple(out.shape), float(out.abs().max()))


def net_case():
    """The unmodified reference Net.forward on a small image with seeded random-init weights: capture what enters
    and leaves the hot path (fc8 logits -> cam_d_norm, f -> f9 -> PCM -> cam_rv_down)."""
    c = NET_CASE
    mod = verbatim.import_reference_net()
    torch.manual_seed(c["seed"])
    net = mod.Net()
    net.eval()
    cap = {}
    orig_pcm = net.PCM

    def spy(cam_d_norm, f):
        cap["cam_d_norm"] = cam_d_norm.detach().clone()
        cap["f"] = f.detach().clone()
        return orig_pcm(cam_d_norm, f)

    net.PCM = spy
    net.fc8.register_forward_hook(lambda m, i, o: cap.__setitem__("cam_logits", o.detach().clone()))
    x = torch.randn(c["n"], 3, *c["size"], generator=synth.gen(c["seed"]))
    with torch.no_grad():
        cam, cam_rv, f_proj, cam_rv_down = net(x)
    np.savez_compressed(os.path.join(OUT, "net_head_72x56.npz"), cam_logits=cap["cam_logits"].numpy(),
                        cam_d_norm=cap["cam_d_norm"].numpy(), f=cap["f"].numpy(), f9_weight=net.f9.weight.detach().numpy(),
                        cam_rv_down=cam_rv_down.numpy(), cam_rv_sample=sample(cam_rv, 13))
    print("net case: grid", tuple(cam_rv_down.shape))


def optim_case():
    c = OPT_CASE
    init, grads = optim_inputs()
    with verbatim.reference_on_path():
        from tool import torchutils                      # the reference's own class
        params = [torch.nn.Parameter(p.clone()) for p in init]
        opt = torchutils.PolyOptimizer(optim_groups(params, c["lr"], c["wt_dec"]), lr=c["lr"], weight_decay=c["wt_dec"],
                                       max_step=c["max_step"])
        out = {"group_momentum": np.array([g["momentum"] for g in opt.param_groups]),
               "group_weight_decay": np.array([g["weight_decay"] for g in opt.param_groups])}
        lrs = []
        for step in range(c["steps"]):
            for p, g in zip(params, grads[step]):
                p.grad = g.clone()
            opt.step()
            lrs.append([g["lr"] for g in opt.param_groups])
            for i, p in enumerate(params):
                out[f"p{i}_step{step}"] = p.detach().numpy().copy()
        out["lrs"] = np.array(lrs)
        out["global_step"] = np.array(opt.global_step)
    np.savez_compressed(os.path.join(OUT, "poly_optimizer.npz"), **out)
    print("optimizer case: lrs", out["lrs"][:, 0])


if __name__ == "__main__":
    if not verbatim.available():
        raise SystemExit("reference checkout not found; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    loss_case()
    pcm_case()
    net_case()
    optim_case()
