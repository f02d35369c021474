"""Localise a K4 gradient mismatch: per-pixel error of grad wrt the loss-grid features."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from wseg_b200 import ops, synth, loss as L
from oracle import reference_fn as R
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
n, hw1, hw2, d = (16, (512, 512), (256, 256), 128) if len(sys.argv) > 1 else (4, (256, 320), (160, 96), 128)
seed = n if len(sys.argv) > 1 else 3
a = {k: v.to(dev) for k, v in synth.head_outputs(n, hw1, seed=seed * 10 + 1, d=d).items()}
b = {k: v.to(dev) for k, v in synth.head_outputs(n, hw2, seed=seed * 10 + 2, d=d).items()}
label = synth.label_with_bg(synth.labels(n, seed)).to(dev)
grid = (hw2[0] // 8, hw2[1] // 8)
f1 = F.interpolate(a["f_proj"], size=grid, mode="bilinear", align_corners=True)
c1 = F.interpolate(a["cam_rv_down"], size=grid, mode="bilinear", align_corners=True)
f2, c2 = b["f_proj"], b["cam_rv_down"]
x1 = f1.clone().requires_grad_(True); x2 = f2.clone().requires_grad_(True)
ref = R.contrast_terms(x1, c1, x2, c2, label.reshape(n, -1, 1, 1), 0.2, sampler=R.ReferenceSampler(7), grid=grid)
ref["loss_nce"].backward()
m = L.contrast_terms(f1, c1, f2, c2, label, 0.2, grid=grid, rng=7, want_mined=True)
P = n * grid[0] * grid[1]
for v, xr, gm in (("1", x1, m["grad_f1"]), ("2", x2, m["grad_f2"])):
    gr = xr.grad
    err = (gr - gm).permute(0, 2, 3, 1).reshape(P, d)
    ref_n = gr.permute(0, 2, 3, 1).reshape(P, d)
    e = err.norm(dim=1); r = ref_n.norm(dim=1)
    worst = torch.argsort(e, descending=True)[:6]
    print(f"view {v}: rel(max) {(err.abs().max() / ref_n.abs().max()).item():.3e}; pixels with err > 1e-3*max: {(e > 1e-3 * r.max()).sum().item()} of {P}")
    pl = m["pseudo_label" + v].long(); plo = m["pseudo_label" + ("2" if v == "1" else "1")].long()
    q = F.normalize((f1 if v == "1" else f2).permute(0, 2, 3, 1).reshape(P, d), dim=-1)
    proto = m["prototypes" + v]
    sc = q @ proto.t()
    for p in worst.tolist():
        order = torch.argsort(sc[p], descending=True, stable=True).tolist()
        srt = torch.sort(sc[p], descending=True)[0]
        cnt_ref = sum(int((rnd == p).sum()) + int((hard == p).sum()) for _, rnd, hard in ref["sets" + v][1])
        print(f"     gaps 2|3 {float(srt[2]-srt[3]):.2e} 12|13 {float(srt[12]-srt[13]):.2e}  ref mined count {cnt_ref}  ratio mine/ref {(gm.permute(0,2,3,1).reshape(P,d)[p].norm()/r[p]).item():.4f}")
        print(f"  px {p}: |err| {e[p].item():.3e} |ref| {r[p].item():.3e} pl {pl[p].item()} pl_other {plo[p].item()} mined {m['mined' + v][p].item()} "
              f"rank of pos {order.index(pl[p].item())} cls count {(pl == pl[p]).sum().item()}  xnorm {(f1 if v == '1' else f2).permute(0,2,3,1).reshape(P,d)[p].norm().item():.3e}")

print("---- mined-set analysis")
for v in "12":
    P_ = m["mined" + v].numel()
    rnd_c = torch.zeros(P_, dtype=torch.int32, device=dev); hard_c = torch.zeros(P_, dtype=torch.int32, device=dev)
    for _, rnd, hard in ref["sets" + v][1]:
        rnd_c.index_add_(0, rnd, torch.ones_like(rnd, dtype=torch.int32))
        hard_c.index_add_(0, hard, torch.ones_like(hard, dtype=torch.int32))
    mine_c = m["mined" + v]
    mism = (rnd_c + hard_c) != mine_c
    pl = ref["pseudo_label" + v]
    print(f"view {v}: mismatches {int(mism.sum())}; by class: {[(c, int((mism & (pl == c)).sum()), int((pl == c).sum())) for c in pl[mism].unique().tolist()]}")
    q = F.normalize((x1 if v == "1" else x2).detach().permute(0, 2, 3, 1).reshape(P_, d), dim=-1)
    sim = ((q * ref["prototypes" + v][pl]).sum(-1) + 1) / 2
    for c in pl[mism].unique().tolist()[:3]:
        mem = (pl == c).nonzero().reshape(-1)
        s_sorted, order = torch.sort(sim[mem], stable=True)
        ln = mem.numel(); kk = int(ln * 0.6); half = ln // 2
        bad = mism[mem][order].nonzero().reshape(-1).tolist()
        print(f"   class {c}: len {ln} kk {kk} half {half} window [{kk - half},{kk}); mismatching sorted positions {bad[:12]}; "
              f"exact-duplicate sims in class: {ln - torch.unique(sim[mem]).numel()}; sum mine {int(mine_c[mem].sum())} sum ref {int((rnd_c + hard_c)[mem].sum())}")
