"""ORACLE — test infrastructure, NOT product code.

CPU (or same-device ATen) restatement, as plain functions, of the Step-1 contrast hot path of
obeychoi0120/wseg.  Every function cites the reference lines it follows (paths relative to the
reference checkout).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product (``wseg_b200``) never does.

Parity status: the reference has no tests or golden vectors for this path, so the
pin is the reference itself: ``oracle/verbatim.py`` executes the reference's own lines on CPU in the
build container and ``oracle/make_golden.py`` stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those stored vectors (and
``tests/test_oracle_verbatim.py`` against a live verbatim run when ``/root/reference`` is mounted).

Tie-breaking.  ``torch.topk`` / ``argmax`` do not specify which of several equal values wins, and the
reference's class-0 CAM plane is the constant ``bg_threshold`` (100 % ties).  The canonical order used
by the oracle, the golden vectors and the CUDA kernels is: value first, then LOWEST flat index.
``topk_fn`` can be overridden (e.g. with ``torch.topk``) to reproduce a stock-ATen run.
"""
from __future__ import annotations

import random as _pyrandom
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

TEMPERATURE = 0.1          # contrast_train.py:261 ff. ("/ 0.1")
NCE_WEIGHT = 0.1           # contrast_train.py:269,281,389
SEMI_HARD_TOPK = 13        # contrast_train.py:293
SEMI_HARD_SKIP = 3         # contrast_train.py:294
HARD_FRACTION = 0.6        # contrast_train.py:321


# ----------------------------------------------------------------------------------------------
# canonical selection primitives
# ----------------------------------------------------------------------------------------------
def canonical_topk(x: torch.Tensor, k: int, dim: int = -1, largest: bool = True,
                   sorted: bool = True):  # noqa: A002 - mirrors torch.topk's signature
    """torch.topk with a defined tie order: value, then lowest index (stable sort)."""
    vals, idx = torch.sort(x, dim=dim, descending=largest, stable=True)
    sl = [slice(None)] * x.dim()
    sl[dim] = slice(0, k)
    return vals[tuple(sl)], idx[tuple(sl)]


TopkFn = Callable[..., Tuple[torch.Tensor, torch.Tensor]]


# ----------------------------------------------------------------------------------------------
# K1a — in-network CAM normalisation            network/resnet38_contrast.py:41-48
# ----------------------------------------------------------------------------------------------
def cam_d_norm(cam: torch.Tensor) -> torch.Tensor:
    """relu, per-(n,c) max-normalise, background = 1 - max foreground, keep only arg-max class."""
    n, c, h, w = cam.shape
    x = F.relu(cam.detach())
    peak = x.reshape(n, c, -1).amax(dim=-1).reshape(n, c, 1, 1) + 1e-5
    y = F.relu(x - 1e-5) / peak
    fg_max = y[:, 1:].amax(dim=1, keepdim=True)
    out = torch.empty_like(y)
    out[:, :1] = 1.0 - fg_max
    fg = y[:, 1:]
    out[:, 1:] = torch.where(fg < fg_max, torch.zeros_like(fg), fg)
    return out


# ----------------------------------------------------------------------------------------------
# K2 — pixel correlation module (after the f9 conv)   network/resnet38_contrast.py:63-75
# ----------------------------------------------------------------------------------------------
def pcm(cam_norm: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """cam_norm [n,C,h,w] (no grad), f [n,K,h,w] (output of f9).  Returns cam_rv [n,C,h,w]."""
    n, k, h, w = f.shape
    cam_flat = cam_norm.reshape(n, -1, h * w)          # F.interpolate to (h,w) is the identity (:66)
    ff = f.reshape(n, k, h * w)
    ff = ff / (torch.norm(ff, dim=1, keepdim=True) + 1e-5)                       # :70
    aff = F.relu(torch.matmul(ff.transpose(1, 2), ff))                           # :71
    aff = aff / (aff.sum(dim=1, keepdim=True) + 1e-5)                            # :72
    return torch.matmul(cam_flat, aff).reshape(n, -1, h, w)                      # :73


# ----------------------------------------------------------------------------------------------
# K1b — full-resolution min/max normalisation, resize, label mask, background channel
#        tool/visualization.py:62-67, contrast_train.py:145-152,157-158,165-166
# ----------------------------------------------------------------------------------------------
def max_norm(p: torch.Tensor, e: float = 1e-5) -> torch.Tensor:
    n, c, h, w = p.shape
    p = F.relu(p)
    flat = p.reshape(n, c, -1)
    hi = flat.max(dim=-1)[0].reshape(n, c, 1, 1)
    lo = flat.min(dim=-1)[0].reshape(n, c, 1, 1)
    return F.relu(p - lo - e) / (hi - lo + e)


def fullres_norm(p: torch.Tensor, label: torch.Tensor,
                 size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """max_norm -> (optional) bilinear align_corners resize -> x label.   label: [N,21,1,1]."""
    y = max_norm(p)
    if size is not None and tuple(size) != tuple(p.shape[-2:]):
        y = F.interpolate(y, size=tuple(size), mode='bilinear', align_corners=True)
    elif size is not None:
        # the reference always calls F.interpolate for view 1 (contrast_train.py:145); same size is
        # an exact identity for align_corners bilinear.
        y = F.interpolate(y, size=tuple(size), mode='bilinear', align_corners=True)
    return y * label


def with_bg_channel(cam: torch.Tensor) -> torch.Tensor:
    """contrast_train.py:165-166 — channel 0 := 1 - max over foreground (out-of-place)."""
    bg = 1.0 - cam[:, 1:].max(dim=1, keepdim=True)[0]
    return torch.cat([bg, cam[:, 1:]], dim=1)


# ----------------------------------------------------------------------------------------------
# K1c — seed labels                                         contrast_train.py:184-197 / 212-226
# ----------------------------------------------------------------------------------------------
def seed_labels(cam_rv_down: torch.Tensor, label: torch.Tensor, bg_threshold: float,
                argmax_of: str = "softmax") -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (norm_cam [N,21,g,g] with channel 0 = bg_threshold, pseudo_label [N*g*g] int64).

    argmax_of="softmax" follows the reference literally (argmax of softmax scores);
    "logits" takes the argmax of ``norm_cam * label`` directly (what the kernel does; identical
    whenever the top-2 gap exceeds fp32 exp rounding, which the tests assert)."""
    x = F.relu(cam_rv_down.detach())
    n, c, h, w = x.shape
    flat = x.reshape(n, c, -1)
    hi = flat.max(dim=-1)[0].reshape(n, c, 1, 1)
    lo = flat.min(dim=-1)[0].reshape(n, c, 1, 1)
    x = torch.where(x < lo + 1e-5, torch.zeros_like(x), x)
    x = (x - lo - 1e-5) / (hi - lo + 1e-5)
    x = x.clone()
    x[:, 0] = bg_threshold
    logits = x * label
    if argmax_of == "softmax":
        pseudo = F.softmax(logits, dim=1).argmax(dim=1)
    else:
        pseudo = logits.argmax(dim=1)
    return x, pseudo.reshape(-1)


# ----------------------------------------------------------------------------------------------
# K3 — per-class top-k + CAM-weighted prototypes            contrast_train.py:199-209 / 229-241
# ----------------------------------------------------------------------------------------------
def prototypes(norm_cam: torch.Tensor, f_proj: torch.Tensor, k: Optional[int] = None,
               topk_fn: TopkFn = canonical_topk
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """norm_cam [N,C,g,g] (not label-masked), f_proj [N,D,g,g] un-normalised.
    Returns (top_values [C,k], top_indices [C,k] into N*g*g, prototypes [C,D] L2-normalised)."""
    n, c, h, w = norm_cam.shape
    d = f_proj.shape[1]
    if k is None:
        k = h * w // 8
    fea = f_proj.detach().permute(0, 2, 3, 1).reshape(-1, d)
    plane = norm_cam.transpose(0, 1).reshape(c, -1)
    tv, ti = topk_fn(plane, k=k, dim=-1)
    proto = torch.zeros(c, d, dtype=fea.dtype, device=fea.device)
    for ci in range(c):
        picked = fea[ti[ci]]
        proto[ci] = (tv[ci].unsqueeze(-1) * picked).sum(dim=0) / tv[ci].sum()
    return tv, ti, F.normalize(proto, dim=-1)


# ----------------------------------------------------------------------------------------------
# K4a — cross-view NCE                                       contrast_train.py:244-281
# ----------------------------------------------------------------------------------------------
def _flat_normalised(f_proj: torch.Tensor) -> torch.Tensor:
    n, d, h, w = f_proj.shape
    return F.normalize(f_proj.permute(0, 2, 3, 1).reshape(n * h * w, d), dim=-1)


def nce_full(q: torch.Tensor, proto: torch.Tensor, pl: torch.Tensor) -> torch.Tensor:
    """mean_p -log( exp(q.proto[pl]/T) / sum_c exp(q.proto_c/T) )   (no max subtraction, as :261-263)"""
    pos = torch.exp((q * proto[pl]).sum(dim=-1) / TEMPERATURE)
    den = torch.exp(torch.matmul(q, proto.t()) / TEMPERATURE).sum(dim=-1)
    return torch.mean(-torch.log(pos / den))


# ----------------------------------------------------------------------------------------------
# K4b — intra-view NCE with semi-hard prototypes and hard-pixel mining   contrast_train.py:283-389
# ----------------------------------------------------------------------------------------------
def semi_hard_denominator(q: torch.Tensor, proto: torch.Tensor, pl: torch.Tensor,
                          topk_fn: TopkFn = canonical_topk) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (A1 = exp(q.proto[pl]/T), A2 = sum over {positive} U {ranks 3..12 of q.proto^T})."""
    pos_dot = (q * proto[pl]).sum(dim=-1)
    a1 = torch.exp(pos_dot / TEMPERATURE)
    scores = torch.matmul(q, proto.t())
    with torch.no_grad():
        _, order = topk_fn(scores, k=SEMI_HARD_TOPK, dim=-1, largest=True)
        semi = order[:, SEMI_HARD_SKIP:]
    picked = torch.gather(scores, 1, semi)
    a2 = torch.exp(torch.cat([pos_dot.unsqueeze(1), picked], dim=1) / TEMPERATURE).sum(dim=-1)
    return a1, a2


def hard_pixel_sets(sim: torch.Tensor, pl: torch.Tensor, n_classes: int,
                    sample_fn: Callable[[int, int], Sequence[int]],
                    topk_fn: TopkFn = canonical_topk
                    ) -> Tuple[int, List[Tuple[int, torch.Tensor, torch.Tensor]]]:
    """For each class present in ``pl``: (class, random-half positions, hard-half positions), the
    positions indexing the class's pixels in ascending flat order (contrast_train.py:304-322).
    Returns (C_present, list) — classes with fewer than 2 pixels count in C_present but yield no set."""
    present = torch.unique(pl).tolist()
    out = []
    c_present = 0
    for ci in range(n_classes):
        if ci not in present:
            continue
        c_present += 1
        members = torch.nonzero(pl == ci).reshape(-1)
        ln = int(members.numel())
        if ln < 2:
            continue
        rnd = torch.as_tensor(list(sample_fn(ln, ln // 2)), dtype=torch.long, device=pl.device)
        kk = int(ln * HARD_FRACTION)
        _, low = topk_fn(sim[members], k=kk, largest=False)
        low = low[kk - ln // 2:]
        out.append((ci, members[rnd], members[low]))
    return c_present, out


def intra_nce(q: torch.Tensor, proto: torch.Tensor, pl: torch.Tensor,
              sample_fn: Callable[[int, int], Sequence[int]],
              topk_fn: TopkFn = canonical_topk,
              return_sets: bool = False):
    a1, a2 = semi_hard_denominator(q, proto, pl, topk_fn)
    sim = ((q * proto[pl]).sum(dim=-1) + 1.0) / 2.0
    c_present, sets = hard_pixel_sets(sim.detach(), pl, proto.shape[0], sample_fn, topk_fn)
    total = torch.zeros(1, dtype=q.dtype, device=q.device)
    for _, rnd_px, hard_px in sets:
        px = torch.cat([rnd_px, hard_px], dim=0)
        total = total + torch.mean(-torch.log(a1[px] / a2[px]))
    loss = total / c_present
    if return_sets:
        return loss, (c_present, sets)
    return loss


# ----------------------------------------------------------------------------------------------
# auxiliary SEAM losses ("next" row) — restated so the whole step can be compared
#                                                     contrast_train.py:16-32,142-144,156,159-174
# ----------------------------------------------------------------------------------------------
def min_pool_loss(x: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x.shape
    k = h * w // 4
    m = x.max(dim=1)[0].reshape(n, -1)
    low = torch.topk(m, k=k, dim=-1, largest=False)[0]
    return F.relu(low).sum() / (k * n)


def keep_argmax_fg(x: torch.Tensor) -> torch.Tensor:
    """max_onehot (contrast_train.py:28-32), out-of-place."""
    fg = x[:, 1:]
    peak = fg.max(dim=1, keepdim=True)[0]
    return torch.cat([x[:, :1], torch.where(fg != peak, torch.zeros_like(fg), fg)], dim=1)


# ----------------------------------------------------------------------------------------------
# the whole loss body                                        contrast_train.py:129-395
# ----------------------------------------------------------------------------------------------
class ReferenceSampler:
    """Reproduces the reference's use of Python's global ``random`` (contrast_train.py:291,316,344,370):
    P dead draws of ``sample(range(21), 10)`` per view, then one ``sample(range(len), len//2)`` per
    mined class, view 1 before view 2."""

    def __init__(self, seed_or_rng=None):
        if isinstance(seed_or_rng, _pyrandom.Random):
            self.rng = seed_or_rng
        else:
            self.rng = _pyrandom.Random(seed_or_rng)

    def burn_dead_draws(self, n_pixels: int, n_classes: int = 21) -> None:
        for _ in range(n_pixels):
            self.rng.sample(range(n_classes), 10)

    def __call__(self, population: int, k: int) -> Sequence[int]:
        return self.rng.sample(range(population), k)


def contrast_terms(f_proj1: torch.Tensor, cam_rv1_down: torch.Tensor,
                   f_proj2: torch.Tensor, cam_rv2_down: torch.Tensor,
                   label: torch.Tensor, bg_threshold: float,
                   sampler: Optional[ReferenceSampler] = None,
                   topk_fn: TopkFn = canonical_topk, argmax_of: str = "softmax",
                   grid: Optional[Tuple[int, int]] = None) -> Dict[str, torch.Tensor]:
    """contrast_train.py:176-392.  ``label`` is [N,21,1,1] with background = 1.
    ``grid`` is the common loss grid (the reference hard-codes view 2's 128//8 = 16)."""
    if sampler is None:
        sampler = ReferenceSampler(0)
    if grid is None:
        grid = tuple(f_proj2.shape[-2:])
    f1 = F.interpolate(f_proj1, size=grid, mode='bilinear', align_corners=True)          # :179
    c1 = F.interpolate(cam_rv1_down, size=grid, mode='bilinear', align_corners=True)     # :180
    f2, c2 = f_proj2, cam_rv2_down
    with torch.no_grad():
        norm1, pl1 = seed_labels(c1, label, bg_threshold, argmax_of)
        tv1, ti1, proto1 = prototypes(norm1, f1, topk_fn=topk_fn)
        norm2, pl2 = seed_labels(c2, label, bg_threshold, argmax_of)
        tv2, ti2, proto2 = prototypes(norm2, f2, topk_fn=topk_fn)
    q1 = _flat_normalised(f1)
    q2 = _flat_normalised(f2)
    cross = NCE_WEIGHT * (nce_full(q1, proto2, pl1) + nce_full(q2, proto1, pl2)) / 2           # :261-269
    cross2 = NCE_WEIGHT * (nce_full(q1, proto1, pl2) + nce_full(q2, proto2, pl1)) / 2          # :272-281
    sampler.burn_dead_draws(q1.shape[0], proto1.shape[0])                                        # :291
    intra1, sets1 = intra_nce(q1, proto1, pl1, sampler, topk_fn, return_sets=True)               # :285-334
    sampler.burn_dead_draws(q2.shape[0], proto2.shape[0])                                        # :344
    intra2, sets2 = intra_nce(q2, proto2, pl2, sampler, topk_fn, return_sets=True)               # :338-387
    intra = NCE_WEIGHT * (intra1 + intra2) / 2                                                   # :389
    return {
        'loss_nce': cross + cross2 + intra, 'loss_intra_nce': intra.reshape(()),
        'loss_cross_nce': cross, 'loss_cross_nce2': cross2,
        'pseudo_label1': pl1, 'pseudo_label2': pl2, 'norm_cam1': norm1, 'norm_cam2': norm2,
        'top_indices1': ti1, 'top_indices2': ti2, 'top_values1': tv1, 'top_values2': tv2,
        'prototypes1': proto1, 'prototypes2': proto2, 'sets1': sets1, 'sets2': sets2,
    }


def contrast_step_loss(cam1, cam_rv1, f_proj1, cam_rv1_down, cam2, cam_rv2, f_proj2, cam_rv2_down,
                       label20: torch.Tensor, bg_threshold: float = 0.20,
                       view2_size: Tuple[int, int] = (128, 128),
                       sampler: Optional[ReferenceSampler] = None,
                       topk_fn: TopkFn = canonical_topk, argmax_of: str = "softmax"
                       ) -> Dict[str, torch.Tensor]:
    """All of contrast_train.py:138-395 given both views' network outputs.  label20: [N,20] multi-hot."""
    n = label20.shape[0]
    label = torch.cat([torch.ones(n, 1, dtype=label20.dtype, device=label20.device), label20], dim=1)
    label = label.reshape(n, -1, 1, 1)
    gap1 = F.adaptive_avg_pool2d(cam1, (1, 1))                                           # :142
    rvmin1 = min_pool_loss((cam_rv1 * label)[:, 1:])                                     # :143
    n1 = fullres_norm(cam1, label, view2_size)                                           # :145-148
    rv1 = fullres_norm(cam_rv1, label, view2_size)                                       # :149-152
    gap2 = F.adaptive_avg_pool2d(cam2, (1, 1))                                           # :155
    rvmin2 = min_pool_loss((cam_rv2 * label)[:, 1:])                                     # :156
    n2 = fullres_norm(cam2, label, None)                                                 # :157
    rv2 = fullres_norm(cam_rv2, label, None)                                             # :158
    cls1 = F.multilabel_soft_margin_loss(gap1[:, 1:], label[:, 1:])                      # :159
    cls2 = F.multilabel_soft_margin_loss(gap2[:, 1:], label[:, 1:])                      # :160
    ns, cs, hs, ws = n2.shape
    loss_er = torch.mean(torch.abs(n1[:, 1:] - n2[:, 1:]))                               # :163
    n1b = with_bg_channel(n1)                                                            # :165
    n2b = with_bg_channel(n2)                                                            # :166
    k_ecr = int(21 * hs * ws * 0.2)
    e1 = torch.abs(keep_argmax_fg(n2b.detach()) - rv1)                                   # :168
    e2 = torch.abs(keep_argmax_fg(n1b.detach()) - rv2)                                   # :169
    loss_ecr = (torch.mean(torch.topk(e1.reshape(ns, -1), k=k_ecr, dim=-1)[0])
                + torch.mean(torch.topk(e2.reshape(ns, -1), k=k_ecr, dim=-1)[0]))        # :170-172
    loss_cls = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2                                 # :174
    grid = (view2_size[0] // 8, view2_size[1] // 8)                                      # :179 (128 // 8)
    terms = contrast_terms(f_proj1, cam_rv1_down, f_proj2, cam_rv2_down, label, bg_threshold,
                           sampler, topk_fn, argmax_of, grid=grid)
    loss = loss_cls + loss_er + loss_ecr + terms['loss_nce']                             # :395
    out = dict(terms)
    out.update({'loss': loss, 'loss_cls': loss_cls, 'loss_er': loss_er, 'loss_ecr': loss_ecr})
    return out


LOSS_KEYS = ('loss', 'loss_cls', 'loss_er', 'loss_ecr', 'loss_nce', 'loss_intra_nce',
             'loss_cross_nce', 'loss_cross_nce2')          # contrast_train.py:111-118
