"""The contrast training-step loss as a function (the reference has it inline, contrast_train.py:129-395).

``contrast_step_loss`` takes what ``Net.forward`` returns for both views plus the image-level labels and
returns the 8 scalars the reference logs (contrast_train.py:111-118).  The hot path — label-masked CAM
normalisation (K1b), seed labels (K1c), per-class top-k prototypes (K3) and the six NCE terms with hard-pixel
mining, forward and backward fused (K4) — runs in the wsegk CUDA kernels, and so do the two bilinear shrinks of view 1
(:179-180) and the auxiliary SEAM losses (:142-144,156,159-174): classification (GAP + multilabel soft
margin), min-pool, ER and ECR, the two top-k's done by value with a radix select.
"""
from __future__ import annotations

import itertools
import random as _pyrandom
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from . import ops

# step number of the device-generated mining keys when the caller does not pass one: advances with every call, so that the
# "random 50 %" of the hard-pixel mining is a fresh draw per training step, as the reference's random.sample calls are
# (contrast_train.py:316,370)
_AUTO_STEP = itertools.count()

LOSS_KEYS = ('loss', 'loss_cls', 'loss_er', 'loss_ecr', 'loss_nce', 'loss_intra_nce',
             'loss_cross_nce', 'loss_cross_nce2')          # contrast_train.py:111-118


def reference_rng_keys(pl1: torch.Tensor, pl2: torch.Tensor, rng: Union[int, _pyrandom.Random, None],
                       n_classes: int = 21) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mining keys that reproduce the reference's Python-RNG draws (contrast_train.py:291,316,344,370): per view,
    P dead ``sample(range(21),10)`` calls, then ``sample(range(len), len//2)`` for each class with >= 2 pixels in
    ascending class order; view 1 before view 2.  Chosen pixels get key 0, the others key 1, so the kernel's rule
    "the len//2 smallest (key, index)" selects exactly the reference's random half.  Needs the pseudo labels on
    the host (one sync) — the default device-hash keys (``ops.hash_keys``) do not."""
    if not hasattr(rng, "sample"):                 # an int seed; a random.Random or the `random` module is used as is
        rng = _pyrandom.Random(rng)
    keys = []
    for pl in (pl1, pl2):
        pl_h = pl.detach().to("cpu", torch.int64)
        for _ in range(pl_h.numel()):
            rng.sample(range(n_classes), 10)
        key = torch.ones(pl_h.numel(), dtype=torch.int32)
        for c in range(n_classes):
            members = torch.nonzero(pl_h == c).reshape(-1)
            ln = int(members.numel())
            if ln < 2:
                continue
            picks = rng.sample(range(ln), ln // 2)
            key[members[torch.as_tensor(picks, dtype=torch.long)]] = 0
        keys.append(key.to(pl.device))
    return keys[0], keys[1]


def contrast_terms(f_proj1: torch.Tensor, cam_rv1_down: torch.Tensor, f_proj2: torch.Tensor,
                   cam_rv2_down: torch.Tensor, label: torch.Tensor, bg_threshold: float,
                   grid: Optional[Tuple[int, int]] = None, rng: Union[int, _pyrandom.Random, None] = None,
                   mining_seed: int = 0, group=None, want_mined: bool = False,
                   keys: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                   mining: str = "shard", step: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """contrast_train.py:176-392.  ``label``: [N,21] (or [N,21,1,1]) with background = 1.
    ``rng``: None -> Philox mining keys generated on the device (no host work, no sync): key = Philox(global pixel index,
    ``step``, ``mining_seed``).  ``step`` is the training-step number — every rank of a data-parallel job must pass the same
    one; None = a process-local counter that advances with every call (identical on all ranks as long as they make the same
    calls), so consecutive steps mine different random halves.  An int / ``random.Random`` ``rng`` -> reproduce the
    reference's Python-RNG draws.  ``keys``: two int32 [P] device tensors to use as mining keys instead (nothing is created
    on the host, so the call can be captured in a CUDA graph; the caller refreshes them between replays, e.g. with
    ``ops.mining_keys(step=...)`` into the same tensors).  ``group``: process group over which prototypes are whole-batch (data parallel).
    ``mining``: "shard" = every rank mines hard pixels inside its own shard; "global" = over the whole batch of ``group`` as
    the single-process reference does (one more NVLink exchange; the reported intra losses are then R x this rank's share,
    so their mean over ranks — what DDP does to the gradients — is the single-process value)."""
    if grid is None:
        grid = tuple(f_proj2.shape[-2:])
    if tuple(f_proj1.shape[-2:]) != tuple(grid):
        f1 = ops.bilinear_resize(f_proj1, grid)                                                   # :179
        c1 = ops.resize_fwd(cam_rv1_down, grid)                                                   # :180 (no grad: seeds)
    else:
        f1, c1 = f_proj1, cam_rv1_down
    f2, c2 = f_proj2, cam_rv2_down
    norm1, pl1 = ops.seed_labels(c1, label, bg_threshold)
    norm2, pl2 = ops.seed_labels(c2, label, bg_threshold)
    proto1, ti1 = ops.prototypes(norm1, f1, group=group, view=0)
    proto2, ti2 = ops.prototypes(norm2, f2, group=group, view=1)
    if keys is not None:            # precomputed device-resident mining keys (e.g. when the step is captured in a CUDA graph)
        key1, key2 = keys
    elif rng is None:
        p = pl1.numel()
        if step is None:
            step = next(_AUTO_STEP)
        if pl1.is_cuda:         # generated on the device: no host RNG, no copy, no sync
            from . import dist as _dist
            index0 = _dist.rank_world(group)[0] * p
            key1 = ops.mining_keys(p, mining_seed, pl1.device, step=step, stream_id=1, index0=index0)
            key2 = ops.mining_keys(p, mining_seed, pl1.device, step=step, stream_id=2, index0=index0)
        else:
            key1 = ops.hash_keys(p, 2 * mining_seed + 1, pl1.device)
            key2 = ops.hash_keys(p, 2 * mining_seed + 2, pl1.device)
    else:
        key1, key2 = reference_rng_keys(pl1, pl2, rng, proto1.shape[0])
    out: Dict[str, torch.Tensor] = {}
    if mining not in ("shard", "global"):
        raise ValueError("mining must be 'shard' or 'global'")
    from . import dist as _dist
    gg = None
    if mining == "global" and _dist.rank_world(group)[1] > 1:
        import torch.distributed as _td
        gg = group if group is not None else _td.group.WORLD
    if want_mined:
        losses, g1, g2, m1, m2 = ops.contrast_loss_raw(f1, f2, proto1, proto2, pl1, pl2, key1, key2, want_mined=True,
                                                       global_group=gg)
        out.update({'grad_f1': g1, 'grad_f2': g2, 'mined1': m1, 'mined2': m2})
    else:
        loss_nce, losses = ops.contrast_loss(f1, f2, proto1, proto2, pl1, pl2, key1, key2, gg)
        out['loss_nce'] = loss_nce                    # the only differentiable entry (its parts are reported values)
    out.update({
        'loss_nce': out.get('loss_nce', losses[9]), 'loss_intra_nce': losses[8], 'loss_cross_nce': losses[6],
        'loss_cross_nce2': losses[7], 'losses_raw': losses,
        'pseudo_label1': pl1, 'pseudo_label2': pl2, 'norm_cam1': norm1, 'norm_cam2': norm2,
        'top_indices1': ti1, 'top_indices2': ti2, 'prototypes1': proto1, 'prototypes2': proto2,
    })
    return out


def contrast_step_loss(cam1, cam_rv1, f_proj1, cam_rv1_down, cam2, cam_rv2, f_proj2, cam_rv2_down,
                       label20: torch.Tensor, bg_threshold: float = 0.20,
                       view2_size: Tuple[int, int] = (128, 128),
                       rng: Union[int, _pyrandom.Random, None] = None, mining_seed: int = 0,
                       group=None, step: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Everything in contrast_train.py:138-395 given both views' network outputs.  ``label20``: [N,20] multi-hot.
    Differentiable w.r.t. the eight network outputs; returns the 8 logged scalars (plus diagnostics).  ``step``: the
    training-step number for the device-generated mining keys (see ``contrast_terms``; None = auto-advancing)."""
    n = label20.shape[0]
    label = torch.cat([torch.ones(n, 1, dtype=torch.float32, device=cam1.device),
                       label20.to(cam1.device, torch.float32)], dim=1)
    cls1 = ops.cls_loss(cam1, label)                                                      # :142,159  GAP + soft margin
    rvmin1 = ops.min_pool_loss(cam_rv1, label)                                            # :143      min-pool
    n1 = ops.fullres_norm(cam1, label, view2_size)                                        # :145-148  (K1b)
    rv1 = ops.fullres_norm(cam_rv1, label, view2_size)                                    # :149-152  (K1b)
    cls2 = ops.cls_loss(cam2, label)                                                      # :155,160
    rvmin2 = ops.min_pool_loss(cam_rv2, label)                                            # :156
    n2 = ops.fullres_norm(cam2, label, None)                                              # :157      (K1b)
    rv2 = ops.fullres_norm(cam_rv2, label, None)                                          # :158      (K1b)
    loss_er, loss_ecr = ops.er_ecr_loss(n1, n2, rv1, rv2)                                 # :163-172  ER + ECR
    loss_cls = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2                                  # :174
    grid = (view2_size[0] // 8, view2_size[1] // 8)                                       # :179
    terms = contrast_terms(f_proj1, cam_rv1_down, f_proj2, cam_rv2_down, label, bg_threshold, grid=grid,
                           rng=rng, mining_seed=mining_seed, group=group, step=step)
    out = dict(terms)
    out.update({'loss': loss_cls + loss_er + loss_ecr + terms['loss_nce'], 'loss_cls': loss_cls,
                'loss_er': loss_er, 'loss_ecr': loss_ecr})
    return out


def contrast_step_loss_lazy(cam1_small, cam_rv1_down, f_proj1, cam2_small, cam_rv2_down, f_proj2,
                            label20: torch.Tensor, image_size: Tuple[int, int], bg_threshold: float = 0.20,
                            view2_size: Tuple[int, int] = (128, 128),
                            rng: Union[int, _pyrandom.Random, None] = None, mining_seed: int = 0,
                            group=None, step: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """``contrast_step_loss`` for a network that does NOT enlarge its maps.

    ``Net.forward`` ends by enlarging the stride-8 ``cam`` and ``cam_rv_down`` x8 (network/resnet38_contrast.py:57-59: four
    tensors of N*21*H*W floats per step, 135 MB each at N=8, 448x448) and the loss reads them back only to pool, select and
    normalise them.  Here the loss takes the SMALL maps (``Net.forward_small``) and every consumer evaluates the enlargement
    on the fly with the same arithmetic: the same 8 scalars (the selections and normalised maps bit-identical, the global
    average pool to fp32 rounding), gradients on the small maps."""
    n = label20.shape[0]
    label = torch.cat([torch.ones(n, 1, dtype=torch.float32, device=cam1_small.device),
                       label20.to(cam1_small.device, torch.float32)], dim=1)
    cls1 = ops.cls_loss_up(cam1_small, label, image_size)                                 # :142,159
    rvmin1 = ops.min_pool_loss_up(cam_rv1_down, label, image_size)                        # :143
    n1 = ops.fullres_norm_up(cam1_small, label, image_size, view2_size)                   # :145-148
    rv1 = ops.fullres_norm_up(cam_rv1_down, label, image_size, view2_size)                # :149-152
    cls2 = ops.cls_loss_up(cam2_small, label, view2_size)                                 # :155,160
    rvmin2 = ops.min_pool_loss_up(cam_rv2_down, label, view2_size)                        # :156
    n2 = ops.fullres_norm_up(cam2_small, label, view2_size, None)                         # :157
    rv2 = ops.fullres_norm_up(cam_rv2_down, label, view2_size, None)                      # :158
    loss_er, loss_ecr = ops.er_ecr_loss(n1, n2, rv1, rv2)                                 # :163-172
    loss_cls = (cls1 + cls2) / 2 + (rvmin1 + rvmin2) / 2                                  # :174
    grid = (view2_size[0] // 8, view2_size[1] // 8)                                       # :179
    terms = contrast_terms(f_proj1, cam_rv1_down, f_proj2, cam_rv2_down, label, bg_threshold, grid=grid,
                           rng=rng, mining_seed=mining_seed, group=group, step=step)
    out = dict(terms)
    out.update({'loss': loss_cls + loss_er + loss_ecr + terms['loss_nce'], 'loss_cls': loss_cls,
                'loss_er': loss_er, 'loss_ecr': loss_ecr})
    return out
