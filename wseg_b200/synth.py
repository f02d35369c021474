"""Seeded synthetic inputs for the contrast hot path (SURVEY.md §8d "Synthetic inputs").

All tensors are made on the CPU with an explicit ``torch.Generator`` so the same seed gives the same
bytes in the build container and on the GPU box; callers move them to the device.  There is no VOC data
and no checkpoint in either place, so every test, the golden vectors and ``bench.py`` draw from here.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

N_CLASSES = 21          # background + 20 VOC classes (network/resnet38_contrast.py:15)
PCM_CHANNELS = 192      # output channels of f9 (network/resnet38_contrast.py:20)
# number of foreground classes per image, from voc12/cls_labels.npy (12 031 images)
_LABEL_COUNT_P = torch.tensor([0.60, 0.29, 0.09, 0.02])


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def labels(n: int, seed: int) -> torch.Tensor:
    """[N,20] float multi-hot image-level labels, 1-4 classes per image with VOC-like frequencies."""
    g = gen(seed)
    counts = torch.multinomial(_LABEL_COUNT_P, n, replacement=True, generator=g) + 1
    out = torch.zeros(n, N_CLASSES - 1)
    for i in range(n):
        picks = torch.randperm(N_CLASSES - 1, generator=g)[: int(counts[i])]
        out[i, picks] = 1.0
    return out


def label_with_bg(label20: torch.Tensor) -> torch.Tensor:
    """[N,20] -> [N,21] with the background entry set to 1 (contrast_train.py:138-139)."""
    n = label20.shape[0]
    return torch.cat([torch.ones(n, 1, dtype=label20.dtype, device=label20.device), label20], dim=1)


def pcm_inputs(n: int, h: int, w: int, seed: int, k: int = PCM_CHANNELS, c: int = N_CLASSES,
               positive: bool = False) -> Dict[str, torch.Tensor]:
    """cam logits [n,c,h,w], PCM features f [n,k,h,w] (output of f9) and an upstream gradient."""
    g = gen(seed)
    cam = 0.05 * torch.randn(n, c, h, w, generator=g)
    f = torch.randn(n, k, h, w, generator=g)
    if positive:
        f = f.abs()
    grad = 1e-3 * torch.randn(n, c, h, w, generator=g)
    return {"cam": cam, "f": f, "grad_out": grad}


def smooth_field(n: int, c: int, h: int, w: int, g: torch.Generator, coarse: int = 4) -> torch.Tensor:
    """Spatially correlated random maps (low-resolution noise, bilinearly enlarged)."""
    hc, wc = max(2, -(-h // coarse)), max(2, -(-w // coarse))
    z = torch.randn(n, c, hc, wc, generator=g)
    return F.interpolate(z, size=(h, w), mode="bilinear", align_corners=True)


def head_outputs(n: int, hw: Tuple[int, int], seed: int, d: int = 128, c: int = N_CLASSES,
                 smooth: bool = True) -> Dict[str, torch.Tensor]:
    """What ``Net.forward`` returns for one view of size ``hw``: cam, cam_rv (full resolution) plus
    f_proj and cam_rv_down on the stride-8 grid.  Statistics per SURVEY.md §8d."""
    g = gen(seed)
    H, W = hw
    h, w = -(-H // 8), -(-W // 8)
    mk = (lambda *s: smooth_field(*s, g)) if smooth else (lambda *s: torch.randn(*s, generator=g))
    cam_small = 0.05 * mk(n, c, h, w) + 0.02 * torch.randn(n, c, h, w, generator=g)
    rv_small = (0.5 * mk(n, c, h, w) + 0.1 * torch.randn(n, c, h, w, generator=g)).abs()
    f_proj = F.relu(0.02 * torch.randn(n, d, h, w, generator=g) + 0.01 * mk(n, d, h, w) + 0.01)
    cam = F.interpolate(cam_small, size=(H, W), mode="bilinear", align_corners=True)
    cam_rv = F.interpolate(rv_small, size=(H, W), mode="bilinear", align_corners=True)
    return {"cam": cam, "cam_rv": cam_rv, "f_proj": f_proj, "cam_rv_down": rv_small}
