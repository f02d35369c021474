"""Multi-scale + flip CAM inference as ``contrast_infer.py:46-80`` does it, with the two variants of every scale (image and
its horizontal flip) in ONE forward of batch 2 instead of two forwards of batch 1: the forward-only PCM
kernel is persistent, so a batch of 2 costs little more than a batch of 1.  Everything after the network — resize to the
original size (align_corners=False, as the script's ``F.upsample``), label mask, un-flip, sum over variants, min/max
normalisation — is the script's arithmetic, on the device.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F


@torch.no_grad()
def msf_cam(model, images: Sequence[torch.Tensor], label20: torch.Tensor, orig_size: Tuple[int, int],
            which: int = 1) -> Tuple[torch.Tensor, Dict[int, torch.Tensor]]:
    """``images``: one CHW (or 1CHW) tensor per SCALE, un-flipped, already normalised (what ``voc12.data.
    VOC12ClsDatasetMSF`` yields at even positions).  ``which``: index of the map to use in the network's output tuple
    (1 = ``cam_rv`` of ``resnet38_contrast.Net`` / ``resnet38_SEAM.Net``, as ``_, cam, f, ff = model(img)`` at :61).
    Returns ``(norm_cam [20,H,W], cam_dict {class: map})`` as :73-85 builds them."""
    dev = next(model.parameters()).device
    label = label20.to(dev, torch.float32).reshape(20, 1, 1)
    total = None
    for img in images:
        x = img.to(dev, torch.float32)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        pair = torch.cat([x, x.flip(-1)], dim=0)                                           # variants 2s, 2s+1 of the script
        cam = model(pair)[which]
        cam = F.interpolate(cam[:, 1:], orig_size, mode='bilinear', align_corners=False) * label
        both = cam[0] + cam[1].flip(-1)                                                    # :64-65 un-flip, then sum
        total = both if total is None else total + both
    total = total.clamp_min(0)                                                             # :74
    cam_max = total.amax(dim=(1, 2), keepdim=True)
    cam_min = total.amin(dim=(1, 2), keepdim=True)
    total = torch.where(total < cam_min + 1e-5, torch.zeros_like(total), total)            # :77
    norm_cam = (total - cam_min - 1e-5) / (cam_max - cam_min + 1e-5)                       # :78
    cam_dict = {i: norm_cam[i] for i in range(20) if float(label20.reshape(-1)[i]) > 1e-5}
    return norm_cam, cam_dict
