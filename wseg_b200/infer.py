"""Multi-scale + flip CAM inference as ``contrast_infer.py:46-80`` does it, with the two variants of every scale (image and
its horizontal flip) in ONE forward of batch 2 instead of two forwards of batch 1: the forward-only PCM
kernel is persistent, so a batch of 2 costs little more than a batch of 1.  The SCALES have different grids (the reference
fans its eight variants over GPUs with a thread pool, ``contrast_infer.py:47,58-71``); here they run on one GPU as
concurrent streams, one per scale (``concurrent=True``): the small scales' forwards — a 0.5x image fills a fraction of
the SMs, and its PCM kernel is prologue-bound — overlap with the large ones instead of queueing before them — and the
PCM of ALL variants is ONE launch of the forward-only kernel (``one_pcm_launch=True``, ``ops.pcm_infer_multi``: every scale
runs on its share of the CTA pairs).  Everything after the network — resize to the original size (align_corners=False, as the script's ``F.upsample``), label mask, un-flip,
sum over variants, min/max normalisation — is the script's arithmetic, on the device.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F


_SCALE_STREAMS: Dict[Tuple[str, int], list] = {}


def _scale_streams(dev: torch.device, count: int):
    """A persistent pool per device: PyTorch's caching allocator and cuDNN keep per-stream state, and fresh streams on every
    call turn each forward into cudaMalloc's and plan look-ups (measured: 5x slower than the sequential loop)."""
    key = (str(dev), count)
    if key not in _SCALE_STREAMS:
        _SCALE_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(count)]
    return _SCALE_STREAMS[key]


@torch.no_grad()
def msf_cam(model, images: Sequence[torch.Tensor], label20: torch.Tensor, orig_size: Tuple[int, int],
            which: int = 1, concurrent: bool = True, one_pcm_launch: bool = True
            ) -> Tuple[torch.Tensor, Dict[int, torch.Tensor]]:
    """``images``: one CHW (or 1CHW) tensor per SCALE, un-flipped, already normalised (what ``voc12.data.
    VOC12ClsDatasetMSF`` yields at even positions).  ``which``: index of the map to use in the network's output tuple
    (1 = ``cam_rv`` of ``resnet38_contrast.Net`` / ``resnet38_SEAM.Net``, as ``_, cam, f, ff = model(img)`` at :61).
    ``concurrent``: one CUDA stream per scale (largest first), joined before the sum; the per-scale results are added in
    the order of ``images`` either way, so the result does not depend on it.
    ``one_pcm_launch`` (for ``which == 1`` and a network with ``pcm_inputs``): the scales' forwards stop in front of the
    PCM, and ONE launch of the forward-only tcgen05 kernel runs the PCM of all variants, every scale on its share of the
    CTA pairs (``ops.pcm_infer_multi``; 8 variants of a 320x480 image: 115 us against 188 us of per-scale launches in
    sequence and 128 us on four streams — the 0.5x scale alone cannot fill the GPU and every launch pays its own
    prologue and tail).
    Returns ``(norm_cam [20,H,W], cam_dict {class: map})`` as :73-85 builds them."""
    from . import ops
    dev = next(model.parameters()).device
    label = label20.to(dev, torch.float32).reshape(20, 1, 1)
    split = one_pcm_launch and which == 1 and hasattr(model, "pcm_inputs") and dev.type == 'cuda' and len(images) <= 8

    def pair_of(img):
        x = img.to(dev, torch.float32)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        return torch.cat([x, x.flip(-1)], dim=0)                                           # variants 2s, 2s+1 of the script

    def finish(cam):
        cam = F.interpolate(cam[:, 1:], orig_size, mode='bilinear', align_corners=False) * label
        return cam[0] + cam[1].flip(-1)                                                    # :64-65 un-flip, then sum

    def one_scale(img):
        pair = pair_of(img)
        if split:                                       # up to the PCM only: (cam logits, f9(f), full-resolution size)
            return model.pcm_inputs(pair) + (tuple(pair.shape[-2:]),)
        return finish(model(pair)[which])

    per_scale = [None] * len(images)
    if concurrent and len(images) > 1 and dev.type == 'cuda':
        main = torch.cuda.current_stream(dev)
        order = sorted(range(len(images)), key=lambda i: -images[i].shape[-1] * images[i].shape[-2])
        streams = _scale_streams(dev, len(order))
        for st, i in zip(streams, order):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                per_scale[i] = one_scale(images[i])
                for t in (per_scale[i] if split else (per_scale[i],)):
                    if torch.is_tensor(t):
                        t.record_stream(main)
        for st in streams:
            main.wait_stream(st)
    else:
        for i, img in enumerate(images):
            per_scale[i] = one_scale(img)
    if split:
        cams = [ops.cam_norm(p[0]) for p in per_scale]                                     # K1a
        rvs = ops.pcm_infer_multi(cams, [p[1] for p in per_scale])                         # K2, all scales in one launch
        per_scale = [finish(ops.resize_fwd(rv, p[2])) for rv, p in zip(rvs, per_scale)]
    total = per_scale[0]
    for both in per_scale[1:]:
        total = total + both
    total = total.clamp_min(0)                                                             # :74
    cam_max = total.amax(dim=(1, 2), keepdim=True)
    cam_min = total.amin(dim=(1, 2), keepdim=True)
    total = torch.where(total < cam_min + 1e-5, torch.zeros_like(total), total)            # :77
    norm_cam = (total - cam_min - 1e-5) / (cam_max - cam_min + 1e-5)                       # :78
    cam_dict = {i: norm_cam[i] for i in range(20) if float(label20.reshape(-1)[i]) > 1e-5}
    return norm_cam, cam_dict
