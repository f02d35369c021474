"""Python operators over the wsegk C ABI (include/wsegk.h).

Every function takes CUDA fp32 tensors, allocates outputs/workspace with torch (the C side never allocates)
and enqueues the kernels on torch's current stream for the tensor's device.  No function here has a PyTorch
or CPU fallback: a missing library or a non-sm_100 device raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"wseg_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"wseg_b200: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


def _workspace(op: int, device, **dims) -> torch.Tensor:
    nbytes = _lib.workspace_bytes(op, **dims)
    # uint8 CUDA allocations from torch's caching allocator are 512-byte aligned; over-allocate to reach 1 KiB
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    shift = (-raw.data_ptr()) % 1024
    return raw[shift:shift + nbytes]


# --------------------------------------------------------------------------------------------- K1a
def cam_norm(cam: torch.Tensor) -> torch.Tensor:
    """network/resnet38_contrast.py:41-48 — no-grad CAM normalisation.  cam [N,C,h,w] -> same shape."""
    cam = _require_cuda_f32("cam", cam.detach())
    n, c, h, w = cam.shape
    out = torch.empty_like(cam)
    with torch.cuda.device(cam.device):
        ws = _workspace(_lib.OP_CAM_NORM, cam.device, N=n, C=c, h=h, w=w)
        check(_lib.lib().wsegk_cam_norm_fwd(_ptr(cam), _ptr(out), n, c, h, w, _ptr(ws), ws.numel(),
                                            _stream(cam)), "wsegk_cam_norm_fwd")
    return out


# --------------------------------------------------------------------------------------------- K2
def pcm_forward(cam_norm_t: torch.Tensor, f: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """network/resnet38_contrast.py:70-73.  Returns (cam_rv [N,C,h,w], workspace kept for the backward)."""
    f = _require_cuda_f32("f", f)
    cam_norm_t = _require_cuda_f32("cam_norm", cam_norm_t)
    n, k, h, w = f.shape
    c = cam_norm_t.shape[1]
    if cam_norm_t.shape[0] != n or tuple(cam_norm_t.shape[2:]) != (h, w):
        raise RuntimeError(f"pcm: cam {tuple(cam_norm_t.shape)} does not match f {tuple(f.shape)}")
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        ws = _workspace(_lib.OP_PCM, f.device, N=n, C=c, K=k, h=h, w=w)
        check(_lib.lib().wsegk_pcm_fwd(_ptr(f), _ptr(cam_norm_t), _ptr(out), n, c, k, h, w, _ptr(ws),
                                       ws.numel(), _stream(f)), "wsegk_pcm_fwd")
    return out, ws


def pcm_backward(f: torch.Tensor, cam_norm_t: torch.Tensor, cam_rv: torch.Tensor, grad_out: torch.Tensor,
                 ws: torch.Tensor) -> torch.Tensor:
    f = _require_cuda_f32("f", f)
    grad_out = _require_cuda_f32("grad_cam_rv", grad_out)
    n, k, h, w = f.shape
    c = cam_norm_t.shape[1]
    grad_f = torch.empty_like(f)
    with torch.cuda.device(f.device):
        check(_lib.lib().wsegk_pcm_bwd(_ptr(f), _ptr(cam_norm_t), _ptr(cam_rv), _ptr(grad_out), _ptr(grad_f),
                                       n, c, k, h, w, _ptr(ws), ws.numel(), _stream(f)), "wsegk_pcm_bwd")
    return grad_f


class _PCMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam_norm_t, f):
        out, ws = pcm_forward(cam_norm_t, f)
        ctx.save_for_backward(f, cam_norm_t, out)
        ctx.ws = ws
        return out

    @staticmethod
    def backward(ctx, grad_out):
        f, cam_norm_t, out = ctx.saved_tensors
        grad_f = pcm_backward(f, cam_norm_t, out, grad_out, ctx.ws)
        ctx.ws = None
        return None, grad_f


def pcm(cam_norm_t: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """Differentiable (w.r.t. ``f``) pixel correlation module; ``f`` is the output of the f9 conv."""
    return _PCMFunction.apply(cam_norm_t.detach(), f)


def pcm_layout(n: int, c: int, k: int, h: int, w: int) -> dict:
    """Test aid: byte offsets of the PCM workspace sections."""
    import ctypes
    buf = (ctypes.c_size_t * 12)()
    got = _lib.lib().wsegk_debug_pcm_layout(n, c, k, h, w, buf, 12)
    names = ["fhat", "camb", "inv_norm", "dsum", "part", "uaug", "vaug", "dfhat", "bytes", "nsplit", "T", "HWp"]
    return {names[i]: int(buf[i]) for i in range(got)}
