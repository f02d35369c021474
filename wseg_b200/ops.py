"""Python operators over the wsegk C ABI (include/wsegk.h).

Every function takes CUDA fp32 tensors, allocates outputs/workspace with torch (the C side never allocates)
and enqueues the kernels on torch's current stream for the tensor's device.  No function here has a PyTorch
or CPU fallback: a missing library or a non-sm_100 device raises.
"""
from __future__ import annotations

import os
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
def _pcm_flags(mode: Optional[str], defer_fixup: bool = False) -> int:
    """Flags of the per-call PCM entry points (include/wsegk.h).  ``mode=None``: the calling thread's default
    (``set_pcm_mode``); the mode travels with the call, so threads cannot change each other's."""
    if mode is None:
        mode = get_pcm_mode()
    if mode not in ("precise", "fast"):
        raise ValueError("mode must be 'precise' or 'fast'")
    flags = _lib.PCM_PRECISE if mode == "precise" else 0
    if defer_fixup and flags:
        flags |= _lib.PCM_DEFER_FIXUP
    return flags


def pcm_forward(cam_norm_t: torch.Tensor, f: torch.Tensor, inference: bool = False, mode: Optional[str] = None,
                defer_fixup: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """network/resnet38_contrast.py:70-73.  Returns (cam_rv [N,C,h,w], workspace kept for the backward).
    ``inference=True``: no backward will follow (no sign mask; workspace not reusable for pcm_backward).
    ``defer_fixup=True`` (precise mode): the caller runs ``pcm_sign_fixup(ws, ...)`` itself before the backward."""
    f = _require_cuda_f32("f", f)
    cam_norm_t = _require_cuda_f32("cam_norm", cam_norm_t)
    n, k, h, w = f.shape
    c = cam_norm_t.shape[1]
    if cam_norm_t.shape[0] != n or tuple(cam_norm_t.shape[2:]) != (h, w):
        raise RuntimeError(f"pcm: cam {tuple(cam_norm_t.shape)} does not match f {tuple(f.shape)}")
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        ws = _workspace(_lib.OP_PCM, f.device, N=n, C=c, K=k, h=h, w=w)
        if inference:
            check(_lib.lib().wsegk_pcm_fwd_infer(_ptr(f), _ptr(cam_norm_t), _ptr(out), n, c, k, h, w, _ptr(ws), ws.numel(),
                                                 _stream(f)), "wsegk_pcm_fwd_infer")
        else:
            check(_lib.lib().wsegk_pcm_fwd_ex(_ptr(f), _ptr(cam_norm_t), None, _ptr(out), n, c, k, h, w, _ptr(ws),
                                              ws.numel(), _pcm_flags(mode, defer_fixup), _stream(f)), "wsegk_pcm_fwd_ex")
    return out, ws


def pcm_infer_multi(cam_norms, fs):
    """Forward-only PCM of several problems of DIFFERENT grids in ONE launch of the tcgen05 kernel (the 4
    scales x {image, flip} of contrast_infer.py:58-71).  ``cam_norms[i]`` [N_i,C,h_i,w_i] (already normalised, K1a),
    ``fs[i]`` [N_i,K,h_i,w_i]; returns the list of ``cam_rv`` maps.  No backward."""
    import ctypes
    n_items = len(fs)
    if n_items == 0 or len(cam_norms) != n_items:
        raise RuntimeError("pcm_infer_multi: need as many cam maps as feature maps (at least one)")
    fs = [_require_cuda_f32("f", t.detach()) for t in fs]
    cams = [_require_cuda_f32("cam_norm", t.detach()) for t in cam_norms]
    dev = fs[0].device
    k, c = fs[0].shape[1], cams[0].shape[1]
    for f, cam in zip(fs, cams):
        if f.shape[1] != k or cam.shape[1] != c or cam.shape[0] != f.shape[0] or tuple(cam.shape[2:]) != tuple(f.shape[2:]) \
                or f.device != dev:
            raise RuntimeError(f"pcm_infer_multi: cam {tuple(cam.shape)} does not match f {tuple(f.shape)}")
    ints = lambda vals: (ctypes.c_int * n_items)(*vals)
    ptrs = lambda ts: (ctypes.c_void_p * n_items)(*[t.data_ptr() for t in ts])
    ns, hs, ws_ = ints([f.shape[0] for f in fs]), ints([f.shape[2] for f in fs]), ints([f.shape[3] for f in fs])
    sizes = (ctypes.c_size_t * n_items)()
    check(_lib.lib().wsegk_pcm_multi_workspace_bytes(n_items, ns, hs, ws_, c, k, sizes), "wsegk_pcm_multi_workspace_bytes")
    outs = [torch.empty_like(cam) for cam in cams]
    with torch.cuda.device(dev):
        spaces = []
        for i in range(n_items):
            raw = torch.empty(int(sizes[i]) + 1024, dtype=torch.uint8, device=dev)
            shift = (-raw.data_ptr()) % 1024
            spaces.append(raw[shift:shift + int(sizes[i])])
        check(_lib.lib().wsegk_pcm_fwd_infer_multi(n_items, ptrs(fs), ptrs(cams), ptrs(outs), ns, hs, ws_, c, k, ptrs(spaces),
                                                   sizes, _stream(fs[0])), "wsegk_pcm_fwd_infer_multi")
    return outs


def pcm_forward_logits(cam_logits: torch.Tensor, f: torch.Tensor, mode: Optional[str] = None, defer_fixup: bool = False):
    """K1a + K2 forward in one call (network/resnet38_contrast.py:41-48 fused into :70-73): the normalisation of the raw
    class logits happens inside the PCM's operand preparation.  Returns (cam_rv, workspace, cam_norm) — ``cam_norm`` is
    bit-identical to ``cam_norm(cam_logits)`` and is what ``pcm_backward`` needs."""
    f = _require_cuda_f32("f", f)
    cam_logits = _require_cuda_f32("cam", cam_logits)
    n, k, h, w = f.shape
    c = cam_logits.shape[1]
    if cam_logits.shape[0] != n or tuple(cam_logits.shape[2:]) != (h, w):
        raise RuntimeError(f"pcm: cam {tuple(cam_logits.shape)} does not match f {tuple(f.shape)}")
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=f.device)
    camn = torch.empty_like(cam_logits)
    with torch.cuda.device(f.device):
        ws = _workspace(_lib.OP_PCM, f.device, N=n, C=c, K=k, h=h, w=w)
        check(_lib.lib().wsegk_pcm_fwd_ex(_ptr(f), _ptr(cam_logits), _ptr(camn), _ptr(out), n, c, k, h, w, _ptr(ws),
                                          ws.numel(), _pcm_flags(mode, defer_fixup), _stream(f)), "wsegk_pcm_fwd_ex")
    return out, ws, camn


def pcm_sign_fixup(ws: torch.Tensor, n: int, c: int, k: int, h: int, w: int) -> None:
    """The sign fix-up of a precise forward that was called with ``defer_fixup=True``, on torch's current stream (the
    caller orders it after the forward and before ``pcm_backward``, e.g. on a side stream off the critical path)."""
    with torch.cuda.device(ws.device):
        check(_lib.lib().wsegk_pcm_sign_fixup(n, c, k, h, w, _ptr(ws), ws.numel(),
                                              torch.cuda.current_stream(ws.device).cuda_stream), "wsegk_pcm_sign_fixup")


def pcm_backward(f: torch.Tensor, cam_norm_t: torch.Tensor, cam_rv: torch.Tensor, grad_out: torch.Tensor,
                 ws: torch.Tensor, mode: Optional[str] = None) -> torch.Tensor:
    """``mode`` must be the mode of the forward that filled ``ws`` (None: the calling thread's default)."""
    f = _require_cuda_f32("f", f)
    grad_out = _require_cuda_f32("grad_cam_rv", grad_out)
    n, k, h, w = f.shape
    c = cam_norm_t.shape[1]
    grad_f = torch.empty_like(f)
    with torch.cuda.device(f.device):
        check(_lib.lib().wsegk_pcm_bwd_ex(_ptr(f), _ptr(cam_norm_t), _ptr(cam_rv), _ptr(grad_out), _ptr(grad_f),
                                          n, c, k, h, w, _ptr(ws), ws.numel(), _pcm_flags(mode), _stream(f)),
              "wsegk_pcm_bwd_ex")
    return grad_f


class _PCMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam_norm_t, f):
        ctx.mode = get_pcm_mode()            # the backward may run on another (autograd) thread: the mode travels in ctx
        out, ws = pcm_forward(cam_norm_t, f, mode=ctx.mode)
        ctx.save_for_backward(f, cam_norm_t, out)
        ctx.ws = ws
        return out

    @staticmethod
    def backward(ctx, grad_out):
        f, cam_norm_t, out = ctx.saved_tensors
        grad_f = pcm_backward(f, cam_norm_t, out, grad_out, ctx.ws, mode=ctx.mode)
        ctx.ws = None
        return None, grad_f


class _PCMFromLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam_logits, f):
        ctx.mode = get_pcm_mode()
        out, ws, camn = pcm_forward_logits(cam_logits, f, mode=ctx.mode)
        ctx.save_for_backward(f, camn, out)
        ctx.ws = ws
        return out

    @staticmethod
    def backward(ctx, grad_out):
        f, camn, out = ctx.saved_tensors
        grad_f = pcm_backward(f, camn, out, grad_out, ctx.ws, mode=ctx.mode)
        ctx.ws = None
        return None, grad_f


def pcm_from_logits(cam_logits: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """``pcm(cam_norm(cam_logits), f)`` with the normalisation fused into the PCM's operand preparation (same values bit for
    bit; no gradient flows to the logits, as in the reference, which normalises under ``torch.no_grad()``)."""
    if not (torch.is_grad_enabled() and f.requires_grad):
        return pcm_forward(cam_norm(cam_logits.detach()), f.detach(), inference=True)[0]
    return _PCMFromLogits.apply(cam_logits.detach(), f)


def pcm(cam_norm_t: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """Differentiable (w.r.t. ``f``) pixel correlation module; ``f`` is the output of the f9 conv.  When no gradient can
    flow (torch.no_grad() as in contrast_infer.py, or f does not require grad) the forward-only kernel is used."""
    if not (torch.is_grad_enabled() and f.requires_grad):
        return pcm_forward(cam_norm_t.detach(), f.detach(), inference=True)[0]
    return _PCMFunction.apply(cam_norm_t.detach(), f)


# Relative Frobenius error bound of grad_f vs the fp32 oracle that the PCM backward is held to in the default
# ("precise") mode (tests/test_gpu_parity.py; BASELINE.json north_star: gradients within 1e-3).
PCM_GRAD_TOL = 1e-3
# "fast" mode: the sign of S is decided from single FP16 products, so entries with |S| < ~1e-4 (about 1e-4 of all
# entries for random-sign features) can flip and each flip adds/drops a whole gradient term.
PCM_GRAD_TOL_FAST = 1e-1


def launch_count() -> int:
    """Kernels enqueued by libwsegk in this process so far (bench.py reports the per-step difference)."""
    return int(_lib.lib().wsegk_debug_launch_count())


def force_radix(on: bool) -> None:
    """Test aid: route K3's top-k and K4's mining through their large-plane (multi-CTA select) paths at any size."""
    check(_lib.lib().wsegk_debug_force_radix(1 if on else 0), "wsegk_debug_force_radix")


def pcm_profile(enable: bool) -> None:
    check(_lib.lib().wsegk_debug_pcm_profile(1 if enable else 0), "wsegk_debug_pcm_profile")


def pcm_fixup_ms() -> float:
    """Duration in ms of the last profiled sign fix-up kernel (precise forward), or -1."""
    import ctypes
    a = ctypes.c_float(-1.0)
    check(_lib.lib().wsegk_debug_pcm_fixup_ms(ctypes.byref(a)), "wsegk_debug_pcm_fixup_ms")
    return float(a.value)


def pcm_main_ms() -> Tuple[float, float]:
    """(forward, backward) duration in ms of the last profiled tcgen05 main kernels (CUDA events, launching stream)."""
    import ctypes
    a, b = ctypes.c_float(-1.0), ctypes.c_float(-1.0)
    check(_lib.lib().wsegk_debug_pcm_main_ms(ctypes.byref(a), ctypes.byref(b)), "wsegk_debug_pcm_main_ms")
    return float(a.value), float(b.value)


def set_pcm_mode(mode: str) -> None:
    """Default mode of the calling THREAD.  'precise' (default): the near-zero entries of S are re-evaluated in fp32, so the
    sign mask the backward consumes is exact; 'fast': signs straight from the FP16 products."""
    if mode not in ("precise", "fast"):
        raise ValueError("mode must be 'precise' or 'fast'")
    check(_lib.lib().wsegk_pcm_set_mode(1 if mode == "precise" else 0), "wsegk_pcm_set_mode")


def set_pcm_pairs(fwd: int = 0, bwd: int = 0) -> None:
    """CTA pairs (of the GPU's 74) the PCM main kernels of the calling thread occupy; 0 = the default (tuned for the whole
    training step: 64 / 52).  72 / 72 is fastest for a PCM call on an otherwise idle GPU.  Set it before sizing workspaces."""
    check(_lib.lib().wsegk_pcm_set_pairs(int(fwd), int(bwd)), "wsegk_pcm_set_pairs")


def get_pcm_mode() -> str:
    return "precise" if _lib.lib().wsegk_pcm_get_mode() == 1 else "fast"


def pcm_layout(n: int, c: int, k: int, h: int, w: int) -> dict:
    """Test aid: byte offsets of the PCM workspace sections and the tiling parameters."""
    import ctypes
    buf = (ctypes.c_size_t * 20)()
    got = _lib.lib().wsegk_debug_pcm_layout(n, c, k, h, w, buf, 20)
    names = ["fa", "fh32", "fb", "nzmask", "camb", "inv_norm", "dsum", "part", "mask", "ua", "vb", "dfhat",
             "bound", "bytes", "nsplit", "PJ", "HWpj", "wb", "NS", "HWpi"]
    return {names[i]: int(buf[i]) for i in range(got)}


# --------------------------------------------------------------------------------------------- K1b
def _label_nc(label: torch.Tensor, n: int, c: int) -> torch.Tensor:
    lab = label.detach().reshape(n, -1).to(torch.float32)
    if lab.shape[1] != c:
        raise RuntimeError(f"label has {lab.shape[1]} classes per image, tensor has {c} channels")
    return lab.contiguous()


def fullres_norm_fwd(p: torch.Tensor, label: torch.Tensor, size: Optional[Tuple[int, int]] = None):
    """Raw K1b forward.  Returns (out [N,C,oh,ow], saved) where saved = (minmax [N,C,4], tapsum [N,C,oh,ow],
    label [N,C] on device) is what the backward needs besides p and out."""
    p = _require_cuda_f32("p", p.detach())
    n, c, H, W = p.shape
    oh, ow = (H, W) if size is None else (int(size[0]), int(size[1]))
    lab = _label_nc(label, n, c).to(p.device)
    out = torch.empty(n, c, oh, ow, dtype=torch.float32, device=p.device)
    tapsum = torch.empty(n, c, oh, ow, dtype=torch.float32, device=p.device)
    minmax = torch.empty(n, c, 4, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        ws = _workspace(_lib.OP_FULLRES_NORM, p.device, N=n, C=c, h=H, w=W)
        check(_lib.lib().wsegk_fullres_norm_fwd(_ptr(p), _ptr(lab), _ptr(out), _ptr(minmax), _ptr(tapsum), n, c, H, W,
                                                oh, ow, _ptr(ws), ws.numel(), _stream(p)), "wsegk_fullres_norm_fwd")
    return out, (minmax, tapsum, lab)


def fullres_norm_bwd(p: torch.Tensor, out: torch.Tensor, saved, grad_out: torch.Tensor) -> torch.Tensor:
    """Raw K1b backward: grad_out [N,C,oh,ow] -> grad_p [N,C,H,W]."""
    minmax, tapsum, lab = saved
    p = _require_cuda_f32("p", p.detach())
    grad_out = _require_cuda_f32("grad_out", grad_out)
    n, c, H, W = p.shape
    oh, ow = out.shape[-2:]
    grad_p = torch.empty_like(p)
    with torch.cuda.device(p.device):
        ws = _workspace(_lib.OP_FULLRES_NORM, p.device, N=n, C=c, h=H, w=W)
        check(_lib.lib().wsegk_fullres_norm_bwd(_ptr(p), _ptr(lab), _ptr(minmax), _ptr(out), _ptr(tapsum),
                                                _ptr(grad_out), _ptr(grad_p), n, c, H, W, oh, ow, _ptr(ws), ws.numel(),
                                                _stream(p)), "wsegk_fullres_norm_bwd")
    return grad_p


class _FullresNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, label, size):
        out, (minmax, tapsum, lab) = fullres_norm_fwd(p, label, size)
        ctx.save_for_backward(p, out, minmax, tapsum, lab)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, out, minmax, tapsum, lab = ctx.saved_tensors
        return fullres_norm_bwd(p, out, (minmax, tapsum, lab), grad_out), None, None


def fullres_norm(p: torch.Tensor, label: torch.Tensor, size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """visualization.max_norm(p) -> bilinear(align_corners) to ``size`` -> * label
    (tool/visualization.py:62-67, contrast_train.py:145-152,157-158).  label: [N,C] or [N,C,1,1].

    If ``p`` is the untouched output of :func:`bilinear_resize` enlarging a small map (what ``Net.forward`` returns as
    ``cam`` / ``cam_rv``, network/resnet38_contrast.py:57-59), the normalisation is evaluated from the SMALL map
    (:func:`fullres_norm_up`: bit-identical values) and the gradient goes straight back to it: the dense full-resolution
    gradient (135 MB per map at N = 8, 448x448: written by this op's backward, read again by the enlargement's) never
    exists.  Any operation on ``p`` in between, in place or not, gives a tensor without the note and the plain path."""
    src = _enlarged_from(p)
    if src is not None:
        return _FullresNormUp.apply(src, label, tuple(p.shape[-2:]), None if size is None else tuple(size))
    return _FullresNorm.apply(p, label, size)


class _FullresNormUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, small, label, up_size, size):
        small = _require_cuda_f32("small", small.detach())
        n, c, h, w = small.shape
        H, W = int(up_size[0]), int(up_size[1])
        oh, ow = (H, W) if size is None else (int(size[0]), int(size[1]))
        lab = _label_nc(label, n, c).to(small.device)
        out = torch.empty(n, c, oh, ow, dtype=torch.float32, device=small.device)
        tapsum = torch.empty(n, c, oh, ow, dtype=torch.float32, device=small.device)
        minmax = torch.empty(n, c, 4, dtype=torch.float32, device=small.device)
        with torch.cuda.device(small.device):
            ws = _workspace(_lib.OP_FULLRES_NORM, small.device, N=n, C=c, h=H, w=W)
            check(_lib.lib().wsegk_fullres_norm_up_fwd(_ptr(small), _ptr(lab), _ptr(out), _ptr(minmax), _ptr(tapsum), n, c,
                                                       h, w, H, W, oh, ow, _ptr(ws), ws.numel(), _stream(small)),
                  "wsegk_fullres_norm_up_fwd")
        ctx.save_for_backward(small, out, minmax, tapsum, lab)
        ctx.dims = (H, W)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        small, out, minmax, tapsum, lab = ctx.saved_tensors
        grad_out = _require_cuda_f32("grad_out", grad_out)
        n, c, h, w = small.shape
        H, W = ctx.dims
        oh, ow = out.shape[-2:]
        grad_small = torch.empty_like(small)
        with torch.cuda.device(small.device):
            ws = _workspace(_lib.OP_LAZY_BWD, small.device, N=n, C=c, h=H, w=W)   # only labelled planes are touched
            check(_lib.lib().wsegk_fullres_norm_up_bwd(_ptr(small), _ptr(lab), _ptr(minmax), _ptr(out), _ptr(tapsum),
                                                       _ptr(grad_out), _ptr(grad_small), n, c, h, w, H, W, oh, ow,
                                                       _ptr(ws), ws.numel(), _stream(small)),
                  "wsegk_fullres_norm_up_bwd")
        return grad_small, None, None, None


def fullres_norm_up(small: torch.Tensor, label: torch.Tensor, up_size: Tuple[int, int],
                    size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """``fullres_norm(bilinear_resize(small, up_size), label, size)`` without the full-resolution tensor: the x8
    enlargement at the end of ``Net.forward`` (network/resnet38_contrast.py:57-59) fused into the normalisation the loss
    applies to it (contrast_train.py:145-152).  Same values bit for bit; the gradient comes back on the small map."""
    return _FullresNormUp.apply(small, label, tuple(up_size), None if size is None else tuple(size))


# --------------------------------------------------------------------------------------------- glue (:179-180)
def resize_fwd(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    x = _require_cuda_f32("x", x.detach())
    n, c, H, W = x.shape
    oh, ow = int(size[0]), int(size[1])
    y = torch.empty(n, c, oh, ow, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().wsegk_bilinear_resize_fwd(_ptr(x), _ptr(y), n * c, H, W, oh, ow, _stream(x)),
              "wsegk_bilinear_resize_fwd")
    return y


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        ctx.in_shape = tuple(x.shape)
        return resize_fwd(x, size)

    @staticmethod
    def backward(ctx, grad_y):
        n, c, H, W = ctx.in_shape
        grad_y = _require_cuda_f32("grad_y", grad_y)
        oh, ow = grad_y.shape[-2:]
        gx = torch.empty(n, c, H, W, dtype=torch.float32, device=grad_y.device)
        with torch.cuda.device(grad_y.device):
            check(_lib.lib().wsegk_bilinear_resize_bwd(_ptr(grad_y), _ptr(gx), n * c, H, W, oh, ow, _stream(grad_y)),
                  "wsegk_bilinear_resize_bwd")
        return gx, None


_REMEMBER_ENLARGEMENT = os.environ.get("WSEG_REMEMBER_ENLARGEMENT", "1") == "1"


def remember_enlargements(on: bool) -> None:
    """Switch the note :func:`bilinear_resize` leaves on an enlarged map (see :func:`fullres_norm`); tests compare both."""
    global _REMEMBER_ENLARGEMENT
    _REMEMBER_ENLARGEMENT = bool(on)


def _enlarged_from(p: torch.Tensor) -> Optional[torch.Tensor]:
    """The small map ``p`` was enlarged from, if ``p`` and that map are still what they were when it was made."""
    note = getattr(p, "_wseg_enlarged", None)
    if note is None or not _REMEMBER_ENLARGEMENT:
        return None
    small, v_small, v_self = note
    if p._version != v_self or small._version != v_small or not p.is_cuda:
        return None
    h, w = small.shape[-2:]
    H, W = p.shape[-2:]
    if h * w * 4 + 2 * (H + W) * 8 + 64 > 200 * 1024:           # the fused kernels keep the small plane in shared memory
        return None
    return small


def bilinear_resize(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """F.interpolate(x, size, mode='bilinear', align_corners=True), differentiable (contrast_train.py:179-180,
    network/resnet38_contrast.py:57-59).  ATen's NCHW kernel loops over N*C inside each thread: 0.3 ms for
    [8,128,56,56] -> 16x16 and 0.5 ms (0.8 ms backward) for [8,21,56,56] -> 448x448 on a B200.
    An ENLARGED map carries a note of the map it came from (a Python attribute, lost by any further operation), which
    lets :func:`fullres_norm` skip the full-resolution gradient."""
    y = _Resize.apply(x, size)
    if _REMEMBER_ENLARGEMENT and x.dim() == 4 and y.shape[-2] >= x.shape[-2] and y.shape[-1] >= x.shape[-1] \
            and x.dtype == torch.float32:
        y._wseg_enlarged = (x, x._version, y._version)
    return y


# --------------------------------------------------------------------------------------------- K1c
def seed_labels(cam_rv_down: torch.Tensor, label: torch.Tensor, bg_threshold: float
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """contrast_train.py:186-197.  Returns (norm_cam [N,C,g,g] with channel 0 = bg_threshold,
    pseudo_label [N*g*g] int32)."""
    x = _require_cuda_f32("cam_rv_down", cam_rv_down.detach())
    n, c, h, w = x.shape
    lab = _label_nc(label, n, c).to(x.device)
    norm = torch.empty_like(x)
    pseudo = torch.empty(n * h * w, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _workspace(_lib.OP_SEED_LABELS, x.device, N=n, C=c, h=h, w=w)
        check(_lib.lib().wsegk_seed_labels_fwd(_ptr(x), _ptr(lab), float(bg_threshold), _ptr(norm), _ptr(pseudo),
                                               n, c, h, w, _ptr(ws), ws.numel(), _stream(x)),
              "wsegk_seed_labels_fwd")
    return norm, pseudo


# --------------------------------------------------------------------------------------------- K3
def proto_topk_local(norm_cam: torch.Tensor, f_proj: torch.Tensor, k: Optional[int] = None,
                     pixel_offset: int = 0):
    """This rank's k best pixels per class: (top_val [C,k], top_idx [C,k] int32 global pixel ids,
    top_fea [C,k,D])."""
    norm_cam = _require_cuda_f32("norm_cam", norm_cam.detach())
    f_proj = _require_cuda_f32("f_proj", f_proj.detach())
    n, c, h, w = norm_cam.shape
    d = f_proj.shape[1]
    if k is None:
        k = h * w // 8
    dev = norm_cam.device
    tv = torch.empty(c, k, dtype=torch.float32, device=dev)
    ti = torch.empty(c, k, dtype=torch.int32, device=dev)
    tf = torch.empty(c, k, d, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(_lib.OP_PROTO, dev, N=n, C=c, D=d, h=h, w=w)
        check(_lib.lib().wsegk_proto_topk_local(_ptr(norm_cam), _ptr(f_proj), int(pixel_offset), int(k), _ptr(tv),
                                                _ptr(ti), _ptr(tf), n, c, d, h, w, _ptr(ws), ws.numel(),
                                                _stream(norm_cam)), "wsegk_proto_topk_local")
    return tv, ti, tf


def proto_finalize(cand_val: torch.Tensor, cand_idx: torch.Tensor, cand_fea: torch.Tensor, k: int):
    """Global top-k among candidates [C,n_cand] (+features [C,n_cand,D]) -> (prototypes [C,D], sel_idx [C,k])."""
    cand_val = _require_cuda_f32("cand_val", cand_val)
    cand_fea = _require_cuda_f32("cand_fea", cand_fea)
    cand_idx = cand_idx.to(torch.int32).contiguous()
    c, n_cand = cand_val.shape
    d = cand_fea.shape[-1]
    dev = cand_val.device
    proto = torch.empty(c, d, dtype=torch.float32, device=dev)
    sel = torch.empty(c, k, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().wsegk_proto_finalize(_ptr(cand_val), _ptr(cand_idx), _ptr(cand_fea), n_cand, int(k),
                                              _ptr(proto), _ptr(sel), c, d, _stream(cand_val)),
              "wsegk_proto_finalize")
    return proto, sel


def prototypes(norm_cam: torch.Tensor, f_proj: torch.Tensor, k: Optional[int] = None, group=None, view: int = 0,
               exchange: Optional[str] = None):
    """contrast_train.py:199-209 with whole-batch semantics.  With a process group the batch is the union of all
    ranks' shards (rank r owns pixels [r*P_local, (r+1)*P_local)): local candidates are exchanged (the only
    collective on the path) and every rank finalises the same global prototypes, bit-identically.
    ``exchange``: "nvlink" = one-shot stores into every peer's symmetric buffer from this rank's kernel (no NCCL
    kernel; ``view`` in {0,1} selects the slot), "nccl" = one all_gather_into_tensor; default: $WSEG_EXCHANGE or
    "nvlink" with a fall back to "nccl" if symmetric memory cannot be set up.
    Returns (prototypes [C,D], top_idx [C,k] int32 global pixel ids)."""
    n, c, h, w = norm_cam.shape
    if k is None:
        k = h * w // 8
    from . import dist as _dist
    rank, world = _dist.rank_world(group)
    tv, ti, tf = proto_topk_local(norm_cam, f_proj, k, pixel_offset=rank * n * h * w)
    if world == 1:
        return proto_finalize(tv, ti, tf, k)
    mode = exchange or os.environ.get("WSEG_EXCHANGE", "nvlink")
    if mode == "nvlink":
        ex = _symm_exchange(group, c, k, tf.shape[-1], norm_cam.device)
        if ex is not None:
            return _prototypes_nvlink(ex, tv, ti, tf, k, view)
    tv, ti, tf = _dist.gather_candidates(tv, ti, tf, group)
    return proto_finalize(tv, ti, tf, k)


_SYMM_MODE = {}            # (group, dims, device) -> SymmExchange, or None = every rank agreed to use NCCL


def _symm_exchange(group, c, k, d, device):
    """The NVLink exchange object, or None if ANY rank of the group could not set it up.  The decision is collective (one
    all_reduce(MIN) of a success flag at set-up time): a rank that fell back to NCCL on its own would sit in an all-gather
    while its peers spin on NVLink flags that never arrive."""
    key = (id(group), c, k, d, str(device))
    if key in _SYMM_MODE:
        return _SYMM_MODE[key]
    import torch.distributed as _td
    from . import dist as _dist
    ex, err = None, None
    try:
        ex = _dist.SymmExchange.get(group, c, k, d, device)
    except Exception as exc:          # no peer access / no symmetric-memory support on this stack
        err = exc
    ok = torch.tensor([1 if ex is not None else 0], dtype=torch.int32, device=device)
    _td.all_reduce(ok, op=_td.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        import warnings
        why = f"{type(err).__name__}: {err}" if err is not None else "a peer rank could not set it up"
        warnings.warn(f"wseg_b200: NVLink symmetric exchange unavailable ({why}); every rank uses NCCL")
        ex = None
    _SYMM_MODE[key] = ex
    return ex


def _prototypes_nvlink(ex, tv, ti, tf, k: int, view: int):
    c, _, d = tf.shape
    dev = tv.device
    proto = torch.empty(c, d, dtype=torch.float32, device=dev)
    sel = torch.empty(c, k, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().wsegk_proto_push(_ptr(tv), _ptr(ti), _ptr(tf), ex.peer_bases, ex.world, ex.rank, int(view), c, k, d,
                                          _ptr(ex.state), _stream(tv)), "wsegk_proto_push")
        check(_lib.lib().wsegk_proto_finalize_symm(_ptr(ex.buf), ex.world, int(view), k, _ptr(proto), _ptr(sel), c, d,
                                                   _ptr(ex.state), _stream(tv)), "wsegk_proto_finalize_symm")
    return proto, sel


# --------------------------------------------------------------------------------------------- K4
LOSS_SLOTS = ("nce_q1_p2_l1", "nce_q2_p1_l2", "nce_q1_p1_l2", "nce_q2_p2_l1", "intra1", "intra2",
              "loss_cross_nce", "loss_cross_nce2", "loss_intra_nce", "loss_nce")


def hash_keys(n_pixels: int, seed: int, device) -> torch.Tensor:
    """Per-pixel uint32 keys for the 'random 50 %' of the hard-pixel mining (stored as int32 bit patterns)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    keys = torch.randint(-2 ** 31, 2 ** 31 - 1, (n_pixels,), dtype=torch.int64, generator=g).to(torch.int32)
    return keys.to(device)


def mining_keys(n_pixels: int, seed: int, device, step: int = 0, stream_id: int = 0, index0: int = 0) -> torch.Tensor:
    """Per-pixel uint32 mining keys generated ON the device (Philox4x32-10, counter = (index0 + i, step, stream_id),
    key = seed), stored as int32 bit patterns.  Nothing is created on the host; ``index0`` = first global pixel of this
    rank's shard, so a pixel gets the same key under any data-parallel sharding."""
    out = torch.empty(int(n_pixels), dtype=torch.int32, device=device)
    with torch.cuda.device(out.device):
        check(_lib.lib().wsegk_mining_keys(_ptr(out), int(n_pixels), int(index0), int(seed) & (2 ** 64 - 1), int(step),
                                           int(stream_id) & 0xFFFF, _stream(out)), "wsegk_mining_keys")
    return out


def philox4x32_10(counter, key):
    """Host reference of the generator (Salmon et al., SC'11) for tests: 4 x uint32 counter, 2 x uint32 key -> 4 x uint32."""
    m0, m1, w0, w1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c, k = [int(x) & mask for x in counter], [int(x) & mask for x in key]
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & mask, (p0 >> 32) ^ c[3] ^ k[1], p0 & mask]
        k = [(k[0] + w0) & mask, (k[1] + w1) & mask]
    return c


def contrast_loss_raw(x1, x2, proto1, proto2, pl1, pl2, key1, key2, grad_scale: Optional[torch.Tensor] = None,
                      want_mined: bool = False, global_group=None, global_exchange=None):
    """One call of wsegk_contrast_loss_fwd_bwd.  Returns (losses [10], grad_x1, grad_x2, mined1, mined2).
    ``global_group``: a process group over which the hard-pixel mining is GLOBAL (the reference's whole-batch semantics
    under data parallelism; one NVLink exchange of 12 bytes per pixel and view; needs symmetric memory).
    ``global_exchange``: the exchange itself (an object with peer_bases, buf, world, rank, state as dist.MineExchange has
    them) instead of a group — lets one process drive several "ranks" whose buffers are plain device memory (tests)."""
    x1 = _require_cuda_f32("x1", x1.detach())
    x2 = _require_cuda_f32("x2", x2.detach())
    n, d, h, w = x1.shape
    if tuple(x2.shape) != (n, d, h, w):
        raise RuntimeError("contrast loss: both views must be on the same grid "
                           f"(got {tuple(x1.shape)} vs {tuple(x2.shape)}; contrast_train.py:179-180)")
    c = proto1.shape[0]
    dev = x1.device
    losses = torch.empty(10, dtype=torch.float32, device=dev)
    g1 = torch.empty_like(x1)
    g2 = torch.empty_like(x2)
    m1 = torch.empty(n * h * w, dtype=torch.int32, device=dev) if want_mined else None
    m2 = torch.empty(n * h * w, dtype=torch.int32, device=dev) if want_mined else None
    args = [_require_cuda_f32("proto1", proto1), _require_cuda_f32("proto2", proto2),
            pl1.to(torch.int32).contiguous(), pl2.to(torch.int32).contiguous(),
            key1.to(torch.int32).contiguous(), key2.to(torch.int32).contiguous()]
    with torch.cuda.device(dev):
        ws = _workspace(_lib.OP_CONTRAST_LOSS, dev, N=n, C=c, D=d, h=h, w=w)
        if global_group is None and global_exchange is None:
            check(_lib.lib().wsegk_contrast_loss_fwd_bwd(
                _ptr(x1), _ptr(x2), *[_ptr(a) for a in args], _ptr(losses), _ptr(g1), _ptr(g2), _ptr(m1), _ptr(m2),
                _ptr(grad_scale), n, c, d, h, w, _ptr(ws), ws.numel(), _stream(x1)), "wsegk_contrast_loss_fwd_bwd")
        else:
            ex = global_exchange
            if ex is None:
                from . import dist as _dist
                ex = _dist.MineExchange.get(global_group, n, h, w, dev)
            check(_lib.lib().wsegk_contrast_loss_fwd_bwd_global(
                _ptr(x1), _ptr(x2), *[_ptr(a) for a in args], _ptr(losses), _ptr(g1), _ptr(g2), _ptr(m1), _ptr(m2),
                _ptr(grad_scale), n, c, d, h, w, _ptr(ws), ws.numel(), ex.peer_bases, _ptr(ex.buf), ex.world, ex.rank,
                _ptr(ex.state), _stream(x1)), "wsegk_contrast_loss_fwd_bwd_global")
    return losses, g1, g2, m1, m2


class _ContrastLoss(torch.autograd.Function):
    """(loss_nce, losses[10]) as a function of (x1, x2).  Only ``loss_nce`` is differentiable — its gradient is produced by
    the same fused call, so backward only rescales it; the 10 reported values are marked non-differentiable, so summing or
    re-weighting the parts instead of using ``loss_nce`` raises instead of silently yielding a zero gradient."""

    @staticmethod
    def forward(ctx, x1, x2, proto1, proto2, pl1, pl2, key1, key2, global_group=None):
        losses, g1, g2, _, _ = contrast_loss_raw(x1, x2, proto1, proto2, pl1, pl2, key1, key2, global_group=global_group)
        ctx.save_for_backward(g1, g2)
        ctx.mark_non_differentiable(losses)
        return losses[9].clone(), losses

    @staticmethod
    def backward(ctx, grad_nce, _grad_losses):
        g1, g2 = ctx.saved_tensors
        return g1 * grad_nce, g2 * grad_nce, None, None, None, None, None, None, None


def contrast_loss(x1, x2, proto1, proto2, pl1, pl2, key1, key2, global_group=None):
    """Returns (loss_nce — differentiable w.r.t. x1, x2 —, losses [10] reported values without gradient)."""
    return _ContrastLoss.apply(x1, x2, proto1, proto2, pl1, pl2, key1, key2, global_group)


# --------------------------------------------------------------------------------------------- auxiliary SEAM losses
def select_kth(x: torch.Tensor, k: int, largest: bool = True):
    """Per row of x [rows,len]: (sel [rows,4] = threshold, sum of the k selected, sum of relu(selected), ties taken;
    idx_cut [rows] int32).  Selection is by value: no sort, no index tensor."""
    x = _require_cuda_f32("x", x.detach())
    rows, ln = x.shape
    sel = torch.empty(rows, 4, dtype=torch.float32, device=x.device)
    cut = torch.empty(rows, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _workspace(_lib.OP_SELECT, x.device, N=rows, h=1, w=1)
        check(_lib.lib().wsegk_select_kth(_ptr(x), rows, ln, int(k), 1 if largest else 0, _ptr(sel), _ptr(cut), _ptr(ws),
                                          ws.numel(), _stream(x)), "wsegk_select_kth")
    return sel, cut


class _ClsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam, label):
        cam = _require_cuda_f32("cam", cam)
        n, c, H, W = cam.shape
        lab = _label_nc(label, n, c).to(cam.device)
        gap = torch.empty(n, c, dtype=torch.float32, device=cam.device)
        dgap = torch.empty_like(gap)
        loss = torch.empty(1, dtype=torch.float32, device=cam.device)
        ws = torch.empty(n * c * 8, dtype=torch.float32, device=cam.device)
        with torch.cuda.device(cam.device):
            check(_lib.lib().wsegk_gap_mlsm_fwd(_ptr(cam), _ptr(lab), _ptr(gap), _ptr(dgap), _ptr(loss), n, c, H, W, _ptr(ws),
                                                ws.numel() * 4, _stream(cam)), "wsegk_gap_mlsm_fwd")
        ctx.save_for_backward(dgap)
        ctx.dims = (n, c, H, W)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        (dgap,) = ctx.saved_tensors
        n, c, H, W = ctx.dims
        g = grad_loss.detach().reshape(1).to(torch.float32).contiguous()
        grad_cam = torch.empty(n, c, H, W, dtype=torch.float32, device=dgap.device)
        with torch.cuda.device(dgap.device):
            check(_lib.lib().wsegk_gap_mlsm_bwd(_ptr(dgap), _ptr(g), _ptr(grad_cam), n, c, H, W, _stream(dgap)),
                  "wsegk_gap_mlsm_bwd")
        return grad_cam, None


def cls_loss(cam: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """multilabel_soft_margin_loss(adaptive_avg_pool2d(cam,1)[:,1:], label[:,1:])  (contrast_train.py:142,159)."""
    return _ClsLoss.apply(cam, label)


class _MinPoolLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label):
        x = _require_cuda_f32("cam_rv", x)
        n, c, H, W = x.shape
        lab = _label_nc(label, n, c).to(x.device)
        m = torch.empty(n, H * W, dtype=torch.float32, device=x.device)
        arg = torch.empty(n, H * W, dtype=torch.int32, device=x.device)
        sel = torch.empty(n, 4, dtype=torch.float32, device=x.device)
        cut = torch.empty(n, dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            ws = _workspace(_lib.OP_SELECT, x.device, N=n, h=1, w=1)
            check(_lib.lib().wsegk_minpool_fwd(_ptr(x), _ptr(lab), _ptr(m), _ptr(arg), _ptr(sel), _ptr(cut), n, c, H, W,
                                               _ptr(ws), ws.numel(), _stream(x)), "wsegk_minpool_fwd")
        ctx.save_for_backward(m, arg, sel, cut, lab)
        ctx.dims = (n, c, H, W)
        k = (H * W) // 4
        return sel[:, 2].sum() / (k * n)

    @staticmethod
    def backward(ctx, grad_loss):
        m, arg, sel, cut, lab = ctx.saved_tensors
        n, c, H, W = ctx.dims
        g = grad_loss.detach().reshape(1).to(torch.float32).contiguous()
        grad_x = torch.empty(n, c, H, W, dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            check(_lib.lib().wsegk_minpool_bwd(_ptr(m), _ptr(arg), _ptr(sel), _ptr(cut), _ptr(lab), _ptr(g), _ptr(grad_x), n, c,
                                               H, W, _stream(m)), "wsegk_minpool_bwd")
        return grad_x, None


def min_pool_loss(cam_rv: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """adaptive_min_pooling_loss((cam_rv * label)[:, 1:])  (contrast_train.py:16-25,143)."""
    return _MinPoolLoss.apply(cam_rv, label)


class _ClsLossUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, small, label, up_size):
        small = _require_cuda_f32("cam", small)
        n, c, h, w = small.shape
        H, W = int(up_size[0]), int(up_size[1])
        lab = _label_nc(label, n, c).to(small.device)
        gap = torch.empty(n, c, dtype=torch.float32, device=small.device)
        dgap = torch.empty_like(gap)
        loss = torch.empty(1, dtype=torch.float32, device=small.device)
        ws = torch.empty(n * c * 8, dtype=torch.float32, device=small.device)
        with torch.cuda.device(small.device):
            check(_lib.lib().wsegk_gap_mlsm_up_fwd(_ptr(small), _ptr(lab), _ptr(gap), _ptr(dgap), _ptr(loss), n, c, h, w, H, W,
                                                   _ptr(ws), ws.numel() * 4, _stream(small)), "wsegk_gap_mlsm_up_fwd")
        ctx.save_for_backward(dgap)
        ctx.dims = (n, c, h, w, H, W)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        (dgap,) = ctx.saved_tensors
        n, c, h, w, H, W = ctx.dims
        g = grad_loss.detach().reshape(1).to(torch.float32).contiguous()
        grad_small = torch.empty(n, c, h, w, dtype=torch.float32, device=dgap.device)
        with torch.cuda.device(dgap.device):
            check(_lib.lib().wsegk_gap_mlsm_up_bwd(_ptr(dgap), _ptr(g), _ptr(grad_small), n, c, h, w, H, W, _stream(dgap)),
                  "wsegk_gap_mlsm_up_bwd")
        return grad_small, None, None


def cls_loss_up(cam_small: torch.Tensor, label: torch.Tensor, up_size: Tuple[int, int]) -> torch.Tensor:
    """``cls_loss(bilinear_resize(cam_small, up_size), label)`` without the full-resolution tensor."""
    return _ClsLossUp.apply(cam_small, label, tuple(up_size))


class _MinPoolLossUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, small, label, up_size):
        small = _require_cuda_f32("cam_rv", small)
        n, c, h, w = small.shape
        H, W = int(up_size[0]), int(up_size[1])
        lab = _label_nc(label, n, c).to(small.device)
        m = torch.empty(n, H * W, dtype=torch.float32, device=small.device)
        arg = torch.empty(n, H * W, dtype=torch.int32, device=small.device)
        sel = torch.empty(n, 4, dtype=torch.float32, device=small.device)
        cut = torch.empty(n, dtype=torch.int32, device=small.device)
        with torch.cuda.device(small.device):
            ws = _workspace(_lib.OP_SELECT, small.device, N=n, h=1, w=1)
            check(_lib.lib().wsegk_minpool_up_fwd(_ptr(small), _ptr(lab), _ptr(m), _ptr(arg), _ptr(sel), _ptr(cut), n, c, h, w,
                                                  H, W, _ptr(ws), ws.numel(), _stream(small)), "wsegk_minpool_up_fwd")
        ctx.save_for_backward(m, arg, sel, cut, lab)
        ctx.dims = (n, c, h, w, H, W)
        k = (H * W) // 4
        return sel[:, 2].sum() / (k * n)

    @staticmethod
    def backward(ctx, grad_loss):
        m, arg, sel, cut, lab = ctx.saved_tensors
        n, c, h, w, H, W = ctx.dims
        g = grad_loss.detach().reshape(1).to(torch.float32).contiguous()
        grad_small = torch.empty(n, c, h, w, dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            ws = _workspace(_lib.OP_LAZY_BWD, m.device, N=n, C=c, h=H, w=W)       # only labelled planes are touched
            check(_lib.lib().wsegk_minpool_up_bwd(_ptr(m), _ptr(arg), _ptr(sel), _ptr(cut), _ptr(lab), _ptr(g),
                                                  _ptr(grad_small), n, c, h, w, H, W, _ptr(ws), ws.numel(), _stream(m)),
                  "wsegk_minpool_up_bwd")
        return grad_small, None, None


def min_pool_loss_up(cam_rv_small: torch.Tensor, label: torch.Tensor, up_size: Tuple[int, int]) -> torch.Tensor:
    """``min_pool_loss(bilinear_resize(cam_rv_small, up_size), label)`` without the full-resolution tensor."""
    return _MinPoolLossUp.apply(cam_rv_small, label, tuple(up_size))


class _ErEcrLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n1, n2, rv1, rv2):
        n1, n2 = _require_cuda_f32("n1", n1), _require_cuda_f32("n2", n2)
        rv1, rv2 = _require_cuda_f32("rv1", rv1), _require_cuda_f32("rv2", rv2)
        n, c, h, w = n1.shape
        for t in (n2, rv1, rv2):
            if tuple(t.shape) != (n, c, h, w):
                raise RuntimeError("er_ecr_loss: the four maps must have the same shape")
        dev = n1.device
        e1 = torch.empty(n, c * h * w, dtype=torch.float32, device=dev)
        e2 = torch.empty_like(e1)
        sel1 = torch.empty(n, 4, dtype=torch.float32, device=dev)
        sel2 = torch.empty_like(sel1)
        cut1 = torch.empty(n, dtype=torch.int32, device=dev)
        cut2 = torch.empty_like(cut1)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _workspace(_lib.OP_ER_ECR, dev, N=n, C=c, h=h, w=w)
            check(_lib.lib().wsegk_er_ecr_fwd(_ptr(n1), _ptr(n2), _ptr(rv1), _ptr(rv2), _ptr(e1), _ptr(e2), _ptr(sel1),
                                              _ptr(cut1), _ptr(sel2), _ptr(cut2), _ptr(losses), n, c, h, w, _ptr(ws),
                                              ws.numel(), _stream(n1)), "wsegk_er_ecr_fwd")
        ctx.save_for_backward(n1, n2, rv1, rv2, sel1, cut1, sel2, cut2)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_er, g_ecr):
        n1, n2, rv1, rv2, sel1, cut1, sel2, cut2 = ctx.saved_tensors
        n, c, h, w = n1.shape
        ge = g_er.detach().reshape(1).to(torch.float32).contiguous()
        gc = g_ecr.detach().reshape(1).to(torch.float32).contiguous()
        outs = [torch.empty_like(n1) for _ in range(4)]
        with torch.cuda.device(n1.device):
            check(_lib.lib().wsegk_er_ecr_bwd(_ptr(n1), _ptr(n2), _ptr(rv1), _ptr(rv2), _ptr(sel1), _ptr(cut1), _ptr(sel2),
                                              _ptr(cut2), _ptr(ge), _ptr(gc), *[_ptr(o) for o in outs], n, c, h, w,
                                              _stream(n1)), "wsegk_er_ecr_bwd")
        return tuple(outs)


def er_ecr_loss(n1: torch.Tensor, n2: torch.Tensor, rv1: torch.Tensor, rv2: torch.Tensor):
    """(loss_er, loss_ecr) of contrast_train.py:163-172 from the four label-masked normalised maps (K1b outputs):
    n1/rv1 of view 1 resized to view 2's resolution, n2/rv2 of view 2."""
    return _ErEcrLoss.apply(n1, n2, rv1, rv2)
