"""ORACLE — test infrastructure, NOT product code.

Runs the reference's OWN source for the hot path on CPU, unmodified, straight from the read-only
checkout (``$WSEG_REFERENCE`` or ``/root/reference``).  Nothing is copied into this repository: the
loop body of ``contrast_train.py`` (lines 129-395, which has no function boundary — SURVEY.md §8b) is
read at run time, dedented and ``exec``'d with

* ``torch.Tensor.cuda`` patched to the identity (the body hard-codes ``.cuda()``; :140,:204,:233,:300,:353),
* ``pydensecrf`` stubbed in ``sys.modules`` (imported at the top of ``tool/visualization.py:5-6`` only
  for CRF drawing helpers the path never calls),
* optionally ``torch.topk`` replaced by the canonical-tie version (value, then lowest index), because
  the reference's class-0 plane is constant and ATen's tie order is unspecified.

The checkout only exists in the build container, so this module is used to (a) pin
``oracle/reference_fn.py`` and (b) generate ``tests/golden`` (``oracle/make_golden.py``).  It is never
needed on the GPU box.
"""
from __future__ import annotations

import contextlib
import os
import random
import sys
import textwrap
import types
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("WSEG_REFERENCE", "/root/reference")
BODY_FIRST, BODY_LAST = 129, 395        # contrast_train.py loop body
HELPERS_FIRST, HELPERS_LAST = 16, 32    # adaptive_min_pooling_loss, max_onehot


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "contrast_train.py"))


def _stub_missing_modules() -> None:
    for name in ("pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []  # type: ignore[attr-defined]
            sys.modules[name] = m
    sys.modules["pydensecrf.utils"].unary_from_softmax = None  # type: ignore[attr-defined]
    sys.modules["pydensecrf.utils"].unary_from_labels = None  # type: ignore[attr-defined]


@contextlib.contextmanager
def reference_on_path():
    _stub_missing_modules()
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REFERENCE_ROOT)


def import_reference_net():
    """The reference's ``network.resnet38_contrast`` module (needs only torch/numpy)."""
    with reference_on_path():
        import importlib
        return importlib.import_module("network.resnet38_contrast")


def _source_lines(rel: str, first: int, last: int) -> str:
    with open(os.path.join(REFERENCE_ROOT, rel)) as fh:
        lines = fh.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


@contextlib.contextmanager
def _patched(canonical_ties: bool):
    from . import reference_fn
    saved_cuda = torch.Tensor.cuda
    saved_topk = torch.topk
    torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]
    if canonical_ties:
        torch.topk = reference_fn.canonical_topk  # type: ignore[assignment]
    try:
        yield
    finally:
        torch.Tensor.cuda = saved_cuda  # type: ignore[assignment]
        torch.topk = saved_topk  # type: ignore[assignment]


def run_loop_body(model: Callable, img1: torch.Tensor, label20: torch.Tensor,
                  bg_threshold: float = 0.20, py_seed: Optional[int] = 0,
                  canonical_ties: bool = True) -> Dict[str, object]:
    """Execute contrast_train.py:129-395 once.  ``model(img) -> (cam, cam_rv, f_proj, cam_rv_down)``.
    Returns the body's local namespace (losses, pseudo labels, prototypes, ...)."""
    with reference_on_path():
        from tool import visualization  # noqa: WPS433 (reference module)
    ns: Dict[str, object] = {
        "torch": torch, "F": F, "np": np, "random": random, "visualization": visualization,
        "model": model, "pack": ("synthetic", img1, label20),
        "args": types.SimpleNamespace(bg_threshold=bg_threshold),
    }
    exec(compile(_source_lines("contrast_train.py", HELPERS_FIRST, HELPERS_LAST),
                 "contrast_train.py[16:32]", "exec"), ns)
    body = _source_lines("contrast_train.py", BODY_FIRST, BODY_LAST)
    if py_seed is not None:
        random.seed(py_seed)
    with _patched(canonical_ties):
        exec(compile(body, "contrast_train.py[129:395]", "exec"), ns)
    return ns


class HeadStub:
    """A stand-in ``model`` that returns pre-made head outputs keyed by input resolution, so the loss
    body can be run on synthetic head tensors without a backbone."""

    def __init__(self):
        self.by_size = {}

    def add(self, size, cam, cam_rv, f_proj, cam_rv_down):
        self.by_size[tuple(size)] = (cam, cam_rv, f_proj, cam_rv_down)

    def __call__(self, img):
        return self.by_size[tuple(img.shape[-2:])]
