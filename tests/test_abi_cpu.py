"""The C-ABI library builds, loads and exports every symbol include/wsegk.h declares; without a GPU every compute
entry point refuses loudly (no CPU fallback).  CPU only — no kernels are launched."""
import ctypes
import os
import re

import pytest
import torch

from wseg_b200 import _lib, build, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "wsegk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wsegk_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 13
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in include/wsegk.h but not exported"
    assert not _lib.MISSING
    assert set(names) == set(_lib.symbols()), "ctypes signature table out of sync with the header"


def test_version_and_workspace_sizes(lib):
    assert lib.wsegk_version() >= 100
    pcm = _lib.workspace_bytes(_lib.OP_PCM, N=8, C=21, K=192, h=56, w=56)
    assert pcm % 1024 == 0 and 60e6 < pcm < 200e6
    lay = ops.pcm_layout(8, 21, 192, 56, 56)
    # 56x56 = 3136 pixels: 13 CTA pairs of 256 rows, B tiles of 56 rows -> 28 stages of 112 columns, no padding along i
    assert (lay["PJ"], lay["HWpj"], lay["wb"], lay["NS"], lay["HWpi"]) == (13, 3328, 56, 28, 3136)
    assert lay["bytes"] == pcm and lay["fa_lo"] - lay["fa_hi"] == 8 * 26 * 128 * 192 * 2
    lay = ops.pcm_layout(1, 21, 192, 64, 64)
    assert (lay["wb"], lay["NS"], lay["HWpi"]) == (64, 32, 4096)
    assert _lib.workspace_bytes(_lib.OP_CONTRAST_LOSS, N=8, C=21, D=128, h=16, w=16) > 2 * 2048 * 4 * 42
    assert _lib.workspace_bytes(99, N=1, C=1, K=16, h=1, w=1) == 0
    assert "unknown op" in _lib.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal path")
def test_compute_refuses_without_gpu(lib):
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.wsegk_cam_norm_fwd(p, p, 1, 21, 1, 1, p, 4096, None)
    assert rc == -5                                            # WSEGK_ERR_DEVICE
    assert "no CPU path" in _lib.last_error() or "sm_100" in _lib.last_error()


def test_ops_reject_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.cam_norm(torch.zeros(1, 21, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.pcm_forward(torch.zeros(1, 21, 4, 4), torch.zeros(1, 192, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.seed_labels(torch.zeros(1, 21, 4, 4), torch.ones(1, 21), 0.2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wseg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
