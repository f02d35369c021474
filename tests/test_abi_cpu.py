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
    assert lay["bytes"] == pcm and lay["fh32"] - lay["fa"] == 8 * 26 * 128 * 192 * 2
    assert lay["fb"] - lay["fh32"] == 8 * 3328 * 192 * 4 and lay["camb"] - lay["nzmask"] == 8 * 28 * 3328 * 16
    lay = ops.pcm_layout(1, 21, 192, 64, 64)
    assert (lay["wb"], lay["NS"], lay["HWpi"]) == (64, 32, 4096)
    assert _lib.workspace_bytes(_lib.OP_CONTRAST_LOSS, N=8, C=21, D=128, h=16, w=16) > 2 * 2048 * 4 * 42
    assert _lib.workspace_bytes(99, N=1, C=1, K=16, h=1, w=1) == 0
    assert "unknown op" in _lib.last_error()


def test_multi_problem_forward_workspace_sizes(lib):
    """wsegk_pcm_multi_workspace_bytes (host only): the forward-only PCM of several grids in one launch splits the grid's CTA
    pairs between the problems and sizes every problem's workspace for its share; bad argument lists are refused."""
    def sizes(shapes):
        n = len(shapes)
        arr = lambda k: (ctypes.c_int * n)(*[s[k] for s in shapes])
        out = (ctypes.c_size_t * max(n, 1))()
        rc = lib.wsegk_pcm_multi_workspace_bytes(n, arr(0), arr(1), arr(2), 21, 192, out)
        return rc, [int(v) for v in out[:n]]
    rc, four = sizes([(2, 20, 30), (2, 40, 60), (2, 60, 90), (2, 80, 120)])     # the MSF variants of a 320x480 image
    assert rc == 0 and all(v % 1024 == 0 and v > 0 for v in four)
    assert four == sorted(four)                                                  # bigger grid, bigger workspace
    rc, one = sizes([(2, 80, 120)])
    assert rc == 0 and one[0] > 0
    rc, _ = sizes([(1, 16, 16)] * 9)                                             # at most 8 problems
    assert rc != 0 and "n_items" in _lib.last_error()
    rc, _ = sizes([(0, 16, 16)])
    assert rc != 0


def _check_schedule(lib, n, h, w):
    """The persistent PCM kernels' schedule (rows + stream-K tail, csrc/common.cuh) as a host-side property: every stage
    of every (image, j pair) unit is run by exactly one CTA pair, the pairs' loads differ by at most one stage, a pair's
    segments are runs of consecutive stages of one unit, and the (unit, slot) pairs are distinct with slots below the
    number of partial slots the workspace holds."""
    lay = ops.pcm_layout(n, 21, 192, h, w)
    units, ns, slots = n * lay["PJ"], lay["NS"], lay["nsplit"]
    seen, loads, used = {}, [], set()
    for g in range(80):
        cap = units * ns
        buf = (ctypes.c_int * (4 * cap))()
        cnt = lib.wsegk_debug_pcm_schedule(n, h, w, g, buf, cap)
        assert cnt >= 0
        if cnt == 0:
            continue
        loads.append(cnt)
        prev = None
        for i in range(cnt):
            u, st, k, slot = buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]
            assert 0 <= u < units and 0 <= st < ns and 0 <= slot < slots
            assert (u, st) not in seen, "stage run twice"
            seen[(u, st)] = g
            if prev is not None and prev[2] == k:
                assert prev[0] == u and prev[1] + 1 == st and prev[3] == slot      # a segment: consecutive stages of a unit
            elif prev is not None:
                assert k == prev[2] + 1
            if prev is None or prev[2] != k:
                assert (u, slot) not in used, "two segments write the same partial slot"
                used.add((u, slot))
            prev = (u, st, k, slot)
    assert len(seen) == units * ns
    assert max(loads) - min(loads) <= 1 or len(loads) < 72
    assert sum(loads) == units * ns


@pytest.mark.parametrize("n,h,w", [(8, 56, 56), (1, 16, 16), (2, 20, 13), (16, 56, 56), (3, 9, 5), (128, 32, 32), (5, 96, 96),
                                   (1, 108, 108)])
def test_pcm_schedule_covers_every_stage_once(lib, n, h, w):
    _check_schedule(lib, n, h, w)


def test_pcm_schedule_random_shapes(lib):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 70), st.integers(1, 70))
    def run(n, h, w):
        _check_schedule(lib, n, h, w)

    run()


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
