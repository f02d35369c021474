#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric   : "contrast_hot_path_mpix_per_s" — stride-8 feature-map pixels of view 1 (N*56*56 per step) pushed through
           the whole Step-1 contrast hot path per second: K1a CAM norm -> K2 PCM fwd -> K1c seeds -> K3 prototypes
           -> K4 NCE losses fwd+bwd -> K1b full-res norm fwd+bwd -> K2 PCM bwd, both views.
workload : configs[1] — batch 8 per GPU @448x448 (56x56 features, K=192, D=128, view 2 = 128x128 -> 16x16 grid),
           head-level synthetic inputs (no backbone), seeded.  N GPUs = N x 8 images (weak scaling), prototypes
           whole-batch via one NCCL all-gather per view.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "contrast_hot_path_mpix_per_s"
UNIT = "Mpix/s"
N_PER_GPU = 8
VIEW1 = (448, 448)
VIEW2 = (128, 128)
K_PCM, C_CLS, D_PROJ = 192, 21, 128
BG_THRESHOLD = 0.20


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ workload
def host_inputs(n: int, seed: int, view1=VIEW1, view2=VIEW2):
    """Head-level synthetic inputs on the host (pinned), per SURVEY.md 8d."""
    from wseg_b200 import synth
    g = synth.gen(seed)
    out = {}
    for v, (H, W) in (("1", view1), ("2", view2)):
        h, w = H // 8, W // 8
        out["cam" + v] = 0.05 * torch.randn(n, C_CLS, h, w, generator=g)
        out["f" + v] = torch.randn(n, K_PCM, h, w, generator=g)
        out["fproj" + v] = F.relu(0.02 * torch.randn(n, D_PROJ, h, w, generator=g) + 0.01)
        out["g_rv_down" + v] = 1e-3 * torch.randn(n, C_CLS, h, w, generator=g)
    for name in ("g_n1", "g_rv1", "g_n2", "g_rv2"):
        out[name] = 1e-3 * torch.randn(n, C_CLS, view2[0], view2[1], generator=g)
    out["label"] = synth.label_with_bg(synth.labels(n, seed + 1))
    return {k: v.contiguous().pin_memory() if torch.cuda.is_available() else v.contiguous() for k, v in out.items()}


class HotPath:
    """One step of the hot path on resident inputs, straight through the C-ABI operators (wseg_b200.ops)."""

    def __init__(self, n: int, dev, seed: int, group=None, view1=VIEW1, view2=VIEW2):
        from wseg_b200 import ops
        self.ops, self.n, self.dev, self.group = ops, n, dev, group
        self.view1, self.view2 = view1, view2
        VIEW1, VIEW2 = view1, view2
        self.host = host_inputs(n, seed, view1, view2)
        self.d = {k: v.to(dev) for k, v in self.host.items()}
        d = self.d
        # full-resolution cam / cam_rv as Net.forward emits them (bilinear x8 of the stride-8 maps), resident
        with torch.no_grad():
            d["cam1_full"] = F.interpolate(d["cam1"], VIEW1, mode="bilinear", align_corners=True)
            d["cam2_full"] = F.interpolate(d["cam2"], VIEW2, mode="bilinear", align_corners=True)
            camn = ops.cam_norm(d["cam1"])
            rv, _ = ops.pcm_forward(camn, d["f1"])
            d["rv1_full"] = F.interpolate(rv, VIEW1, mode="bilinear", align_corners=True)
            camn = ops.cam_norm(d["cam2"])
            rv, _ = ops.pcm_forward(camn, d["f2"])
            d["rv2_full"] = F.interpolate(rv, VIEW2, mode="bilinear", align_corners=True)
        self.grid = (VIEW2[0] // 8, VIEW2[1] // 8)
        self.key1 = ops.hash_keys(n * self.grid[0] * self.grid[1], 2 * seed + 1, dev)
        self.key2 = ops.hash_keys(n * self.grid[0] * self.grid[1], 2 * seed + 2, dev)
        self.pixels = n * (VIEW1[0] // 8) * (VIEW1[1] // 8)

    def step(self):
        """One pass.  Independent chains run on forked streams (captured as parallel branches of the CUDA graph) with
        the dependencies a training step has: view-1 chain K1a -> K2 fwd -> K1b(cam_rv) fwd/bwd -> K2 bwd, the same for
        view 2, the seed/prototype/NCE chain after both K2 forwards, and K1b on the two `cam` maps (which do not depend
        on the PCM).  Every intermediate is returned so nothing is recycled before the join."""
        ops, d = self.ops, self.d
        VIEW1, VIEW2 = self.view1, self.view2
        if not hasattr(self, "_streams"):
            prio = int(os.environ.get("WSEG_BENCH_PRIO", "3"))
            lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
            # the view-1 chain (K1a -> K2 fwd -> K1b -> K2 bwd) is the critical path at one GPU: high priority (prio >= 1).
            # The loss chain (s3: a dozen small latency-bound kernels) ends the step once the prototypes travel between ranks:
            # served first as well (prio >= 2), and so is what gates it — view 2's short forward + seeds + prototypes, on
            # their own stream s6 ahead of everything (prio >= 3) instead of queueing behind the K1b kernels of s2 / s4
            p1 = -1 if prio >= 1 else 0
            p3 = -1 if prio >= 2 else 0
            if prio >= 2 and self.group is not None:
                p3 = max(hi, -2)     # across ranks the loss chain ENDS the step (A/B at 2 GPUs: 0.3195 -> 0.3117 ms; neutral at 1)
            p6 = max(hi, -2) if prio >= 3 else 0
            p1, p3, p6 = (int(os.environ.get(k, v)) for k, v in (("WSEG_BENCH_P1", p1), ("WSEG_BENCH_P3", p3), ("WSEG_BENCH_P6", p6)))
            p2 = int(os.environ.get("WSEG_BENCH_P2", "0"))
            p4 = int(os.environ.get("WSEG_BENCH_P4", "0"))
            p5 = int(os.environ.get("WSEG_BENCH_P5", "0"))
            self._streams = [torch.cuda.Stream(device=self.dev, priority=p1), torch.cuda.Stream(device=self.dev, priority=p2),
                             torch.cuda.Stream(device=self.dev, priority=p3), torch.cuda.Stream(device=self.dev, priority=p4),
                             torch.cuda.Stream(device=self.dev, priority=p5), torch.cuda.Stream(device=self.dev, priority=p6)]
        s1, s2, s3, s4, s5, s6 = self._streams
        outer = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(outer)
        for s in (s1, s2, s3, s4, s5, s6):
            s.wait_event(fork)
        keep = []
        cur = s1
        torch.cuda.set_stream(s1)
        marks = getattr(self, "marks", None)          # tools/step_timeline.py: %globaltimer stamps at the milestones of the step

        def mark(name, stream):
            if marks is not None:
                from wseg_b200 import _lib
                idx = marks["names"].setdefault(name, len(marks["names"]))
                _lib.lib().wsegk_debug_stamp(marks["buf"].data_ptr() + 8 * idx, stream.cuda_stream)
        mark("start", s1)

        def k1b(name, size, gname):
            y, saved = ops.fullres_norm_fwd(d[name], d["label"], size)
            g = ops.fullres_norm_bwd(d[name], y, saved, d[gname])
            keep.extend([y, saved, g])
            return g

        # ---- view 1: K1a + K2 forward
        # K1a fused into the PCM's operand preparation.  In precise mode the sign fix-up of the forward (the fp32
        # re-evaluation of the near-zero entries of S, only the backward needs it) runs on its own stream beside K1b
        defer = ops.get_pcm_mode() == "precise" and os.environ.get("WSEG_BENCH_DEFER_FIXUP", "1") == "1"
        rv1, ws1, camn1 = ops.pcm_forward_logits(d["cam1"], d["f1"], defer_fixup=defer)
        mark("fwd1 done", s1)
        f1_done = torch.cuda.Event()
        f1_done.record(cur)
        fix1_done = torch.cuda.Event()
        with torch.cuda.stream(s5):
            if defer:
                s5.wait_event(f1_done)
                ops.pcm_sign_fixup(ws1, self.n, C_CLS, K_PCM, VIEW1[0] // 8, VIEW1[1] // 8)
            mark("fixup1 done", s5)
            fix1_done.record(s5)
        # ---- view 2 chain: K1a, K2 forward, its seeds and prototypes, then K1b(cam_rv2) and K2 backward
        with torch.cuda.stream(s6):
            rv2, ws2, camn2 = ops.pcm_forward_logits(d["cam2"], d["f2"])
            mark("fwd2 done", s6)
            f2_done = torch.cuda.Event()
            f2_done.record(s6)
            norm2, pl2 = ops.seed_labels(rv2, d["label"], BG_THRESHOLD)
            proto2, ti2 = ops.prototypes(norm2, d["fproj2"], group=self.group, view=1)
            mark("proto2 done", s6)
            v2_done = torch.cuda.Event()
            v2_done.record(s6)
        with torch.cuda.stream(s2):
            # K1b(cam_rv2) and view 2's backward need view 2's forward only (seeds and prototypes carry no gradient)
            s2.wait_event(v2_done if os.environ.get("WSEG_BENCH_S2_AFTER", "fwd2") == "proto2" else f2_done)
            g_rv2 = k1b("rv2_full", None, "g_rv2")
            mark("K1b rv2 done", s2)
            gf2 = ops.pcm_backward(d["f2"], camn2, rv2, d["g_rv_down2"], ws2)
            mark("bwd2 done", s2)
            s2_done = torch.cuda.Event()
            s2_done.record(s2)
        # ---- loss chain: shrink view 1 to the loss grid (contrast_train.py:179-180), K1c, K3 (whole batch across
        # ranks), then K4 forward+backward once view 2's prototypes are there
        with torch.cuda.stream(s3):
            s3.wait_event(f1_done)
            f1 = ops.resize_fwd(d["fproj1"], self.grid)
            c1 = ops.resize_fwd(rv1, self.grid)
            norm1, pl1 = ops.seed_labels(c1, d["label"], BG_THRESHOLD)
            mark("seeds1 done", s3)
            proto1, ti1 = ops.prototypes(norm1, f1, group=self.group, view=0)
            mark("proto1 done", s3)
            s3.wait_event(v2_done)
            losses, g1, g2, _, _ = ops.contrast_loss_raw(f1, d["fproj2"], proto1, proto2, pl1, pl2, self.key1, self.key2)
            keep.extend([f1, c1, norm1, pl1, norm2, pl2, proto1, ti1, proto2, ti2])
            mark("K4 done", s3)
            s3_done = torch.cuda.Event()
            s3_done.record(s3)
        # ---- view 1: K1b on cam_rv (the K2 backward follows below)
        g_rv1 = k1b("rv1_full", VIEW2, "g_rv1")
        mark("K1b rv1 done", s1)
        rv1_k1b_done = torch.cuda.Event()
        rv1_k1b_done.record(s1)
        # ---- K1b on the plain CAMs (independent of the PCM)
        # K1b(cam2) waits until the critical chain's K1b(cam_rv1) is through: both are bound by their dense gradient stores, and
        # cam2's has 100 us of slack (A/B on one box, ms per step: plain 0.2995, gap 0.2960; all of stream 4 behind the event
        # 0.329; only cam1's backward behind it 0.316 — profiles/README.md).  WSEG_BENCH_S4=plain switches it off.
        s4_mode = os.environ.get("WSEG_BENCH_S4", "gap")
        with torch.cuda.stream(s4):
            if os.environ.get("WSEG_BENCH_S4_START", "0") == "fwd2":     # A/B: keep the first 40 us free for the forwards' preparation
                s4.wait_event(f2_done)
            g_n1 = k1b("cam1_full", VIEW2, "g_n1")
            mark("K1b cam1 done", s4)
            if s4_mode == "gap":
                s4.wait_event(rv1_k1b_done)
            g_n2 = k1b("cam2_full", None, "g_n2")
            mark("K1b cam2 done", s4)
            s4_done = torch.cuda.Event()
            s4_done.record(s4)
        s1.wait_event(fix1_done)
        gf1 = ops.pcm_backward(d["f1"], camn1, rv1, d["g_rv_down1"], ws1)
        mark("bwd1 done", s1)
        s1_done = torch.cuda.Event()
        s1_done.record(s1)
        torch.cuda.set_stream(outer)
        for e in (s1_done, s2_done, s3_done, s4_done):
            outer.wait_event(e)
        return losses, gf1, gf2, g1, g2, (g_n1, g_rv1, g_n2, g_rv2), keep, (camn1, camn2, rv1, rv2, ws1, ws2)


class GraphedStep:
    """The resident step captured once into a CUDA graph and replayed (the C ABI only enqueues on the caller's stream
    and never synchronises, so the whole path — NCCL all-gather included — is capturable)."""

    def __init__(self, hp: HotPath):
        self.hp = hp
        side = torch.cuda.Stream(device=hp.dev)
        side.wait_stream(torch.cuda.current_stream(hp.dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                hp.step()
        torch.cuda.current_stream(hp.dev).wait_stream(side)
        torch.cuda.synchronize(hp.dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = hp.ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = hp.step()
        self.launches = hp.ops.launch_count() - l0

    def step(self):
        self.graph.replay()
        return self.out


class E2EPath:
    """The same path through the public, reference-facing API (autograd operators + loss.contrast_terms), from PINNED
    HOST buffers: every step copies a step's inputs host->device, runs Net.forward's tail (K1a, K2, bilinear x8 to full
    resolution), the loss-side operators and their backward, and reads the losses back to the host.

    Input pipeline (round 2): the device holds TWO sets of input buffers.  Step i computes on the set that was filled while
    step i-1 ran, and meanwhile its copy stream brings the next step's inputs into the other set — what a training loop
    with a prefetching loader does — so a step costs max(copies, compute) instead of their (partly overlapped) sum.  One
    H2D copy of all inputs (37 MB) and one D2H read of the losses happen inside every timed step; `prefetch=False` is the
    round-1 arrangement (a step waits for its own copies, operator by operator).  The whole step (copies, public API
    calls with autograd, read-back) is captured in CUDA graphs (one per buffer set) and replayed; `--no-graph` runs the
    same code eagerly."""

    # the path's inputs travel host -> device every step, in this order; the synthetic upstream gradients standing in
    # for the rest of the loss (g_*) stay resident — they would be produced on the device
    NAMES = ["label", "cam1", "cam2", "f2", "fproj2", "f1", "fproj1"]

    def __init__(self, hp: HotPath, use_graph: bool, lazy: bool = False, prefetch: bool = True):
        self.hp = hp
        self.lazy = lazy          # True: ops.fullres_norm_up on the small maps instead of x8 enlargement + ops.fullres_norm
        self.prefetch = prefetch
        self.h2d_bytes = sum(hp.host[k].numel() * hp.host[k].element_size() for k in self.NAMES)
        self.out_host = torch.empty(10, dtype=torch.float32).pin_memory()
        self.d2h_bytes = 40
        self.sets = [{k: torch.empty_like(hp.d[k]) for k in self.NAMES} for _ in range(2 if prefetch else 1)]
        self.static = self.sets[0]
        self.copy_stream = torch.cuda.Stream(device=hp.dev)
        self.side_streams = [torch.cuda.Stream(device=hp.dev), torch.cuda.Stream(device=hp.dev)]
        self.graphs = None
        self.results = None
        self.expect = None
        self.mode = "eager"
        self.turn = 0
        if prefetch:                                   # the first step's inputs: filled before anything is timed
            for k in self.NAMES:
                self.sets[0][k].copy_(hp.host[k], non_blocking=True)
            torch.cuda.synchronize(hp.dev)
        if use_graph:
            try:
                side = torch.cuda.Stream(device=hp.dev)
                side.wait_stream(torch.cuda.current_stream(hp.dev))
                with torch.cuda.stream(side):
                    for i in range(4):
                        eager = self._whole_step(i & 1)
                torch.cuda.current_stream(hp.dev).wait_stream(side)
                torch.cuda.synchronize(hp.dev)
                # what the eager, un-captured run of the same code produced: the replayed graphs must reproduce it
                self.expect = (eager[0].detach().cpu().clone(), [g.detach().abs().sum().item() for g in eager[1]])
                graphs, results = [], []
                for cur in range(len(self.sets)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        results.append(self._whole_step(cur))
                    graphs.append(g)
                self.graphs, self.results, self.mode = graphs, results, "cuda_graph"
                self.turn = 0
            except Exception as exc:
                print(f"bench.py: e2e graph capture failed ({type(exc).__name__}: {exc}); running it eagerly", file=sys.stderr)
                self.graphs = None

    @property
    def graph(self):
        return self.graphs

    @graph.setter
    def graph(self, value):
        self.graphs = value

    def _whole_step(self, cur: int = 0):
        hp = self.hp
        main = torch.cuda.current_stream(hp.dev)
        self.copy_stream.wait_stream(main)
        arrived = {}
        if self.prefetch:
            # compute on set `cur` (complete: the previous step joined its copy stream); fill the other set meanwhile
            self.static = self.sets[cur]
            with torch.cuda.stream(self.copy_stream):
                for k in self.NAMES:
                    self.sets[1 - cur][k].copy_(hp.host[k], non_blocking=True)
            need = lambda stream, *names: None
        else:
            self.static = self.sets[0]
            with torch.cuda.stream(self.copy_stream):
                for k in self.NAMES:
                    self.static[k].copy_(hp.host[k], non_blocking=True)
                    arrived[k] = torch.cuda.Event()
                    arrived[k].record(self.copy_stream)
            need = lambda stream, *names: [stream.wait_event(arrived[k]) for k in names]
        losses, grads = self._device_part(need)
        main.wait_stream(self.copy_stream)
        for side in self.side_streams:
            main.wait_stream(side)
        self.out_host.copy_(losses, non_blocking=True)
        return losses, grads

    def _device_part(self, need_on):
        from wseg_b200 import loss as L
        from wseg_b200 import ops
        hp = self.hp
        x = {k: v.detach() for k, v in self.static.items()}
        leaves = {k: x[k].requires_grad_(True) for k in ("f1", "f2", "fproj1", "fproj2", "cam1", "cam2")}
        parts = []

        def k1b(t, up, size, gname):
            """K1b on the full-resolution map of the small map t: materialised (Net.forward's tail, :57-59, then
            contrast_train.py:145-152) or lazy (bit-identical values)."""
            if self.lazy:
                y = ops.fullres_norm_up(t, x["label"], up, size)
            else:
                y = ops.fullres_norm(ops.bilinear_resize(t, up), x["label"], size)
            parts.append((y * hp.d[gname]).sum())

        # Three independent chains, each on its own stream (autograd runs every backward node on the stream of its
        # forward): (a) what only needs the cam logits of view 1, (b) view 2, (main) view 1's PCM chain, which has to
        # wait for f of view 1 — the largest tensor on the bus — and then the contrast terms, which need everything.
        main = torch.cuda.current_stream(hp.dev)
        sa, sb = self.side_streams
        sa.wait_stream(main)
        sb.wait_stream(main)
        with torch.cuda.stream(sa):
            need_on(sa, "label", "cam1")
            k1b(leaves["cam1"], VIEW1, VIEW2, "g_n1")
        with torch.cuda.stream(sb):
            need_on(sb, "label", "cam2", "f2")
            rv2 = ops.pcm_from_logits(x["cam2"], leaves["f2"])
            k1b(leaves["cam2"], VIEW2, None, "g_n2")
            k1b(rv2, VIEW2, None, "g_rv2")
        need_on(main, "label", "cam1", "f1")
        rv1 = ops.pcm_from_logits(x["cam1"], leaves["f1"])
        k1b(rv1, VIEW1, VIEW2, "g_rv1")
        need_on(main, "fproj2", "fproj1")
        main.wait_stream(sb)
        terms = L.contrast_terms(leaves["fproj1"], rv1, leaves["fproj2"], rv2, x["label"], BG_THRESHOLD, grid=hp.grid,
                                 group=hp.group, keys=(hp.key1, hp.key2))
        main.wait_stream(sa)
        total = terms["loss_nce"]
        for t in parts:
            total = total + t
        grads = torch.autograd.grad(total, list(leaves.values()))
        return terms["losses_raw"].detach(), grads

    def step(self):
        cur = self.turn
        self.turn = (cur + 1) % len(self.sets)
        if self.graphs is not None:
            self.graphs[cur].replay()
            self.last = self.results[cur]
            return self.last[0]
        self.last = self._whole_step(cur)
        return self.last[0]

    def verify(self):
        """After the timed region: the losses the last replay left in the pinned host buffer, and the gradients it left
        on the device, against the eager run of the same code (streams, copies and all) before the capture."""
        if self.graphs is None or self.expect is None:
            return None
        torch.cuda.synchronize(self.hp.dev)
        ok = bool(torch.allclose(self.out_host, self.expect[0], rtol=1e-6, atol=0))
        for g, want in zip(self.last[1], self.expect[1]):
            ok = ok and abs(g.detach().abs().sum().item() - want) <= 1e-5 * abs(want)
        return ok


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def time_steps(fn, steps, warmup, dev, flush, barrier):
    """W warm-ups, then K steps each bracketed by CUDA events on the current stream, with an L2 flush (untimed) before
    each; barrier + synchronize on both sides of the region.  Returns (sum of per-step device ms, wall ms)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    return sum(a.elapsed_time(b) for a, b in evs), wall


def cpu_reference_sample(n_sample: int, seed: int, repeats: int):
    """The oracle port of the reference's algorithm on the host cores: the same per-image work as HotPath.step on
    `n_sample` images.  Returns (best seconds, pixels)."""
    from oracle import reference_fn as R
    from wseg_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    x = host_inputs(n_sample, seed)
    x = {k: v.clone() for k, v in x.items()}
    label4 = x["label"].reshape(n_sample, -1, 1, 1)
    sampler = R.ReferenceSampler(seed)

    def one():
        f1 = x["f1"].clone().requires_grad_(True)
        f2 = x["f2"].clone().requires_grad_(True)
        p1 = x["fproj1"].clone().requires_grad_(True)
        p2 = x["fproj2"].clone().requires_grad_(True)
        cam1 = x["cam1"].clone().requires_grad_(True)
        cam2 = x["cam2"].clone().requires_grad_(True)
        rv1 = R.pcm(R.cam_d_norm(cam1), f1)
        rv2 = R.pcm(R.cam_d_norm(cam2), f2)
        cam1_full = F.interpolate(cam1, VIEW1, mode="bilinear", align_corners=True)
        rv1_full = F.interpolate(rv1, VIEW1, mode="bilinear", align_corners=True)
        cam2_full = F.interpolate(cam2, VIEW2, mode="bilinear", align_corners=True)
        rv2_full = F.interpolate(rv2, VIEW2, mode="bilinear", align_corners=True)
        terms = R.contrast_terms(p1, rv1, p2, rv2, label4, BG_THRESHOLD, sampler=sampler,
                                 grid=(VIEW2[0] // 8, VIEW2[1] // 8))
        total = terms["loss_nce"]
        for t, size, gname in ((cam1_full, VIEW2, "g_n1"), (rv1_full, VIEW2, "g_rv1"), (cam2_full, None, "g_n2"),
                               (rv2_full, None, "g_rv2")):
            total = total + (R.fullres_norm(t, label4, size) * x[gname]).sum()
        total.backward()
        return float(total.detach())

    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        one()
        best = min(best, time.perf_counter() - t0)
    return best, n_sample * (VIEW1[0] // 8) * (VIEW1[1] // 8)


def cpu_config0_step(seed: int = 1234):
    """BASELINE.json configs[0]: ONE whole contrast_train.py step on the host cores — random-init ResNet-38, batch 2 @448x448,
    view 2 = 128x128, forward of both views, the whole loss body (contrast_train.py:129-395 as restated by the oracle) and
    backward to the parameters.  The network is the drop-in's own (same modules and parameter names as the reference's) with
    its head evaluated by the oracle's ATen functions, because the wsegk operators have no CPU path.  Returns seconds."""
    from oracle import reference_fn as R
    from wseg_b200 import synth
    from wseg_b200.network import resnet38_contrast
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    net = resnet38_contrast.Net()
    net.train()
    g = synth.gen(seed)
    img1 = torch.randn(2, 3, 448, 448, generator=g)
    img2 = F.interpolate(img1, size=(128, 128), mode="bilinear", align_corners=True)     # contrast_train.py:131-134
    lab20 = synth.labels(2, seed + 1)

    def forward(x):                                                                       # resnet38_contrast.py:31-61 on CPU
        H, W = x.shape[-2:]
        d = net.forward_as_dict(x)
        fea = net.dropout7(d["conv6"])
        f_proj = F.relu(net.fc_proj(fea))
        cam = net.fc8(fea)
        h, w = cam.shape[-2:]
        with torch.no_grad():
            camn = R.cam_d_norm(cam.detach())
        f = torch.cat([F.interpolate(x, (h, w), mode="bilinear", align_corners=True),
                       F.relu(net.f8_3(d["conv4"].detach())), F.relu(net.f8_4(d["conv5"].detach()))], dim=1)
        rv_down = R.pcm(camn, net.f9(f))
        return (F.interpolate(cam, (H, W), mode="bilinear", align_corners=True),
                F.interpolate(rv_down, (H, W), mode="bilinear", align_corners=True), f_proj, rv_down)

    t0 = time.perf_counter()
    out1, out2 = forward(img1), forward(img2)
    losses = R.contrast_step_loss(*out1, *out2, lab20, BG_THRESHOLD, VIEW2, sampler=R.ReferenceSampler(seed))
    losses["loss"].backward()
    return time.perf_counter() - t0


def config_dict(n_gpus):
    return {"workload": f"configs[1]: contrast hot path, batch {N_PER_GPU}/GPU @448x448 (56x56 stride-8 features, K=192, "
                        f"D=128, 21 classes), view 2 128x128 (16x16 loss grid)",
            "global_batch": N_PER_GPU * n_gpus, "parallelism": f"dp{n_gpus}",
            "l2": "256 MiB flush before every timed step",
            "pcm_mode": os.environ.get("WSEGK_PCM_MODE", "default")}


def pcm_roofline(hp, pk, flush, args):
    """Roofline of the dominant kernel: the PCM main kernels on view 1 of `hp`, timed alone with CUDA events on the launching
    stream (inside the library, around the kernel), L2 flushed before every launch -> burst peak."""
    ops, d, n = hp.ops, hp.d, hp.n
    h, w = hp.view1[0] // 8, hp.view1[1] // 8
    camn1 = ops.cam_norm(d["cam1"])

    def measure():
        fwd_ms, bwd_ms, fix_ms = [], [], []
        ops.pcm_profile(True)
        for _ in range(max(3, args.warmup)):
            rv1, ws1 = ops.pcm_forward(camn1, d["f1"])
            ops.pcm_backward(d["f1"], camn1, rv1, d["g_rv_down1"], ws1)
        for _ in range(max(5, args.steps)):
            flush.zero_()
            rv1, ws1 = ops.pcm_forward(camn1, d["f1"])
            ops.pcm_backward(d["f1"], camn1, rv1, d["g_rv_down1"], ws1)
            a, b = ops.pcm_main_ms()
            fwd_ms.append(a)
            bwd_ms.append(b)
            fix_ms.append(ops.pcm_fixup_ms())
        ops.pcm_profile(False)
        return fwd_ms, bwd_ms, fix_ms

    fwd_ms, bwd_ms, fix_ms = measure()            # as shipped: the CTA-pair counts the step runs with (52 / 52 at 104 units)
    ops.set_pcm_pairs(72, 72)                     # and on 72 / 72 pairs: what the kernels reach with the GPU to themselves
    try:
        fwd72, bwd72, _ = measure()
    finally:
        ops.set_pcm_pairs(0, 0)
    hw = h * w
    fl_fwd = 2.0 * n * hw * hw * (K_PCM + C_CLS)
    fl_bwd = 2.0 * n * hw * hw * (2 * K_PCM + C_CLS)
    avg = lambda v: sum(v) / len(v)
    tf_fwd = fl_fwd / (avg(fwd_ms) * 1e-3) / 1e12
    tf_bwd = fl_bwd / (avg(bwd_ms) * 1e-3) / 1e12
    dom = "pcm_bwd_kernel" if avg(bwd_ms) >= avg(fwd_ms) else "pcm_fwd_kernel"
    ach = tf_bwd if dom == "pcm_bwd_kernel" else tf_fwd
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (n, h, w) == (N_PER_GPU, VIEW1[0] // 8, VIEW1[1] // 8):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops"], "traffic": traffic, "peak_source": pk["source"] + " (burst)",
                "algorithmic_flops_per_launch": fl_bwd if dom == "pcm_bwd_kernel" else fl_fwd,
                "ms_per_launch": avg(bwd_ms) if dom == "pcm_bwd_kernel" else avg(fwd_ms),
                "other": {"pcm_fwd_kernel": {"ms": avg(fwd_ms), "tflops": tf_fwd, "frac": tf_fwd / pk["bf16_tflops"]},
                          "pcm_bwd_kernel": {"ms": avg(bwd_ms), "tflops": tf_bwd, "frac": tf_bwd / pk["bf16_tflops"]},
                          "pcm_fwd+bwd": {"ms": avg(fwd_ms) + avg(bwd_ms),
                                          "frac": (fl_fwd + fl_bwd) / ((avg(fwd_ms) + avg(bwd_ms)) * 1e-3) / 1e12 / pk["bf16_tflops"]},
                          "on_72_of_74_sm_pairs": {
                              "note": "the same kernels with the GPU to themselves (ops.set_pcm_pairs(72, 72)); the step runs them on "
                                      "fewer pairs (at most 64 / 52, snapped to a divisor of the unit count: 52 / 52 at 8 x 56x56) "
                                      "because the SMs they leave free run the other chains (bench A/B: step 0.312 vs 0.330 ms)",
                              "pcm_fwd_kernel": {"ms": avg(fwd72), "frac": fl_fwd / (avg(fwd72) * 1e-3) / 1e12 / pk["bf16_tflops"]},
                              "pcm_bwd_kernel": {"ms": avg(bwd72), "frac": fl_bwd / (avg(bwd72) * 1e-3) / 1e12 / pk["bf16_tflops"]}}}}
    if ops.get_pcm_mode() == "precise":
        # what the tensor cores actually execute over the padded tiles: forward = ONE fp16 S product + a
        # 64-column P V product per stage (the near-zero signs are re-evaluated in fp32 by the fix-up kernel, off the
        # tensor cores); backward = a K=144 split-factor product + the K=192 gradient product
        lay = ops.pcm_layout(n, C_CLS, K_PCM, h, w)
        tiles = 2.0 * n * lay["HWpi"] * lay["HWpj"]
        ex_fwd, ex_bwd = tiles * (K_PCM + 64), tiles * (144 + K_PCM)
        for name, ex, ms in (("pcm_fwd_kernel", ex_fwd, avg(fwd_ms)), ("pcm_bwd_kernel", ex_bwd, avg(bwd_ms))):
            roofline["other"][name]["executed_tflops"] = ex / (ms * 1e-3) / 1e12
            roofline["other"][name]["executed_frac"] = ex / (ms * 1e-3) / 1e12 / pk["bf16_tflops"]
        if fix_ms and fix_ms[-1] > 0:
            roofline["other"]["pcm_sign_fixup_kernel"] = {"ms": avg(fix_ms), "note": "fp32 re-evaluation of the near-zero entries of S "
                                                          "(CUDA cores, L2 gathers); off the critical path in the step (own stream)"}
        roofline["note"] = ("frac = ALGORITHMIC flops (SURVEY 8d) / measured bf16 peak; executed_* counts the MMA flops really "
                            "issued over the padded tiles")
    return roofline, fl_fwd, fl_bwd


def step_roofline_of(hp, pk, ms_per_step, fl_fwd1, fl_bwd1):
    """Whole-step roofline: PCM flops of BOTH views / sustained tensor peak + HBM bytes / measured HBM bandwidth, with the
    bytes counted two ways: (a) SURVEY 8d's formula (every one of the N*C full-resolution planes read and written), (b) the
    bytes the step really has to move — K1b reads only the planes of LABELLED classes (the others are multiplied by a zero
    label: their output and gradient are zero whatever they hold) but still writes the dense outputs and gradients."""
    n = hp.n
    V1, V2 = hp.view1, hp.view2
    hw1, hw2 = (V1[0] // 8) * (V1[1] // 8), (V2[0] // 8) * (V2[1] // 8)
    g = hp.grid[0] * hp.grid[1]
    fl = fl_fwd1 + fl_bwd1 + 2.0 * n * hw2 * hw2 * ((K_PCM + C_CLS) + (2 * K_PCM + C_CLS))
    lab = float(hp.d["label"].reshape(n, -1).sum().item()) / (n * C_CLS)      # share of labelled planes (background included)
    px1, px2 = V1[0] * V1[1], V2[0] * V2[1]

    maps = ((px1, px2), (px1, px2), (px2, px2), (px2, px2))                   # cam1, cam_rv1 (full res -> loss size), cam2, cam_rv2
    # (a) SURVEY 8d: N*C*(H*W + out)*4 per tensor forward, the same again backward
    k1b_survey = sum(2.0 * n * C_CLS * 4 * (src + dst) for src, dst in maps)
    # (b) forward reads the labelled input planes and writes the output; backward reads the labelled input planes, grad_out
    # and out, and writes the DENSE input gradient (zeros for the unlabelled planes)
    k1b_required = sum(n * C_CLS * 4 * ((lab * src + dst) + (lab * src + 2 * dst + src)) for src, dst in maps)
    pcm_io = sum(n * hw * (K_PCM + 2 * C_CLS) * 4 + n * hw * (2 * K_PCM + 3 * C_CLS) * 4 for hw in (hw1, hw2))
    small = 2 * (2 * n * C_CLS * g * 4 + 2 * n * g * D_PROJ * 4 * 2) + 2 * n * C_CLS * (hw1 + hw2) * 4
    t_tensor = fl / (pk["bf16_tflops_sustained"] * 1e12)
    out = {}
    for name, k1b in (("survey_formula", k1b_survey), ("bytes_required", k1b_required)):
        by = k1b + pcm_io + small
        t = t_tensor + by / (pk["hbm_gbs"] * 1e9)
        out[name] = {"t_roof_ms": t * 1e3, "frac": t * 1e3 / ms_per_step, "hbm_mb": by / 1e6}
    out["t_roof_ms"] = out["survey_formula"]["t_roof_ms"]
    out["frac"] = out["survey_formula"]["frac"]
    out["labelled_plane_share"] = lab
    out["note"] = ("t_roof = PCM algorithmic flops (both views) / sustained bf16 peak + HBM bytes / measured HBM bandwidth; "
                   "survey_formula charges K1b with all N*C planes (SURVEY 8d), bytes_required only with the labelled ones it has to read")
    return out


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = N_PER_GPU                     # the same configuration as the GPU arm: one step = one batch of 8 images
    for _ in range(args.warmup):
        cpu_reference_sample(n_sample, 1234, 1)
    t0 = time.perf_counter()
    pix = 0
    for _ in range(args.steps):
        _, p = cpu_reference_sample(n_sample, 1234, 1)
        pix += p
    dt = time.perf_counter() - t0
    value = pix / dt / 1e6
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle/reference_fn.py (restated reference, ATen fp32 on {cores} host threads), "
                                       f"{n_sample} images (the GPU arm's per-GPU batch) per step", "same_config": True},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.config0:
        secs = cpu_config0_step()
        line["configs[0]"] = {"workload": "one whole contrast_train.py step on CPU: random-init ResNet-38, batch 2 @448x448 + 128x128 view, "
                                          "forward, loss body, backward", "seconds_per_step": secs, "cores": torch.get_num_threads(),
                              "mpix_per_s_of_stride8_features": 2 * 56 * 56 / secs / 1e6, "kind": "port (oracle head + drop-in backbone)"}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()

    from wseg_b200 import ops
    hp = HotPath(N_PER_GPU, dev, seed=1234 + rank, group=group)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pk = peaks()

    # -------- device-resident arm
    hp.step()
    torch.cuda.synchronize(dev)
    l0 = ops.launch_count()
    hp.step()
    launches_per_step = ops.launch_count() - l0
    stepper, mode = hp, "eager"
    if not args.no_graph:
        try:
            stepper = GraphedStep(hp)
            launches_per_step, mode = stepper.launches, "cuda_graph"
        except Exception as exc:          # e.g. a collective that cannot be captured on this stack
            if rank == 0:
                print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); timing eagerly", file=sys.stderr)
            stepper, mode = hp, "eager"
    clk = ClockSampler(local)
    clk.__enter__()                                  # sampled through both timed regions (resident and e2e)
    dev_ms, wall_ms = time_steps(stepper.step, args.steps, args.warmup, dev, flush, barrier)
    # main-kernel durations of the last step (view-2 launches overwrite view-1's: profile view 1 separately below)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    total_pix = hp.pixels * world
    value = total_pix / (ms_per_step * 1e-3) / 1e6

    roofline, fl_fwd, fl_bwd = pcm_roofline(hp, pk, flush, args)
    step_roofline = step_roofline_of(hp, pk, ms_per_step, fl_fwd, fl_bwd)

    # -------- end-to-end arm: host buffers, public API
    clk._stop.set()                                  # pause sampling outside the timed regions
    if clk._t is not None:
        clk._t.join()
    e2e = E2EPath(hp, use_graph=not args.no_graph)
    clk._stop.clear()
    if clk.nv is not None:
        clk._t = threading.Thread(target=clk._run, daemon=True)
        clk._t.start()
    e2e_ms, _ = time_steps(e2e.step, args.steps, args.warmup, dev, flush, barrier)
    clk.__exit__()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_pix / (float(t.item()) / args.steps * 1e-3) / 1e6
    e2e_ok = e2e.verify()
    if e2e_ok is False:
        raise RuntimeError("bench.py: the replayed end-to-end graph does not reproduce the eager run of the same step")
    # the round-1 arrangement for comparison: every step waits for its OWN copies (operator by operator)
    e2e_np = E2EPath(hp, use_graph=not args.no_graph, prefetch=False)
    np_ms, _ = time_steps(e2e_np.step, args.steps, args.warmup, dev, flush, barrier)
    tn = torch.tensor([np_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    np_ok = e2e_np.verify()
    e2e_np.graph = None
    # for transparency: the same step with the library's "remembered enlargement" switched off, i.e. fullres_norm reading the
    # materialised x8 maps and autograd carrying the dense full-resolution gradients back through the enlargement (what the
    # public operators did until the last session of round 2); same losses and gradients, only the route differs
    from wseg_b200 import ops as _ops
    _ops.remember_enlargements(False)
    try:
        e2e_dense = E2EPath(hp, use_graph=not args.no_graph)
        dense_ms, _ = time_steps(e2e_dense.step, args.steps, args.warmup, dev, flush, barrier)
        td = torch.tensor([dense_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dense_ok = e2e_dense.verify()
        dense_same = bool(torch.allclose(e2e_dense.out_host, e2e.out_host, rtol=1e-6, atol=0))
        e2e_dense.graph = None
        del e2e_dense
    finally:
        _ops.remember_enlargements(True)
    # the same end-to-end step with the lazy full-resolution maps of the extended API (Net.forward_small +
    # ops.fullres_norm_up; same values, the four N*21*H*W tensors never exist) — reported beside the drop-in number
    e2e_lazy = E2EPath(hp, use_graph=not args.no_graph, lazy=True)
    lazy_ms, _ = time_steps(e2e_lazy.step, args.steps, args.warmup, dev, flush, barrier)
    tl = torch.tensor([lazy_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    lazy_ok = e2e_lazy.verify()
    if lazy_ok is False:
        raise RuntimeError("bench.py: the replayed lazy end-to-end graph does not reproduce its eager run")
    same = bool(torch.allclose(e2e_lazy.out_host, e2e.out_host, rtol=1e-6, atol=0))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 hi+lo (forward) / bf16 hi+lo (backward factors) tensor-core operands, f32 accumulate (PCM, precise mode); f32 elsewhere", "data": "synthetic",
            "config": dict(config_dict(world), launch=mode), "image_mpix_per_s": value * 64,
            "clocks": clk.summary(), "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e.h2d_bytes,
                    "d2h_bytes_per_step": e2e.d2h_bytes, "ms_per_step": float(t.item()) / args.steps, "launch": e2e.mode,
                    "replay_equals_eager": e2e_ok,
                    "with_dense_fullres_gradients": {"ms_per_step": float(td.item()) / args.steps, "replay_equals_eager": dense_ok,
                                                     "losses_equal_e2e": dense_same,
                                                     "note": "ops.remember_enlargements(False): K1b reads the materialised x8 maps, "
                                                             "the dense gradients travel back through the enlargement"},
                    "unpipelined": {"ms_per_step": float(tn.item()) / args.steps, "replay_equals_eager": np_ok,
                                    "note": "the same step waiting for its own copies (round 1's arrangement)"},
                    "input_pipeline": "double-buffered: the next step's inputs are copied while this step computes (one 37 MB H2D copy and one D2H read inside every timed step)",
                    "api": "pinned host inputs (cam logits, f, f_proj, labels; both views) -> ops.pcm_from_logits (K1a + K2) / ops.bilinear_resize x8 / loss.contrast_terms / ops.fullres_norm through autograd, backward, losses -> host; copies on their own stream"},
            "e2e_lazy": {"value": total_pix / (float(tl.item()) / args.steps * 1e-3) / 1e6, "unit": UNIT,
                         "ms_per_step": float(tl.item()) / args.steps, "h2d_bytes_per_step": e2e_lazy.h2d_bytes,
                         "d2h_bytes_per_step": e2e_lazy.d2h_bytes, "launch": e2e_lazy.mode,
                         "replay_equals_eager": lazy_ok, "losses_equal_e2e": same,
                         "api": "as e2e, with ops.fullres_norm_up on the stride-8 maps (Net.forward_small) instead of the x8 "
                                "enlargement + ops.fullres_norm"},
            "roofline": roofline, "step_roofline": step_roofline, "wall_ms_per_step_incl_flush": wall_ms / args.steps}

    if world == 1 and not args.no_extra:
        # configs[3] of BASELINE.json beside the headline: batch 16 @512x512 (64x64 features), view 2 = 0.5x (256x256 ->
        # 32x32 loss grid, P = 16 384 loss pixels), hard-pixel mining on — resident inputs, CUDA graph, same timing rules
        del stepper, e2e, e2e_lazy
        torch.cuda.empty_cache()
        try:
            hp3 = HotPath(16, dev, seed=4321, view1=(512, 512), view2=(256, 256))
            st3 = GraphedStep(hp3) if not args.no_graph else hp3
            ms3, _ = time_steps(st3.step, args.steps, args.warmup, dev, flush, barrier)
            ms3 /= args.steps
            rf3, f3, b3 = pcm_roofline(hp3, pk, flush, args)
            line["extra_configs"] = {"configs[3]": {
                "workload": "batch 16 @512x512 (64x64 stride-8 features, K=192, D=128), view 2 256x256 (0.5x) -> 32x32 loss grid, mining on",
                "value": hp3.pixels / (ms3 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms3,
                "gpu_launches_per_step": int(getattr(st3, "launches", 0)), "roofline": rf3,
                "step_roofline": step_roofline_of(hp3, pk, ms3, f3, b3)}}
            del st3, hp3
        except Exception as exc:
            line["extra_configs"] = {"configs[3]": {"error": f"{type(exc).__name__}: {exc}"}}
        # the inference side of the path (contrast_infer.py:46-80): multi-scale + flip CAMs of one image
        try:
            torch.cuda.empty_cache()
            line["extra_configs"]["msf_inference"] = msf_inference(dev, flush)
        except Exception as exc:
            line["extra_configs"]["msf_inference"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        best, pix = cpu_reference_sample(N_PER_GPU, 1234, 2)
        line["cpu_baseline"] = {"value": pix / best / 1e6, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "same_config": True,
                                "sample": f"oracle/reference_fn.py on host threads, one step of the workload = {N_PER_GPU} images "
                                          "(PCM fwd+bwd at 56x56, K1b at 448x448, losses), best of 2"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing NCCL down while CUDA graphs that captured its kernels are alive hung at exit (r01, 2 GPUs): drop the
        # graph, drain, agree that everyone is done, and leave without the communicator destructor.
        del stepper
        e2e.graph = None
        e2e_lazy.graph = None
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def msf_inference(dev, flush):
    """contrast_infer.py's per-image work on one GPU: a 320x480 image at scales 0.5 / 1 / 1.5 / 2 x {image, flip} through a
    random-init ResNet-38 + head (eval), as the script's per-variant loop and as infer.msf_cam does it; and the PCM of the 8
    variants alone, per scale and as ONE launch (ops.pcm_infer_multi)."""
    from wseg_b200 import infer, ops, synth
    from wseg_b200.network import resnet38_contrast
    H, W = 320, 480
    torch.manual_seed(0)
    net = resnet38_contrast.Net().to(dev).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.randn(3, H, W, generator=g)
    scales = [F.interpolate(img[None], scale_factor=s, mode="bilinear", align_corners=False)[0].to(dev) for s in (0.5, 1.0, 1.5, 2.0)]
    label20 = torch.zeros(20)
    label20[[1, 6, 14]] = 1
    lab = label20.to(dev).view(20, 1, 1)

    def script_loop():
        total = None
        with torch.no_grad():
            for im in scales:
                for flip in (False, True):
                    x = im[None].flip(-1) if flip else im[None]
                    cam = F.interpolate(net(x)[1][:, 1:], (H, W), mode="bilinear", align_corners=False)[0] * lab
                    cam = cam.flip(-1) if flip else cam
                    total = cam if total is None else total + cam
        return total

    def timed(fn, reps, with_flush=False):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            if with_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    t_loop = timed(script_loop, 5)
    t_ours = timed(lambda: infer.msf_cam(net, scales, label20, (H, W)), 5)
    del net
    shapes = [(2, 20, 30), (2, 40, 60), (2, 60, 90), (2, 80, 120)]
    cams, fs = [], []
    for i, (n, h, w) in enumerate(shapes):
        d = synth.pcm_inputs(n, h, w, seed=i)
        cams.append(ops.cam_norm(d["cam"].to(dev)))
        fs.append(d["f"].to(dev))
    t_seq = timed(lambda: [ops.pcm_forward(c, f, inference=True)[0] for c, f in zip(cams, fs)], 9, True)
    t_one = timed(lambda: ops.pcm_infer_multi(cams, fs), 9, True)
    flops = sum(2.0 * n * (h * w) ** 2 * (K_PCM + C_CLS) for n, h, w in shapes)
    return {"workload": "one 320x480 image, scales 0.5/1/1.5/2 x flip (8 variants), random-init ResNet-38 + head, eval",
            "ms_per_image_script_loop": t_loop, "ms_per_image": t_ours,
            "note": "ms_per_image: infer.msf_cam (image + flip as one batch, one stream per scale, the PCM of all scales in one launch); "
                    "script_loop: one forward per variant as contrast_infer.py:58-71, same network and kernels",
            "pcm_of_the_8_variants": {"per_scale_launches_us": t_seq * 1e3, "one_launch_us": t_one * 1e3,
                                      "algorithmic_tflops_one_launch": flops / (t_one * 1e-3) / 1e12,
                                      "grids": "20x30, 40x60, 60x90, 80x120 (batch 2 each)", "timing": "eager, L2 flushed, median of 9"}}


def _watchdog(seconds: float):
    """Hard deadline: a wedged collective must not hold the GPU box."""
    def fire():
        sys.stderr.write(f"bench.py: watchdog fired after {seconds:.0f} s\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def main():
    _watchdog(float(os.environ.get("WSEG_BENCH_DEADLINE_S", "900")))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] line")
    ap.add_argument("--config0", action="store_true", help="--impl reference: also time ONE whole configs[0] step on the CPU")
    ap.add_argument("--no-graph", action="store_true", help="time the resident step eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
