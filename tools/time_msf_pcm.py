"""The PCM of multi-scale + flip inference (contrast_infer.py:58-71) alone: one 320x480 image at scales 0.5/1/1.5/2 x
{image, flip} = four problems of batch 2 on grids 20x30 ... 80x120.  Per-scale calls in sequence, per-scale calls on four
streams, and all four in ONE launch (ops.pcm_infer_multi).  usage: python tools/time_msf_pcm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wseg_b200 import ops, synth

dev = torch.device("cuda:0")
shapes = [(2, 20, 30), (2, 40, 60), (2, 60, 90), (2, 80, 120)]
cams, fs = [], []
for i, (n, h, w) in enumerate(shapes):
    d = synth.pcm_inputs(n, h, w, seed=i)
    cams.append(ops.cam_norm(d["cam"].to(dev)))
    fs.append(d["f"].to(dev))
streams = [torch.cuda.Stream(device=dev) for _ in shapes]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def seq():
    return [ops.pcm_forward(c, f, inference=True)[0] for c, f in zip(cams, fs)]


def par():
    main = torch.cuda.current_stream(dev)
    outs = []
    for st, c, f in zip(streams, reversed(cams), reversed(fs)):      # largest first
        st.wait_stream(main)
        with torch.cuda.stream(st):
            outs.append(ops.pcm_forward(c, f, inference=True)[0])
    for st in streams:
        main.wait_stream(st)
    return outs


def multi():
    return ops.pcm_infer_multi(cams, fs)


def timed(fn, reps=9):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


# as CUDA graphs too: what the kernels cost without the host's launch overhead
def graphed(fn):
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    return (lambda: g.replay()), keep


print("PCM forward of the 8 MSF variants of a 320x480 image (4 problems of batch 2, grids 20x30 .. 80x120), median us:")
for name, fn in (("per-scale calls, in sequence", seq), ("per-scale calls, one stream each", par), ("ONE launch (pcm_infer_multi)", multi)):
    e = timed(fn)
    rep, keep = graphed(fn)
    print(f"  {name:34s} eager {e:7.1f}   as a CUDA graph {timed(rep):7.1f}")
ops.pcm_profile(True)
multi(); torch.cuda.synchronize()
print(f"  main kernel of the single launch: {ops.pcm_main_ms()[0] * 1e3:.1f} us")
small = [(cams[0], fs[0])]
print(f"  the 0.5x scale alone: single call {timed(lambda: ops.pcm_forward(cams[0], fs[0], inference=True)):.1f} us")
