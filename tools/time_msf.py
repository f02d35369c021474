"""Multi-scale + flip inference of one image (contrast_infer.py:46-80): the per-variant loop of the script, the batched
image+flip forward per scale in sequence, and one stream per scale.  usage: python tools/time_msf.py [H W]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from wseg_b200 import infer
from wseg_b200.network import resnet38_contrast

dev = torch.device("cuda:0")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 480)
torch.manual_seed(0)
net = resnet38_contrast.Net().to(dev).eval()
g = torch.Generator().manual_seed(1)
img = torch.randn(3, H, W, generator=g)
scales = [F.interpolate(img[None], scale_factor=s, mode="bilinear", align_corners=False)[0] for s in (0.5, 1.0, 1.5, 2.0)]
label20 = torch.zeros(20)
label20[[1, 6, 14]] = 1


def script_loop():
    total = None
    with torch.no_grad():
        for im in scales:
            for flip in (False, True):
                x = im[None].to(dev)
                x = x.flip(-1) if flip else x
                cam = net(x)[1]
                cam = F.interpolate(cam[:, 1:], (H, W), mode="bilinear", align_corners=False)[0] * label20.to(dev).view(20, 1, 1)
                cam = cam.flip(-1) if flip else cam
                total = cam if total is None else total + cam
    return total


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t0 = timed(script_loop)
t1 = timed(lambda: infer.msf_cam(net, scales, label20, (H, W), concurrent=False, one_pcm_launch=False))
t2 = timed(lambda: infer.msf_cam(net, scales, label20, (H, W), concurrent=True, one_pcm_launch=False))
t3 = timed(lambda: infer.msf_cam(net, scales, label20, (H, W), concurrent=True, one_pcm_launch=True))
print(f"image {H}x{W}, scales 0.5/1/1.5/2 x flip, random-init ResNet-38 (eval): per-variant loop {t0:.1f} ms | image+flip batched per scale, "
      f"sequential {t1:.1f} ms | one stream per scale {t2:.1f} ms | + the PCM of all scales in one launch {t3:.1f} ms")
