"""Drop-in for the reference's ``network/resnet38_contrast.py``: same constructor, parameter names, initialisers,
``forward(x) -> (cam, cam_rv, f_proj, cam_rv_down)`` (``:31-61``), ``PCM(cam, f)`` (``:63-75``) and
``get_parameter_groups()`` (``:77-96``) — but the no-grad CAM normalisation (``:41-48``, K1a) and the pixel
correlation module after the ``f9`` conv (``:70-73``, K2 forward and backward) run in the wsegk sm_100a kernels.
The 1x1 convs, the bilinear resizes and the backbone stay PyTorch/cuDNN.  CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from . import resnet38d


class Net(resnet38d.Net):
    def __init__(self):
        super().__init__()
        self.dropout7 = nn.Dropout2d(0.5)
        self.fc8 = nn.Conv2d(4096, 21, 1, bias=False)
        self.fc_proj = nn.Conv2d(4096, 128, 1, bias=False)
        self.f8_3 = nn.Conv2d(512, 64, 1, bias=False)
        self.f8_4 = nn.Conv2d(1024, 128, 1, bias=False)
        self.f9 = nn.Conv2d(192 + 3, 192, 1, bias=False)
        nn.init.xavier_uniform_(self.fc8.weight)
        nn.init.kaiming_normal_(self.f8_3.weight)
        nn.init.kaiming_normal_(self.f8_4.weight)
        nn.init.xavier_uniform_(self.f9.weight, gain=4)
        nn.init.xavier_uniform_(self.fc_proj.weight)
        self.from_scratch_layers = [self.f8_3, self.f8_4, self.f9, self.fc8, self.fc_proj]
        self.not_training = [self.conv1a, self.b2, self.b2_1, self.b2_2]

    def head(self, x, d, enlarge: bool = True):
        """Everything after the backbone (``:34-61``); split out so benchmarks can feed synthetic backbone taps.
        ``enlarge=False`` skips the two x8 enlargements at the end and returns the small ``cam`` instead (see
        ``forward_small``)."""
        H, W = x.shape[-2:]
        fea = self.dropout7(d['conv6'])
        f_proj = F.relu(self.fc_proj(fea), inplace=True)
        cam = self.fc8(fea)
        h, w = cam.shape[-2:]
        f8_3 = F.relu(self.f8_3(d['conv4'].detach()), inplace=True)
        f8_4 = F.relu(self.f8_4(d['conv5'].detach()), inplace=True)
        x_s = F.interpolate(x, (h, w), mode='bilinear', align_corners=True)
        f = torch.cat([x_s, f8_3, f8_4], dim=1)
        # K1a (:41-48) + K2 (:63-75): self.PCM(cam_d_norm, f) with the normalisation fused into the PCM's operand preparation
        cam_rv_down = ops.pcm_from_logits(cam, self.f9(f))
        if not enlarge:
            return cam, cam_rv_down, f_proj
        cam_rv = ops.bilinear_resize(cam_rv_down, (H, W))                                  # glue (:57-58)
        cam = ops.bilinear_resize(cam, (H, W))                                             # glue (:59)
        return cam, cam_rv, f_proj, cam_rv_down

    def forward(self, x):
        return self.head(x, super().forward_as_dict(x))

    @torch.no_grad()
    def pcm_inputs(self, x):
        """Inference aid: ``(cam_logits, f9(f))`` on the stride-8 grid — the network up to, not including,
        the PCM.  ``infer.msf_cam`` collects these for all scales of an image and runs ONE PCM launch over them
        (``ops.pcm_infer_multi``); ``ops.bilinear_resize(ops.pcm(cam, f), x.shape[-2:])`` of the pair is ``forward(x)[1]``."""
        d = resnet38d.Net.forward_as_dict(self, x)       # (not super(): resnet38_SEAM.Net borrows this method)
        cam = self.fc8(self.dropout7(d['conv6']))
        h, w = cam.shape[-2:]
        f8_3 = F.relu(self.f8_3(d['conv4']), inplace=True)
        f8_4 = F.relu(self.f8_4(d['conv5']), inplace=True)
        x_s = F.interpolate(x, (h, w), mode='bilinear', align_corners=True)
        return cam, self.f9(torch.cat([x_s, f8_3, f8_4], dim=1))

    def forward_small(self, x):
        """Extension: ``(cam_small, cam_rv_down, f_proj)`` — ``forward`` without the two x8 enlargements
        (``:57-59``).  ``loss.contrast_step_loss_lazy`` consumes these and evaluates the enlargements on the fly, so the
        four N*21*H*W tensors of a training step are never written or read."""
        return self.head(x, super().forward_as_dict(x), enlarge=False)

    def PCM(self, cam, f):
        n, c, h, w = f.size()
        if tuple(cam.shape[-2:]) != (h, w):      # identity in Net.forward; kept for callers that pass another size
            cam = F.interpolate(cam, (h, w), mode='bilinear', align_corners=True)
        return ops.pcm(cam, self.f9(f))

    def get_parameter_groups(self):
        groups = ([], [], [], [])
        print('======================================================')
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.modules.normalization.GroupNorm)):
                scratch = any(m is layer for layer in self.from_scratch_layers)
                if m.weight.requires_grad:
                    groups[2 if scratch else 0].append(m.weight)
                if m.bias is not None and m.bias.requires_grad:
                    groups[3 if scratch else 1].append(m.bias)
        return groups
