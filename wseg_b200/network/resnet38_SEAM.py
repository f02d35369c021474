"""Drop-in for the reference's ``network/resnet38_SEAM.py`` (the SEAM baseline the contrast network extends): the same
head as ``resnet38_contrast`` without the projection conv — ``forward(x) -> (cam, cam_rv)`` (``:31-53``), identical
``PCM`` (``:55-67``) — with the CAM normalisation and the pixel correlation module in the wsegk kernels.
CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from . import resnet38d
from .resnet38_contrast import Net as _ContrastNet


class Net(resnet38d.Net):
    def __init__(self):
        super().__init__()
        self.dropout7 = nn.Dropout2d(0.5)
        self.fc8 = nn.Conv2d(4096, 21, 1, bias=False)
        self.f8_3 = nn.Conv2d(512, 64, 1, bias=False)
        self.f8_4 = nn.Conv2d(1024, 128, 1, bias=False)
        self.f9 = nn.Conv2d(192 + 3, 192, 1, bias=False)
        nn.init.xavier_uniform_(self.fc8.weight)
        nn.init.kaiming_normal_(self.f8_3.weight)
        nn.init.kaiming_normal_(self.f8_4.weight)
        nn.init.xavier_uniform_(self.f9.weight, gain=4)
        self.from_scratch_layers = [self.f8_3, self.f8_4, self.f9, self.fc8]
        self.not_training = [self.conv1a, self.b2, self.b2_1, self.b2_2]

    def forward(self, x):
        H, W = x.shape[-2:]
        d = super().forward_as_dict(x)
        cam = self.fc8(self.dropout7(d['conv6']))
        h, w = cam.shape[-2:]
        # the reference runs its whole tail — the final enlargement of `cam` included — under no_grad (:36-52): neither
        # output carries a graph, in train mode too
        with torch.no_grad():
            cam_d_norm = ops.cam_norm(cam.detach())                                        # K1a (:37-43)
            f8_3 = F.relu(self.f8_3(d['conv4'].detach()), inplace=True)
            f8_4 = F.relu(self.f8_4(d['conv5'].detach()), inplace=True)
            x_s = F.interpolate(x, (h, w), mode='bilinear', align_corners=True)
            f = torch.cat([x_s, f8_3, f8_4], dim=1)
            cam_rv = ops.bilinear_resize(self.PCM(cam_d_norm, f), (H, W))                  # K2 (:55-67) + glue (:51)
            cam = ops.bilinear_resize(cam, (H, W))                                         # :52
        return cam, cam_rv

    PCM = _ContrastNet.PCM
    pcm_inputs = _ContrastNet.pcm_inputs
    get_parameter_groups = _ContrastNet.get_parameter_groups
