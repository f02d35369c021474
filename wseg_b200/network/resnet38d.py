"""Wide ResNet-38 backbone (stays PyTorch/cuDNN — out of the optimisation scope).

Written from the architecture description so that ``state_dict()`` keys and shapes equal those of the reference's
``network/resnet38d.py:120-158`` (a ``contrast.pth`` checkpoint loads unchanged); the layer table below is the
whole definition.  ``forward_as_dict`` returns the same four taps (``resnet38d.py:160-189``) and ``train()`` keeps
the reference's freezing rules (``resnet38d.py:192-214``): BatchNorm always in eval mode with frozen affine
parameters, ``not_training`` layers frozen.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class _PreActUnit(nn.Module):
    """BN-ReLU pre-activation residual unit with two 3x3 convs (names follow the checkpoint format)."""

    def __init__(self, cin, cmid, cout, stride=1, first_dilation=None, dilation=1):
        super().__init__()
        first_dilation = dilation if first_dilation is None else first_dilation
        self.same_shape = cin == cout and stride == 1
        self.bn_branch2a = nn.BatchNorm2d(cin)
        self.conv_branch2a = nn.Conv2d(cin, cmid, 3, stride, padding=first_dilation, dilation=first_dilation,
                                       bias=False)
        self.bn_branch2b1 = nn.BatchNorm2d(cmid)
        self.conv_branch2b1 = nn.Conv2d(cmid, cout, 3, padding=dilation, dilation=dilation, bias=False)
        if not self.same_shape:
            self.conv_branch1 = nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x, get_x_bn_relu=False):
        pre = F.relu(self.bn_branch2a(x))
        shortcut = x if self.same_shape else self.conv_branch1(pre)
        y = self.conv_branch2b1(F.relu(self.bn_branch2b1(self.conv_branch2a(pre))))
        y = shortcut + y
        return (y, pre) if get_x_bn_relu else y


class _BottleneckUnit(nn.Module):
    """1x1 -> 3x3 (dilated) -> 1x1 pre-activation bottleneck with channel dropout (b6, b7)."""

    def __init__(self, cin, cout, stride=1, dilation=1, dropout=0.0):
        super().__init__()
        self.same_shape = cin == cout and stride == 1
        self.bn_branch2a = nn.BatchNorm2d(cin)
        self.conv_branch2a = nn.Conv2d(cin, cout // 4, 1, stride, bias=False)
        self.bn_branch2b1 = nn.BatchNorm2d(cout // 4)
        self.dropout_2b1 = nn.Dropout2d(dropout)
        self.conv_branch2b1 = nn.Conv2d(cout // 4, cout // 2, 3, padding=dilation, dilation=dilation, bias=False)
        self.bn_branch2b2 = nn.BatchNorm2d(cout // 2)
        self.dropout_2b2 = nn.Dropout2d(dropout)
        self.conv_branch2b2 = nn.Conv2d(cout // 2, cout, 1, bias=False)
        if not self.same_shape:
            self.conv_branch1 = nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x, get_x_bn_relu=False):
        pre = F.relu(self.bn_branch2a(x))
        shortcut = self.conv_branch1(pre)
        y = self.conv_branch2a(pre)
        y = self.conv_branch2b1(self.dropout_2b1(F.relu(self.bn_branch2b1(y))))
        y = self.conv_branch2b2(self.dropout_2b2(F.relu(self.bn_branch2b2(y))))
        y = shortcut + y
        return (y, pre) if get_x_bn_relu else y


class Normalize:
    """ImageNet mean/std normalisation of an HWC uint8 image (resnet38d.py:104-118)."""

    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        self.mean, self.std = mean, std

    def __call__(self, img):
        arr = np.asarray(img)
        out = np.empty_like(arr, np.float32)
        for ch in range(3):
            out[..., ch] = (arr[..., ch] / 255. - self.mean[ch]) / self.std[ch]
        return out


# (attribute name, constructor, args) in checkpoint order
_STAGES = (
    ("b2", _PreActUnit, dict(cin=64, cmid=128, cout=128, stride=2)),
    ("b2_1", _PreActUnit, dict(cin=128, cmid=128, cout=128)),
    ("b2_2", _PreActUnit, dict(cin=128, cmid=128, cout=128)),
    ("b3", _PreActUnit, dict(cin=128, cmid=256, cout=256, stride=2)),
    ("b3_1", _PreActUnit, dict(cin=256, cmid=256, cout=256)),
    ("b3_2", _PreActUnit, dict(cin=256, cmid=256, cout=256)),
    ("b4", _PreActUnit, dict(cin=256, cmid=512, cout=512, stride=2)),
    ("b4_1", _PreActUnit, dict(cin=512, cmid=512, cout=512)),
    ("b4_2", _PreActUnit, dict(cin=512, cmid=512, cout=512)),
    ("b4_3", _PreActUnit, dict(cin=512, cmid=512, cout=512)),
    ("b4_4", _PreActUnit, dict(cin=512, cmid=512, cout=512)),
    ("b4_5", _PreActUnit, dict(cin=512, cmid=512, cout=512)),
    ("b5", _PreActUnit, dict(cin=512, cmid=512, cout=1024, stride=1, first_dilation=1, dilation=2)),
    ("b5_1", _PreActUnit, dict(cin=1024, cmid=512, cout=1024, dilation=2)),
    ("b5_2", _PreActUnit, dict(cin=1024, cmid=512, cout=1024, dilation=2)),
    ("b6", _BottleneckUnit, dict(cin=1024, cout=2048, stride=1, dilation=4, dropout=0.3)),
    ("b7", _BottleneckUnit, dict(cin=2048, cout=4096, dilation=4, dropout=0.5)),
)
_TAPS = {"b4": "conv3", "b5": "conv4", "b6": "conv5"}      # pre-activation of these units is exported


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1a = nn.Conv2d(3, 64, 3, padding=1, bias=False)
        for name, ctor, kw in _STAGES:
            setattr(self, name, ctor(**kw))
        self.bn7 = nn.BatchNorm2d(4096)
        self.not_training = [self.conv1a]
        self.normalize = Normalize()

    def forward(self, x):
        return self.forward_as_dict(x)['conv6']

    def forward_as_dict(self, x):
        taps = {}
        x = self.conv1a(x)
        for name, _, _ in _STAGES:
            unit = getattr(self, name)
            if name in _TAPS:
                x, taps[_TAPS[name]] = unit(x, get_x_bn_relu=True)
            else:
                x = unit(x)
        taps['conv6'] = F.relu(self.bn7(x))
        return taps

    def train(self, mode=True):
        super().train(mode)
        for layer in self.not_training:
            if isinstance(layer, nn.Conv2d):
                layer.weight.requires_grad = False
            elif isinstance(layer, nn.Module):
                for child in layer.children():
                    child.weight.requires_grad = False
                    if child.bias is not None:
                        child.bias.requires_grad = False
        for layer in self.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()
                layer.bias.requires_grad = False
                layer.weight.requires_grad = False
        return self
