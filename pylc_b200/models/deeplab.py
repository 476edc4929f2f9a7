"""
DeepLabv3+ with an output-stride-16 ResNet-101 encoder -- the only network the PyLC reference can
actually run (reference models/architectures/deeplab.py:17-39,
models/backbone/resnet.py:45-157, models/modules/aspp.py:44-96, models/decoder.py:14-54).

The convolutions stay stock PyTorch / cuDNN tensor-core ops (north_star: "the only dense
contraction").  Parameter and buffer names match the reference module tree, so a PyLC model file's
state_dict loads here unchanged and vice versa:

    backbone.{conv1,bn1,layer1..4.N.{conv1,bn1,conv2,bn2,conv3,bn3,downsample.{0,1}}}
    aspp.{aspp1..4.{atrous_conv,bn},global_avg_pool.{1,2},conv1,bn1}
    decoder.{conv1,bn1,last_conv.{0,1,4,5,8}}
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_RESNET101_BLOCKS = (3, 4, 23)          # layer1..3; layer4 is the 3-block multi-grid unit
_MULTI_GRID = (1, 2, 4)


def _conv(cin, cout, k, stride=1, dilation=1, bias=False):
    pad = dilation * (k // 2)
    return nn.Conv2d(cin, cout, k, stride=stride, padding=pad, dilation=dilation, bias=bias)


class Bottleneck(nn.Module):
    """1x1 -> 3x3 (stride / dilation) -> 1x1 x4 residual block (reference resnet.py:14-58)."""

    def __init__(self, cin, width, stride, dilation, project, norm):
        super().__init__()
        self.conv1 = _conv(cin, width, 1)
        self.bn1 = norm(width)
        self.conv2 = _conv(width, width, 3, stride=stride, dilation=dilation)
        self.bn2 = norm(width)
        self.conv3 = _conv(width, width * 4, 1)
        self.bn3 = norm(width * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if project:
            self.downsample = nn.Sequential(_conv(cin, width * 4, 1, stride=stride), norm(width * 4))

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        y += skip
        return self.relu(y)


class ResNet101(nn.Module):
    """Encoder: returns (stride-16 features [B,2048,h/16,w/16], stride-4 features [B,256,h/4,w/4])."""

    def __init__(self, norm=nn.BatchNorm2d, output_stride=16):
        super().__init__()
        if output_stride == 16:
            strides, dils = (1, 2, 2, 1), (1, 1, 1, 2)
        elif output_stride == 8:
            strides, dils = (1, 2, 1, 1), (1, 1, 2, 4)
        else:
            raise NotImplementedError("output_stride must be 8 or 16")
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = norm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        stages = []
        for i, width in enumerate((64, 128, 256, 512)):
            if i < 3:
                dil_list = [dils[i]] * _RESNET101_BLOCKS[i]
            else:
                dil_list = [g * dils[i] for g in _MULTI_GRID]
            blocks = []
            for j, d in enumerate(dil_list):
                first = j == 0
                blocks.append(Bottleneck(cin, width, strides[i] if first else 1, d,
                                         first and (strides[i] != 1 or cin != width * 4), norm))
                cin = width * 4
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        for m in self.modules():   # reference resnet.py:139-150
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        low = self.layer1(x)
        return self.layer4(self.layer3(self.layer2(low))), low


class _AsppBranch(nn.Module):
    def __init__(self, cin, cout, k, dilation, norm):
        super().__init__()
        self.atrous_conv = _conv(cin, cout, k, dilation=dilation)
        self.bn = norm(cout)
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(self.bn(self.atrous_conv(x)))


class ASPP(nn.Module):
    def __init__(self, norm=nn.BatchNorm2d, output_stride=16, cin=2048):
        super().__init__()
        d = (1, 6, 12, 18) if output_stride == 16 else (1, 12, 24, 36)
        self.aspp1 = _AsppBranch(cin, 256, 1, d[0], norm)
        self.aspp2 = _AsppBranch(cin, 256, 3, d[1], norm)
        self.aspp3 = _AsppBranch(cin, 256, 3, d[2], norm)
        self.aspp4 = _AsppBranch(cin, 256, 3, d[3], norm)
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), _conv(cin, 256, 1), norm(256), nn.ReLU())
        self.conv1 = _conv(5 * 256, 256, 1)
        self.bn1 = norm(256)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)

    def forward(self, x):
        pooled = self.global_avg_pool(x)
        pooled = F.interpolate(pooled, size=x.shape[2:], mode='bilinear', align_corners=True)
        x = torch.cat((self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x), pooled), dim=1)
        return self.dropout(self.relu(self.bn1(self.conv1(x))))


class Decoder(nn.Module):
    def __init__(self, n_classes, norm=nn.BatchNorm2d, low_ch=256):
        super().__init__()
        self.conv1 = _conv(low_ch, 48, 1)
        self.bn1 = norm(48)
        self.relu = nn.ReLU()
        self.last_conv = nn.Sequential(
            _conv(256 + 48, 256, 3), norm(256), nn.ReLU(), nn.Dropout(0.5),
            _conv(256, 256, 3), norm(256), nn.ReLU(), nn.Dropout(0.1),
            nn.Conv2d(256, n_classes, 1))

    def forward(self, x, low):
        low = self.relu(self.bn1(self.conv1(low)))
        x = F.interpolate(x, size=low.shape[2:], mode='bilinear', align_corners=True)
        return self.last_conv(torch.cat((x, low), dim=1))


def _kaiming(module):
    for m in module.modules():   # reference aspp.py:84-96, decoder.py:52-64
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)


class DeepLab(nn.Module):
    """logits[B, n_classes, H, W] = upsample_x4(decoder(aspp(encoder(x)), low-level features))."""

    def __init__(self, n_classes=9, normalizer=nn.BatchNorm2d, output_stride=16, in_channels=3, **_ignored):
        super().__init__()
        self.backbone = ResNet101(normalizer, output_stride)
        self.aspp = ASPP(normalizer, output_stride)
        self.decoder = Decoder(n_classes, normalizer)
        self.in_channels = in_channels
        _kaiming(self.aspp)
        _kaiming(self.decoder)

    def features(self, x):
        """Decoder output at stride 4, before the final bilinear x4 (deeplab.py:36-37)."""
        deep, low = self.backbone(x)
        return self.decoder(self.aspp(deep), low)

    def forward(self, x):
        y = self.features(x)
        return F.interpolate(y, size=x.shape[2:], mode='bilinear', align_corners=True)
