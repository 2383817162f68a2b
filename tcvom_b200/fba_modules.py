"""Parameter containers for the ``vmn_fba`` network with the reference's state_dict layout (203 keys).

Like ``modules.py`` for ``vmn_gca``: the classes hold exactly the reference's parameters under exactly its names
and registration order (``pred_test.py:92`` loads checkpoints with ``strict=True``) and compute nothing themselves;
all arithmetic is done by the sm_100a kernels driven from ``tcvom_b200.fba_engine``.

Name/shape layout restated from (reference checkout, commit f5fa07a):
  encoder  models/FBA/resnet_GN_WS.py:41-128 (Bottleneck x [3,4,6,3], GroupNorm(32), weight-standardised convs),
           models/FBA/models.py:38-63 (11-channel stem), :175-236 (ResnetDilated, output stride 8)
  decoder  models/FBA/models.py:258-311 (pyramid pooling, conv_up1..4), models/VMN/VMN_FBA.py:6-10 (fam)
"""
from __future__ import annotations

from torch import nn

from .modules import TAMParams, _Holder, _seq

RES_LAYERS = (("layer1", 64, 3, 1, 1), ("layer2", 128, 4, 2, 1), ("layer3", 256, 6, 2, 2), ("layer4", 512, 3, 2, 4))
PPM_SCALES = (1, 2, 3, 6)
GN_GROUPS = 32


def block_config(i: int, stride: int, dilate: int):
    """(conv2 stride, conv2 dilation, downsample stride) of block i of a layer after
    ResnetDilated._nostride_dilate (FBA/models.py:204-218)."""
    if dilate == 1:
        s = stride if i == 0 else 1
        return s, 1, s
    if i == 0:
        return 1, dilate // 2, 1
    return 1, dilate, 1


def _gn(c):
    return nn.GroupNorm(GN_GROUPS, c)


class Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, has_down):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _gn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _gn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _gn(planes * 4)
        self.relu = _Holder()
        self.downsample = _seq([nn.Conv2d(inplanes, planes * 4, 1, bias=False), _gn(planes * 4)]) if has_down else None


class FBAEncoderParams(nn.Module):
    """ResnetDilated(l_resnet50 GN+WS, dilate_scale=8) with the 11-channel stem."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(11, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _gn(64)
        self.relu = _Holder()
        self.maxpool = _Holder()
        inplanes = 64
        for name, planes, blocks, stride, dilate in RES_LAYERS:
            layer = _Holder()
            for i in range(blocks):
                layer.add_module(str(i), Bottleneck(inplanes, planes, i == 0))
                inplanes = planes * 4
            self.add_module(name, layer)


class FBADecoderParams(nn.Module):
    """vmn_fba_decoder: pyramid pooling + conv_up1 (per frame), TAM + conv_up2..4 + fusion (per centre frame)."""

    def __init__(self, reduction, window, freeze_backbone=False, batch_norm=False):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("tcvom_b200: the BatchNorm variant of the FBA decoder is not on the built path")
        self.batch_norm = batch_norm
        self.ppm = _Holder()
        for i, _ in enumerate(PPM_SCALES):
            self.ppm.add_module(str(i), _seq([None, nn.Conv2d(2048, 256, 1, bias=True), _gn(256), None]))
        self.conv_up1 = _seq([nn.Conv2d(2048 + 4 * 256, 256, 3, padding=1, bias=True), _gn(256), None,
                              nn.Conv2d(256, 256, 3, padding=1), _gn(256), None])
        self.conv_up2 = _seq([nn.Conv2d(512, 256, 3, padding=1, bias=True), _gn(256), None])
        self.conv_up3 = _seq([nn.Conv2d(256 + 64, 64, 3, padding=1, bias=True), _gn(64), None])
        self.unpool = _Holder()
        self.conv_up4 = _seq([nn.Conv2d(64 + 3 + 3 + 2, 32, 3, padding=1, bias=True), None,
                              nn.Conv2d(32, 16, 3, padding=1, bias=True), None,
                              nn.Conv2d(16, 7, 1, bias=True)])
        self.fam = TAMParams(256, reduction, window)
        self.freeze_backbone = freeze_backbone
