"""Parameter containers for the ``vmn_gca`` network with the reference's state_dict layout.

The reference checkpoints (``NET.state_dict()``, 584 keys for ``vmn_gca``) are part of the
drop-in contract: ``pred_test.py:92`` loads them with ``strict=True`` and
``train_ddp.py:300-305`` resumes optimizer state by parameter order.  The classes here
hold exactly those parameters/buffers under exactly those names and in the same
registration order; they do not compute anything themselves -- all arithmetic is done by
the sm_100a kernels driven from ``tcvom_b200.engine``.

Name/shape layout restated from (reference checkout, commit f5fa07a):
  encoder  models/GCA/encoders/resnet_enc.py:52-127, res_gca_enc.py:10-55
  decoder  models/GCA/decoders/resnet_dec.py:23-124, models/VMN/VMN_GCA.py:10-16
  TAM      models/VMN/VMN_model.py:9-16
  GCA op   models/GCA/ops.py:83-104
  SN       models/GCA/ops.py:56-72  (weight_u, weight_v, weight_bar under ``.module``)
"""
from __future__ import annotations

import math

import torch
from torch import nn

ENC_LAYERS = (("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 4, 2),
              ("layer_bottleneck", 512, 2, 2))
DEC_LAYERS = (("layer1", 256, 2), ("layer2", 128, 3), ("layer3", 64, 3), ("layer4", 32, 2))


def _xavier_uniform_(t: torch.Tensor) -> None:
    # fan_in/fan_out as torch computes them for conv weights: dim1*rf, dim0*rf
    rf = 1
    for s in t.shape[2:]:
        rf *= s
    bound = math.sqrt(6.0 / (t.shape[1] * rf + t.shape[0] * rf))
    with torch.no_grad():
        t.uniform_(-bound, bound)


class _Holder(nn.Module):
    """Parameterless placeholder keeping Sequential indices aligned with the reference
    (ReLU / ReflectionPad / AvgPool / Upsample slots carry no state)."""

    def forward(self, *a, **k):  # pragma: no cover - never used as a compute module
        raise RuntimeError("tcvom_b200 parameter container: compute goes through the engine")


class _SNParams(nn.Module):
    def __init__(self, shape):
        super().__init__()
        h = shape[0]
        wdt = 1
        for s in shape[1:]:
            wdt *= s
        u = torch.randn(h)
        v = torch.randn(wdt)
        self.weight_u = nn.Parameter(u / (u.norm() + 1e-12), requires_grad=False)
        self.weight_v = nn.Parameter(v / (v.norm() + 1e-12), requires_grad=False)
        w = torch.empty(*shape)
        _xavier_uniform_(w)
        self.weight_bar = nn.Parameter(w)


class SpectralNormConv(nn.Module):
    """Holds ``<name>.module.{weight_u,weight_v,weight_bar}``.  ``shape`` is the torch
    weight shape: [Cout,Cin,k,k] for a conv, [Cin,Cout,4,4] for the stride-2 deconv."""

    def __init__(self, shape, transposed=False, stride=1):
        super().__init__()
        self.module = _SNParams(tuple(shape))
        self.transposed = transposed
        self.stride = stride


def _bn(c, gamma=1.0):
    m = nn.BatchNorm2d(c)
    nn.init.constant_(m.weight, gamma)
    nn.init.constant_(m.bias, 0.0)
    return m


def _seq(items):
    s = _Holder()
    for i, it in enumerate(items):
        s.add_module(str(i), it if it is not None else _Holder())
    return s


class EncBlock(nn.Module):
    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = SpectralNormConv((planes, cin, 3, 3), stride=stride)
        self.bn1 = _bn(planes)
        self.activation = _Holder()
        self.conv2 = SpectralNormConv((planes, planes, 3, 3))
        self.bn2 = _bn(planes, 0.0)                       # zero-init last BN (resnet_enc.py:96-98)
        self.downsample = None
        if stride != 1:
            self.downsample = _seq([None, SpectralNormConv((planes, cin, 1, 1)), _bn(planes)])
        self.stride = stride


class DecBlock(nn.Module):
    def __init__(self, cin, planes, up):
        super().__init__()
        if up:
            self.conv1 = SpectralNormConv((cin, cin, 4, 4), transposed=True, stride=2)
        else:
            self.conv1 = SpectralNormConv((cin, cin, 3, 3))
        self.bn1 = _bn(cin)
        self.activation = _Holder()
        self.conv2 = SpectralNormConv((planes, cin, 3, 3))
        self.bn2 = _bn(planes, 0.0)
        self.upsample = None
        if up:
            self.upsample = _seq([None, SpectralNormConv((planes, cin, 1, 1)), _bn(planes)])


class GuidedCxtAttenParams(nn.Module):
    """Parameters of the guided contextual attention block (GCA/ops.py:83-104)."""

    def __init__(self, out_channels, guidance_channels, rate=2):
        super().__init__()
        self.rate = rate
        self.padding = _Holder()
        self.up_sample = _Holder()
        self.guidance_conv = nn.Conv2d(guidance_channels, guidance_channels // 2, 1)
        self.W = nn.Sequential(nn.Conv2d(out_channels, out_channels, 1, bias=False),
                               nn.BatchNorm2d(out_channels))
        nn.init.xavier_uniform_(self.guidance_conv.weight)
        nn.init.constant_(self.guidance_conv.bias, 0)
        nn.init.xavier_uniform_(self.W[0].weight)
        nn.init.constant_(self.W[1].weight, 1e-3)
        nn.init.constant_(self.W[1].bias, 0)


class TAMParams(nn.Module):
    """Parameters of the Temporal Attention Module (VMN/VMN_model.py:9-16)."""

    def __init__(self, input_chn, reduction, window):
        super().__init__()
        out_chn = input_chn // reduction
        self.key_conv = nn.Conv2d(input_chn, out_chn, 3, padding=1)
        self.query_conv = nn.Conv2d(input_chn, out_chn, 3, padding=1)
        self.value_conv = nn.Conv2d(input_chn, out_chn, 3, padding=1)
        self.window = window


class GCAEncoderParams(nn.Module):
    """resnet_gca_encoder_29: ResNet-D BasicBlock x [3,4,4,2] + shortcuts + guidance head + GCA."""

    def __init__(self):
        super().__init__()
        self.conv1 = SpectralNormConv((32, 6, 3, 3), stride=2)
        self.conv2 = SpectralNormConv((32, 32, 3, 3))
        self.conv3 = SpectralNormConv((64, 32, 3, 3), stride=2)
        self.bn1, self.bn2, self.bn3 = _bn(32), _bn(32), _bn(64)
        self.activation = _Holder()
        cin = 64
        for name, planes, blocks, stride in ENC_LAYERS:
            layer = _Holder()
            for i in range(blocks):
                layer.add_module(str(i), EncBlock(cin, planes, stride if i == 0 else 1))
                cin = planes
            self.add_module(name, layer)
        with torch.no_grad():
            self.conv1.module.weight_bar[:, 3:] = 0       # resnet_enc.py:101
        self.shortcut = _Holder()
        for i, (ci, co) in enumerate(((6, 32), (32, 32), (64, 64), (128, 128), (256, 256))):
            self.shortcut.add_module(str(i), _seq([
                SpectralNormConv((co, ci, 3, 3)), None, _bn(co),
                SpectralNormConv((co, co, 3, 3)), None, _bn(co)]))
        self.guidance_head = _seq([
            None, SpectralNormConv((16, 3, 3, 3), stride=2), None, _bn(16),
            None, SpectralNormConv((32, 16, 3, 3), stride=2), None, _bn(32),
            None, SpectralNormConv((128, 32, 3, 3), stride=2), None, _bn(128)])
        self.gca = GuidedCxtAttenParams(128, 128)


class GCADecoderParams(nn.Module):
    """ResGuidedCxtAtten_FAM_Dec: ResNet-D decoder BasicBlock x [2,3,3,2] + GCA + TAM."""

    def __init__(self, reduction, window, freeze_backbone=False):
        super().__init__()
        self.conv1 = SpectralNormConv((32, 32, 4, 4), transposed=True, stride=2)
        self.bn1 = _bn(32)
        self.leaky_relu = _Holder()
        self.conv2 = nn.Conv2d(32, 1, 3, padding=1)
        nn.init.xavier_uniform_(self.conv2.weight)
        self.upsample = _Holder()
        self.tanh = _Holder()
        cin = 512
        for li, (name, planes, blocks) in enumerate(DEC_LAYERS):
            if li == 2:
                cin = int(cin * reduction)                # layer_multi=[1,1,reduction] (VMN_GCA.py:13)
            layer = _Holder()
            for i in range(blocks):
                layer.add_module(str(i), DecBlock(cin, planes, i == 0))
                cin = planes
            self.add_module(name, layer)
        self.gca = GuidedCxtAttenParams(128, 128)
        self.fam = TAMParams(128, reduction, window)
        self.freeze_backbone = freeze_backbone
