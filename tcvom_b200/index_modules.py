"""Parameter containers for the ``vmn_index`` network with the reference's state_dict layout (555 keys).

Like ``modules.py`` for ``vmn_gca``: the classes hold exactly the reference's parameters / buffers under exactly its names
and registration order (``pred_test.py:92`` loads checkpoints with ``strict=True``) and compute nothing themselves; all
arithmetic is done by the sm_100a kernels driven from ``tcvom_b200.index_engine``.

Name/shape layout restated from (reference checkout, commit f5fa07a):
  encoder  models/Index/net.py:25-83 (InvertedResidual), :85-168 (IndexMattingEncoder: MobileNetV2 with every stride moved into
           index pooling, five DepthwiseM2OIndexBlock, ASPP), models/Index/hlindex.py:125-152, models/Index/hlaspp.py:37-118
  decoder  models/Index/net.py:239-257 (seven IndexedUpsamlping 5x5 conv blocks + pred), models/Index/hlconv.py:35-40,
           models/VMN/VMN_Index.py:7-11 (fam = TAM(32))
"""
from __future__ import annotations

from torch import nn

from .modules import TAMParams, _Holder, _seq

# expand_ratio, input_chn, output_chn, num_blocks, dilation  (net.py:107-116; every stride is 1 after :128-135, and with
# output_stride == 32 no block is dilated)
IR_SETTING = ((1, 32, 16, 1), (6, 16, 24, 2), (6, 24, 32, 3), (6, 32, 64, 4), (6, 64, 96, 3), (6, 96, 160, 3), (6, 160, 320, 1))
INDEX_BLOCKS = (("index0", 32), ("index2", 24), ("index3", 32), ("index4", 64), ("index6", 160))
ASPP_DILATIONS = (1, 2, 4, 8)           # hlaspp.py:80-81, output_stride == 32
# (name, input channels (decoder + encoder feature), output channels)  -- net.py:248-254
DEC_LAYERS = (("decoder_layer6", 320, 96), ("decoder_layer5", 192, 64), ("decoder_layer4", 128, 32),
              ("decoder_layer3", 64, 24), ("decoder_layer2", 48, 16), ("decoder_layer1", 32, 32), ("decoder_layer0", 64, 32))


def _bn(c):
    return nn.BatchNorm2d(c)


class InvertedResidualParams(nn.Module):
    def __init__(self, inp, oup, expand_ratio):
        super().__init__()
        hid = round(inp * expand_ratio)
        if expand_ratio == 1:
            mods = [nn.Conv2d(hid, hid, 3, 1, 0, groups=hid, bias=False), _bn(hid), None,
                    nn.Conv2d(hid, oup, 1, bias=False), _bn(oup)]
        else:
            mods = [nn.Conv2d(inp, hid, 1, bias=False), _bn(hid), None,
                    nn.Conv2d(hid, hid, 3, 1, 0, groups=hid, bias=False), _bn(hid), None,
                    nn.Conv2d(hid, oup, 1, bias=False), _bn(oup)]
        self.conv = _seq(mods)


class IndexBlockParams(nn.Module):
    """DepthwiseM2OIndexBlock(inp, use_nonlinear=True, use_context=True): four dense 4x4/s2 conv + BN + ReLU6 + 1x1 branches."""

    def __init__(self, inp):
        super().__init__()
        for i in range(1, 5):
            self.add_module(f"indexnet{i}", _seq([nn.Conv2d(inp, inp, 4, 2, 1, bias=False), _bn(inp), None,
                                                  nn.Conv2d(inp, inp, 1, bias=False)]))


class _AsppBranch(nn.Module):
    def __init__(self, mods):
        super().__init__()
        self.atrous_conv = _seq(mods)


class AsppParams(nn.Module):
    def __init__(self, inp=320, oup=160, planes=256):
        super().__init__()
        self.aspp1 = _AsppBranch([nn.Conv2d(inp, planes, 1, bias=False), _bn(planes), None])
        for i, d in enumerate(ASPP_DILATIONS[1:], start=2):
            self.add_module(f"aspp{i}", _AsppBranch([nn.Conv2d(inp, inp, 3, 1, d, d, groups=inp, bias=False), _bn(inp), None,
                                                     nn.Conv2d(inp, planes, 1, bias=False), _bn(planes), None]))
        self.global_avg_pool = _seq([None, nn.Conv2d(inp, planes, 1, bias=False), _bn(planes), None])
        self.bottleneck_conv = _seq([nn.Conv2d(planes * 5, oup, 1, bias=False), _bn(oup), None])
        self.dropout = _Holder()


class IndexEncoderParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer0 = _seq([nn.Conv2d(4, 32, 3, 1, 1, bias=False), _bn(32), None])
        for li, (t, inp, oup, n) in enumerate(IR_SETTING, start=1):
            blocks = [InvertedResidualParams(inp if i == 0 else oup, oup, t) for i in range(n)]
            self.add_module(f"layer{li}", _seq(blocks))
        for name, c in INDEX_BLOCKS:
            self.add_module(name, IndexBlockParams(c))
        self.dconv_pp = AsppParams()


class _DecBlock(nn.Module):
    def __init__(self, inp, oup):
        super().__init__()
        self.dconv = _seq([nn.Conv2d(inp, oup, 5, 1, 2, bias=False), _bn(oup), None])


class IndexDecoderParams(nn.Module):
    def __init__(self, reduction, window, freeze_backbone=False):
        super().__init__()
        for name, inp, oup in DEC_LAYERS:
            self.add_module(name, _DecBlock(inp, oup))
        self.pred = _seq([_seq([nn.Conv2d(32, 1, 5, 1, 2, bias=False), _bn(1), None]), nn.Conv2d(1, 1, 5, 1, 2, bias=False)])
        self.fam = TAMParams(32, reduction, window)
        self.freeze_backbone = freeze_backbone
