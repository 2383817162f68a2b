"""Parameter containers for the ``vmn_dim`` network with the reference's state_dict layout.

Like ``modules.py`` for ``vmn_gca``: the classes hold exactly the reference's parameters / buffers under exactly its
names and registration order (``pred_test.py:92`` loads checkpoints with ``strict=True``) and compute nothing
themselves; all arithmetic is done by the sm_100a kernels driven from ``tcvom_b200.dim_engine``.

Name/shape layout restated from (reference checkout, commit f5fa07a):
  encoder  models/VMN/VMN_DIM.py:6-46   (VGG-16 with BatchNorm, 2x2 max pooling with indices, 7x7 conv6)
  decoder  models/VMN/VMN_DIM.py:74-99  (1x1 dconv6, five 5x5 dconvs behind max-unpooling, 5x5 alpha head, TAM(256))
"""
from __future__ import annotations

from torch import nn

from .modules import TAMParams, _Holder

# (stage, [(conv name, bn name, cin, cout), ...]) -- VMN_DIM.py:10-44
ENC_STAGES = (
    (1, (("conv11", "bn11", None, 64), ("conv12", "bn12", 64, 64))),
    (2, (("conv21", "bn21", 64, 128), ("conv22", "bn22", 128, 128))),
    (3, (("conv31", "bn31", 128, 256), ("conv32", "bn32", 256, 256), ("conv33", "bn33", 256, 256))),
    (4, (("conv41", "bn41", 256, 512), ("conv42", "bn42", 512, 512), ("conv43", "bn43", 512, 512))),
    (5, (("conv51", "bn51", 512, 512), ("conv52", "bn52", 512, 512), ("conv53", "bn53", 512, 512))),
)
# (name, cin, cout, k) -- VMN_DIM.py:79-97
DEC_CONVS = (("dconv6", 4096, 512, 1), ("dconv5", 512, 512, 5), ("dconv4", 512, 256, 5), ("dconv3", 256, 128, 5),
             ("dconv2", 128, 64, 5), ("dconv1", 64, 64, 5), ("alpha_pred", 64, 1, 5))


class DIMEncoderParams(nn.Module):
    def __init__(self, input_chn: int):
        super().__init__()
        for stage, convs in ENC_STAGES:
            for cname, bname, cin, cout in convs:
                self.add_module(cname, nn.Conv2d(input_chn if cin is None else cin, cout, kernel_size=3, padding=1))
                self.add_module(bname, nn.BatchNorm2d(cout))
            self.add_module(f"pool{stage}", _Holder())
        self.conv6 = nn.Conv2d(512, 4096, kernel_size=7, padding=3)


class DIMDecoderParams(nn.Module):
    def __init__(self, reduction, window, freeze_backbone=False):
        super().__init__()
        self.freeze_backbone = freeze_backbone
        for name, cin, cout, k in DEC_CONVS:
            if name != "dconv6" and name != "alpha_pred":
                self.add_module("unpool" + name[-1], _Holder())
            self.add_module(name, nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2))
        self.fam = TAMParams(256, reduction, window)
