"""CPU oracle for the TCVOM IndexNet+TAM frame-window forward (``vmn_index``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``tcvom_b200/`` imports this module; it is used by ``tests/`` as the checker.

A from-scratch *functional* restatement (plain ``torch`` fp32, driven directly by a ``NET.state_dict()``) of what the
reference computes for ``EvalModel('vmn_index')``.  Each function cites the reference file:line it restates (paths relative
to the reference checkout, commit f5fa07a).

Parity pinning: upstream ships no tests / fixtures / golden vectors, so this oracle is pinned against
outputs of the unmodified reference executed in the build container by ``tests/golden/make_golden_index.py`` and committed as
``tests/golden/index_*.npz`` (``tests/test_oracle_index.py``).  The fixture checkpoint is regenerated from a seed
(``fixture_sd_index``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .vmn_dim_oracle import eval_preprocess            # TRIMAP_CHANNEL == 1 for 'index' as for 'dim' (models/model.py:22-27)
from .vmn_gca_oracle import bn, tam

SD = Dict[str, torch.Tensor]
IR_SETTING = ((1, 32, 16, 1), (6, 16, 24, 2), (6, 24, 32, 3), (6, 32, 64, 4), (6, 64, 96, 3), (6, 96, 160, 3), (6, 160, 320, 1))
ASPP_DILATIONS = (2, 4, 8)


def relu6(x):
    return F.relu6(x)


def conv_bn_relu6(x, sd: SD, p: str, k: int):
    """hlconv.conv_bn -- hlconv.py:35-40 (Conv2d(k, padding=k//2, bias=False) + BatchNorm + ReLU6); Sequential indices 0, 1."""
    return relu6(bn(F.conv2d(x, sd[p + ".0.weight"], None, 1, k // 2), sd, p + ".1"))


def inverted_residual(x, sd: SD, p: str, inp: int, oup: int, t: int):
    """InvertedResidual.forward -- net.py:25-83 with stride 1, dilation 1: fixed_padding(3, 1) = 1 pixel on every side, then
    the padding-0 depthwise conv; residual when inp == oup."""
    c = p + ".conv"
    xp = F.pad(x, (1, 1, 1, 1))
    if t == 1:
        h = relu6(bn(F.conv2d(xp, sd[c + ".0.weight"], None, 1, 0, 1, xp.shape[1]), sd, c + ".1"))
        y = bn(F.conv2d(h, sd[c + ".3.weight"]), sd, c + ".4")
    else:
        # the 1x1 expansion sees the PADDED input (net.py:79-83): its BatchNorm shift makes the border non-zero
        h = relu6(bn(F.conv2d(xp, sd[c + ".0.weight"]), sd, c + ".1"))
        h = relu6(bn(F.conv2d(h, sd[c + ".3.weight"], None, 1, 0, 1, h.shape[1]), sd, c + ".4"))
        y = bn(F.conv2d(h, sd[c + ".6.weight"]), sd, c + ".7")
    return x + y if inp == oup else y


def layer(x, sd: SD, p: str, setting):
    t, inp, oup, n = setting
    for i in range(n):
        x = inverted_residual(x, sd, f"{p}.{i}", inp if i == 0 else oup, oup, t)
    return x


def index_block(x, sd: SD, p: str):
    """DepthwiseM2OIndexBlock.forward -- hlindex.py:125-167 (use_nonlinear, use_context): four branches, sigmoid, softmax over
    the branches, pixel shuffle: branch b lands on sub-pixel (b // 2, b % 2) of the same channel."""
    outs = []
    for i in range(1, 5):
        q = f"{p}.indexnet{i}"
        h = relu6(bn(F.conv2d(x, sd[q + ".0.weight"], None, 2, 1), sd, q + ".1"))
        outs.append(F.conv2d(h, sd[q + ".3.weight"]))
    s = torch.stack(outs, dim=2)                       # [B, C, 4, h/2, w/2]
    y = torch.sigmoid(s)
    z = F.softmax(y, dim=2)
    B, C, _, h2, w2 = s.shape
    return F.pixel_shuffle(z.reshape(B, C * 4, h2, w2), 2), F.pixel_shuffle(y.reshape(B, C * 4, h2, w2), 2)


def aspp(x, sd: SD, p: str):
    """ASPP.forward -- hlaspp.py:106-118 (eval: dropout is the identity)."""
    x1 = conv_bn_relu6(x, sd, p + ".aspp1.atrous_conv", 1)
    branches = [x1]
    for i, d in enumerate(ASPP_DILATIONS, start=2):
        q = f"{p}.aspp{i}.atrous_conv"
        h = relu6(bn(F.conv2d(x, sd[q + ".0.weight"], None, 1, d, d, x.shape[1]), sd, q + ".1"))
        branches.append(relu6(bn(F.conv2d(h, sd[q + ".3.weight"]), sd, q + ".4")))
    g = x.mean(dim=(2, 3), keepdim=True)
    g = relu6(bn(F.conv2d(g, sd[p + ".global_avg_pool.1.weight"]), sd, p + ".global_avg_pool.2"))
    branches.append(g.expand(-1, -1, x.shape[2], x.shape[3]))           # F.interpolate(nearest) of a 1x1 map
    return conv_bn_relu6(torch.cat(branches, dim=1), sd, p + ".bottleneck_conv", 1)


def encoder(x, sd: SD, p="encoder"):
    """IndexMattingEncoder.forward -- net.py:189-228."""
    def pool(t, idx_en):
        return 4 * F.avg_pool2d(idx_en * t, (2, 2), 2)

    l0 = conv_bn_relu6(x, sd, p + ".layer0", 3)
    i0e, i0d = index_block(l0, sd, p + ".index0")
    l0 = i0e * l0
    l1 = layer(4 * F.avg_pool2d(l0, (2, 2), 2), sd, p + ".layer1", IR_SETTING[0])
    l2 = layer(l1, sd, p + ".layer2", IR_SETTING[1])
    i2e, i2d = index_block(l2, sd, p + ".index2")
    l2 = i2e * l2
    l3 = layer(4 * F.avg_pool2d(l2, (2, 2), 2), sd, p + ".layer3", IR_SETTING[2])
    i3e, i3d = index_block(l3, sd, p + ".index3")
    l3 = i3e * l3
    l4 = layer(4 * F.avg_pool2d(l3, (2, 2), 2), sd, p + ".layer4", IR_SETTING[3])
    i4e, i4d = index_block(l4, sd, p + ".index4")
    l4 = i4e * l4
    l5 = layer(4 * F.avg_pool2d(l4, (2, 2), 2), sd, p + ".layer5", IR_SETTING[4])
    l6 = layer(l5, sd, p + ".layer6", IR_SETTING[5])
    i6e, i6d = index_block(l6, sd, p + ".index6")
    l6 = i6e * l6
    l7 = layer(4 * F.avg_pool2d(l6, (2, 2), 2), sd, p + ".layer7", IR_SETTING[6])
    l = aspp(l7, sd, p + ".dconv_pp")
    return [l, l6, i6d, l5, l4, i4d, l3, i3d, l2, i2d, l1, l0, i0d]


def dec_block(l_dec, l_low, idx, sd: SD, p: str):
    """IndexedUpsamlping.forward -- hldecoder.py:121-127."""
    if idx is not None:
        l_dec = idx * F.interpolate(l_dec, size=l_low.shape[2:], mode="nearest")
    return conv_bn_relu6(torch.cat((l_dec, l_low), dim=1), sd, p + ".dconv", 5)


def decoder_head(e, sd: SD, p="decoder"):
    """IndexMattingDecoder_VMN.forward(extract_feature=True) -- VMN_Index.py:15-20: the OS8 feature [B,32,H/8,W/8]."""
    l, l6, i6d, l5, l4, i4d = e[:6]
    t = dec_block(l, l6, i6d, sd, p + ".decoder_layer6")
    t = dec_block(t, l5, None, sd, p + ".decoder_layer5")
    return dec_block(t, l4, i4d, sd, p + ".decoder_layer4")


def decoder_tail(e, sd: SD, x, xb, xf, mask, window=7, p="decoder"):
    """IndexMattingDecoder_VMN.forward(extract_feature=False) -- VMN_Index.py:21-28; pred = net.py:15-22."""
    l3, i3d, l2, i2d, l1, l0, i0d = e[6:]
    t, attb, attf, sm = tam(sd, p + ".fam", x, xb, xf, mask, window)
    t = dec_block(t, l3, i3d, sd, p + ".decoder_layer3")
    t = dec_block(t, l2, i2d, sd, p + ".decoder_layer2")
    t = dec_block(t, l1, None, sd, p + ".decoder_layer1")
    t = dec_block(t, l0, i0d, sd, p + ".decoder_layer0")
    t = conv_bn_relu6(t, sd, p + ".pred.0", 5)
    return F.conv2d(t, sd[p + ".pred.1.weight"], None, 1, 2), attb, attf, sm


def vmn_forward(sd: SD, frames: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], window=7):
    """VMN.forward -- VMN/VMN_model.py:83-113.  frames[i]: [B,4,H,W]; masks[i]: [B,1,H,W]."""
    S = len(frames)
    encs, feats = [], []
    for i in range(S):
        e = encoder(frames[i], sd)
        encs.append(e)
        feats.append(decoder_head(e, sd))
    preds: List[Optional[torch.Tensor]] = [None] * S
    attb: List[Optional[torch.Tensor]] = [None] * S
    attf: List[Optional[torch.Tensor]] = [None] * S
    small: List[Optional[torch.Tensor]] = [None] * S
    for i in range(1, S - 1):
        preds[i], attb[i], attf[i], small[i] = decoder_tail(encs[i], sd, feats[i], feats[i - 1], feats[i + 1], masks[i], window)
    preds[0] = torch.zeros_like(preds[1])
    preds[-1] = torch.zeros_like(preds[-2])
    return preds, attb, attf, small, feats


def eval_forward(sd: SD, imgs, tris, dilate_kernel=None, window=7, return_aux=False):
    """EvalModel.forward for vmn_index -- models/model.py:389-424.  Returns alphas [B,S,1,H,W]."""
    with torch.no_grad():
        x4, trimask = eval_preprocess(imgs, tris, dilate_kernel)
        S = imgs.shape[1]
        frames = [x4[:, i] for i in range(S)]
        masks = [trimask[:, i] for i in range(S)]
        preds, attb, attf, small, feats = vmn_forward(sd, frames, masks, window)
        alphas = []
        for c in range(S):
            if c == 0 or c == S - 1:
                alphas.append(torch.zeros_like(preds[1]))
            else:
                gt = tris[:, c].float() * (1.0 / 255)
                alphas.append(torch.where(trimask[:, c].bool(), preds[c], gt))   # :418
        out = torch.stack(alphas, dim=1)
    if return_aux:
        return out, dict(preds=preds, attb=attb, attf=attf, small_mask=small, feats=feats)
    return out


def fixture_sd_index(shapes, seed: int = 11) -> SD:
    """Deterministic, well-conditioned weights for the reference's 555-key layout (`shapes`: ordered name -> shape, e.g. from
    tests/golden/vmn_index_keys.json): He-normal convolutions, BatchNorm statistics near the identity, a prediction head
    whose output is spread over (0, 1)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("running_var"):
            sd[name] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            gain = 2.0
            if ".indexnet" in name and name.endswith(".3.weight"):
                gain = 8.0                       # index logits: spread the sigmoids
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
        elif name.endswith(".weight"):           # BatchNorm gamma
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                    # BatchNorm beta / conv bias
            sd[name] = 0.1 * torch.randn(shape, generator=g)
    # prediction head: pred.0 = conv + BN(1) + ReLU6, pred.1 = 1 -> 1 5x5 conv
    sd["decoder.pred.0.1.bias"] = torch.tensor([1.0])
    sd["decoder.pred.1.weight"] = torch.full((1, 1, 5, 5), 0.012) + 0.004 * torch.randn((1, 1, 5, 5), generator=g)
    return sd
