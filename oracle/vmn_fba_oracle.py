"""CPU oracle for the TCVOM FBA+TAM frame-window forward (``vmn_fba``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``tcvom_b200/`` imports this module; it is used by
``tests/`` (and may be used by ``bench.py``'s CPU-baseline legs) as the checker.

A from-scratch *functional* restatement (plain ``torch`` fp32 on the CPU, driven directly by a
``NET.state_dict()``) of what the reference computes for ``EvalModel('vmn_fba')``.  Each function
cites the reference file:line it restates (paths relative to the reference checkout, commit f5fa07a).

Parity pinning: upstream ships no tests / fixtures / golden vectors ("parity unpinned" upstream,
SURVEY.md section 8c), so this oracle is pinned against outputs of the unmodified reference executed
in the build container by ``tests/golden/make_golden_fba.py`` and committed as
``tests/golden/fba_*.npz`` (``tests/test_oracle_fba.py``).  The reference computes the trimap distance
transforms with ``cv2.distanceTransform(DIST_L2, DIST_MASK_PRECISE)`` (utils/utils.py:12-23), i.e. the
exact Euclidean transform in float32; here ``scipy.ndimage.distance_transform_edt`` (also exact; equal
to cv2 bit for bit on the golden inputs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .vmn_gca_oracle import IMG_MEAN, IMG_STD, tam

SD = Dict[str, torch.Tensor]
GN_EPS = 1e-5
GN_GROUPS = 32
LEAKY = 0.01                                   # nn.LeakyReLU() default slope, FBA/models.py:266
RES_LAYERS = (("layer1", 64, 3, 1, 1), ("layer2", 128, 4, 2, 1), ("layer3", 256, 6, 2, 2),
              ("layer4", 512, 3, 2, 4))        # (name, planes, blocks, stride, dilate): resnet_GN_WS.py:141-145,
#                                                ResnetDilated(dilate_scale=8) FBA/models.py:186-192
PPM_SCALES = (1, 2, 3, 6)                      # FBA/models.py:259


# --------------------------------------------------------------------------- primitives
def ws_weight(w: torch.Tensor) -> torch.Tensor:
    """Weight standardisation -- FBA/layers_WS.py:13-21 (unbiased variance, eps inside and outside the sqrt)."""
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    w = w - mean
    std = torch.sqrt(torch.var(w.reshape(w.shape[0], -1), dim=1) + 1e-12).reshape(-1, 1, 1, 1) + 1e-5
    return w / std


def ws_conv(x, sd: SD, p: str, stride=1, padding=0, dilation=1):
    return F.conv2d(x, ws_weight(sd[p + ".weight"]), sd.get(p + ".bias"), stride, padding, dilation)


def gn(x, sd: SD, p: str):
    """nn.GroupNorm(32, C) -- FBA/layers_WS.py:26-27, FBA/models.py:239-243."""
    return F.group_norm(x, GN_GROUPS, sd[p + ".weight"], sd[p + ".bias"], GN_EPS)


def bottleneck(x, sd: SD, p: str, stride: int, dilation: int, has_down: bool, down_stride: int):
    """Bottleneck.forward -- FBA/resnet_GN_WS.py:69-91 (conv2 carries stride / dilation)."""
    out = F.relu(gn(ws_conv(x, sd, p + ".conv1"), sd, p + ".bn1"))
    out = F.relu(gn(ws_conv(out, sd, p + ".conv2", stride, dilation, dilation), sd, p + ".bn2"))
    out = gn(ws_conv(out, sd, p + ".conv3"), sd, p + ".bn3")
    idt = x
    if has_down:
        idt = gn(ws_conv(x, sd, p + ".downsample.0", down_stride), sd, p + ".downsample.1")
    return F.relu(out + idt)


def block_config(name: str, i: int, stride: int, dilate: int):
    """(conv2 stride, conv2 dilation, downsample stride) of block i after ResnetDilated._nostride_dilate
    (FBA/models.py:204-218): a stride-2 conv becomes stride 1 (3x3: dilation dilate//2); every other 3x3 gets
    dilation ``dilate``."""
    if dilate == 1:
        return (stride if i == 0 else 1), 1, (stride if i == 0 else 1)
    if i == 0:
        return 1, dilate // 2, 1
    return 1, dilate, 1


def encoder(x, sd: SD, p="encoder") -> List[torch.Tensor]:
    """ResnetDilated.forward -- FBA/models.py:222-236.  x [B,11,H,W] -> conv_out (6 tensors)."""
    conv_out = [x]
    x = F.relu(gn(ws_conv(x, sd, p + ".conv1", 2, 3), sd, p + ".bn1"))
    conv_out.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for name, planes, blocks, stride, dilate in RES_LAYERS:
        for i in range(blocks):
            s, d, ds = block_config(name, i, stride, dilate)
            x = bottleneck(x, sd, f"{p}.{name}.{i}", s, d, i == 0, ds)
        conv_out.append(x)
    return conv_out


def edt(seed_free: np.ndarray) -> np.ndarray:
    """cv2.distanceTransform(x*255, DIST_L2, 0) of utils/utils.py:12-23: distance of every pixel to the nearest
    zero pixel of ``seed_free`` (exact, float32); an image without zeros gives +huge (cv2: 1.8e19)."""
    from scipy import ndimage
    if seed_free.all():
        return np.full(seed_free.shape, 1.8446743e19, np.float32)
    return ndimage.distance_transform_edt(seed_free).astype(np.float32)


def trimap_transform(trimap2: torch.Tensor) -> torch.Tensor:
    """utils/utils.py:25-39.  trimap2 [B,S,2,H,W] (bg, fg one-hot) -> clicks [B,S,6,H,W]."""
    B, S, _, H, W = trimap2.shape
    clicks = torch.zeros((B, S, 6, H, W))
    for k in range(2):
        tk = trimap2[:, :, k]
        if torch.sum(tk != 0) > 0:
            d = np.stack([np.stack([edt((1.0 - tk[b, s]).numpy() != 0) for s in range(S)]) for b in range(B)])
            dt_mask = -torch.from_numpy(d) ** 2
            L = 320
            for j, f in enumerate((0.02, 0.08, 0.16)):
                clicks[:, :, 3 * k + j] = torch.exp(dt_mask / (2 * ((f * L) ** 2)))
    return clicks


def eval_preprocess(imgs, tris, dilate_kernel=None):
    """EvalModel.preprocess with TRIMAP_CHANNEL == 8 -- models/model.py:360-387."""
    scale = 1.0 / 255
    mean = torch.tensor(IMG_MEAN).reshape(1, 1, 3, 1, 1)
    std = torch.tensor(IMG_STD).reshape(1, 1, 3, 1, 1)
    scaled_imgs = imgs.float().flip([2]) * scale
    nimgs = (scaled_imgs - mean) / std
    scaled_tris = tris.float() * scale
    trimask = (scaled_tris > 0) & (scaled_tris < 1)
    if dilate_kernel is not None:
        k = int(dilate_kernel)
        trimask = torch.stack([F.max_pool2d(t.float(), 2 * k + 1, 1, k) for t in trimask]).bool()
    trimap2 = torch.cat([(scaled_tris == 0).float(), (scaled_tris == 1).float()], dim=2)
    x11 = torch.cat([nimgs, trimap_transform(trimap2), trimap2], dim=2)
    return scaled_imgs, x11, trimask.float(), trimap2


def decoder_head(conv_out, sd: SD, p="decoder"):
    """vmn_fba_decoder.forward(extract_feature=True) -- VMN/VMN_FBA.py:20-33."""
    conv5 = conv_out[-1]
    h, w = conv5.shape[2:]
    outs = [conv5]
    for i, s in enumerate(PPM_SCALES):
        t = F.adaptive_avg_pool2d(conv5, s)
        t = F.leaky_relu(gn(ws_conv(t, sd, f"{p}.ppm.{i}.1"), sd, f"{p}.ppm.{i}.2"), LEAKY)
        outs.append(F.interpolate(t, (h, w), mode="bilinear", align_corners=False))
    x = torch.cat(outs, 1)
    x = F.leaky_relu(gn(ws_conv(x, sd, p + ".conv_up1.0", 1, 1), sd, p + ".conv_up1.1"), LEAKY)
    return F.leaky_relu(gn(ws_conv(x, sd, p + ".conv_up1.3", 1, 1), sd, p + ".conv_up1.4"), LEAKY)


def fba_fusion(alpha, img, Fg, Bg):
    """FBA/models.py:246-255."""
    Fg = alpha * img + (1 - alpha ** 2) * Fg - alpha * (1 - alpha) * Bg
    Bg = (1 - alpha) * img + (2 * alpha - alpha ** 2) * Bg - alpha * (1 - alpha) * Fg
    Fg = torch.clamp(Fg, 0, 1)
    Bg = torch.clamp(Bg, 0, 1)
    la = 0.1
    alpha = (alpha * la + torch.sum((img - Bg) * (Fg - Bg), 1, keepdim=True)) / \
        (torch.sum((Fg - Bg) * (Fg - Bg), 1, keepdim=True) + la)
    return torch.clamp(alpha, 0, 1), Fg, Bg


def decoder_tail(conv_out, img, trimap2, sd: SD, x, xb, xf, mask, window=7, p="decoder"):
    """vmn_fba_decoder.forward(extract_feature=False) -- VMN/VMN_FBA.py:34-59."""
    up2 = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    x, attb, attf, sm = tam(sd, p + ".fam", x, xb, xf, mask, window)
    x = torch.cat((up2(x), conv_out[-4]), 1)
    x = F.leaky_relu(gn(ws_conv(x, sd, p + ".conv_up2.0", 1, 1), sd, p + ".conv_up2.1"), LEAKY)
    x = torch.cat((up2(x), conv_out[-5]), 1)
    x = F.leaky_relu(gn(ws_conv(x, sd, p + ".conv_up3.0", 1, 1), sd, p + ".conv_up3.1"), LEAKY)
    x = torch.cat((up2(x), conv_out[-6][:, :3], img, trimap2), 1)
    x = F.leaky_relu(F.conv2d(x, sd[p + ".conv_up4.0.weight"], sd[p + ".conv_up4.0.bias"], 1, 1), LEAKY)
    x = F.leaky_relu(F.conv2d(x, sd[p + ".conv_up4.2.weight"], sd[p + ".conv_up4.2.bias"], 1, 1), LEAKY)
    o = F.conv2d(x, sd[p + ".conv_up4.4.weight"], sd[p + ".conv_up4.4.bias"])
    alpha = torch.clamp(o[:, :1], 0, 1)
    alpha, Fg, Bg = fba_fusion(alpha, img, torch.sigmoid(o[:, 1:4]), torch.sigmoid(o[:, 4:7]))
    return torch.cat((alpha, Fg, Bg), 1), attb, attf, sm


def vmn_forward(sd: SD, frames: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], extras, window=7,
                return_features=False):
    """VMN.forward -- VMN/VMN_model.py:83-113 with the FBA base net.  frames[i] [B,11,H,W]; masks[i] [B,1,H,W];
    extras[i] = (scaled RGB [B,3,H,W], two-channel trimap [B,2,H,W])."""
    S = len(frames)
    couts, feats = [], []
    for i in range(S):
        co = encoder(frames[i], sd)
        couts.append(co)
        feats.append(decoder_head(co, sd))
    preds: List[Optional[torch.Tensor]] = [None] * S
    attb: List[Optional[torch.Tensor]] = [None] * S
    attf: List[Optional[torch.Tensor]] = [None] * S
    small: List[Optional[torch.Tensor]] = [None] * S
    for i in range(1, S - 1):
        preds[i], attb[i], attf[i], small[i] = decoder_tail(couts[i], extras[i][0], extras[i][1], sd, feats[i],
                                                            feats[i - 1], feats[i + 1], masks[i], window)
    preds[0] = torch.zeros_like(preds[1])
    preds[-1] = torch.zeros_like(preds[-2])
    if return_features:
        return preds, attb, attf, small, feats, couts
    return preds, attb, attf, small


def eval_forward(sd: SD, imgs, tris, dilate_kernel=None, window=7, return_aux=False):
    """EvalModel.forward for method 'fba' -- models/model.py:389-446.  imgs [B,S,3,H,W] BGR 0..255, tris [B,S,1,H,W]
    -> (alphas [B,S,1,H,W], Fs [B,S,3,H,W], Bs [B,S,3,H,W]); first / last frame zeros."""
    with torch.no_grad():
        B, S = imgs.shape[:2]
        scaled_imgs, x11, trimask, trimap2 = eval_preprocess(imgs, tris, dilate_kernel)
        frames = [x11[:, i] for i in range(S)]
        masks = [trimask[:, i] for i in range(S)]
        extras = [(scaled_imgs[:, i], trimap2[:, i]) for i in range(S)]
        preds, attb, attf, small = vmn_forward(sd, frames, masks, extras, window)
        alphas = torch.zeros((B, S, 1) + imgs.shape[-2:])
        Fs = torch.zeros((B, S, 3) + imgs.shape[-2:])
        Bs = torch.zeros_like(Fs)
        for c in range(1, S - 1):
            m = trimask[:, c].bool()
            gt = tris[:, c].float() * (1.0 / 255)
            alphas[:, c] = torch.where(m, preds[c][:, :1], gt)
            Fs[:, c] = torch.where(m.repeat(1, 3, 1, 1), preds[c][:, 1:4], scaled_imgs[:, c])
            Bs[:, c] = torch.where(m.repeat(1, 3, 1, 1), preds[c][:, 4:7], scaled_imgs[:, c])
        if return_aux:
            return alphas, Fs, Bs, dict(preds=preds, attb=attb, attf=attf, small=small, x11=x11, trimask=trimask)
        return alphas, Fs, Bs
