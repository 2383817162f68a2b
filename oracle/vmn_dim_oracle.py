"""CPU oracle for the TCVOM DIM+TAM frame-window forward (``vmn_dim``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``tcvom_b200/`` imports this module; it is used by ``tests/`` as the checker.

A from-scratch *functional* restatement (plain ``torch`` fp32, driven directly by a ``NET.state_dict()``) of what the
reference computes for ``EvalModel('vmn_dim')``.  Each function cites the reference file:line it restates (paths
relative to the reference checkout, commit f5fa07a).

Parity pinning: upstream ships no tests / fixtures / golden vectors, so this oracle is pinned
against outputs of the unmodified reference executed in the build container by ``tests/golden/make_golden_dim.py`` and
committed as ``tests/golden/dim_window.npz`` (``tests/test_oracle_dim.py``).  The network's 134 M parameters (conv6 alone
is 4096 x 512 x 7 x 7) are not committed: ``fixture_sd_dim`` regenerates them from a seed.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .vmn_gca_oracle import IMG_MEAN, IMG_STD, bn, tam

SD = Dict[str, torch.Tensor]
# VMN_DIM.py:10-44
ENC_STAGES = ((("conv11", "bn11"), ("conv12", "bn12")),
              (("conv21", "bn21"), ("conv22", "bn22")),
              (("conv31", "bn31"), ("conv32", "bn32"), ("conv33", "bn33")),
              (("conv41", "bn41"), ("conv42", "bn42"), ("conv43", "bn43")),
              (("conv51", "bn51"), ("conv52", "bn52"), ("conv53", "bn53")))


def conv(x, sd: SD, p: str):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], 1, w.shape[-1] // 2)


def encoder(x, sd: SD, p="encoder", force_idx=None, ties=None):
    """DIMEncoder.forward -- VMN_DIM.py:48-72: returns ([idx1p .. idx5p], x6).

    force_idx (test hook): five index tensors in torch's max_pool2d convention that REPLACE the arg-max of each pooling
    stage (the pooled value is gathered at the forced position).  Max-unpooling makes the network discontinuous: two
    correct fp32 evaluations that differ by 1e-6 can route a near-tie to different pixels and then differ by 1e-2 around
    it, so a checker either has to follow the routing of the implementation under test or exclude those neighbourhoods.
    `ties` (list) receives, per stage, the largest gap  max - value_at_forced_position  relative to max(1, |max|): the
    test asserts it is at rounding level, i.e. every routing difference is a genuine near-tie."""
    idxs = []
    for li, stage in enumerate(ENC_STAGES):
        for cname, bname in stage:
            x = F.relu(bn(conv(x, sd, f"{p}.{cname}"), sd, f"{p}.{bname}"))
        pooled, idx = F.max_pool2d(x, (2, 2), 2, return_indices=True)
        if force_idx is not None:
            forced = x.flatten(2).gather(2, force_idx[li].flatten(2)).view_as(pooled)
            if ties is not None:
                ties.append(float(((pooled - forced) / pooled.abs().clamp(min=1.0)).max()))
            pooled, idx = forced, force_idx[li]
        x = pooled
        idxs.append(idx)
    return idxs, F.relu(conv(x, sd, p + ".conv6"))


def window_idx_to_torch(idx_u8: torch.Tensor) -> torch.Tensor:
    """Pooling indices of the native kernels (uint8 NHWC [n,h,w,c], ky*2 + kx inside the 2x2 window) in torch's convention
    (int64 NCHW, flat position in the [2h, 2w] input plane)."""
    k = idx_u8.permute(0, 3, 1, 2).long()
    h, w = k.shape[-2:]
    oy = torch.arange(h, device=k.device).view(1, 1, h, 1) * 2
    ox = torch.arange(w, device=k.device).view(1, 1, 1, w) * 2
    return (oy + k // 2) * (2 * w) + ox + k % 2


def decoder_head(idxs, x6, sd: SD, p="decoder"):
    """DIMDecoder.forward(extract_feature=True) -- VMN_DIM.py:109-119: the OS8 feature [B,256,H/8,W/8]."""
    t = F.relu(conv(x6, sd, p + ".dconv6"))
    t = F.relu(conv(F.max_unpool2d(t, idxs[4], (2, 2), 2), sd, p + ".dconv5"))
    return F.relu(conv(F.max_unpool2d(t, idxs[3], (2, 2), 2), sd, p + ".dconv4"))


def decoder_tail(idxs, sd: SD, x, xb, xf, mask, window=7, p="decoder"):
    """DIMDecoder.forward(extract_feature=False) -- VMN_DIM.py:120-136."""
    t, attb, attf, sm = tam(sd, p + ".fam", x, xb, xf, mask, window)
    t = F.relu(conv(F.max_unpool2d(t, idxs[2], (2, 2), 2), sd, p + ".dconv3"))
    t = F.relu(conv(F.max_unpool2d(t, idxs[1], (2, 2), 2), sd, p + ".dconv2"))
    t = F.relu(conv(F.max_unpool2d(t, idxs[0], (2, 2), 2), sd, p + ".dconv1"))
    return conv(t, sd, p + ".alpha_pred").clamp(0, 1), attb, attf, sm


def vmn_forward(sd: SD, frames: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], window=7, force_idx=None,
                ties=None):
    """VMN.forward -- VMN/VMN_model.py:83-113.  frames[i]: [B,4,H,W]; masks[i]: [B,1,H,W].  force_idx[i]: see `encoder`."""
    S = len(frames)
    idxs, feats = [], []
    for i in range(S):
        ix, x6 = encoder(frames[i], sd, force_idx=None if force_idx is None else force_idx[i], ties=ties)
        idxs.append(ix)
        feats.append(decoder_head(ix, x6, sd))
    preds: List[Optional[torch.Tensor]] = [None] * S
    attb: List[Optional[torch.Tensor]] = [None] * S
    attf: List[Optional[torch.Tensor]] = [None] * S
    small: List[Optional[torch.Tensor]] = [None] * S
    for i in range(1, S - 1):
        preds[i], attb[i], attf[i], small[i] = decoder_tail(idxs[i], sd, feats[i], feats[i - 1], feats[i + 1], masks[i],
                                                            window)
    preds[0] = torch.zeros_like(preds[1])
    preds[-1] = torch.zeros_like(preds[-2])
    return preds, attb, attf, small, feats


def eval_preprocess(imgs, tris, dilate_kernel=None):
    """EvalModel.preprocess -- models/model.py:360-387 with TRIMAP_CHANNEL == 1 (:22-27): the trimap enters the network as
    one channel tri / 255.  Returns (x4 [B,S,4,H,W], trimask float [B,S,1,H,W])."""
    mean = torch.tensor(IMG_MEAN, device=imgs.device).reshape(1, 1, 3, 1, 1)
    std = torch.tensor(IMG_STD, device=imgs.device).reshape(1, 1, 3, 1, 1)
    norm = (imgs.float().flip([2]) * (1.0 / 255) - mean) / std
    st = tris.float() * (1.0 / 255)
    trimask = (st > 0) & (st < 1)
    if dilate_kernel is not None:
        k = dilate_kernel
        tm = trimask.float()
        B, S = tm.shape[:2]
        tm = F.max_pool2d(tm.reshape(B * S, 1, *tm.shape[-2:]), 2 * k + 1, 1, k)
        trimask = tm.reshape(B, S, 1, *tm.shape[-2:]).bool()
    return torch.cat([norm, st], dim=2), trimask.float()


def eval_forward(sd: SD, imgs, tris, dilate_kernel=None, window=7, return_aux=False, force_idx=None, ties=None):
    """EvalModel.forward for vmn_dim -- models/model.py:389-424.  Returns alphas [B,S,1,H,W]."""
    with torch.no_grad():
        x4, trimask = eval_preprocess(imgs, tris, dilate_kernel)
        S = imgs.shape[1]
        frames = [x4[:, i] for i in range(S)]
        masks = [trimask[:, i] for i in range(S)]
        preds, attb, attf, small, feats = vmn_forward(sd, frames, masks, window, force_idx, ties)
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


def fixture_sd_dim(seed: int = 5, device="cpu") -> SD:
    """Deterministic, well-conditioned weights with the reference's state_dict layout (113 keys, VMN_DIM.py): He-normal
    convolutions (activations stay O(1) through 26 ReLU layers), BatchNorm statistics near the identity, an alpha head
    scaled so that the matte is spread over (0, 1) instead of sitting on the clamp."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def conv_p(name, cin, cout, k, gain=2.0, bias=0.05):
        std = math.sqrt(gain / (cin * k * k))
        sd[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * std
        sd[name + ".bias"] = torch.randn((cout,), generator=g) * bias

    chans = {"1": (4, 64), "2": (64, 128), "3": (128, 256), "4": (256, 512), "5": (512, 512)}
    for stage in ENC_STAGES:
        for cname, bname in stage:
            s = cname[4]
            cin = chans[s][0] if cname.endswith("1") else chans[s][1]
            cout = chans[s][1]
            conv_p("encoder." + cname, cin, cout, 3)
            sd[f"encoder.{bname}.weight"] = 1.0 + 0.1 * torch.randn((cout,), generator=g)
            sd[f"encoder.{bname}.bias"] = 0.1 * torch.randn((cout,), generator=g)
            sd[f"encoder.{bname}.running_mean"] = 0.1 * torch.randn((cout,), generator=g)
            sd[f"encoder.{bname}.running_var"] = 1.0 + 0.2 * torch.rand((cout,), generator=g)
            sd[f"encoder.{bname}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    conv_p("encoder.conv6", 512, 4096, 7)
    # max-unpooling leaves 3 of 4 positions zero: the convolution behind it sees a quarter of the energy
    for name, cin, cout, k, gain in (("dconv6", 4096, 512, 1, 2.0), ("dconv5", 512, 512, 5, 8.0), ("dconv4", 512, 256, 5, 8.0),
                                     ("dconv3", 256, 128, 5, 8.0), ("dconv2", 128, 64, 5, 8.0), ("dconv1", 64, 64, 5, 8.0)):
        conv_p("decoder." + name, cin, cout, k, gain)
    conv_p("decoder.alpha_pred", 64, 1, 5, gain=0.15, bias=0.0)
    sd["decoder.alpha_pred.bias"] = torch.tensor([0.45])
    for c in ("key_conv", "query_conv", "value_conv"):
        conv_p("decoder.fam." + c, 256, 256, 3, gain=1.0)
    # registration order of the reference: encoder convs/bns interleaved, conv6, decoder convs, fam
    order: List[str] = []
    for stage in ENC_STAGES:
        for cname, bname in stage:
            order += [f"encoder.{cname}.weight", f"encoder.{cname}.bias"]
            order += [f"encoder.{bname}.{k}" for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    order += ["encoder.conv6.weight", "encoder.conv6.bias"]
    for name in ("dconv6", "dconv5", "dconv4", "dconv3", "dconv2", "dconv1", "alpha_pred"):
        order += [f"decoder.{name}.weight", f"decoder.{name}.bias"]
    for c in ("key_conv", "query_conv", "value_conv"):
        order += [f"decoder.fam.{c}.weight", f"decoder.fam.{c}.bias"]
    return {k: sd[k].to(device) for k in order}
