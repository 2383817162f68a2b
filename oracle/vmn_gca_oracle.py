"""CPU oracle for the TCVOM GCA+TAM frame-window hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tcvom_b200/`` imports this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker / the timed CPU baseline.

This is a from-scratch *functional* restatement (plain ``torch`` fp32 on the CPU, driven
directly by a ``NET.state_dict()``) of what the reference computes for ``vmn_gca``.
Each function cites the reference file:line it restates (paths relative to the
reference checkout, commit f5fa07a).

Parity pinning: the upstream repository ships no tests, fixtures or golden vectors
("parity unpinned" upstream).  This oracle is therefore pinned
against outputs of the reference itself, executed in the build container by
``tests/golden/make_golden.py`` (which imports the unmodified reference modules) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
function below against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5
IMG_MEAN = (0.485, 0.456, 0.406)
IMG_STD = (0.229, 0.224, 0.225)

ENC_LAYERS = (("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 4, 2),
              ("layer_bottleneck", 512, 2, 2))          # encoders/__init__.py:21-24
DEC_LAYERS = (("layer1", 256, 2), ("layer2", 128, 3), ("layer3", 64, 3), ("layer4", 32, 2))  # VMN_GCA.py:10


# --------------------------------------------------------------------------- primitives
def sn_sigma(sd: SD, p: str) -> torch.Tensor:
    """sigma = u^T W v with the stored u, v (no update) -- GCA/ops.py:38-45."""
    w = sd[p + ".module.weight_bar"]
    h = w.shape[0]
    return sd[p + ".module.weight_u"].dot(w.reshape(h, -1).mv(sd[p + ".module.weight_v"]))


class TrainMode:
    """Context manager switching the primitives below to the reference's ``.train()`` semantics:
    SpectralNorm runs one power-iteration step per call and stores u, v back into ``sd``
    (GCA/ops.py:25-36,74-80); BatchNorm2d uses batch statistics and updates ``running_mean`` /
    ``running_var`` / ``num_batches_tracked`` in ``sd`` (momentum 0.1).  ``sd`` values that are leaves with
    ``requires_grad`` receive gradients through ordinary autograd (u, v are constants, like ``.data``)."""
    active = False

    def __enter__(self):
        TrainMode.active = True
        return self

    def __exit__(self, *exc):
        TrainMode.active = False
        return False


def sn_weight(sd: SD, p: str) -> torch.Tensor:
    """SpectralNorm weight: W_bar / sigma.  Eval: stored u, v (GCA/ops.py:38-45,74-80); train
    (``TrainMode``): one power iteration first, u/v updated in place (GCA/ops.py:25-36)."""
    if TrainMode.active:
        w = sd[p + ".module.weight_bar"]
        u, v = sd[p + ".module.weight_u"], sd[p + ".module.weight_v"]
        h = w.shape[0]
        with torch.no_grad():
            wm = w.detach().reshape(h, -1)
            nv = wm.t().mv(u)
            v = nv / (nv.norm() + 1e-12)
            nu = wm.mv(v)
            u = nu / (nu.norm() + 1e-12)
            # rebind (like the reference's ``u.data = ...``): autograd keeps the tensors of earlier calls
            sd[p + ".module.weight_u"], sd[p + ".module.weight_v"] = u, v
        sigma = u.dot(w.reshape(h, -1).mv(v))
        return w / sigma
    return sd[p + ".module.weight_bar"] / sn_sigma(sd, p)


def bn(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """BatchNorm2d: running statistics in eval mode, batch statistics (+ running-stat update) under
    ``TrainMode``."""
    if TrainMode.active:
        if p + ".num_batches_tracked" in sd:
            sd[p + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                            sd[p + ".weight"], sd[p + ".bias"], True, 0.1, BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def bn_affine(sd: SD, p: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scale, shift) such that eval BN(x) = x*scale + shift."""
    s = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    return s, sd[p + ".bias"] - sd[p + ".running_mean"] * s


def sn_conv(x, sd, p, stride=1, padding=1):
    return F.conv2d(x, sn_weight(sd, p), None, stride, padding)


# --------------------------------------------------------------------------- encoder
def enc_block(x, sd, p, stride, has_down):
    """encoders/resnet_enc.py:33-49 (BasicBlock) and :104-115 (ResNet-D downsample)."""
    out = torch.relu(bn(sn_conv(x, sd, p + ".conv1", stride), sd, p + ".bn1"))
    out = bn(sn_conv(out, sd, p + ".conv2"), sd, p + ".bn2")
    idt = x
    if has_down:
        idt = F.avg_pool2d(x, 2, stride)
        idt = bn(sn_conv(idt, sd, p + ".downsample.1", 1, 0), sd, p + ".downsample.2")
    return torch.relu(out + idt)


def shortcut(x, sd, p):
    """res_gca_enc.py:47-55 -- note conv -> ReLU -> BN order."""
    x = bn(torch.relu(sn_conv(x, sd, p + ".0")), sd, p + ".2")
    return bn(torch.relu(sn_conv(x, sd, p + ".3")), sd, p + ".5")


def guidance_head(rgb, sd, p):
    """res_gca_enc.py:20-33 -- three reflect-padded stride-2 convs, ReLU then BN."""
    x = rgb
    for ci, bi in ((1, 3), (5, 7), (9, 11)):
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        x = bn(torch.relu(sn_conv(x, sd, f"{p}.{ci}", 2, 0)), sd, f"{p}.{bi}")
    return x


def gca_patches(x, kernel, stride):
    """GuidedCxtAtten.extract_patches (ops.py:231-238) flattened to [B, P, C*k*k]
    with (c, kh, kw) fastest order, P = patch positions in raster order."""
    left = (kernel - stride + 1) // 2
    right = (kernel - stride) // 2
    xp = F.pad(x, (left, right, left, right), mode="reflect")
    cols = F.unfold(xp, kernel, stride=stride)          # [B, C*k*k, P], (c,kh,kw) order
    return cols.transpose(1, 2).contiguous()


def gca_attention(sd: SD, p: str, im_fea, feat, unknown, return_attn=False):
    """GuidedCxtAtten.forward, GCA/ops.py:106-229, written as plain attention + fold.

    im_fea  [B,128,h,w]  guidance feature (OS8)
    feat    [B,128,h,w]  alpha feature (OS8)
    unknown [B,1,h,w]    unknown-region indicator (OS8)
    """
    B, C, h, w = feat.shape
    g = F.conv2d(im_fea, sd[p + ".guidance_conv.weight"], sd[p + ".guidance_conv.bias"])  # :108
    g = g[:, :, ::2, ::2]                                # nearest 1/2 (:121)
    u = unknown[:, :, ::2, ::2]                          # :137
    hh, ww = g.shape[2:]
    P = hh * ww
    Q = gca_patches(g, 3, 1)                             # [B,P,576]  (:125-130)
    nrm = torch.sqrt((Q * Q).sum(-1, keepdim=True))     # l2_norm (:249-259)
    K = Q / torch.clamp(nrm, min=1e-4)                   # :174-175
    um = u.mean(dim=[2, 3])                              # [B,1]   (:139)
    km = 1.0 - um
    s_u = torch.clamp(torch.sqrt(um / km), 0.1, 10)      # :141
    s_k = torch.clamp(torch.sqrt(km / um), 0.1, 10)      # :142
    mm = (gca_patches(u, 3, 1).mean(-1) > 0).float()     # [B,P]   (:148-156)
    V = gca_patches(feat, 4, 2)                          # [B,P,2048] (:112-118)
    outs = []
    attn = []
    eye = torch.eye(P)
    for b in range(B):
        S = Q[b] @ K[b].t()                              # S[q,p]  (:177)
        scale = s_u[b, 0] * mm[b] + s_k[b, 0] * (1 - mm[b])   # per key p (:186)
        S = S * scale[None, :] + eye * (-1e4) * mm[b][None, :]  # :188
        A = torch.softmax(S, dim=1)                      # over keys (:190)
        if return_attn:
            attn.append(A)
        O = A @ V[b]                                     # [P(q), 2048]
        Y = F.fold(O.t().unsqueeze(0), (h, w), 4, padding=1, stride=2) / 4.0   # :204
        outs.append(Y)
    y = torch.cat(outs, 0)
    y = bn(F.conv2d(y, sd[p + ".W.0.weight"]), sd, p + ".W.1") + feat          # :227
    if return_attn:
        return y, attn
    return y


def encoder(x, sd: SD, p="encoder"):
    """ResGuidedCxtAtten.forward -- encoders/res_gca_enc.py:57-90."""
    out = torch.relu(bn(sn_conv(x, sd, p + ".conv1", 2), sd, p + ".bn1"))
    x1 = torch.relu(bn(sn_conv(out, sd, p + ".conv2", 1), sd, p + ".bn2"))
    out = torch.relu(bn(sn_conv(x1, sd, p + ".conv3", 2), sd, p + ".bn3"))
    im_fea = guidance_head(x[:, :3], sd, p + ".guidance_head")
    unknown = x[:, 4:5, ::8, ::8]                        # nearest 1/8 of one-hot channel 1 (:71)
    feats = []
    cur = out
    for name, planes, blocks, stride in ENC_LAYERS:
        if name == "layer3":
            cur = gca_attention(sd, p + ".gca", im_fea, cur, unknown)   # :77
            feats[-1] = cur                              # x3 is the post-GCA tensor
        for i in range(blocks):
            cur = enc_block(cur, sd, f"{p}.{name}.{i}", stride if i == 0 else 1,
                            i == 0 and stride != 1)
        feats.append(cur)
    x2, x3, x4, emb = feats
    fea = (shortcut(x, sd, p + ".shortcut.0"), shortcut(x1, sd, p + ".shortcut.1"),
           shortcut(x2, sd, p + ".shortcut.2"), shortcut(x3, sd, p + ".shortcut.3"),
           shortcut(x4, sd, p + ".shortcut.4"))
    return emb, dict(shortcut=fea, image_fea=im_fea, unknown=unknown)


# --------------------------------------------------------------------------- decoder
def dec_block(x, sd, p, up):
    """decoders/resnet_dec.py:43-59 (BasicBlock) and :107-124 (upsample path)."""
    if up:
        out = F.conv_transpose2d(x, sn_weight(sd, p + ".conv1"), None, 2, 1)
    else:
        out = sn_conv(x, sd, p + ".conv1")
    out = F.leaky_relu(bn(out, sd, p + ".bn1"), 0.2)
    out = bn(sn_conv(out, sd, p + ".conv2"), sd, p + ".bn2")
    idt = x
    if up:
        idt = F.interpolate(x, scale_factor=2, mode="nearest")
        idt = bn(sn_conv(idt, sd, p + ".upsample.1", 1, 0), sd, p + ".upsample.2")
    return F.leaky_relu(out + idt, 0.2)


def dec_layer(x, sd, p, blocks):
    for i in range(blocks):
        x = dec_block(x, sd, f"{p}.{i}", i == 0)
    return x


def decoder_head(emb, mid, sd: SD, p="decoder"):
    """ResGuidedCxtAtten_FAM_Dec.forward(extract_feature=True) -- VMN_GCA.py:27-34."""
    fea1, fea2, fea3, fea4, fea5 = mid["shortcut"]
    x = dec_layer(emb, sd, p + ".layer1", 2) + fea5
    x = dec_layer(x, sd, p + ".layer2", 3) + fea4
    return gca_attention(sd, p + ".gca", mid["image_fea"], x, mid["unknown"])


def tam(sd: SD, p: str, x, xb, xf, mask, window=7):
    """FeatureAggregationModule.forward -- VMN/VMN_model.py:18-68, as dense-then-mask.

    Returns (feat [B,C,H,W], attb [B,w*w,H*W], attf, mask_small bool [B,1,H,W])."""
    B, C, H, W = x.shape
    m = F.interpolate(mask, size=(H, W), mode="nearest").bool()       # :22
    q = F.conv2d(x, sd[p + ".query_conv.weight"], sd[p + ".query_conv.bias"], 1, 1)
    v = F.conv2d(x, sd[p + ".value_conv.weight"], sd[p + ".value_conv.bias"], 1, 1)
    mf = m.reshape(B, 1, H * W).float()

    def attend(t):
        k = F.conv2d(t, sd[p + ".key_conv.weight"], sd[p + ".key_conv.bias"], 1, 1)
        ku = F.unfold(k, window, padding=window // 2).reshape(B, C, window * window, H * W)
        logit = (q.reshape(B, C, 1, H * W) * ku).sum(1) / math.sqrt(C)   # [B,w2,N]  (:46)
        att = torch.softmax(logit, dim=1)                                 # :50
        agg = (att.unsqueeze(1) * ku).sum(2)                              # keys aggregated (:53)
        return (agg * mf).reshape(B, C, H, W), logit * mf                 # zero outside mask (:47-55)

    ab, lb = attend(xb)
    af, lf = attend(xf)
    return v + ab + af, lb, lf, m


def decoder_tail(mid, sd: SD, x, xb, xf, mask, window=7, p="decoder"):
    """ResGuidedCxtAtten_FAM_Dec.forward(extract_feature=False) -- VMN_GCA.py:35-49."""
    fea1, fea2, fea3, fea4, fea5 = mid["shortcut"]
    x, attb, attf, sm = tam(sd, p + ".fam", x, xb, xf, mask, window)
    x = dec_layer(x, sd, p + ".layer3", 3) + fea3
    x = dec_layer(x, sd, p + ".layer4", 2) + fea2
    x = F.conv_transpose2d(x, sn_weight(sd, p + ".conv1"), None, 2, 1)
    x = F.leaky_relu(bn(x, sd, p + ".bn1"), 0.2) + fea1
    x = F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], 1, 1)
    return (torch.tanh(x) + 1.0) / 2.0, attb, attf, sm


def vmn_forward(sd: SD, frames: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], window=7, freeze_backbone=False):
    """VMN.forward -- VMN/VMN_model.py:83-113.  frames[i]: [B,6,H,W]; masks[i]: [B,1,H,W].

    ``freeze_backbone`` (VMN_model.py:77-81,99-103, VMN_GCA.py:18-24): the encoder and the decoder's feature-extraction half
    (layer1, layer2, gca) run in eval mode (running-statistics BatchNorm, spectral norm without a power iteration) under
    no_grad, whatever mode the rest of the network is in."""
    S = len(frames)
    mids, feats = [], []
    was_train = TrainMode.active
    for i in range(S):
        if freeze_backbone:
            TrainMode.active = False
            try:
                with torch.no_grad():
                    emb, mid = encoder(frames[i], sd)
                    feat = decoder_head(emb, mid, sd)
            finally:
                TrainMode.active = was_train
        else:
            emb, mid = encoder(frames[i], sd)
            feat = decoder_head(emb, mid, sd)
        mids.append(mid)
        feats.append(feat)
    preds: List[Optional[torch.Tensor]] = [None] * S
    attb: List[Optional[torch.Tensor]] = [None] * S
    attf: List[Optional[torch.Tensor]] = [None] * S
    small: List[Optional[torch.Tensor]] = [None] * S
    for i in range(1, S - 1):
        preds[i], attb[i], attf[i], small[i] = decoder_tail(
            mids[i], sd, feats[i], feats[i - 1], feats[i + 1], masks[i], window)
    preds[0] = torch.zeros_like(preds[1])
    preds[-1] = torch.zeros_like(preds[-2])
    return preds, attb, attf, small, feats


# --------------------------------------------------------------------------- task wrapper
def eval_preprocess(imgs, tris, dilate_kernel=None):
    """EvalModel.preprocess -- models/model.py:360-387 (TRIMAP_CHANNEL == 3 branch).

    imgs [B,S,3,H,W] BGR 0..255; tris [B,S,1,H,W] in {0,128,255}.
    Returns (x6 [B,S,6,H,W], trimask float [B,S,1,H,W])."""
    mean = torch.tensor(IMG_MEAN, device=imgs.device).reshape(1, 1, 3, 1, 1)
    std = torch.tensor(IMG_STD, device=imgs.device).reshape(1, 1, 3, 1, 1)
    scaled = imgs.float().flip([2]) * (1.0 / 255)
    norm = (scaled - mean) / std
    st = tris.float() * (1.0 / 255)
    trimask = (st > 0) & (st < 1)
    if dilate_kernel is not None:
        k = dilate_kernel
        tm = trimask.float()
        B, S = tm.shape[:2]
        tm = F.max_pool2d(tm.reshape(B * S, 1, *tm.shape[-2:]), 2 * k + 1, 1, k)
        trimask = tm.reshape(B, S, 1, *tm.shape[-2:]).bool()
    cls = torch.where(trimask, torch.ones_like(st), 2 * st).long()     # 0 bg, 1 unknown, 2 fg
    onehot = F.one_hot(cls.squeeze(2), 3).permute(0, 1, 4, 2, 3).float()
    return torch.cat([norm, onehot], dim=2), trimask.float()


def eval_forward(sd: SD, imgs, tris, dilate_kernel=None, window=7, return_aux=False):
    """EvalModel.forward for vmn_gca -- models/model.py:389-424.  Returns alphas [B,S,1,H,W]."""
    with torch.no_grad():
        x6, trimask = eval_preprocess(imgs, tris, dilate_kernel)
        S = imgs.shape[1]
        frames = [x6[:, i] for i in range(S)]
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


# --------------------------------------------------------------------------- FullModel_VMD (eval-mode forward)
def l1_mask(x, y, mask, epsilon=1.001e-5):
    """utils/loss_func.py:9-22 (normalize=True, mask given)."""
    res = torch.abs(x - y) * mask
    b, c, h, w = y.shape
    safe = torch.sum((mask > epsilon).float()).clamp(epsilon, b * c * h * w + 1)
    return torch.sum(res) / safe


def train_preprocess(a, fg, bg, radii, eps=0.0):
    """FullModel.preprocess + make_trimap, TRIMAP_CHANNEL == 3 -- models/model.py:54-92.
    a [B,S,1,H,W], fg/bg [B,S,3,H,W] (BGR 0..255); radii: dilation radius per sample (the reference draws
    torch.randint(0, 26) per sample when DILATION_KERNEL is None, model.py:62)."""
    mean = torch.tensor(IMG_MEAN, device=a.device).reshape(1, 1, 3, 1, 1)
    std = torch.tensor(IMG_STD, device=a.device).reshape(1, 1, 3, 1, 1)
    gts = a * (1.0 / 255)
    fgs = fg.flip([2]) * (1.0 / 255)
    bgs = bg.flip([2]) * (1.0 / 255)
    imgs = fgs * gts + bgs * (1.0 - gts)
    alpha = torch.where(gts < eps, torch.zeros_like(gts), gts)
    alpha = torch.where(alpha > 1 - eps, torch.ones_like(alpha), alpha)
    raw = ((alpha > 0) & (alpha < 1.0)).float()
    tm = []
    for i in range(a.shape[0]):
        r = int(radii[i])
        tm.append(F.max_pool2d(raw[i], kernel_size=2 * r + 1, stride=1, padding=r))
    trimask = torch.stack(tm)
    cls = torch.where(trimask > 0.5, torch.ones_like(alpha), 2 * alpha).long()
    onehot = F.one_hot(cls.squeeze(2), 3).permute(0, 1, 4, 2, 3).float()
    norm = (imgs - mean) / std
    return dict(imgs=imgs, fgs=fgs, bgs=bgs, gts=gts, tris=onehot, trimask=trimask, x6=torch.cat([norm, onehot], 2))


def vmd_losses(pp, preds, attb, attf, small, window=7, att_thres=0.3, label_smooth=0.2, rows=None):
    """The loss half of FullModel_VMD.forward -- models/model.py:94-127 (L_im), :285-323 (L_af), :326-345 (L_tc) --
    on the batch rows ``rows`` (a slice; default: the whole batch).  Returns (L_alpha, L_dt, L_att, alphas, comps)."""
    sl = rows if rows is not None else slice(None)
    gts, trimask = pp["gts"][sl], pp["trimask"][sl]
    fgs, bgs = pp["fgs"][sl], pp["bgs"][sl]
    B, S = gts.shape[:2]
    preds = [p[sl] for p in preds]
    alphas: List[Optional[torch.Tensor]] = [None] * S
    comps: List[Optional[torch.Tensor]] = [None] * S
    L_alpha = []
    for c in range(1, S - 1):
        m = trimask[:, c].float()
        refine = torch.where(m.bool(), preds[c], gts[:, c])
        alphas[c] = refine
        comps[c] = fgs[:, c] * refine + bgs[:, c] * (1.0 - refine)
        L_alpha.append(l1_mask(refine, gts[:, c], m))
    L_alpha = sum(L_alpha) / float(len(L_alpha))
    for i in (0, S - 1):
        alphas[i] = torch.zeros_like(alphas[1])
        comps[i] = torch.zeros_like(comps[1])
    alphas_t = torch.stack(alphas, 1).clamp(0, 1)
    comps_t = torch.stack(comps, 1).clamp(0, 1)
    # attention-map loss
    H8, W8 = gts.shape[-2] // 8, gts.shape[-1] // 8
    L_att = []
    bce = torch.nn.BCEWithLogitsLoss(reduction="mean")
    for c in range(1, S - 1):
        bgt = F.avg_pool2d(gts[:, c - 1], 8, 8)
        fgt = F.avg_pool2d(gts[:, c + 1], 8, 8)
        cgt = F.avg_pool2d(gts[:, c], 8, 8)
        m = small[c][sl].reshape(B, -1)
        if m.float().sum() == 0:
            L_att.append(torch.zeros_like(L_alpha))
            continue
        bb = attb[c][sl].reshape(B, -1, H8 * W8).permute(1, 0, 2)[:, m]
        ff = attf[c][sl].reshape(B, -1, H8 * W8).permute(1, 0, 2)[:, m]
        bu = F.unfold(bgt, window, padding=window // 2).reshape(B, -1, H8 * W8).permute(1, 0, 2)[:, m]
        fu = F.unfold(fgt, window, padding=window // 2).reshape(B, -1, H8 * W8).permute(1, 0, 2)[:, m]
        cg = cgt.reshape(B, 1, H8 * W8).permute(1, 0, 2)[:, m]
        tb = (torch.abs(cg - bu) < att_thres).float() * (1 - label_smooth)
        tf = (torch.abs(cg - fu) < att_thres).float() * (1 - label_smooth)
        L_att.append((bce(bb, tb) + bce(ff, tf)) / 2.0)
    L_att = sum(L_att) / float(len(L_att))
    if S >= 5:
        L_dt = []
        for c in range(1, S - 2):
            L_dt.append(l1_mask(alphas_t[:, c] - alphas_t[:, c + 1], gts[:, c] - gts[:, c + 1], trimask[:, c]))
        L_dt = sum(L_dt) / float(len(L_dt))
    else:
        L_dt = torch.zeros_like(L_att)
    return L_alpha, L_dt, L_att, alphas_t, comps_t


def full_vmd_forward(sd: SD, a, fg, bg, radii, window=7, att_thres=0.3, label_smooth=0.2, eps=0.0, train=False,
                     rank_rows=None, freeze_backbone=False):
    """FullModel_VMD.forward for vmn_gca -- models/model.py:258-357 (single_image_loss :94-127, L_att :285-323,
    _dtSSD :326-345).  Returns the reference's 12-list.  ``train=False``: network in eval mode under no_grad
    (pred_vmn.py:107-116); ``train=True``: ``.train()`` semantics with autograd enabled (train_ddp.py:52-65) --
    ``sd`` is updated in place (u, v, running statistics) and its ``requires_grad`` leaves get gradients when
    the caller backpropagates the returned losses.

    ``rank_rows`` (list of batch slices) emulates DistributedDataParallel + SyncBatchNorm (train_ddp.py:270-280):
    the network sees the whole batch (global BatchNorm statistics) while every rank computes the losses on its own
    rows; the first five entries are then lists with one loss per rank (DDP averages the ranks' gradients, i.e.
    differentiates the mean of the per-rank totals)."""
    with torch.no_grad():
        pp = train_preprocess(a, fg, bg, radii, eps)
    import contextlib
    with (TrainMode() if train else contextlib.nullcontext()), torch.set_grad_enabled(bool(train)):
        S = a.shape[1]
        frames = [pp["x6"][:, i] for i in range(S)]
        masks = [pp["trimask"][:, i] for i in range(S)]
        preds, attb, attf, small, _ = vmn_forward(sd, frames, masks, window, freeze_backbone=freeze_backbone and train)
        gts, trimask = pp["gts"], pp["trimask"]
        tris_vis = torch.where(trimask.bool(), torch.ones_like(gts) * 128 * (1.0 / 255), gts)
        if rank_rows is not None:
            per = [vmd_losses(pp, preds, attb, attf, small, window, att_thres, label_smooth, rows=r) for r in rank_rows]
            zero = [torch.zeros_like(p[0]) for p in per]
            return [[p[0] for p in per], zero, [z.clone() for z in zero], [p[1] for p in per], [p[2] for p in per],
                    pp["imgs"], tris_vis, torch.cat([p[3] for p in per]), torch.cat([p[4] for p in per]), gts,
                    pp["fgs"], pp["bgs"]]
        L_alpha, L_dt, L_att, alphas_t, comps_t = vmd_losses(pp, preds, attb, attf, small, window, att_thres,
                                                             label_smooth)
        zero = torch.zeros_like(L_alpha)
    return [L_alpha, zero, zero.clone(), L_dt, L_att, pp["imgs"], tris_vis, alphas_t, comps_t, gts, pp["fgs"], pp["bgs"]]
