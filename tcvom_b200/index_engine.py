"""Host-side engine for the ``vmn_index`` frame-window forward.

Same machinery as ``engine.GcaVmnEngine`` (weight cache, recorded per-shape plans replayed as one CUDA graph, the tcgen05 /
CUDA-core convolution dispatcher, the TAM operator); this subclass adds the IndexNet kernel program:

  per-frame part   VMN_model.py:93-98 -> models/Index/net.py:189-228 (MobileNetV2 encoder whose strides are index pooling,
                   five index blocks, ASPP), VMN_Index.py:15-20 (decoder_layer6 / 5 / 4: the OS8 feature the TAM reads)
  per-centre part  VMN_model.py:107-110 -> VMN_Index.py:21-28 (TAM(32), decoder_layer3..0, pred)

Channel counts that are not multiples of 32 (16, 24, 144, 48; the 1-channel prediction) are zero-padded in the packed
weights and in the folded BatchNorm affines, so that every dense convolution qualifies for the tensor-core kernels and a
padded channel carries exact zeros (ReLU6(0 * s + 0) = 0).  5x5 convolutions run as chains of <= 3x3 tap groups
(partial sums through the residual input of the epilogue; the BatchNorm scale is applied by every launch, the shift by
the last one).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi
from ._cabi import ACT_NONE, ACT_RELU6, PAD_ZERO
from .engine import BN_EPS, Act, GcaVmnEngine
from .index_modules import ASPP_DILATIONS, DEC_LAYERS, INDEX_BLOCKS, IR_SETTING


def _pad32(c: int) -> int:
    return (c + 31) // 32 * 32


class IndexVmnEngine(GcaVmnEngine):
    """Owns derived device state for one ``vmn_index`` VMN module on one device."""

    def __init__(self, window: int):
        super().__init__(window)
        self.s2d_stride2 = False
        self.dw: Dict[str, torch.Tensor] = {}          # depthwise weights [9][c_pad]
        self.border: Dict[str, torch.Tensor] = {}      # relu6(BatchNorm shift) of a 1x1 expansion (see tcv_dwconv3x3)

    # ------------------------------------------------------------------ weights
    def refresh_weights(self, net: torch.nn.Module, force=False) -> None:
        if net is not self.net:
            self.net = net
            self._tensors = None
        named = self._named()
        dev = next(iter(named.values())).device
        self._check_device(dev)
        if self.device is not None and dev != self.device:
            self.w.clear(); self.aff.clear(); self.bias.clear(); self.dw.clear(); self.border.clear(); self.plans.clear()
            self._tensors = None
        self.device = dev
        fp = self._current_fingerprint()
        if not force and fp == self._fingerprint:
            return
        L = _cabi.lib()
        st = self._stream_ptr()
        for name, t in named.items():
            if t.dtype.is_floating_point and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"tcvom_b200: parameter {name} must be contiguous fp32")
        with torch.no_grad():
            for name, t in named.items():
                if name.endswith(".weight") and t.dim() == 4:
                    p = name[: -len(".weight")]
                    if p.startswith("decoder.fam.") or p == "encoder.layer0.0":
                        self._pack(L, st, p, t, None, None, None, transposed=False)       # 32 -> 32 (+ bias) / folded 4 -> 32
                        b = named.get(p + ".bias")
                        if b is not None:
                            self._own_bias(p, b)
                    elif t.shape[1] == 1 and t.shape[0] > 1:
                        c = t.shape[0]
                        own = self.dw.get(p)
                        if own is None or own.device != dev:
                            own = self.dw[p] = torch.zeros((9, _pad32(c)), dtype=torch.float32, device=dev)
                        own[:, :c].copy_(t.reshape(c, 9).t())
                    else:
                        self._pack_dense(L, st, p, t)
                elif name.endswith(".running_var"):
                    p = name[: -len(".running_var")]
                    c = t.numel()
                    cp = _pad32(c) if c > 1 else 8
                    if p not in self.aff or self.aff[p][0].device != dev:
                        self.aff[p] = (torch.zeros(cp, dtype=torch.float32, device=dev),
                                       torch.zeros(cp, dtype=torch.float32, device=dev))
                    s, b = self.aff[p]
                    _cabi.check(L.tcv_bn_fold(named[p + ".weight"].data_ptr(), named[p + ".bias"].data_ptr(),
                                              named[p + ".running_mean"].data_ptr(), t.data_ptr(), BN_EPS, c,
                                              s.data_ptr(), b.data_ptr(), st), "bn_fold")
            for p, (s, b) in self.aff.items():          # what the depthwise conv behind a 1x1 expansion sees in its border
                own = self.border.get(p)
                if own is None or own.device != dev:
                    own = self.border[p] = torch.empty_like(b)
                torch.clamp(b, 0.0, 6.0, out=own)
        self._fingerprint = fp

    def _pack_dense(self, L, st, p: str, w: torch.Tensor) -> None:
        cout, cin, kh, kw = w.shape
        cin_pad = _pad32(cin) if cin > 1 else 8
        cout_pad = _pad32(cout) if cout > 1 else 8
        ent = self.w.get(p)
        if ent is None or ent["w"].device != w.device:
            ent = self.w[p] = dict(w=torch.empty((kh * kw, cin_pad, cout_pad), dtype=torch.float32, device=w.device),
                                   cout=cout_pad, cout_real=cout, cin=cin_pad, cin_real=cin, k=kh, transposed=False)
        _cabi.check(L.tcv_ws_pack(w.data_ptr(), cout, cin, kh, kw, 0, cin_pad, cout_pad, ent["w"].data_ptr(), st), "ws_pack")
        if cin_pad % 32 == 0 and cout_pad % 32 == 0:
            if "w_tc" not in ent:
                ent["w_tc"] = torch.empty((2, kh * kw, cout_pad, cin_pad), dtype=torch.bfloat16, device=w.device)
            _cabi.check(L.tcv_pack_weight_tc(ent["w"].data_ptr(), kh * kw, cin_pad, cout_pad, ent["w_tc"].data_ptr(), st),
                        "pack_weight_tc")

    # ------------------------------------------------------------------ operators
    def cbr(self, x: Act, wkey: str, bnkey: Optional[str], act=ACT_RELU6, res1: Optional[Act] = None, stride=1) -> Act:
        """k x k convolution (padding k // 2; 4x4: stride 2, padding 1) + eval BatchNorm + activation (+ residual before the
        activation slot is unused here: the reference adds the block input after the linear 1x1)."""
        ent = self.w[wkey]
        k, cout = ent["k"], ent["cout"]
        assert ent["cin"] == x.c, (wkey, ent["cin"], x.c)
        if k == 4:
            assert stride == 2
            taps = [(ky - 1, kx - 1) for ky in range(4) for kx in range(4)]
            oh, ow = x.h // 2, x.w // 2
            groups = [(taps, list(range(16)))]
        else:
            assert stride == 1
            r = k // 2
            offs = list(range(-r, r + 1))
            cuts = [offs[i:i + 3] for i in range(0, k, 3)]
            groups = [([(dy, dx) for dy in gy for dx in gx], [(dy + r) * k + (dx + r) for dy in gy for dx in gx])
                      for gy in cuts for gx in cuts]
            oh, ow = x.h, x.w
        part: Optional[Act] = None
        for gi, (taps, wtap) in enumerate(groups):
            last = gi == len(groups) - 1
            y = self._act(x.n, oh, ow, cout)
            d = self._desc(x, ent["w"].data_ptr(), taps, stride, PAD_ZERO, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey, bnkey,
                           False, act if last else ACT_NONE, None, 0, None, None, 0, wtap=wtap)
            if not last:
                d.b1 = None                              # the shift once, the scale on every partial sum
            if part is not None:
                d.res1, d.res1_plane, d.res1_shift = part.ptr, part.plane, 0
            elif last and res1 is not None:
                assert (res1.n, res1.h, res1.w, res1.c) == (x.n, oh, ow, cout), wkey
                d.res1, d.res1_plane, d.res1_shift = res1.ptr, res1.plane, 0
            assert not (res1 is not None and len(groups) > 1)
            self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey, x, k, stride))
            part = y
        return part

    def dwconv(self, x: Act, wkey: str, bnkey: str, dil: int, border: Optional[str]) -> Act:
        w = self.dw[wkey]
        assert w.shape[1] == x.c and x.plane == x.n * x.img_elems, (wkey, w.shape, x.c)
        s, b = self.aff[bnkey]
        y = self._act(x.n, x.h, x.w, x.c)
        self._call("tcv_dwconv3x3", x.ptr, x.n, x.h, x.w, x.c, dil, w.data_ptr(), s.data_ptr(), b.data_ptr(),
                   self.border[border].data_ptr() if border is not None else None, ACT_RELU6, y.ptr,
                   meta=dict(kind="tcv_dwconv3x3", bytes=8 * x.n * x.img_elems, flops=18 * x.n * x.img_elems))
        return y

    def inverted_residual(self, x: Act, p: str, inp: int, oup: int, t: int) -> Act:
        c = p + ".conv"
        if t == 1:
            h = self.dwconv(x, c + ".0", c + ".1", 1, None)
            return self.cbr(h, c + ".3", c + ".4", act=ACT_NONE, res1=x if inp == oup else None)
        h = self.cbr(x, c + ".0", c + ".1")
        h = self.dwconv(h, c + ".3", c + ".4", 1, c + ".1")
        return self.cbr(h, c + ".6", c + ".7", act=ACT_NONE, res1=x if inp == oup else None)

    def layer(self, x: Act, p: str, setting) -> Act:
        t, inp, oup, n = setting
        for i in range(n):
            x = self.inverted_residual(x, f"{p}.{i}", inp if i == 0 else oup, oup, t)
        return x

    def index_pool(self, x: Act, p: str) -> Tuple[Act, Act, Act]:
        """index block + index pooling: returns (idx_en * x, 4 * avg_pool2(idx_en * x), idx_de)."""
        outs = []
        for i in range(1, 5):
            q = f"{p}.indexnet{i}"
            h = self.cbr(x, q + ".0", q + ".1", stride=2)
            outs.append(self.cbr(h, q + ".3", None, act=ACT_NONE))
        idx_en = self._act(x.n, x.h, x.w, x.c)
        idx_de = self._act(x.n, x.h, x.w, x.c)
        self._call("tcv_index_finish", outs[0].ptr, outs[1].ptr, outs[2].ptr, outs[3].ptr, x.n, x.h // 2, x.w // 2, x.c,
                   idx_en.ptr, idx_de.ptr, meta=dict(kind="tcv_index_finish", bytes=12 * x.n * x.img_elems))
        masked = self._act(x.n, x.h, x.w, x.c)
        pooled = self._act(x.n, x.h // 2, x.w // 2, x.c)
        self._call("tcv_index_pool", x.ptr, idx_en.ptr, x.n, x.h, x.w, x.c, masked.ptr, pooled.ptr,
                   meta=dict(kind="tcv_index_pool", bytes=13 * x.n * x.img_elems))
        return masked, pooled, idx_de

    def aspp(self, x: Act, p: str) -> Act:
        n, h, w = x.n, x.h, x.w
        cat = self._act(n, h, w, 5 * 256)
        b1 = self.cbr(x, p + ".aspp1.atrous_conv.0", p + ".aspp1.atrous_conv.1")
        self._copy(b1, 256, cat, 0)
        for i, d in enumerate(ASPP_DILATIONS[1:], start=2):
            q = f"{p}.aspp{i}.atrous_conv"
            hd = self.dwconv(x, q + ".0", q + ".1", d, None)
            self._copy(self.cbr(hd, q + ".3", q + ".4"), 256, cat, 256 * (i - 1))
        g = self._act(n, 1, 1, x.c)
        self._call("tcv_adaptive_avgpool", x.ptr, x.plane, n, h, w, x.c, x.c, 0, 1, g.ptr)
        g = self.cbr(g, p + ".global_avg_pool.1", p + ".global_avg_pool.2")
        # F.interpolate(nearest) of a 1 x 1 map = the same vector at every pixel (a bilinear resize of one sample is that too)
        self._call("tcv_bilinear", g.ptr, n, 1, 1, g.c, cat.ptr, cat.plane, h, w, cat.c, 4 * 256)
        return self.cbr(cat, p + ".bottleneck_conv.0", p + ".bottleneck_conv.1")

    def _copy(self, x: Act, c: int, out: Act, off: int) -> None:
        self._call("tcv_copy_channels", x.ptr, x.plane, x.c, 0, out.ptr, out.plane, out.c, off, c, x.n * x.h * x.w)

    def dec_block(self, dec: Act, dec_real: int, low: Act, low_real: int, idx: Optional[Act], p: str) -> Act:
        """IndexedUpsamlping: cat(idx * nearest_up(dec), low) -> 5x5 conv + BatchNorm + ReLU6."""
        cat_c = _pad32(dec_real + low_real)
        cat = self._act(low.n, low.h, low.w, cat_c)
        up = 1 if idx is not None else 0
        assert (dec.h << up, dec.w << up) == (low.h, low.w) and dec.n == low.n, p
        assert dec.plane == dec.n * dec.img_elems
        self._call("tcv_index_upcat", dec.ptr, dec.c, dec_real, up, idx.ptr if idx is not None else None,
                   idx.c if idx is not None else 0, idx.plane if idx is not None else 0, low.ptr, low.c, low.plane,
                   low_real, low.n, low.h, low.w, cat_c, cat.ptr,
                   meta=dict(kind="tcv_index_upcat", bytes=8 * low.n * low.h * low.w * cat_c))
        return self.cbr(cat, p + ".dconv.0", p + ".dconv.1")

    # ------------------------------------------------------------------ network program
    def per_frame(self, x8: Act) -> dict:
        e, d = "encoder", "decoder"
        l0 = self.conv(x8, e + ".layer0.0", bn=e + ".layer0.1", act=ACT_RELU6)
        l0, l0p, i0d = self.index_pool(l0, e + ".index0")
        l1 = self.layer(l0p, e + ".layer1", IR_SETTING[0])
        l2 = self.layer(l1, e + ".layer2", IR_SETTING[1])
        l2, l2p, i2d = self.index_pool(l2, e + ".index2")
        l3 = self.layer(l2p, e + ".layer3", IR_SETTING[2])
        l3, l3p, i3d = self.index_pool(l3, e + ".index3")
        l4 = self.layer(l3p, e + ".layer4", IR_SETTING[3])
        l4, l4p, i4d = self.index_pool(l4, e + ".index4")
        l5 = self.layer(l4p, e + ".layer5", IR_SETTING[4])
        l6 = self.layer(l5, e + ".layer6", IR_SETTING[5])
        l6, l6p, i6d = self.index_pool(l6, e + ".index6")
        l7 = self.layer(l6p, e + ".layer7", IR_SETTING[6])
        l = self.aspp(l7, e + ".dconv_pp")
        t = self.dec_block(l, 160, l6, 160, i6d, d + ".decoder_layer6")
        t = self.dec_block(t, 96, l5, 96, None, d + ".decoder_layer5")
        feat = self.dec_block(t, 64, l4, 64, i4d, d + ".decoder_layer4")
        return dict(feat=feat, l3=l3, i3d=i3d, l2=l2, i2d=i2d, l1=l1, l0=l0, i0d=i0d)

    def tail(self, pf: dict, n0: int, ncen: int, mask_ptr: int, mask_stride: int, H: int, W: int, pred_ptr: int,
             attb_ptr: int, attf_ptr: int, sm_ptr: int) -> None:
        d = "decoder"
        feat: Act = pf["feat"]
        x = feat.slice(n0 + 1, n0 + 1 + ncen)
        xb = feat.slice(n0, n0 + ncen)
        xf = feat.slice(n0 + 2, n0 + 2 + ncen)
        c = {k: pf[k].slice(n0 + 1, n0 + 1 + ncen) for k in ("l3", "i3d", "l2", "i2d", "l1", "l0", "i0d")}
        t = self.tam(d + ".fam", x, xb, xf, mask_ptr, mask_stride, H, W, attb_ptr, attf_ptr, sm_ptr)
        t = self.dec_block(t, 32, c["l3"], 32, c["i3d"], d + ".decoder_layer3")
        t = self.dec_block(t, 24, c["l2"], 24, c["i2d"], d + ".decoder_layer2")
        t = self.dec_block(t, 16, c["l1"], 16, None, d + ".decoder_layer1")
        t = self.dec_block(t, 32, c["l0"], 32, c["i0d"], d + ".decoder_layer0")
        t = self.cbr(t, d + ".pred.0.0", d + ".pred.0.1")                     # 32 -> 1 (8 with the padding) + BN + ReLU6
        z = self.cbr(t, d + ".pred.1", None, act=ACT_NONE)                    # 1 -> 1, 5x5, no bias
        self._call("tcv_split_to_nchw", z.ptr, z.n, 1, z.h, z.w, z.c, z.plane, pred_ptr,
                   meta=dict(kind="tcv_split_to_nchw", bytes=z.n * z.h * z.w * 8))

    def window_program(self, x8: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        ncen = S - 2
        N8 = (H // 8) * (W // 8)
        w2 = self.window * self.window
        pred = self._empty((B, ncen, 1, H, W))
        attb = self._empty((B, ncen, w2, N8))
        attf = self._empty((B, ncen, w2, N8))
        sm = self._empty((B, ncen, 1, H // 8, W // 8), torch.uint8)
        pf = self.per_frame(x8)
        for b in range(B):
            n0 = b * S
            self.tail(pf, n0, ncen, trimask.data_ptr() + 4 * (n0 + 1) * H * W, H * W, H, W,
                      pred[b].data_ptr(), attb[b].data_ptr(), attf[b].data_ptr(), sm[b].data_ptr())
        return dict(pred=pred, attb=attb, attf=attf, small_mask=sm, feat=pf["feat"], pf=pf)
