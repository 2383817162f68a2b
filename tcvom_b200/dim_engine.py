"""Host-side engine for the ``vmn_dim`` frame-window forward.

Same machinery as ``engine.GcaVmnEngine`` (weight cache, recorded per-shape plans replayed as one CUDA graph, the
tcgen05 / CUDA-core convolution dispatcher, the TAM operator); this subclass adds the DIM kernel program:

  per-frame part   VMN_model.py:93-98 -> VMN_DIM.py:48-72 (VGG-16/BN encoder, 2x2 max pooling with indices, 7x7 conv6),
                   VMN_DIM.py:109-119 (dconv6, unpool5 + dconv5, unpool4 + dconv4: the OS8 feature the TAM reads)
  per-centre part  VMN_model.py:107-110 -> VMN_DIM.py:120-136 (TAM(256), unpool3..1 + dconv3..1, 5x5 alpha head, clamp)

Layer semantics: encoder convolutions carry a bias AND an eval BatchNorm; both fold into one affine of the conv
epilogue, ``(conv + c) * s + b = conv * s + (b + c * s)``.  The 5x5 / 7x7 convolutions run on the <= 3x3-tap tcgen05
kernels as chains of tap groups: every launch adds its taps to the partial sum of the previous one through the
epilogue's residual input, the last one adds the bias and applies the activation (``conv_k``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch

from . import _cabi
from ._cabi import ACT_CLAMP01, ACT_NONE, ACT_RELU, PAD_ZERO
from .dim_modules import ENC_STAGES
from .engine import Act, GcaVmnEngine

FOLDED = "#b"          # BatchNorm affine with the convolution bias folded in
C11 = "encoder.conv11#c32"
HEAD = "decoder.alpha_pred"


class DimVmnEngine(GcaVmnEngine):
    """Owns derived device state for one ``vmn_dim`` VMN module on one device."""

    def __init__(self, window: int):
        super().__init__(window)
        self.s2d_stride2 = False          # no stride-2 convolutions in this network
        self._hold: list = []
        import os
        # alpha head as one HBM-bound kernel (default) or as the zero-padded 32-output-channel tensor-core conv chain
        self.head_direct = os.environ.get("TCV_HEAD_DIRECT", "1") == "1"
        # conv11 (4 -> 64) reads a 32-channel zero-padded copy of the input so that it runs on the CTA-pair tcgen05 kernel
        # instead of the CUDA-core one (3.1 -> ~0.5 ms per 1088x1920 window); "0": 8-channel input on the CUDA cores
        self.conv11_pad32 = os.environ.get("TCV_DIM_CONV11_PAD32", "1") == "1"

    # ------------------------------------------------------------------ weights
    def refresh_weights(self, net: torch.nn.Module, force=False) -> None:
        before = self._fingerprint
        super().refresh_weights(net, force)
        if force or self._fingerprint is not before:
            self._derive_dim()

    def _derive_dim(self) -> None:
        """Bias-folded BatchNorm affines of the encoder and the zero-padded 32-output-channel alpha head, engine-owned
        and updated in place (recorded plans keep their pointers)."""
        named = self._named()
        dev = self.device
        with torch.no_grad():
            for _, convs in ENC_STAGES:
                for cname, bname, _, _ in convs:
                    s, b = self.aff["encoder." + bname]
                    key = "encoder." + bname + FOLDED
                    own = self.aff.get(key)
                    if own is None or own[0].device != dev:
                        own = self.aff[key] = (torch.empty_like(s), torch.empty_like(b))
                    own[0].copy_(s)
                    torch.addcmul(b, named[f"encoder.{cname}.bias"], s, out=own[1])
        w, b = named[HEAD + ".weight"], named[HEAD + ".bias"]
        L, st = _cabi.lib(), self._stream_ptr()
        self._pack_head32(L, st, HEAD, w, b)
        # conv11 with its input channels zero-padded to 32 (weights [9][32][64] + the tensor-core copy), updated in place
        w11 = named["encoder.conv11.weight"]
        ent = self.w.get(C11)
        if ent is None or ent["w"].device != dev:
            ent = self.w[C11] = dict(w=torch.empty((9, 32, 64), dtype=torch.float32, device=dev),
                                     w_tc=torch.empty((2, 9, 64, 32), dtype=torch.bfloat16, device=dev),
                                     cout=64, cin=32, cin_real=w11.shape[1], k=3, transposed=False)
        _cabi.check(L.tcv_sn_fold_pack(w11.data_ptr(), None, None, 64, w11.shape[1], 3, 3, 0, 32, ent["w"].data_ptr(), None,
                                       st), "sn_fold_pack")
        _cabi.check(L.tcv_pack_weight_tc(ent["w"].data_ptr(), 9, 32, 64, ent["w_tc"].data_ptr(), st), "pack_weight_tc")

    # ------------------------------------------------------------------ operators
    def maxpool2(self, x: Act) -> Tuple[Act, torch.Tensor]:
        assert x.plane == x.n * x.img_elems
        y = self._act(x.n, x.h // 2, x.w // 2, x.c)
        idx = self._empty((x.n, x.h // 2, x.w // 2, x.c), torch.uint8)
        self._call("tcv_maxpool2_idx", x.ptr, x.n, x.h, x.w, x.c, y.ptr, idx.data_ptr(),
                   meta=dict(kind="tcv_maxpool2_idx", bytes=x.n * x.img_elems * 5 + x.n * x.img_elems // 4))
        return y, idx

    def unpool2(self, x: Act, idx_ptr: int) -> Act:
        assert x.plane == x.n * x.img_elems
        y = self._act(x.n, 2 * x.h, 2 * x.w, x.c)
        self._call("tcv_maxunpool2", x.ptr, idx_ptr, y.n, y.h, y.w, y.c, y.ptr,
                   meta=dict(kind="tcv_maxunpool2", bytes=x.n * x.img_elems * 5 + 4 * y.n * y.img_elems))
        return y

    def conv_k(self, x: Act, wkey: str, *, act=ACT_NONE) -> Act:
        """k x k convolution (k = 1, 3, 5, 7; zero padding k//2; bias) + activation.  k <= 3: one launch.  Otherwise the
        tap grid is cut into groups of at most 3 x 3 taps; launch i computes  partial_i = taps_i(x) + partial_{i-1}  and
        the last launch adds the bias and applies the activation."""
        ent = self.w[wkey]
        k, cout = ent["k"], ent["cout"]
        if k <= 3:
            return self.conv(x, wkey, bias=True, act=act)
        r = k // 2
        offs = list(range(-r, r + 1))
        groups = [offs[i:i + 3] for i in range(0, k, 3)]
        launches = [(gy, gx) for gy in groups for gx in groups]
        part = None
        for li, (gy, gx) in enumerate(launches):
            last = li == len(launches) - 1
            taps = [(dy, dx) for dy in gy for dx in gx]
            wtap = [(dy + r) * k + (dx + r) for dy in gy for dx in gx]
            y = self._act(x.n, x.h, x.w, cout)
            d = self._desc(x, ent["w"].data_ptr(), taps, 1, PAD_ZERO, y, x.h, x.w, cout, x.h, x.w, 1, 0, 1, 0, wkey, None,
                           last, act if last else ACT_NONE, part, 0, None, None, 0, wtap=wtap)
            meta = self._conv_meta(d, wkey, x, k, 1)
            self._call("tcv_conv2d", C.byref(d), meta=meta)
            part = y
        return part

    # ------------------------------------------------------------------ network program
    def per_frame(self, x8: Act) -> dict:
        """DIMEncoder.forward + DIMDecoder.forward(extract_feature=True) for all frames at once."""
        e, d = "encoder", "decoder"
        x = x8
        idxs: List[torch.Tensor] = []
        for _, convs in ENC_STAGES:
            for cname, bname, _, _ in convs:
                if cname == "conv11" and self.conv11_pad32 and self.use_tc_conv:
                    x32 = self._act(x8.n, x8.h, x8.w, 32)
                    x32.buf.zero_()              # channels 8..31 stay zero: written once, no recorded call touches them
                    self._hold.append(x32.buf)   # (a pooled plan must not hand this block to a later buffer)
                    self._call("tcv_copy_channels", x8.ptr, x8.plane, 8, 0, x32.ptr, x32.plane, 32, 0, 8,
                               x8.n * x8.h * x8.w, meta=dict(kind="tcv_copy_channels", bytes=x8.n * x8.h * x8.w * 64))
                    x = self.conv(x32, C11, bn=f"{e}.{bname}{FOLDED}", act=ACT_RELU)
                    continue
                x = self.conv(x, f"{e}.{cname}", bn=f"{e}.{bname}{FOLDED}", act=ACT_RELU)
            x, idx = self.maxpool2(x)
            idxs.append(idx)
        x6 = self.conv_k(x, e + ".conv6", act=ACT_RELU)
        t = self.conv_k(x6, d + ".dconv6", act=ACT_RELU)
        t = self.conv_k(self.unpool2(t, idxs[4].data_ptr()), d + ".dconv5", act=ACT_RELU)       # OS16
        t = self.conv_k(self.unpool2(t, idxs[3].data_ptr()), d + ".dconv4", act=ACT_RELU)       # OS8
        return dict(feat=t, idxs=idxs)

    def tail(self, pf: dict, n0: int, ncen: int, mask_ptr: int, mask_stride: int, H: int, W: int, pred_ptr: int,
             attb_ptr: int, attf_ptr: int, sm_ptr: int) -> None:
        """DIMDecoder.forward(extract_feature=False) for `ncen` consecutive centre frames starting at image n0+1."""
        d = "decoder"
        feat: Act = pf["feat"]
        x = feat.slice(n0 + 1, n0 + 1 + ncen)
        xb = feat.slice(n0, n0 + ncen)
        xf = feat.slice(n0 + 2, n0 + 2 + ncen)

        def idx_ptr(level: int) -> int:           # pooling indices of the centre frames (uint8, one byte per element)
            t = pf["idxs"][level]
            return t.data_ptr() + (n0 + 1) * t[0].numel()

        t = self.tam(d + ".fam", x, xb, xf, mask_ptr, mask_stride, H, W, attb_ptr, attf_ptr, sm_ptr)
        t = self.conv_k(self.unpool2(t, idx_ptr(2)), d + ".dconv3", act=ACT_RELU)                # OS4
        t = self.conv_k(self.unpool2(t, idx_ptr(1)), d + ".dconv2", act=ACT_RELU)                # OS2
        t = self.conv_k(self.unpool2(t, idx_ptr(0)), d + ".dconv1", act=ACT_RELU)                # OS1
        hw_ = self.w[HEAD]
        if self.head_direct and hw_["cin"] == 64 and hw_["cout"] == 1 and hw_["k"] == 5:
            # one HBM-bound pass: 5x5 conv to the single alpha channel + clamp (the padded tensor-core form below writes and
            # re-reads 32-channel full-resolution partial sums four times for one real channel)
            self._call("tcv_head_conv5_clamp01", t.ptr, t.plane, t.n, t.h, t.w, hw_["w"].data_ptr(),
                       self.bias[HEAD].data_ptr(), pred_ptr,
                       meta=dict(kind="tcv_head_conv5_clamp01", bytes=t.n * t.h * t.w * (4 * 64 + 4),
                                 flops=2 * t.n * t.h * t.w * 25 * 64))
            return
        z = self.conv_k(t, HEAD + self.HEAD32, act=ACT_CLAMP01)      # 64 -> 1 as 64 -> 32 (zero weights), channel 0 = alpha
        self._call("tcv_split_to_nchw", z.ptr, z.n, 1, z.h, z.w, z.c, z.plane, pred_ptr,
                   meta=dict(kind="tcv_split_to_nchw", bytes=z.n * z.h * z.w * 8))

    def window_program(self, x8: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        """Runs (and records) the whole VMN forward on preprocessed input.  trimask fp32 [B*S,H,W]."""
        ncen = S - 2
        N8 = (H // 8) * (W // 8)
        w2 = self.window * self.window
        pred = self._empty((B, ncen, 1, H, W))
        attb = self._empty((B, ncen, w2, N8))
        attf = self._empty((B, ncen, w2, N8))
        sm = self._empty((B, ncen, 1, H // 8, W // 8), torch.uint8)
        self._hold = []
        pf = self.per_frame(x8)
        pf["hold"] = self._hold
        for b in range(B):
            n0 = b * S
            self.tail(pf, n0, ncen, trimask.data_ptr() + 4 * (n0 + 1) * H * W, H * W, H, W,
                      pred[b].data_ptr(), attb[b].data_ptr(), attf[b].data_ptr(), sm[b].data_ptr())
        return dict(pred=pred, attb=attb, attf=attf, small_mask=sm, feat=pf["feat"], pf=pf)
