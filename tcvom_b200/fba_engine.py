"""Host-side engine for the ``vmn_fba`` frame-window forward.

Same machinery as ``engine.GcaVmnEngine`` (weight cache, recorded per-shape plans replayed as one CUDA graph, the
tcgen05 / CUDA-core convolution dispatcher, the TAM operator); this subclass adds the FBA kernel program:

  per-frame part   VMN_model.py:93-98 -> FBA/models.py:222-236 (ResnetDilated), VMN_FBA.py:20-33 (pyramid pooling)
  per-centre part  VMN_model.py:107-110 -> VMN_FBA.py:34-59 (TAM, three bilinear ups, 7-channel head, fusion)

Layer semantics: every encoder / pyramid / conv_up1..3 convolution is weight-standardised (layers_WS.py:13-23; folded
into the packed weights by ``tcv_ws_pack`` whenever the parameters change) and followed by GroupNorm(32)
(``tcv_gn_stats`` -> ``tcv_gn_finalize`` -> ``tcv_gn_apply`` with the activation, the residual add and the write into
a concatenation buffer fused into the last pass).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _cabi
from ._cabi import ACT_LEAKY001, ACT_NONE, ACT_RELU, PAD_ZERO
from .engine import Act, GcaVmnEngine
from .fba_modules import GN_GROUPS, PPM_SCALES, RES_LAYERS, block_config

GN_EPS = 1e-5
# decoder convolutions that are layers_WS.Conv2d (FBA/models.py:263-292); conv_up4.* and fam.* are plain nn.Conv2d
_WS_DECODER = ("decoder.ppm.", "decoder.conv_up1.", "decoder.conv_up2.", "decoder.conv_up3.")


# channel padding beyond the kernels' minimum: conv_up4.2 (32 -> 16) is computed as 32 -> 32 with zero weights / bias
# for the extra outputs so that it runs on the narrow-layer tcgen05 kernel instead of the CUDA cores; conv_up4.4 then
# reads 32 channels (zero weights for the padding)
_PAD_OVERRIDE = {"decoder.conv_up4.2": dict(cout_pad=32), "decoder.conv_up4.4": dict(cin_pad=32)}
STEM = "encoder.conv1"
STEM_S2D = STEM + "#s2d"


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class FbaVmnEngine(GcaVmnEngine):
    """Owns derived device state for one ``vmn_fba`` VMN module on one device."""

    def __init__(self, window: int):
        super().__init__(window)
        self.gn_params: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        # GroupNorm statistics from the epilogue of the producing convolution (tcv_conv_desc.stats) were built and measured
        # on B200 (1088x1920 window): the pass they remove costs 4.1 ms, the convolutions got 7.7 ms slower with one
        # accumulator copy (same-address fp64 atomics) and 4.6 ms with 32 copies + 1.2 ms for the wider finalize -- these
        # layers (1x1 bottleneck convs, 5-6 tiles per CTA pair) are epilogue-bound, the 62 shuffles + 2 atomics per
        # 32-channel chunk are not hidden behind the MMAs.  The engine therefore keeps the separate tcv_gn_stats pass; the
        # C-ABI option stays (tests/test_gpu_kernels.py::test_conv_epilogue_statistics).

    # ------------------------------------------------------------------ weights
    def refresh_weights(self, net: torch.nn.Module, force=False) -> None:
        """Standardises / packs the convolution weights when parameters changed (load_state_dict, device move).
        Packed buffers are updated in place so recorded plans stay valid."""
        if net is not self.net:
            self.net = net
            self._tensors = None
        named = self._named()
        dev = next(iter(named.values())).device
        self._check_device(dev)
        if self.device is not None and dev != self.device:
            self.w.clear(); self.bias.clear(); self.gn_params.clear(); self.plans.clear()
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
        for name, t in named.items():
            if not name.endswith(".weight"):
                continue
            p = name[: -len(".weight")]
            if t.dim() == 4:
                ws = p.startswith("encoder.") or p.startswith(_WS_DECODER)
                self._pack_fba(L, st, p, t, ws)
                b = named.get(p + ".bias")
                if b is not None:
                    cout_pad = self.w[p]["cout"]
                    if cout_pad != b.numel():
                        pads = self.__dict__.setdefault("_bias_pads", {})
                        if p not in pads or pads[p].device != dev:
                            pads[p] = torch.zeros(cout_pad, dtype=torch.float32, device=dev)
                        with torch.no_grad():
                            pads[p][: b.numel()].copy_(b)
                        self.bias[p] = pads[p]
                    else:
                        self._own_bias(p, b)
            elif t.dim() == 1:
                # engine-owned copies (see GcaVmnEngine._own_bias): plans record these pointers
                own = self.gn_params.get(p)
                gb = named[p + ".bias"]
                if own is None or own[0].device != dev or own[0].shape != t.shape:
                    own = self.gn_params[p] = (torch.empty_like(t), torch.empty_like(gb))
                with torch.no_grad():
                    own[0].copy_(t)
                    own[1].copy_(gb)
        self._fingerprint = fp

    def _pack_fba(self, L, st, p, w, standardize: bool) -> None:
        cout, cin, kh, kw = w.shape
        cin_pad = _round_up(cin, 8) if cin <= 32 else _round_up(cin, 32)
        cout_pad = _round_up(cout, 8)
        ov = _PAD_OVERRIDE.get(p, {})
        cin_pad, cout_pad = ov.get("cin_pad", cin_pad), ov.get("cout_pad", cout_pad)
        ent = self.w.get(p)
        if ent is None:
            ent = self.w[p] = dict(w=torch.empty((kh * kw, cin_pad, cout_pad), dtype=torch.float32, device=w.device),
                                   cout=cout_pad, cout_real=cout, cin=cin_pad, cin_real=cin, k=kh, transposed=False)
        _cabi.check(L.tcv_ws_pack(w.data_ptr(), cout, cin, kh, kw, 1 if standardize else 0, cin_pad, cout_pad,
                                  ent["w"].data_ptr(), st), "ws_pack")
        if cin_pad % 32 == 0 and cout_pad % 32 == 0:
            if "w_tc" not in ent:
                ent["w_tc"] = torch.empty((2, kh * kw, cout_pad, cin_pad), dtype=torch.bfloat16, device=w.device)
            _cabi.check(L.tcv_pack_weight_tc(ent["w"].data_ptr(), kh * kw, cin_pad, cout_pad, ent["w_tc"].data_ptr(), st),
                        "pack_weight_tc")
        if p == STEM and kh == 7:
            # the stem as a 16-tap 4x4 convolution over the 2x2 space-to-depth input (tcgen05 path)
            e2 = self.w.get(STEM_S2D)
            if e2 is None:
                e2 = self.w[STEM_S2D] = dict(
                    w=torch.empty((16, 4 * cin_pad, cout_pad), dtype=torch.float32, device=w.device),
                    w_tc=torch.empty((2, 16, cout_pad, 4 * cin_pad), dtype=torch.bfloat16, device=w.device),
                    cout=cout_pad, cout_real=cout, cin=4 * cin_pad, cin_real=4 * cin_pad, k=4, transposed=False)
            _cabi.check(L.tcv_s2d_pack_stem(ent["w"].data_ptr(), cin_pad, cout_pad, e2["w"].data_ptr(), st), "s2d_pack_stem")
            _cabi.check(L.tcv_pack_weight_tc(e2["w"].data_ptr(), 16, 4 * cin_pad, cout_pad, e2["w_tc"].data_ptr(), st),
                        "pack_weight_tc")

    # ------------------------------------------------------------------ operators
    def convf(self, x: Act, wkey: str, *, stride=1, dilation=1, bias=False, act=ACT_NONE, split_rows=False) -> Act:
        """k x k convolution (k in {1, 3}), padding = dilation * (k // 2), optional bias / activation epilogue.

        split_rows: one launch per filter row, chained through the residual input of the epilogue.  The tensor cores
        accumulate in fp32 with truncation, an error that grows linearly with the reduction length (measured on
        B200: 4e-5 at K = 4608, 3.5e-4 at K = 27648 for unit-variance outputs); conv_up1.0 (K = 9 * 3072) is therefore
        reduced in three K = 9216 pieces whose partial sums are added in the (round-to-nearest) epilogue."""
        ent = self.w[wkey]
        k, cout = ent["k"], ent["cout"]
        assert ent["cin"] == x.c, (wkey, ent["cin"], x.c)
        r = k // 2
        oh, ow = (x.h - 1) // stride + 1, (x.w - 1) // stride + 1
        rows = [[ky] for ky in range(k)] if (split_rows and k > 1) else [list(range(k))]
        prev: Optional[Act] = None
        for i, kys in enumerate(rows):
            last = i == len(rows) - 1
            taps = [((ky - r) * dilation, (kx - r) * dilation) for ky in kys for kx in range(k)]
            wtap = [ky * k + kx for ky in kys for kx in range(k)]
            y = self._act(x.n, oh, ow, cout)
            d = self._desc(x, ent["w"].data_ptr(), taps, stride, PAD_ZERO, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey,
                           None, bias and last, act if last else ACT_NONE, prev, 0, None, None, 0, wtap=wtap)
            self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey, x, k, stride))
            prev = y
        return prev

    def stem_s2d(self, x: Act, wkey: str = STEM) -> Act:
        """The 7x7 / stride-2 / pad-3 stem (resnet_GN_WS.py:98) on the tensor cores: 2x2 space-to-depth of the
        16-channel input, then ONE 16-tap (4x4, offsets -2..1) stride-1 convolution with K = 16 * 64."""
        key = wkey + "#s2d"
        ent = self.w[key]
        assert x.h % 2 == 0 and x.w % 2 == 0 and ent["cin"] == 4 * x.c
        oh, ow, cout = x.h // 2, x.w // 2, ent["cout"]
        xs = self._act(x.n, oh, ow, 4 * x.c)
        self._call("tcv_space_to_depth2", x.ptr, x.plane, x.n, x.h, x.w, x.c, xs.ptr,
                   meta=dict(kind="tcv_space_to_depth2", bytes=8 * x.n * x.img_elems))
        taps = [(ty, tx) for ty in range(-2, 2) for tx in range(-2, 2)]
        y = self._act(x.n, oh, ow, cout)
        d = self._desc(xs, ent["w"].data_ptr(), taps, 1, PAD_ZERO, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, key, None,
                       False, ACT_NONE, None, 0, None, None, 0)
        meta = self._conv_meta(d, key, xs, 4, 1)
        meta.update(flops=2 * x.n * oh * ow * 49 * self.w[wkey]["cin_real"] * self.w[wkey]["cout_real"], layer=wkey)
        self._call("tcv_conv2d", C.byref(d), meta=meta)
        return y

    def conv7x7s2(self, x: Act, wkey: str) -> Act:
        """The same stem as four partial CUDA-core convolutions of <= 14 taps each (tcv_conv2d takes at most 16 taps),
        chained through the residual input of the epilogue.  Kept as the cross-check of stem_s2d (TCV_FBA_STEM=direct)."""
        ent = self.w[wkey]
        cout = ent["cout"]
        assert ent["k"] == 7 and ent["cin"] == x.c
        oh, ow = (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1
        prev: Optional[Act] = None
        for r0 in range(0, 7, 2):
            rows = range(r0, min(r0 + 2, 7))
            taps = [(ky - 3, kx - 3) for ky in rows for kx in range(7)]
            wtap = [ky * 7 + kx for ky in rows for kx in range(7)]
            y = self._act(x.n, oh, ow, cout)
            d = self._desc(x, ent["w"].data_ptr(), taps, 2, PAD_ZERO, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey, None,
                           False, ACT_NONE, prev, 0, None, None, 0, wtap=wtap)
            self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey, x, 7, 2))
            prev = y
        return prev

    def gn(self, z: Act, p: str, act, res: Optional[Act] = None, out: Optional[Act] = None, out_off=0) -> Act:
        """GroupNorm(32) + optional residual + activation; `out`/`out_off`: write into a channel slice of a wider
        (concatenation) tensor instead of a fresh one."""
        gamma, beta = self.gn_params[p]
        n, pixels, c = z.n, z.h * z.w, z.c
        assert gamma.numel() == c, (p, gamma.numel(), c)
        scale = self._empty((n, c))
        shift = self._empty((n, c))
        nbytes = 4 * n * pixels * c
        sums = self._empty((n, c, 2), torch.float64)
        self._call("tcv_gn_stats", z.ptr, z.plane, n, pixels, c, sums.data_ptr(),
                   meta=dict(kind="tcv_gn_stats", bytes=nbytes, layer=p))
        self._call("tcv_gn_finalize", sums.data_ptr(), n, pixels, c, GN_GROUPS, gamma.data_ptr(), beta.data_ptr(),
                   GN_EPS, scale.data_ptr(), shift.data_ptr())
        y = out if out is not None else self._act(n, z.h, z.w, c)
        assert (y.n, y.h, y.w) == (n, z.h, z.w) and out_off + c <= y.c
        if res is not None:
            assert (res.n, res.h, res.w, res.c) == (n, z.h, z.w, c), p
        self._call("tcv_gn_apply", z.ptr, z.plane, n, pixels, c, scale.data_ptr(), shift.data_ptr(),
                   res.ptr if res is not None else None, res.plane if res is not None else 0, act, y.ptr, y.plane,
                   y.c, out_off, meta=dict(kind="tcv_gn_apply", bytes=nbytes * (3 if res is not None else 2), layer=p))
        return y

    def maxpool(self, x: Act) -> Act:
        assert x.plane == x.n * x.img_elems
        y = self._act(x.n, (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1, x.c)
        self._call("tcv_maxpool3s2", x.ptr, x.n, x.h, x.w, x.c, y.ptr,
                   meta=dict(kind="tcv_maxpool3s2", bytes=4 * x.n * (x.img_elems + y.img_elems)))
        return y

    def bilinear(self, x: Act, oh: int, ow: int, out: Act, out_off: int) -> None:
        assert x.plane == x.n * x.img_elems and (out.n, out.h, out.w) == (x.n, oh, ow)
        self._call("tcv_bilinear", x.ptr, x.n, x.h, x.w, x.c, out.ptr, out.plane, oh, ow, out.c, out_off,
                   meta=dict(kind="tcv_bilinear", bytes=4 * x.n * (x.img_elems + oh * ow * x.c)))

    def copy_channels(self, x: Act, x_off: int, c: int, out: Act, out_off: int) -> None:
        assert (out.n, out.h, out.w) == (x.n, x.h, x.w)
        self._call("tcv_copy_channels", x.ptr, x.plane, x.c, x_off, out.ptr, out.plane, out.c, out_off, c,
                   x.n * x.h * x.w, meta=dict(kind="tcv_copy_channels", bytes=8 * x.n * x.h * x.w * c))

    # ------------------------------------------------------------------ network program
    def _bottleneck(self, x: Act, p: str, stride: int, dil: int, has_down: bool, dstride: int,
                    out: Optional[Act] = None) -> Act:
        """Bottleneck.forward (resnet_GN_WS.py:69-91)."""
        o = self.gn(self.convf(x, p + ".conv1"), p + ".bn1", ACT_RELU)
        o = self.gn(self.convf(o, p + ".conv2", stride=stride, dilation=dil), p + ".bn2", ACT_RELU)
        z = self.convf(o, p + ".conv3")
        idt = x
        if has_down:
            idt = self.gn(self.convf(x, p + ".downsample.0", stride=dstride), p + ".downsample.1", ACT_NONE)
        return self.gn(z, p + ".bn3", ACT_RELU, res=idt, out=out)

    def per_frame(self, x16: Act) -> dict:
        """encoder + pyramid-pooling head for all frames at once (VMN_model.py:93-98)."""
        e, d = "encoder", "decoder"
        stem = self.conv7x7s2 if os.environ.get("TCV_FBA_STEM", "s2d") == "direct" else self.stem_s2d
        c1 = self.gn(stem(x16, e + ".conv1"), e + ".bn1", ACT_RELU)                      # conv_out[1], OS2, 64 ch
        x = self.maxpool(c1)
        l1 = None
        cat = None
        for name, planes, blocks, stride, dilate in RES_LAYERS:
            for i in range(blocks):
                s, dl, ds = block_config(i, stride, dilate)
                out = None
                if name == "layer4" and i == blocks - 1:
                    # conv5 goes straight into the pyramid-pooling concatenation (VMN_FBA.py:24-31)
                    cat = out = self._act(x.n, x.h, x.w, 2048 + 256 * len(PPM_SCALES))
                x = self._bottleneck(x, f"{e}.{name}.{i}", s, dl, i == 0, ds, out=out)
            if name == "layer1":
                l1 = x                                                                    # conv_out[-4], OS4, 256 ch
        h8, w8 = cat.h, cat.w
        for i, s in enumerate(PPM_SCALES):
            pooled = self._act(cat.n, s, s, 2048)
            self._call("tcv_adaptive_avgpool", cat.ptr, cat.plane, cat.n, h8, w8, 2048, cat.c, 0, s, pooled.ptr,
                       meta=dict(kind="tcv_adaptive_avgpool", bytes=4 * cat.n * h8 * w8 * 2048))
            t = self.gn(self.convf(pooled, f"{d}.ppm.{i}.1", bias=True), f"{d}.ppm.{i}.2", ACT_LEAKY001)
            self.bilinear(t, h8, w8, cat, 2048 + 256 * i)
        split = os.environ.get("TCV_FBA_SPLIT_K", "1") == "1"
        x = self.gn(self.convf(cat, d + ".conv_up1.0", bias=True, split_rows=split), d + ".conv_up1.1", ACT_LEAKY001)
        feat = self.gn(self.convf(x, d + ".conv_up1.3", bias=True), d + ".conv_up1.4", ACT_LEAKY001)
        return dict(feat=feat, l1=l1, c1=c1, x16=x16)

    def tail(self, pf: dict, n0: int, ncen: int, mask_ptr: int, mask_stride: int, H: int, W: int, pred_ptr: int,
             attb_ptr: int, attf_ptr: int, sm_ptr: int) -> None:
        """decoder tail for `ncen` consecutive centre frames starting at image n0+1 (VMN_FBA.py:34-59)."""
        d = "decoder"
        feat: Act = pf["feat"]
        sl = lambda a: a.slice(n0 + 1, n0 + 1 + ncen)
        t = self.tam(d + ".fam", sl(feat), feat.slice(n0, n0 + ncen), feat.slice(n0 + 2, n0 + 2 + ncen), mask_ptr,
                     mask_stride, H, W, attb_ptr, attf_ptr, sm_ptr)
        l1, c1, x16 = sl(pf["l1"]), sl(pf["c1"]), sl(pf["x16"])
        cat2 = self._act(ncen, l1.h, l1.w, 512)
        self.bilinear(t, l1.h, l1.w, cat2, 0)
        self.copy_channels(l1, 0, 256, cat2, 256)
        x = self.gn(self.convf(cat2, d + ".conv_up2.0", bias=True), d + ".conv_up2.1", ACT_LEAKY001)
        cat3 = self._act(ncen, c1.h, c1.w, 320)
        self.bilinear(x, c1.h, c1.w, cat3, 0)
        self.copy_channels(c1, 0, 64, cat3, 256)
        x = self.gn(self.convf(cat3, d + ".conv_up3.0", bias=True), d + ".conv_up3.1", ACT_LEAKY001)
        cat4 = self._act(ncen, H, W, 96)                       # 64 + 3 + 3 + 2 = 72 channels, zero-padded to 96
        self.bilinear(x, H, W, cat4, 0)
        self._call("tcv_fba_cat_inputs", x16.ptr, x16.plane, ncen * H * W, cat4.ptr, cat4.plane, cat4.c, 64,
                   meta=dict(kind="tcv_fba_cat_inputs", bytes=4 * ncen * H * W * 40))
        x = self.convf(cat4, d + ".conv_up4.0", bias=True, act=ACT_LEAKY001)
        x = self.convf(x, d + ".conv_up4.2", bias=True, act=ACT_LEAKY001)
        o8 = self.convf(x, d + ".conv_up4.4", bias=True)
        self._call("tcv_fba_fusion", o8.ptr, x16.ptr, x16.plane, x16.img_elems, ncen, H, W, pred_ptr,
                   meta=dict(kind="tcv_fba_fusion", bytes=ncen * H * W * (32 + 12 + 28)))

    def window_program(self, x16: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        """Runs (and records) the whole VMN forward on the encoded input.  trimask fp32 [B*S,H,W]."""
        ncen = S - 2
        N8 = (H // 8) * (W // 8)
        w2 = self.window * self.window
        pred = self._empty((B, ncen, 7, H, W))
        attb = self._empty((B, ncen, w2, N8))
        attf = self._empty((B, ncen, w2, N8))
        sm = self._empty((B, ncen, 1, H // 8, W // 8), torch.uint8)
        pf = self.per_frame(x16)
        for b in range(B):
            n0 = b * S
            self.tail(pf, n0, ncen, trimask.data_ptr() + 4 * (n0 + 1) * H * W, H * W, H, W,
                      pred[b].data_ptr(), attb[b].data_ptr(), attf[b].data_ptr(), sm[b].data_ptr())
        return dict(pred=pred, attb=attb, attf=attf, small_mask=sm, feat=pf["feat"], pf=pf)

    def encode_inputs(self, imgs: torch.Tensor, tris: torch.Tensor, frames: int, H: int, W: int, x16: Act) -> None:
        """EvalModel.preprocess for 'fba' (models/model.py:366-386): colour channels, two-channel trimap and the six
        distance-transform channels of every frame."""
        u8 = 1 if imgs.dtype == torch.uint8 else 0
        g = self._empty((frames, 2, H, W), torch.int32)
        self._call("tcv_fba_encode_inputs", imgs.data_ptr(), tris.data_ptr(), u8, frames, H, W, x16.ptr,
                   meta=dict(kind="tcv_fba_encode_inputs", bytes=frames * H * W * (4 * (1 if u8 else 4) + 40)))
        self._call("tcv_fba_edt_cols", x16.ptr, frames, H, W, g.data_ptr())
        self._call("tcv_fba_edt_rows", g.data_ptr(), frames, H, W, x16.ptr,
                   meta=dict(kind="tcv_fba_edt_rows", bytes=frames * H * W * 2 * (4 + 12)))

    def eval_program(self, B: int, S: int, H: int, W: int, dilate: int, u8: bool) -> dict:
        """EvalModel.forward for method 'fba' (models/model.py:389-446) on static buffers: trimask (+ optional
        dilation, the kernel shared with vmn_gca), input encoding, the VMN program, the where()-tail.  Returns the
        io tensors (inputs ``imgs`` / ``tris`` to be filled before a replay; outputs ``alphas`` / ``Fs`` / ``Bs``)."""
        in_dt = torch.uint8 if u8 else torch.float32
        sfx = "_u8" if u8 else ""
        imgs = self._empty((B, S, 3, H, W), in_dt)
        tris = self._empty((B, S, 1, H, W), in_dt)
        x8 = self._act(B * S, H, W, 8)                          # by-product of the shared trimask kernel (unused)
        x16 = self._act(B * S, H, W, 16)
        trimask = self._empty((B * S, H, W))
        tmp = self._empty((2 * B * S * H * W,), torch.uint8)
        alphas = self._empty((B, S, 1, H, W))
        Fs = self._empty((B, S, 3, H, W))
        Bs = self._empty((B, S, 3, H, W))
        imgs.zero_(); tris.zero_()                              # recording runs the kernels once: valid inputs
        self._call("tcv_preprocess_eval" + sfx, imgs.data_ptr(), tris.data_ptr(), B * S, H, W, dilate, x8.ptr,
                   trimask.data_ptr(), tmp.data_ptr())
        self.encode_inputs(imgs, tris, B * S, H, W, x16)
        out = self.window_program(x16, trimask, B, S, H, W)
        self._call("tcv_postprocess_eval_fba", out["pred"].data_ptr(), imgs.data_ptr(), tris.data_ptr(),
                   1 if u8 else 0, trimask.data_ptr(), B, S, H, W, alphas.data_ptr(), Fs.data_ptr(), Bs.data_ptr())
        io = dict(imgs=imgs, tris=tris, alphas=alphas, Fs=Fs, Bs=Bs, trimask=trimask, x16=x16.buf,
                  feat=out["feat"].buf, **{k: out[k] for k in ("pred", "attb", "attf", "small_mask")})
        return io
