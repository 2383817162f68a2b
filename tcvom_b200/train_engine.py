"""Host-side engine for one ``vmn_gca`` TRAINING step (train_ddp.py:52-65): train-mode forward
(SpectralNorm power iteration per call, batch-statistics BatchNorm per frame, optional SyncBatchNorm),
the losses, and the full backward, all on the sm_100a kernels behind the C ABI.

What the reference's autograd graph does is replayed by an explicit tape: every forward operator appends a
closure that computes its input / parameter gradients from the gradient of its output.  torch is used for
device memory, the NCCL all-reduce of the BatchNorm statistics and as the autograd *boundary*
(``tcvom_b200.model._TrainStepFn``), never for arithmetic on activations.

Program restated from (reference checkout):
  VMN.forward                    models/VMN/VMN_model.py:83-113  (per-frame loop, per-centre-frame loop)
  ResGuidedCxtAtten.forward      models/GCA/encoders/res_gca_enc.py:57-90
  ResGuidedCxtAtten_FAM_Dec      models/VMN/VMN_GCA.py:26-49
  SpectralNorm._update_u_v       models/GCA/ops.py:25-36
  FullModel_VMD.forward (losses) models/model.py:94-127, 285-345
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional

import torch

from . import _cabi
from ._cabi import (ACT_LEAKY02, ACT_NONE, ACT_RELU, ACT_TANH01, PAD_REFLECT, PAD_ZERO, BnDesc, ConvDesc, SnDesc)
from .engine import BN_EPS, Act, GcaVmnEngine, _k, named_tensors
from .modules import DEC_LAYERS, ENC_LAYERS

BN_MOMENTUM = 0.1


class TAct:
    """Activation on the tape: value ``a`` (split-bf16 NHWC) + gradient slot ``g``.  ``g_owned`` tells whether
    the gradient buffer may be accumulated into in place (a buffer aliased from another tensor's gradient is
    copied before the first accumulation)."""
    __slots__ = ("a", "g", "g_owned", "needs_grad", "groups")

    def __init__(self, a: Act, groups: int, needs_grad=True):
        self.a, self.g, self.g_owned, self.needs_grad, self.groups = a, None, False, needs_grad, groups


# the part of vmn_gca that ``freeze_backbone=True`` keeps in eval mode under no_grad: the whole encoder (VMN_model.py:77-81,
# 99-103) and the decoder's feature-extraction half (VMN_GCA.py:18-24,26-34)
FROZEN_PREFIXES = ("encoder.", "decoder.layer1.", "decoder.layer2.", "decoder.gca.")


class FrozenBackboneEngine(GcaVmnEngine):
    """Inference engine over the frozen part only: its folded / packed weights are re-derived when a BACKBONE tensor changes,
    not on every optimizer step of the trainable tail."""

    def _named(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in super()._named().items() if k.startswith(FROZEN_PREFIXES)}


class TrainEngine(GcaVmnEngine):
    """Owns the derived device state of one training step for one ``VMN`` module on one device."""

    def __init__(self, window: int):
        super().__init__(window)
        self.tape: List[Callable[[], None]] = []
        self.sn: Dict[str, dict] = {}
        self.dw: Dict[str, torch.Tensor] = {}
        self.dbias: Dict[str, torch.Tensor] = {}
        self.dbn: Dict[str, tuple] = {}
        self.sync_bn = False
        # TAM pre-training (get_VMN_models(freeze_backbone=True)): per-frame features from `backbone` (eval-mode program,
        # running-statistics BatchNorm, stored spectral-norm u / v), tape and gradients for the decoder tail only
        self.freeze_backbone = False
        self.backbone: Optional[FrozenBackboneEngine] = None
        self._peer = None
        self._peer_tried = False
        self.process_group = None
        self.world = 1
        # weight gradients of the stride-1 convs / deconv phases on the tensor cores (split-K tcgen05 GEMM over
        # channel-major copies); TCV_TC_WGRAD=0 keeps the CUDA-core fp32 kernel everywhere (exact cross-check)
        import os
        # small accumulators (weight / bias / BatchNorm-parameter gradients, sigma dot products) are carved out of ONE
        # zero-filled buffer per step: they were ~700 separate fill kernels (2.6 ms) per step
        self._arena = None
        self._arena_off = 0
        self._arena_need = 0
        self._nbt: List = []
        self.use_head32 = os.environ.get("TCV_HEAD32", "1") == "1"
        # training GCA in the shift-sum form (value GEMM and its two backward GEMMs 3.8x smaller); "0": round-1 fold form
        self.gca_shift_sum_train = os.environ.get("TCV_GCA_SHIFT_SUM_TRAIN", "1") != "0"
        # attention-backward GEMMs with MN-major operands (tcv_gemm_tc_ex) instead of transposed copies; "0": transposes
        self.gca_mn_gemm = os.environ.get("TCV_GCA_MN_GEMM", "1") != "0"
        self.use_tc_wgrad = os.environ.get("TCV_TC_WGRAD", "1") != "0"
        # "2": the NHWC-direct kernel (MN-major operands, no channel-major copies); "1": the split-K GEMM over copies
        self.wgrad_nhwc = os.environ.get("TCV_TC_WGRAD", "2") == "2"

    # ------------------------------------------------------------------ per-step weight state
    def refresh_weights(self, net: torch.nn.Module, force=False) -> None:
        """Packs the RAW weights (W_bar, not W_bar/sigma: the per-call 1/sigma is applied by the BatchNorm kernels
        that follow every spectral-norm conv) in the forward and in the data-gradient (transposed) layout."""
        self.net = net
        named = self._named()
        dev = next(iter(named.values())).device
        if dev.type != "cuda":
            raise RuntimeError("tcvom_b200: the module must live on a CUDA device (no CPU fallback)")
        self.device = dev
        self.named = named
        L = _cabi.lib()
        st = torch.cuda.current_stream(dev).cuda_stream
        for name, t in named.items():
            if t.dtype.is_floating_point and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"tcvom_b200: parameter {name} must be contiguous fp32")
        for name, t in named.items():
            if name.endswith(".module.weight_bar"):
                p = name[: -len(".module.weight_bar")]
                self._pack(L, st, p, t, None, None, None, transposed=(t.shape[2] == 4))
                self.w[p]["param"] = name
                self.w[p]["sn"] = True
            elif name.endswith(".weight") and t.dim() == 4:
                p = name[: -len(".weight")]
                self._pack(L, st, p, t, None, None, None, transposed=False)
                self.w[p]["param"] = name
                self.w[p]["sn"] = False
                b = named.get(p + ".bias")
                if b is not None:
                    self.bias[p] = b
                if self.use_head32 and t.shape[0] == 1 and t.shape[1] % 32 == 0 and t.shape[2] == 3:
                    # alpha head as a zero-padded 32-channel conv: its forward and both gradients run on tcgen05
                    self._pack_head32(L, st, p, t, b)
                    hk = p + self.HEAD32
                    self.w[hk].update(param=name, sn=False, bias_param=p + ".bias")
                    self.w[p].pop("param")
        for p, ent in list(self.w.items()):
            if p.endswith("#T"):
                continue
            self._pack_transposed(L, st, p, ent)

    def _pack_transposed(self, L, st, p, ent):
        """wT[tap][co_pad][ci]: the weight of the data-gradient conv (its input channels are the forward's output
        channels, padded to 8 for the 1-channel alpha head)."""
        taps = ent["k"] * ent["k"]
        cin_f, cout_f = ent["cin"], ent["cout"]
        cout_pad = (cout_f + 7) // 8 * 8
        tk = p + "#T"
        te = self.w.get(tk)
        if te is None:
            te = self.w[tk] = dict(w=torch.empty((taps, cout_pad, cin_f), dtype=torch.float32, device=self.device),
                                   cout=cin_f, cin=cout_pad, cin_real=cout_f, k=ent["k"], transposed=False)
        _cabi.check(L.tcv_transpose_packed(ent["w"].data_ptr(), taps, cin_f, cout_f, cout_pad, te["w"].data_ptr(), st),
                    "transpose_packed")
        if cout_pad % 32 == 0 and cin_f % 32 == 0:
            if "w_tc" not in te:
                te["w_tc"] = torch.empty((2, taps, cin_f, cout_pad), dtype=torch.bfloat16, device=self.device)
            _cabi.check(L.tcv_pack_weight_tc(te["w"].data_ptr(), taps, cout_pad, cin_f, te["w_tc"].data_ptr(), st),
                        "pack_weight_tc")

    def _sn_calls(self, p: str, S: int, ncen: int) -> int:
        tail = p.startswith(("decoder.layer3", "decoder.layer4", "decoder.conv1"))
        return ncen if tail else S

    def spectral_norm_step(self, S: int, ncen: int) -> None:
        """One power iteration per forward call of every spectral-norm layer (ops.py:25-36,74-80), all layers in
        one launch; u, v of the module are updated in place like the reference's ``u.data = ...``."""
        L = _cabi.lib()
        keys = [p for p, e in self.w.items() if e.get("sn")]
        if self.freeze_backbone:          # eval-mode layers keep u / v (ops.py:38-45,76-80)
            keys = [p for p in keys if not (p + ".").startswith(FROZEN_PREFIXES)]
        descs = (SnDesc * len(keys))()
        for i, p in enumerate(keys):
            wbar = self.named[p + ".module.weight_bar"]
            u, v = self.named[p + ".module.weight_u"], self.named[p + ".module.weight_v"]
            rows = wbar.shape[0]
            cols = wbar.numel() // rows
            calls = self._sn_calls(p, S, ncen)
            s = self.sn.get(p)
            if s is None or s["calls"] != calls:
                s = self.sn[p] = dict(calls=calls, rows=rows, cols=cols,
                                      u_hist=torch.empty((calls, rows), dtype=torch.float32, device=self.device),
                                      v_hist=torch.empty((calls, cols), dtype=torch.float32, device=self.device),
                                      sigma=torch.empty((calls,), dtype=torch.float32, device=self.device),
                                      isig=torch.empty((calls,), dtype=torch.float32, device=self.device),
                                      zdot=None)
            assert rows <= 512 and cols <= 8192, (p, rows, cols)      # cluster kernel's shared-memory slices
            s["zdot"] = self._zeros((calls,), torch.float64)
            d = descs[i]
            d.w_bar, d.rows, d.cols, d.u, d.v, d.calls = wbar.data_ptr(), rows, cols, u.data_ptr(), v.data_ptr(), calls
            d.u_hist, d.v_hist = s["u_hist"].data_ptr(), s["v_hist"].data_ptr()
            d.sigma, d.inv_sigma = s["sigma"].data_ptr(), s["isig"].data_ptr()
        raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(self.device)
        self._sn_descs = raw
        st = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(L.tcv_sn_power_iter(raw.data_ptr(), len(keys), st), "sn_power_iter")
        for p in keys:                       # u / v were updated in place by the kernel: tell torch (version counters)
            for sfx in (".module.weight_u", ".module.weight_v"):
                torch.autograd.graph.increment_version(self.named[p + sfx])

    # ------------------------------------------------------------------ zero arena
    def _arena_begin(self) -> None:
        need = self._arena_need
        self._arena = torch.zeros((need,), dtype=torch.uint8, device=self.device) if need else None
        self._arena_off, self._arena_need = 0, 0

    def _zeros(self, shape, dtype=torch.float32) -> torch.Tensor:
        n = 1
        for v in shape:
            n *= v
        nbytes = (n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256
        self._arena_need += nbytes
        if self._arena is not None and self._arena_off + nbytes <= self._arena.numel():
            t = self._arena[self._arena_off:self._arena_off + nbytes].view(dtype)[:n].view(shape)
            self._arena_off += nbytes
            return t
        return torch.zeros(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ tape helpers
    def _acc(self, t: Optional[TAct], g: Act, owned: bool) -> None:
        if t is None or not t.needs_grad:
            return
        if t.g is None:
            t.g, t.g_owned = g, owned
            return
        if not t.g_owned:
            buf = t.g.buf.clone()
            t.g = Act(buf, t.g.n, t.g.h, t.g.w, t.g.c, buf.data_ptr(), t.g.plane)
            t.g_owned = True
        assert g.plane == t.g.plane and g.n * g.img_elems == t.g.n * t.g.img_elems
        self._call("tcv_add_split", g.ptr, g.plane, t.g.ptr, t.g.plane, g.n * g.img_elems)

    def _zeros_act(self, n, h, w, c) -> Act:
        buf = torch.zeros((2, n, h, w, c), dtype=torch.bfloat16, device=self.device)
        return Act(buf, n, h, w, c, buf.data_ptr(), n * h * w * c)

    def _dw(self, wkey: str, cout_pad: Optional[int] = None) -> torch.Tensor:
        t = self.dw.get(wkey)
        if t is None:
            ent = self.w[wkey]
            t = self.dw[wkey] = self._zeros((ent["k"] * ent["k"], ent["cin"], cout_pad or ent["cout"]))
        return t

    # ------------------------------------------------------------------ weight gradient
    def _transpose_pad(self, a: Act, mul: int, oy: int, ox: int, row: int, shift: int, ktot: int) -> torch.Tensor:
        t = torch.empty((2, a.c, ktot), dtype=torch.bfloat16, device=self.device)
        self._call("tcv_transpose_pad", a.ptr, a.plane, a.n, a.h, a.w, a.c, mul, oy, ox, row, shift, t.data_ptr(),
                   a.c * ktot, ktot, meta=dict(tag=f"c{a.c} K{ktot}"))
        return t

    def _wgrad(self, d: ConvDesc, xa: Act, dz: Act, dz_c: int, dw: torch.Tensor, mul=1, oy=0, ox=0, xt=None):
        """dw += weight gradient of the conv described by the forward descriptor ``d``.  Stride-1 zero-padded
        taps within one pixel run on the tensor cores; everything else (stride-2, reflect) on CUDA cores.
        Returns the channel-major copy of x (reusable by the other phases of a deconv)."""
        if self.use_tc_wgrad and self.wgrad_nhwc and d.pad_mode == PAD_ZERO and dz_c % 8 == 0:
            self._call("tcv_conv2d_wgrad_nhwc_tc", C.byref(d), dz.ptr, dz.plane, dz_c, dw.data_ptr(),
                       meta=dict(tag=f"cin{xa.c} cout{dz_c} px{xa.n * d.gh * d.gw} taps{d.ntaps} s{d.stride}"))
            return None
        taps_ok = all(abs(d.dy[t]) <= 1 and abs(d.dx[t]) <= 1 for t in range(d.ntaps))
        if not (self.use_tc_wgrad and d.stride == 1 and d.pad_mode == PAD_ZERO and taps_ok and dz_c % 8 == 0):
            self._call("tcv_conv2d_wgrad", C.byref(d), dz.ptr, dz.plane, dz_c, dw.data_ptr())
            return None
        gh, gw = d.gh, d.gw
        assert (gh, gw) == (xa.h, xa.w) and (dz.h, dz.w) == (gh * mul, gw * mul) and dz.c == dz_c
        row = (gw + 2 + 7) // 8 * 8                  # 16-byte aligned rows: vertical tap shifts stay TMA-legal
        ktot = xa.n * (gh + 2) * row
        if xt is None:
            xt = self._transpose_pad(xa, 1, 0, 0, row, 0, ktot)
        cin = xa.c
        tiles = ((cin + 127) // 128) * ((dz_c + 127) // 128)
        nsplit = max(1, min(ktot // 2048, (2 * 148 + tiles - 1) // tiles))
        partial = torch.empty((nsplit, cin, 3 * dz_c), dtype=torch.float32, device=self.device)
        for sx in sorted({d.dx[t] for t in range(d.ntaps)}):
            # horizontal offset baked into a shifted copy of dz: sum_p x[p + dy*row + sx] z[p] = sum_q x[q + dy*row] z[q - sx]
            zt = self._transpose_pad(dz, mul, oy, ox, row, sx, ktot)
            ts = [t for t in range(d.ntaps) if d.dx[t] == sx]
            IntArr = C.c_int * len(ts)
            dy, dx = IntArr(*[d.dy[t] for t in ts]), IntArr(*([0] * len(ts)))
            wt = IntArr(*[d.wtap[t] for t in ts])
            self._call("tcv_wgrad_tc", xt.data_ptr(), cin * ktot, zt.data_ptr(), dz_c * ktot, cin, dz_c, ktot, row, len(ts),
                       dy, dx, wt, partial.data_ptr(), nsplit, dw.data_ptr(), dz_c,
                       meta=dict(tag=f"cin{cin} cout{dz_c} K{ktot} taps{len(ts)} split{nsplit}"))
        return xt

    def _ctag(self, d: ConvDesc, what: str) -> dict:
        if getattr(self, "_prof", None) is None:
            return {}
        path = {0: "direct", 1: "tc", 2: "tc2", 3: "tc3", 4: "tc2p"}[_cabi.lib().tcv_conv2d_path(C.byref(d))]
        return dict(tag=f"{what} {path} {d.cin}->{d.cout} t{d.ntaps} s{d.stride} px{d.n * d.gh * d.gw}")

    # ------------------------------------------------------------------ convolution (raw, no BatchNorm)
    def _fwd_geometry(self, x: Act, k: int, stride: int, prepadded: bool):
        if k == 3 and prepadded:
            taps = [(ky, kx) for ky in range(3) for kx in range(3)]
            oh, ow = (x.h - 3) // stride + 1, (x.w - 3) // stride + 1
        elif k == 3:
            taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
            oh, ow = (x.h + 2 - 3) // stride + 1, (x.w + 2 - 3) // stride + 1
        else:
            taps = [(0, 0)]
            oh, ow = (x.h - 1) // stride + 1, (x.w - 1) // stride + 1
        return taps, oh, ow

    def conv_op(self, x: TAct, wkey: str, *, stride=1, pad=PAD_ZERO, prepadded=False, bias=False, act=ACT_NONE,
                f32_out: Optional[torch.Tensor] = None) -> TAct:
        """z = conv(x, W_raw) (+ bias) (+ tanh01 for the alpha head).  Backward: weight gradient, bias gradient,
        data gradient (tcv_conv2d on the transposed weights)."""
        ent = self.w[wkey]
        k, cout = ent["k"], ent["cout"]
        xa = x.a
        taps, oh, ow = self._fwd_geometry(xa, k, stride, prepadded)
        head = f32_out is not None                     # 1-channel alpha head: fp32 output only
        y = None if head else self._act(xa.n, oh, ow, cout)
        d = self._desc(xa, ent["w"].data_ptr(), taps, stride, pad, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey, None, bias,
                       act, None, 0, None, None, f32_out.data_ptr() if head else 0)
        self._call("tcv_conv2d", C.byref(d), meta=self._ctag(d, "fwd"))
        z = TAct(y, x.groups)
        cpad = (cout + 7) // 8 * 8

        def backward():
            dz = z.g
            if dz is None:
                return
            L = _cabi.lib()
            self._wgrad(d, xa, dz, cpad, self._dw(wkey, cpad))
            if bias:
                db = self.dbias.get(wkey)
                if db is None:
                    db = self.dbias[wkey] = self._zeros((cpad,))
                self._call("tcv_channel_sum", dz.ptr, dz.plane, dz.n * dz.h * dz.w, cpad, db.data_ptr())
            if x.needs_grad:
                self._acc(x, self._dgrad(dz, wkey, xa, k, stride, prepadded), True)
            z.g = None
        self.tape.append(backward)
        return z

    def _dgrad(self, dz: Act, wkey: str, xa: Act, k: int, stride: int, prepadded: bool) -> Act:
        """dL/dx of z = conv(x): a gather-form conv of dz with the transposed weights."""
        tk = wkey + "#T"
        ent = self.w[tk]
        cin_f = ent["cout"]                      # channels of x
        assert ent["cin"] == dz.c, (wkey, ent["cin"], dz.c)
        if stride == 1:
            dx = self._act(xa.n, xa.h, xa.w, cin_f)
            if k == 3:
                assert not prepadded
                taps = [(1 - ky, 1 - kx) for ky in range(3) for kx in range(3)]
            else:
                taps = [(0, 0)]
            d = self._desc(dz, ent["w"].data_ptr(), taps, 1, PAD_ZERO, dx, xa.h, xa.w, cin_f, xa.h, xa.w, 1, 0, 1, 0, tk,
                           None, False, ACT_NONE, None, 0, None, None, 0)
            self._call("tcv_conv2d", C.byref(d), meta=self._ctag(d, "dgrad"))
            return dx
        assert stride == 2 and xa.h % 2 == 0 and xa.w % 2 == 0
        if k == 1:
            dx = self._zeros_act(xa.n, xa.h, xa.w, cin_f)
            d = self._desc(dz, ent["w"].data_ptr(), [(0, 0)], 1, PAD_ZERO, dx, xa.h, xa.w, cin_f, dz.h, dz.w, 2, 0, 2, 0,
                           tk, None, False, ACT_NONE, None, 0, None, None, 0)
            self._call("tcv_conv2d", C.byref(d))
            return dx
        dx = self._act(xa.n, xa.h, xa.w, cin_f)
        for py in range(2):
            for px in range(2):
                if prepadded:       # x_pad[r] is read by output oy with 2*oy + ky = r
                    rows = [(0, 0), (2, -1)] if py == 0 else [(1, 0)]
                    cols = [(0, 0), (2, -1)] if px == 0 else [(1, 0)]
                else:               # 2*oy + ky - 1 = r
                    rows = [(1, 0)] if py == 0 else [(0, 1), (2, 0)]
                    cols = [(1, 0)] if px == 0 else [(0, 1), (2, 0)]
                taps = [(oy, ox) for ky, oy in rows for kx, ox in cols]
                wtap = [ky * 3 + kx for ky, oy in rows for kx, ox in cols]
                d = self._desc(dz, ent["w"].data_ptr(), taps, 1, PAD_ZERO, dx, xa.h, xa.w, cin_f, xa.h // 2, xa.w // 2,
                               2, py, 2, px, tk, None, False, ACT_NONE, None, 0, None, None, 0, wtap=wtap)
                self._call("tcv_conv2d", C.byref(d), meta=self._ctag(d, "dgrad-s2phase"))
        return dx

    def deconv_op(self, x: TAct, wkey: str) -> TAct:
        """ConvTranspose2d(k=4, s=2, p=1) as 4 sub-pixel phases (raw); backward: 4 phase weight gradients and a
        16-tap stride-2 data-gradient conv."""
        ent = self.w[wkey]
        assert ent["transposed"] and ent["cin"] == x.a.c
        cout = ent["cout"]
        xa = x.a
        oh, ow = 2 * xa.h, 2 * xa.w
        y = self._act(xa.n, oh, ow, cout)
        descs = []
        for py in range(2):
            for px in range(2):
                kys = [(1, 0), (3, -1)] if py == 0 else [(0, 1), (2, 0)]
                kxs = [(1, 0), (3, -1)] if px == 0 else [(0, 1), (2, 0)]
                taps = [(dy, dx) for ky, dy in kys for kx, dx in kxs]
                wtap = [ky * 4 + kx for ky, dy in kys for kx, dx in kxs]
                d = self._desc(xa, ent["w"].data_ptr(), taps, 1, PAD_ZERO, y, oh, ow, cout, xa.h, xa.w, 2, py, 2, px,
                               wkey, None, False, ACT_NONE, None, 0, None, None, 0, wtap=wtap)
                self._call("tcv_conv2d", C.byref(d), meta=self._ctag(d, "fwd-deconv"))
                descs.append(d)
        z = TAct(y, x.groups)

        def backward():
            dz = z.g
            if dz is None:
                return
            xt = None
            for i, d in enumerate(descs):
                xt = self._wgrad(d, xa, dz, cout, self._dw(wkey), mul=2, oy=i // 2, ox=i % 2, xt=xt)
            if x.needs_grad:
                te = self.w[wkey + "#T"]
                dx = self._act(xa.n, xa.h, xa.w, xa.c)
                taps = [(ky - 1, kx - 1) for ky in range(4) for kx in range(4)]
                dd = self._desc(dz, te["w"].data_ptr(), taps, 2, PAD_ZERO, dx, xa.h, xa.w, xa.c, xa.h, xa.w, 1, 0, 1, 0,
                                wkey + "#T", None, False, ACT_NONE, None, 0, None, None, 0)
                self._call("tcv_conv2d", C.byref(dd), meta=self._ctag(dd, "dgrad-deconv"))
                self._acc(x, dx, True)
            z.g = None
        self.tape.append(backward)
        return z

    def _allreduce_sums(self, t: torch.Tensor) -> None:
        """Cross-rank sum of one BatchNorm statistic block (nn.SyncBatchNorm, train_ddp.py:273): the peer-memory kernel on
        the compute stream when the node's GPUs can map each other (tcvom_b200/peer.py), else NCCL."""
        if self._peer is None and not self._peer_tried:
            self._peer_tried = True
            from .peer import make_peer_reducer
            self._peer = make_peer_reducer(self.process_group, self.device)
        if self._peer is not None:
            self._peer.allreduce_(t, self._stream_ptr())
        else:
            torch.distributed.all_reduce(t, group=self.process_group)

    # ------------------------------------------------------------------ train-mode BatchNorm (+ 1/sigma, act, residuals)
    def bn_op(self, z: TAct, bnkey: str, *, mode=1, act=ACT_NONE, snkey: Optional[str] = None,
              res1: Optional[TAct] = None, res1_shift=0, res2: Optional[TAct] = None, unbias_mul=1) -> TAct:
        za = z.a
        groups = z.groups
        c = za.c
        dev = self.device
        gamma, beta = self.named[bnkey + ".weight"], self.named[bnkey + ".bias"]
        mean = torch.empty((groups, c), dtype=torch.float32, device=dev)
        invstd = torch.empty((groups, c), dtype=torch.float32, device=dev)
        sums = torch.empty((groups, c, 2), dtype=torch.float64, device=dev)
        y = self._act(za.n, za.h, za.w, c)
        d = BnDesc()
        d.z, d.z_plane = za.ptr, za.plane
        d.n, d.h, d.w, d.c, d.groups = za.n, za.h, za.w, c, groups
        sn = self.sn[snkey] if snkey is not None else None
        d.inv_sigma = sn["isig"].data_ptr() if sn is not None else None
        d.mode, d.act = mode, act
        d.gamma, d.beta, d.mean, d.invstd = gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), invstd.data_ptr()
        if res1 is not None:
            ra = res1.a
            assert ra.c == c and ra.n == za.n and ra.h == za.h >> res1_shift and ra.w == za.w >> res1_shift, bnkey
            d.res1, d.res1_plane, d.res1_shift = ra.ptr, ra.plane, res1_shift
        if res2 is not None:
            assert res2.a.c == c and res2.a.n == za.n and res2.a.h == za.h and res2.a.w == za.w, bnkey
            d.res2, d.res2_plane = res2.a.ptr, res2.a.plane
        d.y, d.y_plane = y.ptr, y.plane
        tag = None
        if getattr(self, "_prof", None) is not None:
            tag = dict(tag=f"c{c} px{za.n * za.h * za.w} m{mode}{'r1' if res1 is not None else ''}{'r2' if res2 is not None else ''}")
        self._call("tcv_bn_stats", C.byref(d), sums.data_ptr(), meta=tag)
        count = float((za.n // groups) * za.h * za.w)
        if self.sync_bn:
            self._allreduce_sums(sums)
            count *= self.world
        self._call("tcv_bn_finalize", sums.data_ptr(), count, count * unbias_mul, groups, c, BN_EPS, BN_MOMENTUM,
                   mean.data_ptr(), invstd.data_ptr(), self.named[bnkey + ".running_mean"].data_ptr(),
                   self.named[bnkey + ".running_var"].data_ptr())
        for sfx in (".running_mean", ".running_var"):       # written by the kernel: tell torch (version counters)
            torch.autograd.graph.increment_version(self.named[bnkey + sfx])
        nbt = self.named.get(bnkey + ".num_batches_tracked")
        if nbt is not None:
            self._nbt.append((nbt, groups))
        self._call("tcv_bn_apply", C.byref(d), meta=tag)
        out = TAct(y, groups)
        keep = (mean, invstd, gamma, beta)

        def backward():
            dy = out.g
            if dy is None:
                return
            _ = keep
            bsums = torch.empty((groups, c, 2), dtype=torch.float64, device=dev)
            e_is_dy = 0
            if mode == 1 and res1 is not None:
                e = self._act(za.n, za.h, za.w, c)
                self._call("tcv_bn_bwd_reduce", C.byref(d), dy.ptr, dy.plane, e.ptr, e.plane, bsums.data_ptr(), meta=tag)
            else:
                # mode 2, or mode 1 without a residual input: nothing but tcv_bn_bwd_apply needs e = dy * act'(...), which it
                # recomputes from dy (one full-tensor write less)
                e = dy
                e_is_dy = 1 if mode == 1 else 0
                self._call("tcv_bn_bwd_reduce", C.byref(d), dy.ptr, dy.plane, None, 0, bsums.data_ptr(), meta=tag)
            dg, db = self.dbn.get(bnkey, (None, None))
            if dg is None:
                dg, db = self._zeros((c,)), self._zeros((c,))
                self.dbn[bnkey] = (dg, db)
            self._call("tcv_bn_param_grads", bsums.data_ptr(), groups, c, dg.data_ptr(), db.data_ptr())
            if self.sync_bn:
                self._allreduce_sums(bsums)
            dz = self._act(za.n, za.h, za.w, c)
            self._call("tcv_bn_bwd_apply", C.byref(d), e.ptr, e.plane, bsums.data_ptr(), count, dz.ptr, dz.plane,
                       sn["zdot"].data_ptr() if sn is not None else None, e_is_dy, meta=tag)
            self._acc(z, dz, True)
            if res1 is not None and res1.needs_grad:
                if res1_shift:
                    assert res1_shift == 1
                    pooled = self._act(za.n, za.h // 2, za.w // 2, c)
                    self._call("tcv_pool2_scaled", e.ptr, za.n, za.h, za.w, c, 1.0, pooled.ptr)
                    self._acc(res1, pooled, True)
                else:
                    self._acc(res1, e, True)
            if res2 is not None and res2.needs_grad:
                self._acc(res2, dy, out.g_owned and mode == 1)
            out.g = None
        self.tape.append(backward)
        return out

    def conv_bn(self, x: TAct, wkey: str, bnkey: str, *, stride=1, pad=PAD_ZERO, prepadded=False, mode=1, act=ACT_NONE,
                res1=None, res1_shift=0, res2=None, unbias_mul=1) -> TAct:
        z = self.conv_op(x, wkey, stride=stride, pad=pad, prepadded=prepadded)
        return self.bn_op(z, bnkey, mode=mode, act=act, snkey=wkey if self.w[wkey].get("sn") else None, res1=res1,
                          res1_shift=res1_shift, res2=res2, unbias_mul=unbias_mul)

    # ------------------------------------------------------------------ small structural ops
    def avgpool_op(self, x: TAct) -> TAct:
        xa = x.a
        y = self._act(xa.n, xa.h // 2, xa.w // 2, xa.c)
        self._call("tcv_pool2_scaled", xa.ptr, xa.n, xa.h, xa.w, xa.c, 0.25, y.ptr)
        out = TAct(y, x.groups)

        def backward():
            if out.g is None or not x.needs_grad:
                return
            dx = self._act(xa.n, xa.h, xa.w, xa.c)
            self._call("tcv_upsample2_scaled", out.g.ptr, xa.n, xa.h // 2, xa.w // 2, xa.c, 0.25, dx.ptr)
            self._acc(x, dx, True)
            out.g = None
        self.tape.append(backward)
        return out

    def pad_reflect_op(self, x: TAct) -> TAct:
        xa = x.a
        y = self._act(xa.n, xa.h + 2, xa.w + 2, xa.c)
        self._call("tcv_pad_reflect1", xa.ptr, xa.n, xa.h, xa.w, xa.c, y.ptr)
        out = TAct(y, x.groups, needs_grad=x.needs_grad)

        def backward():
            if out.g is None or not x.needs_grad:
                return
            dx = self._act(xa.n, xa.h, xa.w, xa.c)
            self._call("tcv_pad_reflect1_bwd", out.g.ptr, xa.n, xa.h, xa.w, xa.c, dx.ptr)
            self._acc(x, dx, True)
            out.g = None
        self.tape.append(backward)
        return out

    def gather_op(self, x: TAct, B: int, S: int, ncen: int, off: int) -> TAct:
        """frames (b, off + j), j < ncen, of every sample as one batch (decoder-tail inputs, VMN_model.py:107-110)."""
        xa = x.a
        y = self._act(B * ncen, xa.h, xa.w, xa.c)
        self._call("tcv_copy_images", xa.ptr, xa.plane, y.ptr, y.plane, xa.img_elems, B * ncen, ncen, S, off, ncen, 0, 0)
        out = TAct(y, ncen, needs_grad=x.needs_grad)

        def backward():
            if out.g is None or not x.needs_grad:
                return
            if x.g is None:
                x.g, x.g_owned = self._zeros_act(xa.n, xa.h, xa.w, xa.c), True
            elif not x.g_owned:
                buf = x.g.buf.clone()
                x.g = Act(buf, xa.n, xa.h, xa.w, xa.c, buf.data_ptr(), x.g.plane)
                x.g_owned = True
            self._call("tcv_copy_images", out.g.ptr, out.g.plane, x.g.ptr, x.g.plane, xa.img_elems, B * ncen, ncen,
                       ncen, 0, S, off, 1)
            out.g = None
        self.tape.append(backward)
        return out

    # ------------------------------------------------------------------ guided contextual attention
    def gca_op(self, p: str, im_fea: TAct, feat: TAct, unknown: torch.Tensor) -> TAct:
        fa = feat.a
        n, h, w = fa.n, fa.h, fa.w
        assert fa.c == 128 and im_fea.a.c == 128 and h % 2 == 0 and w % 2 == 0
        g = self.conv_op(im_fea, _k(p, "guidance_conv"), stride=2, bias=True)
        ga = g.a
        P = (h // 2) * (w // 2)
        P_pad = (P + 63) // 64 * 64
        dev = self.device
        f32 = torch.float32
        mm = torch.empty((n, P), dtype=f32, device=dev)
        scales = torch.empty((n, 2), dtype=f32, device=dev)
        O = torch.empty((n, P, 2048), dtype=f32, device=dev)
        if self.use_tc_attn and self.gca_shift_sum_train:
            return self._gca_shift_sum(p, feat, g, unknown, mm, scales)
        Q = torch.empty((2, n, P, 576), dtype=torch.bfloat16, device=dev)
        Kn = torch.empty_like(Q)
        self._call("tcv_gca_prep", ga.ptr, unknown.data_ptr(), n, h, w, Q.data_ptr(), Kn.data_ptr(), mm.data_ptr(),
                   scales.data_ptr(), 2)
        Vt = torch.empty((2, n, 2048, P_pad), dtype=torch.bfloat16, device=dev)
        self._call("tcv_gca_values", fa.ptr, n, h, w, Vt.data_ptr(), 2)
        Sm = torch.empty((n, P, P_pad), dtype=f32, device=dev)
        self._call("tcv_gemm_tn_tc", Q.data_ptr(), n * P * 576, Kn.data_ptr(), n * P * 576, Sm.data_ptr(), P, P, 576,
                   P_pad, P * P_pad, n, 3, 0, 0)
        Pb = torch.empty((2, n, P, P_pad), dtype=torch.bfloat16, device=dev)
        self._call("tcv_gca_softmax", Sm.data_ptr(), mm.data_ptr(), n, P, P_pad, Pb.data_ptr(), 2)
        del Sm
        self._call("tcv_gemm_tn_tc", Pb.data_ptr(), n * P * P_pad, Vt.data_ptr(), n * 2048 * P_pad, O.data_ptr(), P, 2048,
                   P_pad, 2048, P * 2048, n, 3, 0, 0)
        if not self.use_tc_attn:
            Q = Kn = Vt = None
        Ya = self._act(n, h, w, 128)
        self._call("tcv_gca_fold", O.data_ptr(), n, h, w, Ya.ptr)
        Y = TAct(Ya, feat.groups)

        bf16 = torch.bfloat16

        def transpose(src, rows, cols, ld_in, ld_out):
            """split-bf16 [2][n][rows][ld_in] -> [2][n][cols][ld_out] (K-major operand of the next GEMM)"""
            out = torch.empty((2, n, cols, ld_out), dtype=bf16, device=dev)
            self._call("tcv_transpose_planes", src.data_ptr(), n * rows * ld_in, rows, cols, ld_in, rows * ld_in,
                       out.data_ptr(), n * cols * ld_out, ld_out, cols * ld_out, n)
            return out

        def gemm_tc(A_, a_rows, B_, b_rows, K_, C_, ldc):
            """C[n][a_rows][ldc] = A[n][a_rows][K] . B[n][b_rows][K]^T on the tensor cores (bf16x3)"""
            self._call("tcv_gemm_tn_tc", A_.data_ptr(), n * a_rows * K_, B_.data_ptr(), n * b_rows * K_, C_.data_ptr(),
                       a_rows, b_rows, K_, ldc, a_rows * ldc, n, 3, 0, 0)

        def backward():
            dY = Y.g
            if dY is None:
                return
            tc = self.use_tc_attn
            dO = torch.empty((n, P, 2048), dtype=f32, device=dev)
            dO_s = torch.empty((2, n, P, 2048), dtype=bf16, device=dev) if tc else None
            delta = torch.empty((n, P), dtype=f32, device=dev)
            self._call("tcv_gca_fold_bwd", dY.ptr, O.data_ptr(), n, h, w, dO.data_ptr(), delta.data_ptr(),
                       dO_s.data_ptr() if tc else None)
            A = torch.empty((n, P, P_pad), dtype=f32, device=dev)
            self._call("tcv_split_to_f32", Pb.data_ptr(), n * P * P_pad, n * P * P_pad, A.data_ptr())
            dA = torch.empty((n, P, P_pad), dtype=f32, device=dev)
            if tc:
                # dA[q,p] = sum_d dO[q,d] V[p,d]
                V_s = transpose(Vt, 2048, P, P_pad, 2048)
                gemm_tc(dO_s, P, V_s, P, 2048, dA, P_pad)
                del V_s
                dS_s = torch.empty((2, n, P, P_pad), dtype=bf16, device=dev)
                self._call("tcv_gca_softmax_bwd", A.data_ptr(), dA.data_ptr(), delta.data_ptr(), n, P, P_pad,
                           dS_s.data_ptr())
                if feat.needs_grad:
                    # dV[p,d] = sum_q A[q,p] dO[q,d]
                    dV = torch.empty((n, P, 2048), dtype=f32, device=dev)
                    gemm_tc(transpose(Pb, P, P, P_pad, P_pad), P, transpose(dO_s, P, 2048, 2048, P_pad), 2048, P_pad, dV, 2048)
                    dfeat = self._act(n, h, w, 128)
                    self._call("tcv_gca_values_bwd", dV.data_ptr(), n, h, w, dfeat.ptr)
                    self._acc(feat, dfeat, True)
                Q32 = torch.empty((n, P, 576), dtype=f32, device=dev)
                self._call("tcv_split_to_f32", Q.data_ptr(), n * P * 576, n * P * 576, Q32.data_ptr())
                dQ = torch.empty((n, P, 576), dtype=f32, device=dev)
                dKn = torch.empty((n, P, 576), dtype=f32, device=dev)
                # dQ[q,c] = sum_p dS[q,p] Kn[p,c] ; dKn[p,c] = sum_q dS[q,p] Q[q,c]
                gemm_tc(dS_s, P, transpose(Kn, P, 576, 576, P_pad), 576, P_pad, dQ, 576)
                gemm_tc(transpose(dS_s, P, P, P_pad, P_pad), P, transpose(Q, P, 576, 576, P_pad), 576, P_pad, dKn, 576)
            else:
                V32 = torch.empty((n, 2048, P_pad), dtype=f32, device=dev)
                self._call("tcv_gca_values", fa.ptr, n, h, w, V32.data_ptr(), 0)
                self._call("tcv_gemm_f32_strided", dO.data_ptr(), 2048, 1, V32.data_ptr(), 1, P_pad, dA.data_ptr(), P_pad,
                           P, P, 2048, P * 2048, 2048 * P_pad, P * P_pad, n, 0)
                self._call("tcv_gca_softmax_bwd", A.data_ptr(), dA.data_ptr(), delta.data_ptr(), n, P, P_pad, None)
                dS = dA
                if feat.needs_grad:
                    dV = torch.empty((n, P, 2048), dtype=f32, device=dev)
                    self._call("tcv_gemm_f32_strided", A.data_ptr(), 1, P_pad, dO.data_ptr(), 1, 2048, dV.data_ptr(), 2048,
                               P, 2048, P, P * P_pad, P * 2048, P * 2048, n, 0)
                    dfeat = self._act(n, h, w, 128)
                    self._call("tcv_gca_values_bwd", dV.data_ptr(), n, h, w, dfeat.ptr)
                    self._acc(feat, dfeat, True)
                Q32 = torch.empty((n, P, 576), dtype=f32, device=dev)
                K32 = torch.empty((n, P, 576), dtype=f32, device=dev)
                mm2 = torch.empty((n, P), dtype=f32, device=dev)
                sc2 = torch.empty((n, 2), dtype=f32, device=dev)
                self._call("tcv_gca_prep", ga.ptr, unknown.data_ptr(), n, h, w, Q32.data_ptr(), K32.data_ptr(),
                           mm2.data_ptr(), sc2.data_ptr(), 0)
                dQ = torch.empty((n, P, 576), dtype=f32, device=dev)
                dKn = torch.empty((n, P, 576), dtype=f32, device=dev)
                self._call("tcv_gemm_f32_strided", dS.data_ptr(), P_pad, 1, K32.data_ptr(), 1, 576, dQ.data_ptr(), 576,
                           P, 576, P, P * P_pad, P * 576, P * 576, n, 0)
                self._call("tcv_gemm_f32_strided", dS.data_ptr(), 1, P_pad, Q32.data_ptr(), 1, 576, dKn.data_ptr(), 576,
                           P, 576, P, P * P_pad, P * 576, P * 576, n, 0)
            dg = self._act(n, h // 2, w // 2, 64)
            self._call("tcv_gca_prep_bwd", dQ.data_ptr(), dKn.data_ptr(), Q32.data_ptr(), mm.data_ptr(),
                       scales.data_ptr(), n, h, w, dg.ptr)
            self._acc(g, dg, True)
            Y.g = None
        self.tape.append(backward)
        self.last_gca_scales = scales
        return self.conv_bn(Y, _k(p, "W.0"), _k(p, "W.1"), res1=feat)

    def _gca_shift_sum(self, p: str, feat: TAct, g: TAct, unknown: torch.Tensor, mm: torch.Tensor,
                       scales: torch.Tensor) -> TAct:
        """Attention and its backward in the shift-sum form on the padded (hh+1) x (ww+1) key grid -- the same forward
        kernels as the inference engine (engine.py:_gca_shift_sum) plus csrc/gca_train2.cu: with
        A2[m][p'] = sum_a A[m-a][p'-a], fold(A.V)/4 = unfold_parity(A2.F), so the value GEMM and its two backward GEMMs are
        [Pk x 512 x Pk] instead of [P x 2048 x P], and shift-add / gather are linear column offsets on the grid.
        Reference: GCA/ops.py:106-229 (scores, masked softmax, conv_transpose2d aggregation) and its autograd."""
        fa, ga = feat.a, g.a
        n, h, w = fa.n, fa.h, fa.w
        hh, ww = h // 2, w // 2
        P = hh * ww
        P_pad = (P + 63) // 64 * 64
        Pk = (hh + 1) * (ww + 1)
        ld = (Pk + 63) // 64 * 64
        dev, f32, bf16 = self.device, torch.float32, torch.bfloat16

        def transpose(src, rows, cols, ld_in, ld_out):
            out = torch.empty((2, n, cols, ld_out), dtype=bf16, device=dev)
            self._call("tcv_transpose_planes", src.data_ptr(), n * rows * ld_in, rows, cols, ld_in, rows * ld_in,
                       out.data_ptr(), n * cols * ld_out, ld_out, cols * ld_out, n)
            return out

        def gemm_tc(A_, a_rows, B_, b_rows, K_, C_, ldc):
            self._call("tcv_gemm_tn_tc", A_.data_ptr(), n * a_rows * K_, B_.data_ptr(), n * b_rows * K_, C_.data_ptr(),
                       a_rows, b_rows, K_, ldc, a_rows * ldc, n, 3, 0, 0)

        def gemm_ex(A_, a_mn, B_, b_mn, M_, N_, K_, C_, ldc):
            """C[M x N] = A . B^T on the CTA-pair kernel; an operand [2][n][r][ld] is K-major (r = its rows, K_ <= ld) or
            MN-major (r = K_ rows of the reduction index, its rows <= ld): no transposed copies (tcv_gemm_tc_ex)."""
            ra, la = A_.shape[2], A_.shape[3]
            rb, lb = B_.shape[2], B_.shape[3]
            self._call("tcv_gemm_tc_ex", A_.data_ptr(), n * ra * la, la, ra * la, int(a_mn), B_.data_ptr(), n * rb * lb, lb,
                       rb * lb, int(b_mn), C_.data_ptr(), M_, N_, K_, ldc, M_ * ldc, n)

        # the backward GEMMs read their operands as the forward left them when the CTA-pair kernel applies
        mn_ok = self.gca_mn_gemm and P >= 512

        Q = torch.empty((2, n, P, 576), dtype=bf16, device=dev)
        Kn = torch.empty((2, n, Pk, 576), dtype=bf16, device=dev)       # zero rows at the pad keys
        self._call("tcv_gca_prep_grid", ga.ptr, unknown.data_ptr(), n, h, w, Q.data_ptr(), Kn.data_ptr(), mm.data_ptr(),
                   scales.data_ptr())
        A = torch.empty((n, P, ld), dtype=f32, device=dev)              # logits, then probabilities (kept for backward)
        gemm_tc(Q, P, Kn, Pk, 576, A, ld)
        stats = torch.empty((n, P, 2), dtype=f32, device=dev)
        self._call("tcv_gca_rowstats", A.data_ptr(), mm.data_ptr(), n, h, w, ld, stats.data_ptr(), 1)
        A2 = torch.empty((2, n, Pk, ld), dtype=bf16, device=dev)
        self._call("tcv_gca_shift_add", A.data_ptr(), n, h, w, ld, A2.data_ptr())
        Ft = torch.empty((2, n, 512, ld), dtype=bf16, device=dev)
        self._call("tcv_gca_values_parity", fa.ptr, n, h, w, ld, Ft.data_ptr())
        O2 = torch.empty((n, Pk, 512), dtype=f32, device=dev)
        gemm_tc(A2, Pk, Ft, 512, ld, O2, 512)
        Ya = self._act(n, h, w, 128)
        self._call("tcv_gca_unfold_parity", O2.data_ptr(), n, h, w, Ya.ptr)
        del O2, stats
        Y = TAct(Ya, feat.groups)

        def backward():
            dY = Y.g
            if dY is None:
                return
            dO2 = torch.empty((2, n, Pk, 512), dtype=bf16, device=dev)
            self._call("tcv_gca_unfold_parity_bwd", dY.ptr, n, h, w, dO2.data_ptr())
            # dA2[m,p'] = sum_d dO2[m,d] F[p',d]
            dA2 = torch.empty((n, Pk, ld), dtype=f32, device=dev)
            if mn_ok:
                gemm_ex(dO2, False, Ft, True, Pk, Pk, 512, dA2, ld)           # F^T: Ft [512][ld] is F MN-major
            else:
                gemm_tc(dO2, Pk, transpose(Ft, 512, Pk, ld, 512), Pk, 512, dA2, ld)
            # dS = A * (gather(dA2) - <A, gather(dA2)>): gather, row dot and softmax backward in one kernel
            dS_s = torch.empty((2, n, P, ld), dtype=bf16, device=dev)
            self._call("tcv_gca_softmax_bwd_grid", A.data_ptr(), dA2.data_ptr(), n, h, w, ld, dS_s.data_ptr())
            del dA2
            if feat.needs_grad:
                # dF[p',d] = sum_m A2[m,p'] dO2[m,d]
                dF = torch.empty((n, Pk, 512), dtype=f32, device=dev)
                if mn_ok:
                    gemm_ex(A2, True, dO2, True, Pk, 512, Pk, dF, 512)
                else:
                    gemm_tc(transpose(A2, Pk, Pk, ld, ld), Pk, transpose(dO2, Pk, 512, 512, ld), 512, ld, dF, 512)
                dfeat = self._act(n, h, w, 128)
                self._call("tcv_gca_values_parity_bwd", dF.data_ptr(), n, h, w, dfeat.ptr)
                self._acc(feat, dfeat, True)
            Q32 = torch.empty((n, P, 576), dtype=f32, device=dev)
            self._call("tcv_split_to_f32", Q.data_ptr(), n * P * 576, n * P * 576, Q32.data_ptr())
            dQ = torch.empty((n, P, 576), dtype=f32, device=dev)
            dKn = torch.empty((n, Pk, 576), dtype=f32, device=dev)
            # dQ[q,c] = sum_p' dS[q,p'] Kn[p',c] ; dKn[p',c] = sum_q dS[q,p'] Q[q,c]
            if mn_ok:
                gemm_ex(dS_s, False, Kn, True, P, 576, Pk, dQ, 576)
                gemm_ex(dS_s, True, Q, True, Pk, 576, P, dKn, 576)
            else:
                gemm_tc(dS_s, P, transpose(Kn, Pk, 576, 576, ld), 576, ld, dQ, 576)
                gemm_tc(transpose(dS_s, P, Pk, ld, P_pad), Pk, transpose(Q, P, 576, 576, P_pad), 576, P_pad, dKn, 576)
            dg = self._act(n, h // 2, w // 2, 64)
            self._call("tcv_gca_prep_bwd_grid", dQ.data_ptr(), dKn.data_ptr(), Q32.data_ptr(), mm.data_ptr(),
                       scales.data_ptr(), n, h, w, dg.ptr)
            self._acc(g, dg, True)
            Y.g = None
        self.tape.append(backward)
        self.last_gca_scales = scales
        return self.conv_bn(Y, _k(p, "W.0"), _k(p, "W.1"), res1=feat)

    # ------------------------------------------------------------------ temporal attention module
    def tam_op(self, p: str, x: TAct, xb: TAct, xf: TAct, mask: torch.Tensor, H: int, W: int, attb: torch.Tensor,
               attf: torch.Tensor, sm: torch.Tensor, datt: dict) -> TAct:
        q = self.conv_op(x, _k(p, "query_conv"), bias=True)
        v = self.conv_op(x, _k(p, "value_conv"), bias=True)
        kb = self.conv_op(xb, _k(p, "key_conv"), bias=True)
        kf = self.conv_op(xf, _k(p, "key_conv"), bias=True)
        xa = x.a
        oa = self._act(xa.n, xa.h, xa.w, xa.c)
        self._call("tcv_tam_attend", q.a.ptr, v.a.ptr, kb.a.ptr, kf.a.ptr, mask.data_ptr(), H * W, H, W, xa.n, xa.h,
                   xa.w, xa.c, self.window, oa.ptr, attb.data_ptr(), attf.data_ptr(), sm.data_ptr())
        out = TAct(oa, x.groups)

        def backward():
            do = out.g
            if do is None:
                return
            dq = self._act(xa.n, xa.h, xa.w, xa.c)
            dkb32 = torch.empty((xa.n, xa.h, xa.w, xa.c), dtype=torch.float32, device=self.device)
            dkf32 = torch.empty_like(dkb32)
            db_, df_ = datt.get("b"), datt.get("f")
            self._call("tcv_tam_attend_bwd", q.a.ptr, kb.a.ptr, kf.a.ptr, mask.data_ptr(), H * W, H, W, xa.n, xa.h,
                       xa.w, xa.c, self.window, do.ptr, db_.data_ptr() if db_ is not None else None,
                       df_.data_ptr() if df_ is not None else None, dq.ptr, dkb32.data_ptr(), dkf32.data_ptr())
            self._acc(q, dq, True)
            self._acc(v, do, out.g_owned)
            for t32, t in ((dkb32, kb), (dkf32, kf)):
                ga = self._act(xa.n, xa.h, xa.w, xa.c)
                self._call("tcv_f32_to_split", t32.data_ptr(), t32.numel(), ga.ptr, ga.plane)
                self._acc(t, ga, True)
            out.g = None
        self.tape.append(backward)
        return out

    # ------------------------------------------------------------------ network program (train mode)
    def _enc_block_t(self, x: TAct, p: str, stride: int) -> TAct:
        o = self.conv_bn(x, p + ".conv1", p + ".bn1", stride=stride, act=ACT_RELU)
        idt = x
        if stride != 1:
            idt = self.conv_bn(self.avgpool_op(x), p + ".downsample.1", p + ".downsample.2")
        return self.conv_bn(o, p + ".conv2", p + ".bn2", res1=idt, act=ACT_RELU)

    def _shortcut_t(self, x: TAct, p: str) -> TAct:
        o = self.conv_bn(x, p + ".0", p + ".2", mode=2, act=ACT_RELU)
        return self.conv_bn(o, p + ".3", p + ".5", mode=2, act=ACT_RELU)

    def _dec_layer_t(self, x: TAct, p: str, blocks: int, res2: Optional[TAct]) -> TAct:
        for i in range(blocks):
            bp = f"{p}.{i}"
            last = res2 if i == blocks - 1 else None
            if i == 0:
                o = self.bn_op(self.deconv_op(x, bp + ".conv1"), bp + ".bn1", act=ACT_LEAKY02, snkey=bp + ".conv1")
                # 1x1 conv + BN evaluated before the nearest x2 upsample they commute with (statistics are identical;
                # only the unbiased-variance count of the running statistics sees the 4x larger tensor)
                idt = self.conv_bn(x, bp + ".upsample.1", bp + ".upsample.2", unbias_mul=4)
                x = self.conv_bn(o, bp + ".conv2", bp + ".bn2", res1=idt, res1_shift=1, act=ACT_LEAKY02, res2=last)
            else:
                o = self.conv_bn(x, bp + ".conv1", bp + ".bn1", act=ACT_LEAKY02)
                x = self.conv_bn(o, bp + ".conv2", bp + ".bn2", res1=x, act=ACT_LEAKY02, res2=last)
        return x

    def per_frame_t(self, x8: TAct) -> dict:
        e = "encoder"
        c1 = self.conv_bn(x8, e + ".conv1", e + ".bn1", stride=2, act=ACT_RELU)
        x1 = self.conv_bn(c1, e + ".conv2", e + ".bn2", act=ACT_RELU)
        c3 = self.conv_bn(x1, e + ".conv3", e + ".bn3", stride=2, act=ACT_RELU)
        g = x8
        for ci, bi in ((1, 3), (5, 7), (9, 11)):
            # reflect border materialised once (ReflectionPad2d(1), res_gca_enc.py:20-33): the stride-2 conv and both of
            # its gradients then run on the zero-padding-free "pre-padded" tensor-core paths
            g = self.conv_bn(self.pad_reflect_op(g), f"{e}.guidance_head.{ci}", f"{e}.guidance_head.{bi}", stride=2,
                             prepadded=True, mode=2, act=ACT_RELU)
        im_fea = g
        xa = x8.a
        unknown = torch.empty((xa.n, xa.h // 8, xa.w // 8), dtype=torch.float32, device=self.device)
        self._call("tcv_unknown_os8", xa.ptr, xa.n, xa.h, xa.w, unknown.data_ptr())
        feats = []
        cur = c3
        for name, planes, blocks, stride in ENC_LAYERS:
            if name == "layer3":
                cur = self.gca_op(e + ".gca", im_fea, cur, unknown)
                feats[-1] = cur
            for i in range(blocks):
                cur = self._enc_block_t(cur, f"{e}.{name}.{i}", stride if i == 0 else 1)
            feats.append(cur)
        x2, x3, x4, emb = feats
        fea = [self._shortcut_t(t, f"{e}.shortcut.{i}") for i, t in enumerate((x8, x1, x2, x3, x4))]
        d = self._dec_layer_t(emb, "decoder.layer1", DEC_LAYERS[0][2], fea[4])
        d = self._dec_layer_t(d, "decoder.layer2", DEC_LAYERS[1][2], fea[3])
        feat = self.gca_op("decoder.gca", im_fea, d, unknown)
        return dict(fea=fea, feat=feat)

    def begin_operator_step(self) -> None:
        """Fresh tape and gradient accumulators for ONE standalone operator call (the TAM / GCA drop-in modules in train
        mode, model._TamTrainFn / _GcaTrainFn); the whole-network step does the same at the top of `_train_forward`."""
        self.tape = []
        self.step_id = getattr(self, "step_id", 0) + 1
        self.dw.clear(); self.dbias.clear(); self.dbn.clear()
        self._arena_begin()
        self._nbt = []

    def end_operator_forward(self) -> None:
        if self._nbt:
            torch._foreach_add_([t_ for t_, _ in self._nbt], [int(g_) for _, g_ in self._nbt])
            self._nbt = []

    def run_tape(self) -> None:
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    def train_forward(self, x8a: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        """VMN.forward in train mode on preprocessed input; records the tape.  trimask fp32 [B,S,1,H,W]."""
        with self.stream_scope():
            return self._train_forward(x8a, trimask, B, S, H, W)

    def _train_forward(self, x8a: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        self.tape = []
        self.step_id = getattr(self, "step_id", 0) + 1       # stamps the autograd node (model._TrainStepFn / _VMNTrainFn)
        self.dw.clear(); self.dbias.clear(); self.dbn.clear()
        self._arena_begin()
        self._nbt = []
        ncen = S - 2
        self.spectral_norm_step(S, ncen)
        N8 = (H // 8) * (W // 8)
        w2 = self.window * self.window
        dev = self.device
        if self.freeze_backbone:
            # VMN_model.py:99-101: feature extraction under no_grad, in eval mode; no tape entries, no statistics updates
            if self.backbone is None:
                raise RuntimeError("tcvom_b200: freeze_backbone step without a backbone engine")
            # shortcut branches 0..2 feed the tail of CENTRE frames only (VMN_GCA.py:38-44): evaluated for those frames
            # (the reference computes them for the end frames as well and drops the result); 3 and 4 end in the head
            pfi = self.backbone.per_frame(x8a, shortcuts="head")
            feat = TAct(pfi["feat"], S, needs_grad=False)
            fea = []
            for i in range(3):
                src = self.gather_op(TAct(pfi["shortcut_src"][i], S, needs_grad=False), B, S, ncen, 1)
                fea.append(TAct(self.backbone._shortcut(src.a, f"encoder.shortcut.{i}"), ncen, needs_grad=False))
        else:
            pf = self.per_frame_t(TAct(x8a, S, needs_grad=False))
            feat = pf["feat"]
            fea = [self.gather_op(f, B, S, ncen, 1) for f in pf["fea"][:3]]
        x = self.gather_op(feat, B, S, ncen, 1)
        xb = self.gather_op(feat, B, S, ncen, 0)
        xf = self.gather_op(feat, B, S, ncen, 2)
        mask = trimask.reshape(B, S, H, W)[:, 1:S - 1].contiguous()            # centre-frame unknown masks
        pred = torch.empty((B, ncen, 1, H, W), dtype=torch.float32, device=dev)
        attb = torch.empty((B, ncen, w2, N8), dtype=torch.float32, device=dev)
        attf = torch.empty_like(attb)
        sm = torch.empty((B, ncen, 1, H // 8, W // 8), dtype=torch.uint8, device=dev)
        self.datt = {}
        t = self.tam_op("decoder.fam", x, xb, xf, mask, H, W, attb, attf, sm, self.datt)
        t = self._dec_layer_t(t, "decoder.layer3", DEC_LAYERS[2][2], fea[2])
        t = self._dec_layer_t(t, "decoder.layer4", DEC_LAYERS[3][2], fea[1])
        t = self.bn_op(self.deconv_op(t, "decoder.conv1"), "decoder.bn1", act=ACT_LEAKY02, snkey="decoder.conv1",
                       res2=fea[0])
        self.head_in = t
        hk = "decoder.conv2" + self.HEAD32
        if self.use_head32 and hk in self.w:
            self.head = self.conv_op(t, hk, bias=True)
            ha = self.head.a
            self._call("tcv_head_tanh01", ha.ptr, ha.plane, ha.n * ha.h * ha.w, ha.c, pred.data_ptr())
        else:
            self.head = self.conv_op(t, "decoder.conv2", bias=True, act=ACT_TANH01, f32_out=pred)
        self.pred = pred
        if self._nbt:                                   # num_batches_tracked += calls, one fused update for all layers
            torch._foreach_add_([t_ for t_, _ in self._nbt], [int(g_) for _, g_ in self._nbt])
            self._nbt = []
        return dict(pred=pred, attb=attb, attf=attf, small_mask=sm)

    def train_backward(self, dpred: torch.Tensor, dattb: Optional[torch.Tensor], dattf: Optional[torch.Tensor]) -> None:
        """Runs the tape in reverse from dL/dpred (fp32 [B,ncen,1,H,W]) and the TAM-logit gradients of L_af."""
        with self.stream_scope():
            self._train_backward(dpred, dattb, dattf)

    def _train_backward(self, dpred: torch.Tensor, dattb: Optional[torch.Tensor], dattf: Optional[torch.Tensor]) -> None:
        n = dpred.numel()
        hin = self.head_in.a
        if self.head.a is not None:              # padded 32-channel head
            dz = self._act(hin.n, hin.h, hin.w, self.head.a.c)
            self._call("tcv_head_tanh01_bwd", self.pred.data_ptr(), dpred.data_ptr(), n, dz.c, dz.ptr)
        else:
            dz = self._act(hin.n, hin.h, hin.w, 8)
            self._call("tcv_tanh01_bwd", self.pred.data_ptr(), dpred.data_ptr(), n, dz.ptr)
        self.head.g, self.head.g_owned = dz, True
        self.datt["b"], self.datt["f"] = dattb, dattf
        for fn in reversed(self.tape):
            fn()
        self.tape = []

    def collect_grads(self, names: List[str]) -> List[Optional[torch.Tensor]]:
        """Gradients in the torch parameter layout for the given parameter names (state_dict naming)."""
        out: List[Optional[torch.Tensor]] = []
        by_param = {e["param"]: (p, e) for p, e in self.w.items() if "param" in e}
        for name in names:
            prm = self.named[name]
            if name in by_param:
                p, ent = by_param[name]
                dw = self.dw.get(p)
                if dw is None:
                    out.append(torch.zeros_like(prm)); continue
                grad = torch.empty_like(prm)
                sn = self.sn.get(p) if ent.get("sn") else None
                cout, cin = ent.get("cout_real", ent["cout"]), ent["cin_real"]
                self._call("tcv_weight_grad_unpack", dw.data_ptr(), cout, cin, ent["k"], ent["k"],
                           1 if ent["transposed"] else 0, ent["cin"], dw.shape[2],
                           sn["u_hist"].data_ptr() if sn else None, sn["v_hist"].data_ptr() if sn else None,
                           sn["sigma"].data_ptr() if sn else None, sn["zdot"].data_ptr() if sn else None,
                           sn["calls"] if sn else 0, grad.data_ptr())
                out.append(grad)
            elif name.endswith(".bias") and name[: -len(".bias")] in self.bias:
                bkey = name[: -len(".bias")]
                if self.use_head32 and bkey + self.HEAD32 in self.w:
                    bkey = bkey + self.HEAD32
                db = self.dbias.get(bkey)
                out.append(db[: prm.numel()].clone() if db is not None else torch.zeros_like(prm))
            elif name.endswith((".weight", ".bias")) and name.rsplit(".", 1)[0] in self.dbn:
                dg, db = self.dbn[name.rsplit(".", 1)[0]]
                out.append(dg if name.endswith(".weight") else db)
            else:
                out.append(torch.zeros_like(prm) if prm.requires_grad else None)
        return out
