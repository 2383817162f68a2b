"""Host-side mirror of the reference's operator surface for the GCA+TAM path.

Same names, argument meaning and error behaviour as the reference (checkout f5fa07a):

  get_VMN_models(arch, agg_window, agg_reduction=1, freeze_backbone=False)   models/VMN/__init__.py:11-29
  VMN.forward(images, masks, extras=None)                                   models/VMN/VMN_model.py:83-113
  FeatureAggregationModule(input_chn, reduction, window).forward(x,b,f,mask) models/VMN/VMN_model.py:9-68
  GuidedCxtAtten(out_channels, guidance_channels, rate=2).forward(f, alpha, unknown)  models/GCA/ops.py:83-229
  EvalModel(model, dilate_kernel=None, eps=0, agg_window=...).forward(imgs, tris)      models/model.py:359-424

All compute goes through the C ABI (include/tcvom_b200.h).  There is no PyTorch/CPU fallback:
tensors must be on a CUDA device and the native library must be built, otherwise a
RuntimeError is raised.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _cabi
from .engine import Act, GcaVmnEngine, Plan, named_tensors
from .dim_engine import DimVmnEngine
from .dim_modules import DIMDecoderParams, DIMEncoderParams
from .fba_engine import FbaVmnEngine
from .index_engine import IndexVmnEngine
from .index_modules import IndexDecoderParams, IndexEncoderParams
from .fba_modules import FBADecoderParams, FBAEncoderParams
from .train_engine import FROZEN_PREFIXES, FrozenBackboneEngine, TAct, TrainEngine
from .modules import GCADecoderParams, GCAEncoderParams, GuidedCxtAttenParams, TAMParams


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"tcvom_b200: {what} must be a CUDA tensor (there is no CPU fallback)")


_ENGINE_LOCK = threading.Lock()


def _engine_for(module: nn.Module, window: int, engine_cls=GcaVmnEngine) -> GcaVmnEngine:
    """One engine per (module, device).  nn.DataParallel replicas share the module __dict__ (and so
    this table) but run one thread per device, so a per-device engine is never used concurrently."""
    dev = next(iter(named_tensors(module).values())).device
    if dev.type != "cuda":
        raise RuntimeError("tcvom_b200: the module must live on a CUDA device (no CPU fallback)")
    with _ENGINE_LOCK:
        table = module.__dict__.get("_engines")
        if table is None:
            table = module.__dict__["_engines"] = {}
        eng = table.get(dev.index)
        if eng is None:
            eng = table[dev.index] = engine_cls(window)
    eng.refresh_weights(module)
    return eng


def _train_engine_for(module: nn.Module, window: int) -> TrainEngine:
    """One training engine per (module, device); weights are re-packed from the parameters on every step."""
    dev = next(iter(named_tensors(module).values())).device
    if dev.type != "cuda":
        raise RuntimeError("tcvom_b200: the module must live on a CUDA device (no CPU fallback)")
    with _ENGINE_LOCK:
        table = module.__dict__.get("_train_engines")
        if table is None:
            table = module.__dict__["_train_engines"] = {}
        eng = table.get(dev.index)
        if eng is None:
            eng = table[dev.index] = TrainEngine(window)
    eng.refresh_weights(module)
    eng.freeze_backbone = bool(getattr(module, "freeze_backbone", False))
    if eng.freeze_backbone:
        if eng.backbone is None:
            eng.backbone = FrozenBackboneEngine(window)
        eng.backbone.refresh_weights(module)
    sync = any(isinstance(m, nn.SyncBatchNorm) for m in module.modules())
    dist = torch.distributed
    eng.sync_bn = bool(sync and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    eng.world = dist.get_world_size() if eng.sync_bn else 1
    return eng


def _trainable_names(net: nn.Module, named) -> List[str]:
    """Parameters the native step returns gradients for.  With ``freeze_backbone`` the frozen part runs under no_grad in the
    reference, so its parameters keep ``grad is None`` (an optimizer with weight decay must not touch them)."""
    names = [n for n, t in named.items() if isinstance(t, nn.Parameter) and t.requires_grad]
    if getattr(net, "freeze_backbone", False):
        names = [n for n in names if not n.startswith(FROZEN_PREFIXES)]
    return names


class _TrainStepFn(torch.autograd.Function):
    """Autograd boundary of the native training step: forward = train-mode network + losses on the sm_100a
    kernels, backward = the native tape; the trainable parameters are the differentiable inputs, so
    ``loss.backward()`` / DistributedDataParallel see ordinary ``.grad`` accumulation (train_ddp.py:63-64)."""

    @staticmethod
    def forward(ctx, wrapper, a, fg, bg, *params):
        st = wrapper._train_forward(a, fg, bg)
        ctx.wrapper, ctx.st = wrapper, st
        ctx.step_id = st["eng"].step_id
        vis = tuple(st[k] for k in ("imgs", "tris_vis", "alphas", "comps", "gts", "fgs", "bgs"))
        ctx.mark_non_differentiable(*vis)
        return (st["losses"],) + vis

    @staticmethod
    def backward(ctx, gl, *unused):
        if ctx.st is None:
            raise RuntimeError("tcvom_b200: the native training step was already back-propagated once "
                               "(retain_graph / double backward are not supported)")
        eng = ctx.st["eng"]
        if eng.step_id != ctx.step_id or not eng.tape:
            # the engine keeps ONE tape / set of gradient accumulators: a second train-mode forward (another micro-batch,
            # or a no_grad forward in train mode) replaced the state this node needs.  Plain autograd would support the
            # pattern; silently back-propagating the other step's tape would not be the same thing.
            raise RuntimeError("tcvom_b200: another train-mode forward ran on this module before backward(); call "
                               "loss.backward() after every forward (gradient accumulation over micro-batches works "
                               "that way: forward, backward, forward, backward, optimizer.step())")
        grads = ctx.wrapper._train_backward(ctx.st, gl)
        ctx.st = None
        return (None, None, None, None) + tuple(grads)


class _VMNTrainFn(torch.autograd.Function):
    """Autograd boundary at the plugin seam: VMN.forward in train mode (the reference's own FullModel_VMD computes the
    losses on top with torch).  Differentiable outputs: preds of the centre frames and the TAM logits."""

    @staticmethod
    def forward(ctx, net, x8, trimask, B, S, H, W, *params):
        eng = _train_engine_for(net, int(net.decoder.fam.window))
        out = eng.train_forward(x8, trimask, B, S, H, W)
        ctx.eng, ctx.names = eng, net.__dict__["_train_param_names"]
        ctx.step_id = eng.step_id
        ctx.mark_non_differentiable(out["small_mask"])
        return out["pred"], out["attb"], out["attf"], out["small_mask"]

    @staticmethod
    def backward(ctx, dpred, dattb, dattf, _):
        eng = ctx.eng
        if eng is None or not eng.tape or eng.step_id != ctx.step_id:
            raise RuntimeError("tcvom_b200: the native VMN step was already back-propagated once, or another train-mode "
                               "forward ran on this module before backward() (retain_graph / double backward / two "
                               "forwards before one backward are not supported)")
        z = lambda g, like: g.contiguous().float() if g is not None else torch.zeros_like(like)
        eng.train_backward(z(dpred, eng.pred), dattb.contiguous().float() if dattb is not None else None,
                           dattf.contiguous().float() if dattf is not None else None)
        return (None,) * 7 + tuple(eng.collect_grads(ctx.names))


def _nchw_to_act(t: torch.Tensor) -> Act:
    n, c, h, w = t.shape
    a = Act.empty(n, h, w, c, t.device)
    _cabi.check(_cabi.lib().tcv_nchw_to_split(t.detach().contiguous().float().data_ptr(), n, c, h, w, c, a.ptr, 0,
                                              _stream(t.device)), "nchw_to_split")
    return a


def _act_to_nchw(a: Optional[Act], like: torch.Tensor) -> torch.Tensor:
    if a is None:
        return torch.zeros_like(like, dtype=torch.float32)
    y = torch.empty((a.n, a.c, a.h, a.w), dtype=torch.float32, device=like.device)
    _cabi.check(_cabi.lib().tcv_split_to_nchw(a.ptr, a.n, a.c, a.h, a.w, a.c, 0, y.data_ptr(), _stream(like.device)),
                "split_to_nchw")
    return y


def _op_param_names(mod: nn.Module) -> List[str]:
    return [n for n, t in named_tensors(mod).items() if isinstance(t, nn.Parameter) and t.requires_grad]


def _check_op_tape(ctx):
    eng = ctx.eng
    if eng is None or not eng.tape or eng.step_id != ctx.step_id:
        raise RuntimeError("tcvom_b200: this operator call was already back-propagated, or the same module ran another "
                           "train-mode forward before backward() (one tape per module: forward, backward, forward, ...)")
    return eng


class _NoCtx:
    """Stand-in for the autograd context when an operator's train-mode forward runs under no_grad."""

    def mark_non_differentiable(self, *a):
        pass


class _TamTrainFn(torch.autograd.Function):
    """FeatureAggregationModule.forward in train mode (VMN_model.py:18-68) on the training engine's TAM operator: the
    three 3x3 convolutions and the windowed attention forward, their data / weight / bias gradients and the gradient of
    the returned attention logits backward."""

    @staticmethod
    def forward(ctx, mod, x, b, f, mask, *params):
        eng = _train_engine_for(mod, int(mod.window))
        B, Cc, H, W = x.shape
        w2 = int(mod.window) ** 2
        with eng.stream_scope():
            eng.begin_operator_step()
            ts = [TAct(_nchw_to_act(t), 1) for t in (x, b, f)]
            m = mask.detach().contiguous().float()
            mh, mw = m.shape[-2:]
            attb = torch.empty((B, w2, H * W), dtype=torch.float32, device=x.device)
            attf = torch.empty_like(attb)
            sm = torch.empty((B, 1, H, W), dtype=torch.uint8, device=x.device)
            ctx.datt = {}
            out = eng.tam_op("", ts[0], ts[1], ts[2], m, mh, mw, attb, attf, sm, ctx.datt)
            y = _act_to_nchw(out.a, x)
        ctx.eng, ctx.ts, ctx.out, ctx.mask = eng, ts, out, m
        ctx.step_id, ctx.names, ctx.like = eng.step_id, _op_param_names(mod), (x, b, f)
        smb = sm.bool()
        ctx.mark_non_differentiable(smb)
        return y, attb, attf, smb

    @staticmethod
    def backward(ctx, dy, dattb, dattf, _):
        eng = _check_op_tape(ctx)
        with eng.stream_scope():
            ctx.out.g = _nchw_to_act(dy if dy is not None else torch.zeros_like(ctx.like[0]))
            ctx.out.g_owned = True
            ctx.datt["b"] = dattb.contiguous().float() if dattb is not None else None
            ctx.datt["f"] = dattf.contiguous().float() if dattf is not None else None
            eng.run_tape()
            gin = [_act_to_nchw(t.g, like) for t, like in zip(ctx.ts, ctx.like)]
            pg = eng.collect_grads(ctx.names)
        ctx.eng = None
        return (None, gin[0], gin[1], gin[2], None) + tuple(pg)


class _GcaTrainFn(torch.autograd.Function):
    """GuidedCxtAtten.forward in train mode (GCA/ops.py:106-229): guidance 1x1 conv, guided attention, W = 1x1 conv +
    batch-statistics BatchNorm (running statistics updated), residual; gradients for both inputs and all parameters."""

    @staticmethod
    def forward(ctx, mod, f, alpha, unknown, *params):
        eng = _train_engine_for(mod, 1)
        B, Cc, H, W = alpha.shape
        with eng.stream_scope():
            eng.begin_operator_step()
            tf, ta = TAct(_nchw_to_act(f), 1), TAct(_nchw_to_act(alpha), 1)
            unk = unknown.detach().contiguous().float().reshape(B, H, W)
            out = eng.gca_op("", tf, ta, unk)
            eng.end_operator_forward()
            y = _act_to_nchw(out.a, alpha)
            scales = eng.last_gca_scales.clone()
        ctx.eng, ctx.ts, ctx.out, ctx.unk = eng, (tf, ta), out, unk
        ctx.step_id, ctx.names, ctx.like = eng.step_id, _op_param_names(mod), (f, alpha)
        ctx.mark_non_differentiable(scales)
        return y, scales

    @staticmethod
    def backward(ctx, dy, _):
        eng = _check_op_tape(ctx)
        with eng.stream_scope():
            ctx.out.g = _nchw_to_act(dy if dy is not None else torch.zeros_like(ctx.like[1]))
            ctx.out.g_owned = True
            eng.run_tape()
            gin = [_act_to_nchw(t.g, like) for t, like in zip(ctx.ts, ctx.like)]
            pg = eng.collect_grads(ctx.names)
        ctx.eng = None
        return (None, gin[0], gin[1], None) + tuple(pg)


class _OpEngineMixin:
    """Gives a standalone operator module (TAM / GCA) its own small engine over its parameters."""

    def _op_engine(self, window=1) -> GcaVmnEngine:
        return _engine_for(self, window)


class FeatureAggregationModule(TAMParams, _OpEngineMixin):
    """Temporal Attention Module (VMN_model.py:9-68) as a drop-in operator: NCHW fp32 in/out."""

    def forward(self, x, b, f, mask):
        for t, nme in ((x, "x"), (b, "b"), (f, "f"), (mask, "mask")):
            _require_cuda(t, nme)
        B, Cc, H, W = x.shape
        assert b.shape == x.shape and f.shape == x.shape          # VMN_model.py:26
        if torch.is_grad_enabled() and (any(t.requires_grad for t in (x, b, f)) or
                                        any(p.requires_grad for p in self.parameters())):
            # autograd-connected call (the module has no train/eval difference, VMN_model.py:18-68): the training
            # engine's TAM operator -- lets the reference's own VMN_DIM / VMN_Index / VMN_FBA train with the native TAM
            named = named_tensors(self)
            return _TamTrainFn.apply(self, x, b, f, mask, *[named[n] for n in _op_param_names(self)])
        eng = self._op_engine(self.window)
        L = _cabi.lib()
        st = _stream(x.device)
        acts = []
        for t in (x, b, f):
            a = Act.empty(B, H, W, Cc, x.device)
            _cabi.check(L.tcv_nchw_to_split(t.contiguous().float().data_ptr(), B, Cc, H, W, Cc, a.ptr, 0, st), "nchw_to_split")
            acts.append(a)
        w2 = self.window * self.window
        attb = torch.empty((B, w2, H * W), dtype=torch.float32, device=x.device)
        attf = torch.empty_like(attb)
        sm = torch.empty((B, 1, H, W), dtype=torch.uint8, device=x.device)
        m = mask.contiguous().float()
        mh, mw = m.shape[-2:]
        out = eng.tam("", acts[0], acts[1], acts[2], m.data_ptr(), mh * mw, mh, mw, attb.data_ptr(),
                      attf.data_ptr(), sm.data_ptr())
        y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
        _cabi.check(L.tcv_split_to_nchw(out.ptr, B, Cc, H, W, Cc, 0, y.data_ptr(), st), "split_to_nchw")
        return y, attb, attf, sm.bool()


class GuidedCxtAtten(GuidedCxtAttenParams, _OpEngineMixin):
    """Guided contextual attention (GCA/ops.py:83-229) as a drop-in operator: NCHW fp32 in/out.
    Returns (y, (offsets, softmax_scale)); ``offsets`` (an argmax visualisation the VMN path
    drops, VMN_GCA.py:33) is not computed and returned as None."""

    def forward(self, f, alpha, unknown=None, ksize=3, stride=1, fuse_k=3, softmax_scale=1., training=True):
        _require_cuda(f, "f"); _require_cuda(alpha, "alpha")
        if unknown is None:
            raise NotImplementedError("tcvom_b200: GuidedCxtAtten without an unknown map is not on the VMN path")
        B, Cc, H, W = alpha.shape
        assert unknown.shape[2] == H and unknown.shape[3] == W, "mask should have same size as f at dim 2,3"
        if self.training:
            # batch-statistics BatchNorm in W (ops.py:97-101): the training engine's GCA operator, with or without autograd
            named = named_tensors(self)
            if torch.is_grad_enabled():
                y, scales = _GcaTrainFn.apply(self, f, alpha, unknown, *[named[n] for n in _op_param_names(self)])
            else:
                y, scales = _GcaTrainFn.forward(_NoCtx(), self, f, alpha, unknown)
                _train_engine_for(self, 1).tape = []
            return y, (None, scales)
        if torch.is_grad_enabled() and (f.requires_grad or alpha.requires_grad):
            raise NotImplementedError("tcvom_b200: gradients through the eval-mode GCA operator (running-statistics "
                                      "BatchNorm) are not built; call .train() or wrap the call in torch.no_grad()")
        eng = self._op_engine()
        L = _cabi.lib()
        st = _stream(f.device)
        a_f = Act.empty(B, H, W, f.shape[1], f.device)
        a_al = Act.empty(B, H, W, Cc, f.device)
        _cabi.check(L.tcv_nchw_to_split(f.contiguous().float().data_ptr(), B, f.shape[1], H, W, f.shape[1], a_f.ptr, 0, st), "nchw_to_split")
        _cabi.check(L.tcv_nchw_to_split(alpha.contiguous().float().data_ptr(), B, Cc, H, W, Cc, a_al.ptr, 0, st), "nchw_to_split")
        unk = unknown.contiguous().float().reshape(B, H, W)
        out = eng.gca("", a_f, a_al, unk)
        y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=f.device)
        _cabi.check(L.tcv_split_to_nchw(out.ptr, B, Cc, H, W, Cc, 0, y.data_ptr(), st), "split_to_nchw")
        return y, (None, eng.last_gca_scales.clone())


class _GCADecoder(GCADecoderParams):
    def __init__(self, reduction, window, freeze_backbone=False):
        super().__init__(reduction, window, freeze_backbone)

    def train(self, mode=True):
        super().train(mode)
        if self.freeze_backbone:                                  # VMN_GCA.py:18-24
            print('Set GCA decoder feature extraction part in eval() mode.')
            self.layer1.eval(); self.layer2.eval(); self.gca.eval()
        return self


class VMN(nn.Module):
    """Video matting network (VMN_model.py:70-113) with the GCA base net, native forward."""

    def __init__(self, encoder, decoder, freeze_backbone=False):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.freeze_backbone = freeze_backbone
        # per-device engine tables, created HERE so that nn.DataParallel replicas (shallow __dict__ copies made on every
        # forward) always share them with the wrapped module instead of re-recording plans per call
        self.__dict__["_engines"] = {}
        self.__dict__["_train_engines"] = {}

    def train(self, mode=True):
        super().train(mode)
        if self.freeze_backbone:
            print('Set VMN encoder to eval() mode.')
            self.encoder.eval()
        return self

    # -- engine plumbing -------------------------------------------------------------
    def engine(self) -> GcaVmnEngine:
        """Engine for this module on its current device, with weights refreshed."""
        return _engine_for(self, self.decoder.fam.window)

    def __deepcopy__(self, memo):
        # engines hold device pointers; never copy them with the module (DataParallel.replicate,
        # copy.deepcopy): the copy derives its own state lazily.
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy
        for k, v in self.__dict__.items():
            if k in ("_engines", "_train_engines"):
                new.__dict__[k] = {}
                continue
            if k == "_train_param_names":
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _check_mode(self):
        if self.training:
            raise NotImplementedError(
                "tcvom_b200: this entry point is inference-only; the training step goes through "
                "FullModel_VMD / FullModel / VMN.forward in .train() mode.")

    def _forward_train(self, images, masks):
        """Train-mode VMN.forward (VMN_model.py:83-113) on the native training engine, autograd-connected."""
        S = len(images)
        for i in range(S):
            images[i] = images[i].squeeze(1)
        x0 = images[0]
        _require_cuda(x0, "images")
        B, Cin, H, W = x0.shape
        assert Cin == 6, "vmn_gca takes 3 image + 3 trimap channels"
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32")
        L = _cabi.lib()
        st = _stream(x0.device)
        x8 = Act.empty(B * S, H, W, 8, x0.device)
        trimask = torch.empty((B, S, 1, H, W), dtype=torch.float32, device=x0.device)
        for i in range(S):
            xi = images[i].detach().contiguous().float()
            for b in range(B):
                n = b * S + i
                _cabi.check(L.tcv_nchw_to_split(xi[b].data_ptr(), 1, 6, H, W, 8, x8.slice(n, n + 1).ptr, x8.plane, st),
                            "nchw_to_split")
            trimask[:, i] = masks[i].reshape(B, 1, H, W).float()
        named = named_tensors(self)
        names = _trainable_names(self, named)
        self.__dict__["_train_param_names"] = names
        params = [named[n] for n in names]
        if torch.is_grad_enabled():
            pred, attb_t, attf_t, sm = _VMNTrainFn.apply(self, x8, trimask, B, S, H, W, *params)
        else:
            eng = _train_engine_for(self, int(self.decoder.fam.window))
            out = eng.train_forward(x8, trimask, B, S, H, W)
            eng.tape = []
            pred, attb_t, attf_t, sm = out["pred"], out["attb"], out["attf"], out["small_mask"]
        preds = [None] * S; attb = [None] * S; attf = [None] * S; small = [None] * S
        for i in range(1, S - 1):
            preds[i] = pred[:, i - 1]
            attb[i] = attb_t[:, i - 1]
            attf[i] = attf_t[:, i - 1]
            small[i] = sm[:, i - 1].bool()
        preds[0] = torch.zeros_like(preds[1])
        preds[-1] = torch.zeros_like(preds[-2])
        return preds, attb, attf, small

    def forward(self, images: List[torch.Tensor], masks: Sequence[torch.Tensor], extras=None):
        """images: list of S tensors [B,1,6,H,W] (normalised RGB + one-hot trimap); masks: S tensors
        [B,1,1,H,W].  Returns (preds, attb, attf, small_mask) exactly like VMN_model.py:113."""
        if extras is not None:
            raise NotImplementedError("tcvom_b200: `extras` is only used by the FBA base network")
        if self.training:
            return self._forward_train(images, masks)
        S = len(images)
        for i in range(S):
            images[i] = images[i].squeeze(1)                        # VMN_model.py:94 (in-place list update)
        x0 = images[0]
        _require_cuda(x0, "images")
        B, Cin, H, W = x0.shape
        assert Cin == 6, "vmn_gca takes 3 image + 3 trimap channels"
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32 (pred_test.py pads to 32)")
        eng = self.engine()
        L = _cabi.lib()
        st = _stream(x0.device)
        x8 = Act.empty(B * S, H, W, 8, x0.device)
        trimask = torch.empty((B * S, H, W), dtype=torch.float32, device=x0.device)
        for i in range(S):
            xi = images[i].contiguous().float()
            mi = masks[i].reshape(B, H, W).float()
            for b in range(B):
                n = b * S + i
                _cabi.check(L.tcv_nchw_to_split(xi[b].data_ptr(), 1, 6, H, W, 8, x8.slice(n, n + 1).ptr, x8.plane,
                                                st), "nchw_to_split")
                trimask[n].copy_(mi[b])
        out = eng.window_program(x8, trimask, B, S, H, W)
        preds: List[Optional[torch.Tensor]] = [None] * S
        attb: List[Optional[torch.Tensor]] = [None] * S
        attf: List[Optional[torch.Tensor]] = [None] * S
        small: List[Optional[torch.Tensor]] = [None] * S
        for i in range(1, S - 1):
            preds[i] = out["pred"][:, i - 1]
            attb[i] = out["attb"][:, i - 1]
            attf[i] = out["attf"][:, i - 1]
            small[i] = out["small_mask"][:, i - 1].bool()
        preds[0] = torch.zeros_like(preds[1])
        preds[-1] = torch.zeros_like(preds[-2])
        return preds, attb, attf, small


class _FBADecoder(FBADecoderParams):
    def train(self, mode=True):
        super().train(mode)
        if self.freeze_backbone:                                  # VMN_FBA.py:12-16
            print('Set FBA decoder feature extraction part in eval() mode.')
            self.conv_up1.eval()
        return self


class VMN_FBA(VMN):
    """Video matting network (VMN_model.py:70-113) with the FBA base net (models/VMN/__init__.py:18-21), native
    inference forward.  ``extras[i] = [RGB/255 [B,3,H,W], two-channel trimap [B,2,H,W]]`` (models/model.py:404-405)."""

    def engine(self) -> FbaVmnEngine:
        return _engine_for(self, self.decoder.fam.window, FbaVmnEngine)

    def forward(self, images: List[torch.Tensor], masks: Sequence[torch.Tensor], extras=None):
        if self.training:
            raise NotImplementedError("tcvom_b200: vmn_fba is built for inference (BASELINE configs[4]); call .eval()")
        if extras is None:
            raise ValueError("tcvom_b200: vmn_fba needs `extras` (scaled image and two-channel trimap per frame)")
        S = len(images)
        for i in range(S):
            images[i] = images[i].squeeze(1)                        # VMN_model.py:94 (in-place list update)
        x0 = images[0]
        _require_cuda(x0, "images")
        B, Cin, H, W = x0.shape
        assert Cin == 11, "vmn_fba takes 3 image + 6 transformed-trimap + 2 trimap channels"
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32 (pred_test.py pads to 32)")
        eng = self.engine()
        L = _cabi.lib()
        st = _stream(x0.device)
        x16 = Act.empty(B * S, H, W, 16, x0.device)
        x16.buf.zero_()
        trimask = torch.empty((B * S, H, W), dtype=torch.float32, device=x0.device)
        for i in range(S):
            # channels 0..10 = the network input, 11..13 = the decoder's `img` extra (zero stem weights); the
            # two-channel trimap extra equals input channels 9..10 (models/model.py:382-385, 404-405)
            xi = torch.cat([images[i].float(), extras[i][0].float()], dim=1).contiguous()
            mi = masks[i].reshape(B, H, W).float()
            for b in range(B):
                n = b * S + i
                _cabi.check(L.tcv_nchw_to_split(xi[b].data_ptr(), 1, 14, H, W, 16, x16.slice(n, n + 1).ptr, x16.plane,
                                                st), "nchw_to_split")
                trimask[n].copy_(mi[b])
        out = eng.window_program(x16, trimask, B, S, H, W)
        preds: List[Optional[torch.Tensor]] = [None] * S
        attb: List[Optional[torch.Tensor]] = [None] * S
        attf: List[Optional[torch.Tensor]] = [None] * S
        small: List[Optional[torch.Tensor]] = [None] * S
        for i in range(1, S - 1):
            preds[i] = out["pred"][:, i - 1]
            attb[i] = out["attb"][:, i - 1]
            attf[i] = out["attf"][:, i - 1]
            small[i] = out["small_mask"][:, i - 1].bool()
        preds[0] = torch.zeros_like(preds[1])
        preds[-1] = torch.zeros_like(preds[-2])
        return preds, attb, attf, small


class _DIMDecoder(DIMDecoderParams):
    def train(self, mode=True):
        super().train(mode)
        if self.freeze_backbone:                                  # VMN_DIM.py:101-107
            print('Set DIM decoder feature extraction part in eval() mode.')
            self.dconv6.eval(); self.dconv5.eval(); self.dconv4.eval()
        return self


class _FourChannelVMN(VMN):
    """Shared inference forward of the base networks that read normalised RGB + the one-channel trimap (TRIMAP_CHANNEL == 1,
    models/model.py:22-27: 'dim', 'index')."""
    ENGINE = None
    NAME = ""

    def engine(self):
        return _engine_for(self, self.decoder.fam.window, self.ENGINE)

    def forward(self, images: List[torch.Tensor], masks: Sequence[torch.Tensor], extras=None):
        if self.training:
            raise NotImplementedError(f"tcvom_b200: {self.NAME} is built for inference; call .eval()")
        if extras is not None:
            raise NotImplementedError("tcvom_b200: `extras` is only used by the FBA base network")
        S = len(images)
        for i in range(S):
            images[i] = images[i].squeeze(1)                        # VMN_model.py:94 (in-place list update)
        x0 = images[0]
        _require_cuda(x0, "images")
        B, Cin, H, W = x0.shape
        assert Cin == 4, f"{self.NAME} takes 3 image channels + the one-channel trimap"
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32 (pred_test.py pads to 32)")
        eng = self.engine()
        L = _cabi.lib()
        st = _stream(x0.device)
        x8 = Act.empty(B * S, H, W, 8, x0.device)
        trimask = torch.empty((B * S, H, W), dtype=torch.float32, device=x0.device)
        for i in range(S):
            xi = images[i].contiguous().float()
            mi = masks[i].reshape(B, H, W).float()
            for b in range(B):
                n = b * S + i
                _cabi.check(L.tcv_nchw_to_split(xi[b].data_ptr(), 1, 4, H, W, 8, x8.slice(n, n + 1).ptr, x8.plane, st),
                            "nchw_to_split")
                trimask[n].copy_(mi[b])
        out = eng.window_program(x8, trimask, B, S, H, W)
        preds: List[Optional[torch.Tensor]] = [None] * S
        attb: List[Optional[torch.Tensor]] = [None] * S
        attf: List[Optional[torch.Tensor]] = [None] * S
        small: List[Optional[torch.Tensor]] = [None] * S
        for i in range(1, S - 1):
            preds[i] = out["pred"][:, i - 1]
            attb[i] = out["attb"][:, i - 1]
            attf[i] = out["attf"][:, i - 1]
            small[i] = out["small_mask"][:, i - 1].bool()
        preds[0] = torch.zeros_like(preds[1])
        preds[-1] = torch.zeros_like(preds[-2])
        return preds, attb, attf, small


class VMN_Index(_FourChannelVMN):
    """Video matting network (VMN_model.py:70-113) with the IndexNet base net (models/VMN/__init__.py:22-24, VMN_Index.py,
    models/Index/net.py), native inference forward."""
    ENGINE = IndexVmnEngine
    NAME = "vmn_index"


class VMN_DIM(VMN):
    """Video matting network (VMN_model.py:70-113) with the Deep-Image-Matting base net (models/VMN/__init__.py:15-17,
    VMN_DIM.py), native inference forward.  images: S tensors [B,1,4,H,W] (normalised RGB + trimap / 255,
    models/model.py:366-368,392 with TRIMAP_CHANNEL == 1)."""

    def engine(self) -> DimVmnEngine:
        return _engine_for(self, self.decoder.fam.window, DimVmnEngine)

    def forward(self, images: List[torch.Tensor], masks: Sequence[torch.Tensor], extras=None):
        if self.training:
            raise NotImplementedError("tcvom_b200: vmn_dim is built for inference; call .eval()")
        if extras is not None:
            raise NotImplementedError("tcvom_b200: `extras` is only used by the FBA base network")
        S = len(images)
        for i in range(S):
            images[i] = images[i].squeeze(1)                        # VMN_model.py:94 (in-place list update)
        x0 = images[0]
        _require_cuda(x0, "images")
        B, Cin, H, W = x0.shape
        assert Cin == 4, "vmn_dim takes 3 image channels + the one-channel trimap"
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32 (pred_test.py pads to 32)")
        eng = self.engine()
        L = _cabi.lib()
        st = _stream(x0.device)
        x8 = Act.empty(B * S, H, W, 8, x0.device)
        trimask = torch.empty((B * S, H, W), dtype=torch.float32, device=x0.device)
        for i in range(S):
            xi = images[i].contiguous().float()
            mi = masks[i].reshape(B, H, W).float()
            for b in range(B):
                n = b * S + i
                _cabi.check(L.tcv_nchw_to_split(xi[b].data_ptr(), 1, 4, H, W, 8, x8.slice(n, n + 1).ptr, x8.plane, st),
                            "nchw_to_split")
                trimask[n].copy_(mi[b])
        out = eng.window_program(x8, trimask, B, S, H, W)
        preds: List[Optional[torch.Tensor]] = [None] * S
        attb: List[Optional[torch.Tensor]] = [None] * S
        attf: List[Optional[torch.Tensor]] = [None] * S
        small: List[Optional[torch.Tensor]] = [None] * S
        for i in range(1, S - 1):
            preds[i] = out["pred"][:, i - 1]
            attb[i] = out["attb"][:, i - 1]
            attf[i] = out["attf"][:, i - 1]
            small[i] = out["small_mask"][:, i - 1].bool()
        preds[0] = torch.zeros_like(preds[1])
        preds[-1] = torch.zeros_like(preds[-2])
        return preds, attb, attf, small


def _trimap_transform_impl(trimap: torch.Tensor, stream_ptr: int) -> torch.Tensor:
    B, S, C2, H, W = trimap.shape
    assert C2 == 2, "trimap: tensor [B, S, 2, H, W] (background, foreground one-hot)"
    L = _cabi.lib()
    F_ = B * S
    dev = trimap.device
    # the distance-transform kernels read their seeds from channels 9 / 10 of the 16-channel FBA input tensor and
    # write the six features into channels 3..8 (include/tcvom_b200.h, tcv_fba_edt_*)
    x = torch.zeros((F_, 16, H, W), dtype=torch.float32, device=dev)
    # seeds = pixels where cv2 sees a zero: ((1 - tk) * 255).astype(uint8) == 0   (utils/utils.py:21,32)
    x[:, 9:11] = (((1.0 - trimap.reshape(F_, 2, H, W)) * 255).to(torch.uint8) == 0).float()
    x16 = Act.empty(F_, H, W, 16, dev)
    g = torch.empty((F_, 2, H, W), dtype=torch.int32, device=dev)
    out = torch.empty((F_, 16, H, W), dtype=torch.float32, device=dev)
    _cabi.check(L.tcv_nchw_to_split(x.data_ptr(), F_, 16, H, W, 16, x16.ptr, 0, stream_ptr), "nchw_to_split")
    _cabi.check(L.tcv_fba_edt_cols(x16.ptr, F_, H, W, g.data_ptr(), stream_ptr), "fba_edt_cols")
    _cabi.check(L.tcv_fba_edt_rows(g.data_ptr(), F_, H, W, x16.ptr, stream_ptr), "fba_edt_rows")
    _cabi.check(L.tcv_split_to_nchw(x16.ptr, F_, 16, H, W, 16, 0, out.data_ptr(), stream_ptr), "split_to_nchw")
    return out[:, 3:9].reshape(B, S, 6, H, W).contiguous()


def trimap_transform(trimap: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``utils.utils.trimap_transform`` (utils/utils.py:25-39): trimap [B,S,2,H,W] (one-hot background /
    foreground) -> clicks [B,S,6,H,W] = exp(-d^2 / (2 (f*320)^2)), f in (0.02, 0.08, 0.16), d = Euclidean distance to the
    nearest background resp. foreground pixel.  The reference moves every frame to the host for ``cv2.distanceTransform``;
    here the exact transform runs on the GPU (``tcv_fba_edt_cols`` / ``tcv_fba_edt_rows``).  A channel without any seed
    gives zeros (like the reference: cv2 returns +huge there, and :31 skips a kind that is absent from the whole batch);
    values below 1e-12 are written as exact zeros."""
    _require_cuda(trimap, "trimap")
    return _trimap_transform_impl(trimap.float(), _stream(trimap.device))


def get_VMN_models(arch, agg_window, agg_reduction=1, freeze_backbone=False, **kwargs):
    """Plugin seam of the reference (models/VMN/__init__.py:11-29)."""
    if arch not in ('vmn_gca', 'vmn_fba', 'vmn_dim', 'vmn_index'):
        raise ValueError
    if agg_reduction != 1:
        raise NotImplementedError("tcvom_b200: agg_reduction != 1 is not supported")
    if arch == 'vmn_fba':
        e = FBAEncoderParams()
        d = _FBADecoder(agg_reduction, int(agg_window), freeze_backbone=freeze_backbone)
        d.fam = FeatureAggregationModule(256, agg_reduction, int(agg_window))
        return VMN_FBA(encoder=e, decoder=d, freeze_backbone=freeze_backbone)
    if arch == 'vmn_index':
        e = IndexEncoderParams()
        d = IndexDecoderParams(agg_reduction, int(agg_window), freeze_backbone=freeze_backbone)
        d.fam = FeatureAggregationModule(32, agg_reduction, int(agg_window))
        return VMN_Index(encoder=e, decoder=d, freeze_backbone=freeze_backbone)
    if arch == 'vmn_dim':
        e = DIMEncoderParams(4)
        d = _DIMDecoder(agg_reduction, int(agg_window), freeze_backbone=freeze_backbone)
        d.fam = FeatureAggregationModule(256, agg_reduction, int(agg_window))
        return VMN_DIM(encoder=e, decoder=d, freeze_backbone=freeze_backbone)
    e = GCAEncoderParams()
    d = _GCADecoder(agg_reduction, int(agg_window), freeze_backbone=freeze_backbone)
    d.fam = FeatureAggregationModule(128, agg_reduction, int(agg_window))
    d.gca = GuidedCxtAtten(128, 128)
    e.gca = GuidedCxtAtten(128, 128)
    return VMN(encoder=e, decoder=d, freeze_backbone=freeze_backbone)


class EvalModel(nn.Module):
    """Inference wrapper (models/model.py:359-424): raw BGR frames + trimaps -> alpha mattes.

    forward(imgs [B,S,3,H,W] BGR 0..255, tris [B,S,1,H,W] in {0,128,255}) -> alphas [B,S,1,H,W]
    (first and last frame of each sample are zeros, model.py:419-421)."""

    def __init__(self, model, dilate_kernel=None, eps=0, **kwargs):
        super().__init__()
        if not model.startswith('vmn'):
            raise NotImplementedError("tcvom_b200: only the VMN (video) architectures are on the built path")
        self.DILATION_KERNEL = dilate_kernel
        self.EPS = eps
        self.IMG_SCALE = 1. / 255
        self.register_buffer('IMG_MEAN', torch.tensor([0.485, 0.456, 0.406]).reshape([1, 1, 3, 1, 1]).float())
        self.register_buffer('IMG_STD', torch.tensor([0.229, 0.224, 0.225]).reshape([1, 1, 3, 1, 1]).float())
        self.model_name = model
        self.NET = get_VMN_models(arch=model, **kwargs)
        self.window = kwargs['agg_window']
        self.method = model[model.rfind('_') + 1:]
        self.TRIMAP_CHANNEL = {'fba': 8, 'dim': 1, 'index': 1}.get(self.method, 3)      # models/model.py:22-27

    # -- plan handling ---------------------------------------------------------------
    def _plan_fba(self, B, S, H, W, u8=False) -> Plan:
        """EvalModel.forward for method 'fba' (models/model.py:389-446) as one recorded plan: trimask (+ dilation),
        input encoding incl. the distance transforms, the VMN program, the where()-tail."""
        eng = self.NET.engine()
        dil = -1 if self.DILATION_KERNEL is None else int(self.DILATION_KERNEL)
        self.__dict__["_eng"] = eng
        key = ("eval_fba", B, S, H, W, dil, u8)
        plan = eng.get_plan(key)
        if plan is not None:
            return plan
        plan = Plan()
        with eng.recording(plan):
            n0 = _cabi.launch_count()
            plan.io = eng.eval_program(B, S, H, W, dil, u8)
            plan.n_launch = _cabi.launch_count() - n0
        eng.put_plan(key, plan)
        return plan

    def _plan(self, B, S, H, W, dev, u8=False) -> Plan:
        eng = self.NET.engine()
        dil = -1 if self.DILATION_KERNEL is None else int(self.DILATION_KERNEL)
        self.__dict__["_eng"] = eng
        key = ("eval", B, S, H, W, dil, u8, self.method)
        plan = eng.get_plan(key)
        if plan is not None:
            return plan
        plan = Plan()
        with eng.recording(plan):
            in_dt = torch.uint8 if u8 else torch.float32
            sfx = "_u8" if u8 else ""
            imgs = eng._empty((B, S, 3, H, W), in_dt)
            tris = eng._empty((B, S, 1, H, W), in_dt)
            x8 = eng._act(B * S, H, W, 8)
            trimask = eng._empty((B * S, H, W))
            tmp = eng._empty((2 * B * S * H * W,), torch.uint8)
            alphas = eng._empty((B, S, 1, H, W))
            # inputs must hold valid data while recording runs the kernels once
            imgs.zero_(); tris.zero_()
            n0 = _cabi.launch_count()
            eng._call("tcv_preprocess_eval" + sfx, imgs.data_ptr(), tris.data_ptr(), B * S, H, W, dil, x8.ptr,
                      trimask.data_ptr(), tmp.data_ptr())
            if self.method in ('dim', 'index'):
                # TRIMAP_CHANNEL == 1 (models/model.py:368,392): the network reads the raw trimap / 255 as its 4th channel
                eng._call("tcv_dim_fix_inputs", tris.data_ptr(), 1 if u8 else 0, B * S, H, W, x8.ptr)
            out = eng.window_program(x8, trimask, B, S, H, W)
            eng._call("tcv_postprocess_eval" + sfx, out["pred"].data_ptr(), tris.data_ptr(), trimask.data_ptr(), B, S, H, W,
                      alphas.data_ptr())
            plan.n_launch = _cabi.launch_count() - n0
            plan.io = dict(imgs=imgs, tris=tris, alphas=alphas, trimask=trimask, **{k: out[k] for k in
                                                                                   ("pred", "attb", "attf", "small_mask")})
            plan.io["feat"] = out["feat"].buf
            if self.method == 'dim':
                plan.io["pool_idx"] = out["pf"]["idxs"]          # max-pooling routing (tests follow it, see oracle/vmn_dim_oracle.py)
                plan.io["x8"] = x8.buf
        eng.put_plan(key, plan)
        return plan

    def run_plan(self, plan: Plan) -> None:
        """Replays the recorded kernel sequence on the current stream (CUDA graph when enabled)."""
        eng = self._eng
        dev = eng.device
        if torch.cuda.current_device() != dev.index:       # model moved with .to('cuda:N') without set_device(N)
            with torch.cuda.device(dev):
                return self.run_plan(plan)
        if eng.use_graphs:
            if plan.graph is None:
                st = torch.cuda.Stream(dev)
                st.wait_stream(torch.cuda.current_stream(dev))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(st):
                    plan.replay(st.cuda_stream)                     # warm-up outside capture
                    st.synchronize()
                    # thread_local: nn.DataParallel replicas capture concurrently from one thread per device
                    with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                        plan.replay(torch.cuda.current_stream(dev).cuda_stream)
                torch.cuda.current_stream(dev).wait_stream(st)
                plan.graph = g
            plan.graph.replay()
        else:
            plan.replay(torch.cuda.current_stream(dev).cuda_stream)

    def forward(self, imgs, tris):
        _require_cuda(imgs, "imgs")
        self.NET._check_mode()
        B, S, Cc, H, W = imgs.shape
        assert Cc == 3 and S >= 3
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32 (pred_test.py pads to 32)")
        # uint8 frames/trimaps are consumed as such (the reference casts with .float(), model.py:366-368)
        u8 = imgs.dtype == torch.uint8 and tris.dtype == torch.uint8
        if self.method == 'fba':
            plan = self._plan_fba(B, S, H, W, u8)
            plan.io["imgs"].copy_(imgs, non_blocking=True)
            plan.io["tris"].copy_(tris, non_blocking=True)
            self.run_plan(plan)
            return plan.io["alphas"].clone(), plan.io["Fs"].clone(), plan.io["Bs"].clone()
        plan = self._plan(B, S, H, W, imgs.device, u8)
        plan.io["imgs"].copy_(imgs, non_blocking=True)
        plan.io["tris"].copy_(tris, non_blocking=True)
        self.run_plan(plan)
        return plan.io["alphas"].clone()


class FullModel_VMD(nn.Module):
    """Task wrapper with losses (models/model.py:248-357, and :15-246 for the shared parts).

    forward(a [B,S,1,H,W] 0..255, fg, bg [B,S,3,H,W] BGR 0..255) -> the reference's 12-list
    [L_alpha, L_comp, L_grad, L_dt, L_att, scaled_imgs, tris_vis, alphas, comps, scaled_gts, Fs, Bs].

    Eval mode (pred_vmn.py:107-116: ``net.eval()`` + ``no_grad``): one recorded plan, loss values from fused forward
    kernels.  Train mode (train_ddp.py:52-65): the native training step behind ONE autograd node (`_TrainStepFn`), for
    ``vmn_gca`` with or without ``freeze_backbone``; other base networks raise instead of falling back to PyTorch."""
    TAM_OS = 8
    FBA_L_ATT_MULTIPLIER = 1
    ARCH_DICT = {'gca': None, 'dim': None, 'fba': None, 'index': None}     # pred_vmn.py:29 lists the keys
    TRIMAP_CHANNEL_DICT = {'gca': 3, 'dim': 1, 'index': 1, 'fba': 8}
    _with_att = True

    def __init__(self, model, att_thres=0.3, label_smooth=0.2, dilate_kernel=None, eps=0, **kwargs):
        super().__init__()
        assert model.startswith('vmn'), "FullModel_VMD only support VMN arch"
        if model != 'vmn_gca':
            raise NotImplementedError("tcvom_b200: the loss wrappers are built for vmn_gca; vmn_fba is built for "
                                      "inference through EvalModel / the VMN seam (BASELINE configs[4])")
        self.att_thres = att_thres
        self.label_smooth = label_smooth
        self.DILATION_KERNEL = dilate_kernel
        self.EPS = eps
        self.IMG_SCALE = 1. / 255
        self.register_buffer('IMG_MEAN', torch.tensor([0.485, 0.456, 0.406]).reshape([1, 1, 3, 1, 1]).float())
        self.register_buffer('IMG_STD', torch.tensor([0.229, 0.224, 0.225]).reshape([1, 1, 3, 1, 1]).float())
        self.model_name = model
        self.NET = get_VMN_models(arch=model, **kwargs)
        self.window = kwargs['agg_window']
        self.method = model[model.rfind('_') + 1:]
        self.TRIMAP_CHANNEL = self.TRIMAP_CHANNEL_DICT[self.method]

    def _plan(self, B, S, H, W) -> Plan:
        eng = self.NET.engine()
        self.__dict__["_eng"] = eng
        key = ("vmd" if self._with_att else "full", B, S, H, W, float(self.EPS), float(self.att_thres),
               float(self.label_smooth), int(self.window))      # every value the recorded tcv_losses_vmd call bakes in
        plan = eng.get_plan(key)
        if plan is not None:
            return plan
        plan = Plan()
        with eng.recording(plan):
            a = eng._empty((B, S, 1, H, W)); fg = eng._empty((B, S, 3, H, W)); bg = eng._empty((B, S, 3, H, W))
            radii = eng._empty((B,), torch.int32)
            a.zero_(); fg.zero_(); bg.zero_(); radii.zero_()
            x8 = eng._act(B * S, H, W, 8)
            trimask = eng._empty((B, S, 1, H, W)); gts = eng._empty((B, S, 1, H, W)); tris_vis = eng._empty((B, S, 1, H, W))
            fgs = eng._empty((B, S, 3, H, W)); bgs = eng._empty((B, S, 3, H, W)); imgs = eng._empty((B, S, 3, H, W))
            tmp = eng._empty((2 * B * S * H * W,), torch.uint8)
            alphas = eng._empty((B, S, 1, H, W)); comps = eng._empty((B, S, 3, H, W))
            gt8 = eng._empty((B, S, H // 8, W // 8)); acc = eng._empty((6 * S,), torch.float64)
            losses = eng._empty((5,))
            n0 = _cabi.launch_count()
            eng._call("tcv_preprocess_train", a.data_ptr(), fg.data_ptr(), bg.data_ptr(), B, S, H, W, float(self.EPS),
                      radii.data_ptr(), x8.ptr, trimask.data_ptr(), gts.data_ptr(), fgs.data_ptr(), bgs.data_ptr(),
                      imgs.data_ptr(), tris_vis.data_ptr(), tmp.data_ptr())
            out = eng.window_program(x8, trimask.reshape(B * S, H, W), B, S, H, W)
            att = self._with_att
            eng._call("tcv_losses_vmd", out["pred"].data_ptr(), trimask.data_ptr(), gts.data_ptr(), fgs.data_ptr(),
                      bgs.data_ptr(), out["attb"].data_ptr() if att else None, out["attf"].data_ptr() if att else None,
                      out["small_mask"].data_ptr() if att else None, B, S, H, W, int(self.window),
                      float(self.att_thres), float(self.label_smooth), float(self.FBA_L_ATT_MULTIPLIER if self.method == 'fba' else 1),
                      alphas.data_ptr(), comps.data_ptr(), gt8.data_ptr(), acc.data_ptr(), losses.data_ptr())
            plan.n_launch = _cabi.launch_count() - n0
            plan.io = dict(a=a, fg=fg, bg=bg, radii=radii, losses=losses, imgs=imgs, tris_vis=tris_vis, alphas=alphas,
                           comps=comps, gts=gts, fgs=fgs, bgs=bgs, trimask=trimask,
                           **{k: out[k] for k in ("pred", "attb", "attf", "small_mask")})
        eng.put_plan(key, plan)
        return plan

    run_plan = EvalModel.run_plan

    # -- training step (train_ddp.py:52-65) --------------------------------------------
    def _train_forward(self, a, fg, bg) -> dict:
        """Train-mode forward on the native kernels: preprocess, VMN (batch-statistics BatchNorm, SpectralNorm
        power iteration), losses.  Returns the tensors the backward needs plus the visualisation outputs."""
        if self.method != 'gca':
            raise NotImplementedError("tcvom_b200: training is built for vmn_gca only")
        B, S = a.shape[:2]
        H, W = a.shape[-2:]
        dev = a.device
        eng = _train_engine_for(self.NET, int(self.window))
        f32 = torch.float32
        rad = [int(torch.randint(0, 26, size=())) if self.DILATION_KERNEL is None else int(self.DILATION_KERNEL)
               for _ in range(B)]
        radii = torch.tensor(rad, dtype=torch.int32).to(dev)
        a = a.contiguous().float(); fg = fg.contiguous().float(); bg = bg.contiguous().float()
        x8 = Act.empty(B * S, H, W, 8, dev)
        mk = lambda c: torch.empty((B, S, c, H, W), dtype=f32, device=dev)
        trimask, gts, tris_vis = mk(1), mk(1), mk(1)
        fgs, bgs, imgs = mk(3), mk(3), mk(3)
        tmp = torch.empty((2 * B * S * H * W,), dtype=torch.uint8, device=dev)
        eng._call("tcv_preprocess_train", a.data_ptr(), fg.data_ptr(), bg.data_ptr(), B, S, H, W, float(self.EPS),
                  radii.data_ptr(), x8.ptr, trimask.data_ptr(), gts.data_ptr(), fgs.data_ptr(), bgs.data_ptr(),
                  imgs.data_ptr(), tris_vis.data_ptr(), tmp.data_ptr())
        out = eng.train_forward(x8, trimask, B, S, H, W)
        alphas, comps = mk(1), mk(3)
        gt8 = torch.empty((B, S, H // 8, W // 8), dtype=f32, device=dev)
        acc = torch.empty((6 * S,), dtype=torch.float64, device=dev)
        losses = torch.empty((5,), dtype=f32, device=dev)
        att = self._with_att
        mult = float(self.FBA_L_ATT_MULTIPLIER if self.method == 'fba' else 1)
        eng._call("tcv_losses_vmd", out["pred"].data_ptr(), trimask.data_ptr(), gts.data_ptr(), fgs.data_ptr(),
                  bgs.data_ptr(), out["attb"].data_ptr() if att else None, out["attf"].data_ptr() if att else None,
                  out["small_mask"].data_ptr() if att else None, B, S, H, W, int(self.window), float(self.att_thres),
                  float(self.label_smooth), mult, alphas.data_ptr(), comps.data_ptr(), gt8.data_ptr(), acc.data_ptr(),
                  losses.data_ptr())
        return dict(eng=eng, B=B, S=S, H=H, W=W, losses=losses, imgs=imgs, tris_vis=tris_vis, alphas=alphas,
                    comps=comps, gts=gts, fgs=fgs, bgs=bgs, trimask=trimask, gt8=gt8, acc=acc, mult=mult, **out)

    def _train_backward(self, st: dict, gl: torch.Tensor):
        eng: TrainEngine = st["eng"]
        B, S, H, W = st["B"], st["S"], st["H"], st["W"]
        att = self._with_att
        gl = gl.contiguous().float()
        dpred = torch.empty_like(st["pred"])
        dattb = torch.empty_like(st["attb"]) if att else None
        dattf = torch.empty_like(st["attf"]) if att else None
        eng._call("tcv_losses_vmd_bwd", st["pred"].data_ptr(), st["trimask"].data_ptr(), st["gts"].data_ptr(),
                  st["attb"].data_ptr() if att else None, st["attf"].data_ptr() if att else None,
                  st["small_mask"].data_ptr() if att else None, st["gt8"].data_ptr(), st["acc"].data_ptr(),
                  gl.data_ptr(), B, S, H, W, int(self.window), float(self.att_thres), float(self.label_smooth),
                  st["mult"], dpred.data_ptr(), dattb.data_ptr() if att else None, dattf.data_ptr() if att else None)
        eng.train_backward(dpred, dattb, dattf)
        return eng.collect_grads(self._train_param_names)

    def _train_step(self, a, fg, bg):
        named = named_tensors(self.NET)
        names = _trainable_names(self.NET, named)
        self.__dict__["_train_param_names"] = names
        params = [named[n] for n in names]
        if torch.is_grad_enabled() and not names:
            # e.g. an nn.DataParallel replica (its parameters are plain tensors): gradients could not flow back
            raise RuntimeError("tcvom_b200: no trainable nn.Parameter reachable from this module -- training under "
                               "nn.DataParallel is not supported, use DistributedDataParallel (train_ddp.py:275-280)")
        if torch.is_grad_enabled():
            res = _TrainStepFn.apply(self, a, fg, bg, *params)
            L, vis = res[0], list(res[1:])
        else:
            st = self._train_forward(a, fg, bg)
            st["eng"].tape = []
            L = st["losses"]
            vis = [st[k] for k in ("imgs", "tris_vis", "alphas", "comps", "gts", "fgs", "bgs")]
        outs = [L[0], L[1], L[2]]
        if self._with_att:
            outs += [L[3], L[4]]
        return outs + vis

    def forward(self, a, fg, bg, wb=None, wf=None):
        _require_cuda(a, "a")
        if self.NET.training:
            if a.shape[-2] % 32 or a.shape[-1] % 32:
                raise ValueError("tcvom_b200: H and W must be multiples of 32")
            assert a.shape[1] >= 3
            return self._train_step(a, fg, bg)
        B, S = a.shape[:2]
        H, W = a.shape[-2:]
        assert S >= 3
        if H % 32 or W % 32:
            raise ValueError("tcvom_b200: H and W must be multiples of 32")
        plan = self._plan(B, S, H, W)
        # trimap width: the reference draws one radius per sample on the host (model.py:62)
        rad = [int(torch.randint(0, 26, size=())) if self.DILATION_KERNEL is None else int(self.DILATION_KERNEL)
               for _ in range(B)]
        plan.io["radii"].copy_(torch.tensor(rad, dtype=torch.int32), non_blocking=True)
        plan.io["a"].copy_(a, non_blocking=True)
        plan.io["fg"].copy_(fg, non_blocking=True)
        plan.io["bg"].copy_(bg, non_blocking=True)
        self.run_plan(plan)
        io = plan.io
        L = io["losses"].clone()
        outs = [L[0], L[1], L[2]]
        if self._with_att:
            outs += [L[3], L[4]]
        return outs + [io["imgs"].clone(), io["tris_vis"].clone(), io["alphas"].clone(), io["comps"].clone(),
                       io["gts"].clone(), io["fgs"].clone(), io["bgs"].clone()]


class FullModel(FullModel_VMD):
    """models/model.py:15-246 for the VMN archs: [L_alpha, L_comp, L_grad, imgs, tris_vis, alphas, comps, gts, Fs, Bs]."""
    _with_att = False

    def __init__(self, model, dilate_kernel=None, eps=0, **kwargs):
        if not model.startswith('vmn'):
            raise NotImplementedError("tcvom_b200: only the VMN (video) architectures are on the built path")
        super().__init__(model, dilate_kernel=dilate_kernel, eps=eps, **kwargs)
