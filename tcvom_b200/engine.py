"""Host-side engine for the ``vmn_gca`` frame-window forward: weight folding/packing cache,
static per-shape execution plans, and the kernel program (what runs, in which order, on which
buffers).  All arithmetic is done by the sm_100a kernels behind the C ABI; torch is used for
device memory, streams and (optionally) CUDA-graph capture only.

Program restated from (reference checkout):
  per-frame part   VMN_model.py:93-98 -> res_gca_enc.py:57-90, VMN_GCA.py:27-34
  per-centre part  VMN_model.py:107-110 -> VMN_GCA.py:35-49
"""
from __future__ import annotations

import collections
import contextlib
import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi
from ._cabi import ACT_LEAKY02, ACT_NONE, ACT_RELU, ACT_TANH01, PAD_REFLECT, PAD_ZERO, ConvDesc
from .modules import DEC_LAYERS, ENC_LAYERS

BN_EPS = 1e-5


class Act:
    """split-bf16 NHWC activation: ``buf[0]`` = hi plane, ``buf[1]`` = lo plane."""
    __slots__ = ("buf", "n", "h", "w", "c", "ptr", "plane")

    def __init__(self, buf, n, h, w, c, ptr, plane):
        self.buf, self.n, self.h, self.w, self.c, self.ptr, self.plane = buf, n, h, w, c, ptr, plane

    @staticmethod
    def empty(n, h, w, c, device) -> "Act":
        buf = torch.empty((2, n, h, w, c), dtype=torch.bfloat16, device=device)
        return Act(buf, n, h, w, c, buf.data_ptr(), n * h * w * c)

    def slice(self, n0: int, n1: int) -> "Act":
        return Act(self.buf, n1 - n0, self.h, self.w, self.c, self.ptr + 2 * n0 * self.h * self.w * self.c,
                   self.plane)

    @property
    def img_elems(self) -> int:
        return self.h * self.w * self.c

    def float(self) -> torch.Tensor:
        """[n,h,w,c] fp32 view of the full parent tensor (debug / tests)."""
        return self.buf[0].float() + self.buf[1].float()


class Plan:
    """A recorded, replayable list of C-ABI calls on static buffers for one input shape."""

    def __init__(self):
        self.calls: List[Tuple] = []
        self.meta: List[dict] = []
        self.keep: List = []
        self.io: Dict[str, torch.Tensor] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.n_launch = 0
        # private memory pool the plan was recorded in (GcaVmnEngine.recording): dead intermediates are reused by later
        # allocations of the same recording, and the pool keeps the address ranges reserved for the replays
        self.pool = None
        self.pool_device = None

    def __del__(self):
        # The pool goes back to a free list instead of being destroyed here: a MemPool destructor empties its cache, which
        # the caching allocator refuses (abort) while ANOTHER recording is routing allocations to a pool -- and the garbage
        # collector may run this finaliser at exactly such a moment.  The next recording on the device reuses the pool (and
        # its cached blocks); release_idle_pools() frees them.
        pool = self.__dict__.get("pool")
        if pool is not None:
            _IDLE_POOLS.setdefault(self.pool_device, []).append(pool)
            self.pool = None

    def replay(self, stream_ptr: int) -> None:
        for fn, args, what in self.calls:
            rc = fn(*args, stream_ptr)
            if rc != 0:
                _cabi.check(rc, what)

    def replay_timed(self, device) -> List[float]:
        """Eager replay on the current stream with a CUDA-event pair around every call; returns
        milliseconds per call (same order as ``calls`` / ``meta``).  Measurement helper for bench.py."""
        st = torch.cuda.current_stream(device)
        evs = []
        for fn, args, what in self.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = fn(*args, st.cuda_stream)
            e1.record(st)
            if rc != 0:
                _cabi.check(rc, what)
            evs.append((e0, e1))
        st.synchronize()
        return [a.elapsed_time(b) for a, b in evs]


_IDLE_POOLS: Dict[Optional[int], list] = {}
_RECORDING = [0]


def release_idle_pools() -> int:
    """Destroys the memory pools of plans that no longer exist (their device memory returns to the driver).  Returns the
    number of pools released; a no-op while a plan is being recorded."""
    if _RECORDING[0]:
        return 0
    n = 0
    for dev, pools in list(_IDLE_POOLS.items()):
        n += len(pools)
        pools.clear()
    if n and torch.cuda.is_available():
        torch.cuda.empty_cache()
    return n


def named_tensors(net: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """state_dict-style name -> tensor for parameters and buffers.  Also works on nn.DataParallel replicas,
    whose parameters are plain attributes recorded in ``_former_parameters`` (torch/nn/parallel/replicate.py)
    and therefore invisible to ``named_parameters()``."""
    out: Dict[str, torch.Tensor] = {}
    for mname, m in net.named_modules():
        pre = mname + "." if mname else ""
        for k, v in m._parameters.items():
            if v is not None:
                out[pre + k] = v
        for k, v in getattr(m, "_former_parameters", {}).items():
            out[pre + k] = v
        for k, v in m._buffers.items():
            if v is not None:
                out[pre + k] = v
    return out


def _k(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


class GcaVmnEngine:
    """Owns derived device state for one ``VMN`` module on one device."""

    def __init__(self, window: int):
        self.net = None
        self.window = window
        self.device: Optional[torch.device] = None
        self.w: Dict[str, dict] = {}
        self.aff: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.bias: Dict[str, torch.Tensor] = {}
        self._fingerprint = None
        self._tensors = None
        self.plans: "collections.OrderedDict[Tuple, Plan]" = collections.OrderedDict()
        # recorded plans own all their activation buffers (10.5 GB for a 1080p window): keep only the most
        # recently used shapes so that a folder of differently sized clips cannot exhaust HBM
        self.max_plans = int(os.environ.get("TCV_MAX_PLANS", "2"))
        self._rec: Optional[Plan] = None
        self._stream_cache: Optional[int] = None
        self.use_graphs = os.environ.get("TCV_GRAPHS", "1") == "1"
        self.plan_pool = os.environ.get("TCV_PLAN_POOL", "1") == "1"
        # tcgen05 paths (default on); the CUDA-core fp32 paths stay available as the exact cross-check
        self.use_tc_conv = os.environ.get("TCV_TC_CONV", "1") == "1"
        self.use_tc_attn = os.environ.get("TCV_TC_ATTN", "1") == "1"
        if "TCV_CONV_TC_VERSION" in os.environ:
            _cabi.lib().tcv_set_conv_tc_version(int(os.environ["TCV_CONV_TC_VERSION"]))
        # operand format of the P.V aggregation GEMM: "bf16x3" (3 MMAs/step; measured 3.7e-4 max-abs alpha
        # error at 384x512), "fp16" (1 MMA; 1.3e-3, over the 1e-3 bar), "bf16" (1 MMA; 8.8e-3)
        self.pv_mode = os.environ.get("TCV_PV_MODE", "bf16x3")
        # bf16 planes per operand of the scores GEMM Q.Kn^T: 2 (three MMAs per K step, shipped) or 3 (six MMAs,
        # fp32-grade logits; measured on B200: no parity gain on any fixture, +0.66 ms per 1080p window)
        self.score_planes = int(os.environ.get("TCV_SCORE_PLANES", "2"))
        # aggregation GEMM in its shift-sum form (default; TCV_GCA_SHIFT_SUM=0: the unfold-values + overlap-add form,
        # 3.8x the FLOPs, kept as the cross-check)
        # windowed forward: shortcut branches 0..2 only for the centre frames that consume them (see per_frame)
        self.centre_shortcuts = os.environ.get("TCV_CENTRE_SHORTCUTS", "1") == "1"
        # softmax folded into the shift-add kernel (one pass over S, 4 exponentials per output) instead of normalise-in-place
        # + shift-add: measured slower on B200 (0.77 + 0.25 vs 0.34 + 0.50 ms per launch), kept as the cross-check
        self.gca_softmax_in_consumer = os.environ.get("TCV_GCA_SOFTMAX_IN_CONSUMER", "0") == "1"
        self.gca_shift_sum = os.environ.get("TCV_GCA_SHIFT_SUM", "1") == "1" and self.pv_mode == "bf16x3"
        # the stride-2 layers with 8 / 16 input channels (encoder.conv1, guidance_head.1 / .5) either on the CUDA-core kernel
        # or as stride-1 2x2-tap convolutions over the 2x2 space-to-depth image on the tcgen05 kernels (same rewrite as the
        # FBA stem, tcv_s2d_pack): TCV_S2D_STRIDE2 = "conv1" (default), "1" (all three), "0" (none)
        # Measured on B200 (round 2, 1088x1920 x 3): encoder.conv1 309 us on the CUDA cores -> 71 (space-to-depth copy) + 107 us;
        # guidance_head.1 / .5 170 / 163 us -> 240 / 256 us (reflect border + copy + conv).  Default: conv1 only.
        mode = os.environ.get("TCV_S2D_STRIDE2", "conv1")
        self.s2d_stride2 = mode in ("1", "conv1")
        self.s2d_guidance = mode == "1"

    # ------------------------------------------------------------------ weights
    def _named(self) -> Dict[str, torch.Tensor]:
        return named_tensors(self.net)

    def _current_fingerprint(self):
        if self._tensors is None:
            self._tensors = list(self._named().values())
        return (tuple(t._version for t in self._tensors), tuple(t.data_ptr() for t in self._tensors))

    def refresh_weights(self, net: torch.nn.Module, force=False) -> None:
        """(Re)derives folded/packed weights when parameters changed (load_state_dict, optimizer step,
        device move, a fresh DataParallel replica).  Packed buffers are updated in place so
        recorded plans stay valid."""
        if net is not self.net:
            self.net = net
            self._tensors = None
        named = self._named()
        dev = next(iter(named.values())).device
        self._check_device(dev)
        if dev.type == "cuda" and torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.refresh_weights(net, force)
        if self.device is not None and dev != self.device:
            self.w.clear(); self.aff.clear(); self.bias.clear(); self.plans.clear()
            self._tensors = None
        self.device = dev
        fp = self._current_fingerprint()
        if not force and fp == self._fingerprint:
            return
        L = _cabi.lib()
        st = self._stream_ptr()
        sig = getattr(self, "_sigma_ws", None)
        if sig is None or sig.device != dev:
            sig = self._sigma_ws = torch.empty(256, dtype=torch.float32, device=dev)
        i_sig = 0
        for name, t in named.items():
            if t.dtype.is_floating_point and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError(f"tcvom_b200: parameter {name} must be contiguous fp32")
        for name, t in named.items():
            if name.endswith(".module.weight_bar"):
                p = name[: -len(".module.weight_bar")]
                u, v = named[p + ".module.weight_u"], named[p + ".module.weight_v"]
                self._pack(L, st, p, t, u, v, sig[i_sig:i_sig + 1], transposed=(t.shape[2] == 4))
                i_sig += 1
            elif name.endswith(".weight") and t.dim() == 4:
                p = name[: -len(".weight")]
                self._pack(L, st, p, t, None, None, None, transposed=False)
                b = named.get(p + ".bias")
                if b is not None:
                    self._own_bias(p, b)
                if t.shape[0] == 1 and t.shape[1] % 32 == 0 and t.shape[2] == 3:
                    self._pack_head32(L, st, p, t, b)
            elif name.endswith(".running_var"):
                p = name[: -len(".running_var")]
                c = t.numel()
                if p not in self.aff:
                    self.aff[p] = (torch.empty(c, dtype=torch.float32, device=dev),
                                   torch.empty(c, dtype=torch.float32, device=dev))
                s, b = self.aff[p]
                _cabi.check(L.tcv_bn_fold(named[p + ".weight"].data_ptr(), named[p + ".bias"].data_ptr(),
                                          named[p + ".running_mean"].data_ptr(), t.data_ptr(), BN_EPS, c,
                                          s.data_ptr(), b.data_ptr(), st), "bn_fold")
        if self.s2d_stride2:
            self._derive_s2d(L, st)
        self._fingerprint = fp

    def _own_bias(self, p: str, b: torch.Tensor) -> None:
        """Engine-owned copy of a conv bias, updated in place: recorded plans / CUDA graphs bake in device pointers, and a
        module parameter's storage is not stable (nn.DataParallel broadcasts fresh replica tensors on every forward)."""
        own = self.bias.get(p)
        if own is None or own.device != b.device or own.shape != b.shape:
            own = self.bias[p] = torch.empty_like(b, memory_format=torch.contiguous_format)
        with torch.no_grad():
            own.copy_(b)

    def invalidate(self) -> None:
        """Forces the next call to re-derive every folded / packed weight.  Needed only after parameter updates that
        bypass torch's version counter (``p.data.copy_()``, ``vector_to_parameters``, EMA swaps through ``.data``)."""
        self._fingerprint = None

    def _pack(self, L, st, p, wbar, u, v, sig, transposed):
        if transposed:
            cin, cout, kh, kw = wbar.shape
        else:
            cout, cin, kh, kw = wbar.shape
        cin_pad = (cin + 7) // 8 * 8
        ent = self.w.get(p)
        if ent is None:
            ent = self.w[p] = dict(w=torch.empty((kh * kw, cin_pad, cout), dtype=torch.float32, device=wbar.device),
                                   cout=cout, cin=cin_pad, cin_real=cin, k=kh, transposed=transposed)
        _cabi.check(L.tcv_sn_fold_pack(wbar.data_ptr(), u.data_ptr() if u is not None else None,
                                       v.data_ptr() if v is not None else None, cout, cin, kh, kw,
                                       1 if transposed else 0, cin_pad, ent["w"].data_ptr(),
                                       sig.data_ptr() if sig is not None else None, st), "sn_fold_pack")
        if cin_pad == 8 and kh == 3 and cout % 32 == 0 and not transposed:
            if "w_tc_fold" not in ent:
                ent["w_tc_fold"] = torch.empty((2, 3, cout, 32), dtype=torch.bfloat16, device=wbar.device)
            _cabi.check(L.tcv_pack_weight_fold(ent["w"].data_ptr(), cout, ent["w_tc_fold"].data_ptr(), st),
                        "pack_weight_fold")
        if cin_pad % 32 == 0 and cout % 32 == 0:
            if "w_tc" not in ent:
                ent["w_tc"] = torch.empty((2, kh * kw, cout, cin_pad), dtype=torch.bfloat16, device=wbar.device)
            _cabi.check(L.tcv_pack_weight_tc(ent["w"].data_ptr(), kh * kw, cin_pad, cout, ent["w_tc"].data_ptr(), st),
                        "pack_weight_tc")

    HEAD32 = "#head32"

    def _pack_head32(self, L, st, p, w, b):
        """The 1-channel alpha head (decoder.conv2) as a 32-output-channel conv with zero-padded weights, so that it
        (and, in training, both of its gradients) runs on the tensor-core conv kernels; tcv_head_tanh01 reads channel 0."""
        key = p + self.HEAD32
        pads = self.__dict__.setdefault("_head32_pads", {})
        if key not in pads or pads[key][0].device != w.device:
            pads[key] = (torch.zeros((32,) + tuple(w.shape[1:]), dtype=torch.float32, device=w.device),
                         torch.zeros((32,), dtype=torch.float32, device=w.device))
        wpad, bpad = pads[key]
        with torch.no_grad():
            wpad[0].copy_(w[0])
            if b is not None:
                bpad[:1].copy_(b)
        self._pack(L, st, key, wpad, None, None, None, transposed=False)      # in place: recorded plans stay valid
        self.w[key]["cout_real"] = 1
        self.bias[key] = bpad

    S2D = "#s2d"
    # (layer, zero padding of the 3x3 / stride-2 conv: 1 = zero-padded input, 0 = reflect border materialised first,
    #  input channels per sub-pixel, output channels incl. padding, BatchNorm whose affine needs the same padding)
    S2D_LAYERS = (("encoder.conv1", 1, 8, 32, None),
                  ("encoder.guidance_head.1", 0, 8, 32, "encoder.guidance_head.3"),
                  ("encoder.guidance_head.5", 0, 32, 32, None))

    def _derive_s2d(self, L, st) -> None:
        """Space-to-depth forms of the stride-2 layers (opt-in TCV_S2D_STRIDE2): weights [T*T][4*cin][cout] + their
        tensor-core copies, zero-padded BatchNorm affines.  Updated in place like every other derived buffer."""
        for key, pad, cin_dst, cout_dst, bn in self.S2D_LAYERS:
            if key not in self.w:
                continue
            src = self.w[key]
            assert src["k"] == 3 and src["cin"] <= cin_dst and src["cout"] <= cout_dst, key
            ent = self.w.get(key + self.S2D)
            dev = src["w"].device
            if ent is None:
                ent = self.w[key + self.S2D] = dict(
                    w=torch.empty((4, 4 * cin_dst, cout_dst), dtype=torch.float32, device=dev),
                    w_tc=torch.empty((2, 4, cout_dst, 4 * cin_dst), dtype=torch.bfloat16, device=dev),
                    cout=cout_dst, cin=4 * cin_dst, cin_real=4 * src.get("cin_real", src["cin"]), k=2, transposed=False,
                    t0=-1 if pad else 0)
            _cabi.check(L.tcv_s2d_pack(src["w"].data_ptr(), 3, pad, src["cin"], src["cout"], cin_dst, cout_dst,
                                       ent["w"].data_ptr(), st), "s2d_pack")
            _cabi.check(L.tcv_pack_weight_tc(ent["w"].data_ptr(), 4, 4 * cin_dst, cout_dst, ent["w_tc"].data_ptr(), st),
                        "pack_weight_tc")
            if bn is not None and bn in self.aff:
                sc, sh = self.aff[bn]
                pk = bn + self.S2D
                if pk not in self.aff or self.aff[pk][0].device != dev:
                    self.aff[pk] = (torch.zeros(cout_dst, dtype=torch.float32, device=dev),
                                    torch.zeros(cout_dst, dtype=torch.float32, device=dev))
                with torch.no_grad():
                    self.aff[pk][0][: sc.numel()].copy_(sc)
                    self.aff[pk][1][: sh.numel()].copy_(sh)

    def conv_s2d(self, x: Act, wkey: str, *, bn: Optional[str] = None, act=ACT_NONE, bn2: Optional[str] = None) -> Act:
        """3x3 / stride-2 convolution of `x` through its space-to-depth form: x [n,h,w,c] -> [n,h/2,w/2,4c], then a
        stride-1 2x2-tap convolution (taps -1..0 for a zero-padded conv; 0..1 when `x` already carries its border)."""
        ent = self.w[wkey + self.S2D]
        assert x.h % 2 == 0 and x.w % 2 == 0 and ent["cin"] == 4 * x.c, (wkey, ent["cin"], x.c)
        assert x.plane == x.n * x.img_elems
        xs = self._act(x.n, x.h // 2, x.w // 2, 4 * x.c)
        self._call("tcv_space_to_depth2", x.ptr, x.plane, x.n, x.h, x.w, x.c, xs.ptr,
                   meta=dict(kind="tcv_space_to_depth2", bytes=8 * x.n * x.img_elems))
        t0 = ent["t0"]
        taps = [(ty, tx) for ty in (t0, t0 + 1) for tx in (t0, t0 + 1)]
        oh, ow = (xs.h, xs.w) if t0 < 0 else (xs.h - 1, xs.w - 1)
        cout = ent["cout"]
        y = self._act(x.n, oh, ow, cout)
        d = self._desc(xs, ent["w"].data_ptr(), taps, 1, PAD_ZERO, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey + self.S2D,
                       bn, False, act, None, 0, bn2, None, 0)
        self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey + self.S2D, xs, 2, 1))
        return y

    def get_plan(self, key) -> Optional[Plan]:
        plan = self.plans.get(key)
        if plan is not None:
            self.plans.move_to_end(key)
        return plan

    def put_plan(self, key, plan: Plan) -> None:
        self.plans[key] = plan
        self.plans.move_to_end(key)
        while len(self.plans) > max(self.max_plans, 1):
            self.plans.popitem(last=False)          # drops the buffers (and CUDA graph) of the oldest shape

    # ------------------------------------------------------------------ call recording
    @staticmethod
    def _check_device(dev: torch.device) -> None:
        if dev.type != "cuda":
            raise RuntimeError("tcvom_b200: the module must live on a CUDA device (no CPU fallback)")

    def _stream_ptr(self) -> int:
        st = self._stream_cache
        if st is not None:
            return st
        return torch.cuda.current_stream(self.device).cuda_stream

    @contextlib.contextmanager
    def stream_scope(self):
        """Pins the stream handle for the calls issued inside the block (one torch.cuda.current_stream() lookup instead of one
        per C-ABI call).  The caller must not switch streams inside."""
        prev = self._stream_cache
        self._stream_cache = torch.cuda.current_stream(self.device).cuda_stream
        try:
            yield
        finally:
            self._stream_cache = prev

    def _call(self, fn_name: str, *args, meta: Optional[dict] = None):
        # (the device context manager costs ~5 us per call and a training step makes 1 200 of them: enter it only when the
        # module's device is not the current one)
        dev = self.device
        if dev is None or dev.type != "cuda" or torch.cuda.current_device() == dev.index:
            self._call_on_device(fn_name, *args, meta=meta)
            return
        with self._device_guard():
            self._call_on_device(fn_name, *args, meta=meta)

    def _device_guard(self):
        """Kernels, memsets and tensor-map setup run on the CURRENT device while the stream comes from the module's device:
        make them agree for callers that moved the model with .to('cuda:1') without torch.cuda.set_device(1)."""
        if self.device is not None and self.device.type == "cuda":
            return torch.cuda.device(self.device)
        import contextlib
        return contextlib.nullcontext()

    def _call_on_device(self, fn_name: str, *args, meta: Optional[dict] = None):
        fn = getattr(_cabi.lib(), fn_name)
        st = self._stream_ptr()
        prof = getattr(self, "_prof", None)
        if prof is not None:                      # per-call CUDA-event timing (measurement helper, tools/train_bench.py)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
            _cabi.check(fn(*args, st), fn_name)
            e1.record(torch.cuda.current_stream(self.device))
            prof.append((fn_name, (meta or {}).get("tag", ""), e0, e1))
            return
        _cabi.check(fn(*args, st), fn_name)
        if self._rec is not None:
            self._rec.calls.append((fn, args, fn_name))
            m = dict(kind=fn_name, flops=0, bytes=0)
            if meta:
                m.update(meta)
            self._rec.meta.append(m)

    def _keep(self, *objs):
        if self._rec is not None:
            if self._rec.pool is not None:
                # pooled recording: device tensors live exactly as long as the program references them (their memory is
                # then reused by later buffers of the same plan); only host-side objects (descriptors) are pinned
                objs = [o for o in objs if not isinstance(o, torch.Tensor)]
            self._rec.keep.extend(objs)

    @contextlib.contextmanager
    def recording(self, plan: "Plan"):
        """Records the C-ABI calls issued inside the block into `plan`.  On a CUDA device the block allocates from a private
        ``torch.cuda.MemPool`` owned by the plan (TCV_PLAN_POOL=0: every buffer is kept for the plan's lifetime, the round-1
        behaviour): an intermediate whose last consumer has been issued is freed by Python's reference counting and its
        memory serves a later buffer of the same plan -- all kernels of a plan run in issue order on one stream, so the
        aliasing is safe, and replays see the same addresses.  A 1088x1920 GCA window needs its true peak liveness instead
        of the sum of all its buffers (10.5 GB)."""
        self._rec = plan
        cm = contextlib.nullcontext()
        if self.plan_pool and self.device is not None and self.device.type == "cuda":
            idle = _IDLE_POOLS.get(self.device.index)
            plan.pool = idle.pop() if idle else torch.cuda.MemPool()
            plan.pool_device = self.device.index
            cm = torch.cuda.use_mem_pool(plan.pool, device=self.device)
        _RECORDING[0] += 1
        try:
            with cm:
                yield plan
        finally:
            _RECORDING[0] -= 1
            self._rec = None

    def _empty(self, shape, dtype=torch.float32):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self._keep(t)
        return t

    def _act(self, n, h, w, c) -> Act:
        a = Act.empty(n, h, w, c, self.device)
        self._keep(a.buf)
        return a

    # ------------------------------------------------------------------ operators
    def pad_reflect1(self, x: Act) -> Act:
        y = self._act(x.n, x.h + 2, x.w + 2, x.c)
        assert x.plane == x.n * x.img_elems
        self._call("tcv_pad_reflect1", x.ptr, x.n, x.h, x.w, x.c, y.ptr,
                   meta=dict(kind="tcv_pad_reflect1", bytes=8 * x.n * x.img_elems))
        return y

    def conv(self, x: Act, wkey: str, *, stride=1, pad=PAD_ZERO, prepadded=False, bn: Optional[str] = None, bias=False,
             act=ACT_NONE, res1: Optional[Act] = None, res1_shift=0, bn2: Optional[str] = None,
             res2: Optional[Act] = None, f32_out: Optional[torch.Tensor] = None, f32_ptr: int = 0,
             want_split=True) -> Optional[Act]:
        ent = self.w[wkey]
        cout, k = ent["cout"], ent["k"]
        assert ent["cin"] == x.c, (wkey, ent["cin"], x.c)
        assert not ent["transposed"]
        if k == 3 and prepadded:
            # x already carries its 1-pixel border: taps are unshifted and never leave the tensor
            taps = [(ky, kx) for ky in range(3) for kx in range(3)]
            oh, ow = (x.h - 3) // stride + 1, (x.w - 3) // stride + 1
        elif k == 3:
            taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
            oh, ow = (x.h + 2 - 3) // stride + 1, (x.w + 2 - 3) // stride + 1
        else:
            taps = [(0, 0)]
            oh, ow = (x.h - 1) // stride + 1, (x.w - 1) // stride + 1
        y = self._act(x.n, oh, ow, cout) if want_split else None
        d = self._desc(x, ent["w"].data_ptr(), taps, stride, pad, y, oh, ow, cout, oh, ow, 1, 0, 1, 0, wkey, bn, bias,
                       act, res1, res1_shift, bn2, res2, f32_ptr if f32_ptr else
                       (f32_out.data_ptr() if f32_out is not None else 0))
        self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey, x, k, stride))
        return y

    def deconv4x4s2(self, x: Act, wkey: str, *, bn: str, act, res2: Optional[Act] = None) -> Act:
        """ConvTranspose2d(k=4, s=2, p=1) as 4 sub-pixel phases of 2x2 taps each."""
        ent = self.w[wkey]
        assert ent["transposed"] and ent["cin"] == x.c
        cout = ent["cout"]
        oh, ow = 2 * x.h, 2 * x.w
        y = self._act(x.n, oh, ow, cout)
        for py in range(2):
            for px in range(2):
                # oy = 2*iy - 1 + ky: phase py uses kernel rows (ky, input row offset dy)
                kys = [(1, 0), (3, -1)] if py == 0 else [(0, 1), (2, 0)]
                kxs = [(1, 0), (3, -1)] if px == 0 else [(0, 1), (2, 0)]
                taps = [(dy, dx) for ky, dy in kys for kx, dx in kxs]
                wtap = [ky * 4 + kx for ky, dy in kys for kx, dx in kxs]
                d = self._desc(x, ent["w"].data_ptr(), taps, 1, PAD_ZERO, y, oh, ow, cout, x.h, x.w, 2, py, 2, px,
                               wkey, bn, False, act, None, 0, None, res2, 0, wtap=wtap)
                self._call("tcv_conv2d", C.byref(d), meta=self._conv_meta(d, wkey, x, 4, 2))
        return y

    def _conv_meta(self, d: ConvDesc, wkey: str, x: Act, k: int, stride: int) -> dict:
        """Algorithmic work of one conv launch: 2*MACs over the real (unpadded) channels; bytes = every
        input/output/residual element moved once in split-bf16 (4 B/element) + the weights once."""
        cin_real = self.w[wkey].get("cin_real", d.cin)
        px = d.n * d.gh * d.gw
        flops = 2 * px * d.ntaps * cin_real * d.cout
        in_px = d.n * d.ih * d.iw if d.oy_mul == 1 else px
        nbytes = 4 * (in_px * d.cin + px * d.cout) + 4 * d.ntaps * d.cin * d.cout
        if d.res1:
            nbytes += 4 * (px >> (2 * d.res1_shift)) * d.cout
        if d.res2:
            nbytes += 4 * px * d.cout
        path = {0: "direct", 1: "tc", 2: "tc2", 3: "tc3", 4: "tc2p"}[_cabi.lib().tcv_conv2d_path(C.byref(d))]
        return dict(kind=f"conv_{path}", layer=wkey, flops=flops, bytes=nbytes,
                    shape=f"{d.cin}->{d.cout} k{k} s{stride} @{d.gh}x{d.gw} n{d.n}")

    def _desc(self, x: Act, wptr, taps, stride, pad, y: Optional[Act], oh, ow, cout, gh, gw, oy_mul, oy_off, ox_mul,
              ox_off, wkey, bn, bias, act, res1, res1_shift, bn2, res2, f32_ptr, wtap=None) -> ConvDesc:
        d = ConvDesc()
        d.x, d.x_plane, d.x_img_stride = x.ptr, x.plane, x.img_elems
        d.n, d.ih, d.iw, d.cin = x.n, x.h, x.w, x.c
        d.w, d.ntaps = wptr, len(taps)
        ent = self.w[wkey]
        if self.use_tc_conv and "w_tc" in ent:
            d.w_tc, d.w_tc_taps = ent["w_tc"].data_ptr(), ent["k"] * ent["k"]
        if self.use_tc_conv and "w_tc_fold" in ent:
            d.w_tc_fold = ent["w_tc_fold"].data_ptr()
        for i, (dy, dx) in enumerate(taps):
            d.dy[i], d.dx[i] = dy, dx
            d.wtap[i] = wtap[i] if wtap is not None else i
        d.stride, d.pad_mode = stride, pad
        d.y = y.ptr if y is not None else None
        d.y_f32 = f32_ptr or None
        d.oh, d.ow, d.cout, d.gh, d.gw = oh, ow, cout, gh, gw
        d.oy_mul, d.oy_off, d.ox_mul, d.ox_off = oy_mul, oy_off, ox_mul, ox_off
        if bn is not None:
            s, b = self.aff[bn]
            d.s1, d.b1 = s.data_ptr(), b.data_ptr()
        elif bias:
            d.b1 = self.bias[wkey].data_ptr()
        if res1 is not None:
            assert res1.c == cout and res1.n == x.n and res1.h == oh >> res1_shift and res1.w == ow >> res1_shift, wkey
            d.res1, d.res1_plane, d.res1_shift = res1.ptr, res1.plane, res1_shift
        d.act = act
        if bn2 is not None:
            s, b = self.aff[bn2]
            d.s2, d.b2 = s.data_ptr(), b.data_ptr()
        if res2 is not None:
            assert res2.c == cout and res2.n == x.n and res2.h == oh and res2.w == ow, wkey
            d.res2, d.res2_plane = res2.ptr, res2.plane
        self._keep(d)
        return d

    def avgpool2(self, x: Act) -> Act:
        y = self._act(x.n, x.h // 2, x.w // 2, x.c)
        assert x.plane == x.n * x.img_elems
        self._call("tcv_avgpool2", x.ptr, x.n, x.h, x.w, x.c, y.ptr)
        return y

    def gca(self, p: str, im_fea: Act, feat: Act, unknown: torch.Tensor) -> Act:
        """GuidedCxtAtten.forward (GCA/ops.py:106-229)."""
        n, h, w = feat.n, feat.h, feat.w
        assert feat.c == 128 and im_fea.c == 128 and (im_fea.h, im_fea.w) == (h, w)
        assert h % 2 == 0 and w % 2 == 0
        g = self.conv(im_fea, _k(p, "guidance_conv"), stride=2, bias=True)       # 1x1, then [::2, ::2]
        P = (h // 2) * (w // 2)
        P_pad = (P + 63) // 64 * 64
        mm = self._empty((n, P))
        scales = self._empty((n, 2))
        O = self._empty((n, P, 2048))
        if self.use_tc_attn and self.gca_shift_sum:
            return self._gca_shift_sum(p, g, feat, unknown, mm, scales)
        if self.use_tc_attn:
            # tcgen05 path: scores in bf16x3 (fp32-accurate logits), probabilities and values in bf16
            sp = self.score_planes
            Q = self._empty((sp, n, P, 576), torch.bfloat16)
            Kn = self._empty((sp, n, P, 576), torch.bfloat16)
            self._call("tcv_gca_prep", g.ptr, unknown.data_ptr(), n, h, w, Q.data_ptr(), Kn.data_ptr(),
                       mm.data_ptr(), scales.data_ptr(), sp)
            pv = {"bf16": (1, 1, 1, 0), "bf16x3": (2, 2, 3, 0), "fp16": (3, 1, 1, 1)}[self.pv_mode]
            mode, planes, nsplit, fp16 = pv
            Vt = self._empty((planes, n, 2048, P_pad), torch.bfloat16)
            self._call("tcv_gca_values", feat.ptr, n, h, w, Vt.data_ptr(), mode)
            Sm = self._empty((n, P, P_pad))
            self._call("tcv_gemm_tn_tc", Q.data_ptr(), n * P * 576, Kn.data_ptr(), n * P * 576, Sm.data_ptr(), P, P,
                       576, P_pad, P * P_pad, n, 6 if sp == 3 else 3, 0, 0,
                       meta=dict(kind="gca_scores_gemm_tc", flops=2 * n * P * P * 576,
                                 bytes=n * (2 * 2 * sp * P * 576 + 4 * P * P)))
            Pb = self._empty((planes, n, P, P_pad), torch.bfloat16)
            self._call("tcv_gca_softmax", Sm.data_ptr(), mm.data_ptr(), n, P, P_pad, Pb.data_ptr(), mode,
                       meta=dict(kind="tcv_gca_softmax", bytes=n * P * P * (4 + 2 * planes)))
            self._call("tcv_gemm_tn_tc", Pb.data_ptr(), n * P * P_pad, Vt.data_ptr(), n * 2048 * P_pad, O.data_ptr(), P,
                       2048, P_pad, 2048, P * 2048, n, nsplit, 0, fp16,
                       meta=dict(kind="gca_pv_gemm_tc", flops=2 * n * P * P * 2048,
                                 bytes=n * (2 * planes * (P * P + 2048 * P) + 4 * P * 2048)))
        else:
            Q = self._empty((n, P, 576))
            Kn = self._empty((n, P, 576))
            self._call("tcv_gca_prep", g.ptr, unknown.data_ptr(), n, h, w, Q.data_ptr(), Kn.data_ptr(),
                       mm.data_ptr(), scales.data_ptr(), 0)
            Vt = self._empty((n, 2048, P_pad))
            self._call("tcv_gca_values", feat.ptr, n, h, w, Vt.data_ptr(), 0)
            Sm = self._empty((n, P, P_pad))
            self._call("tcv_gemm_tn_f32", Q.data_ptr(), Kn.data_ptr(), Sm.data_ptr(), P, P, 576, 576, 576, P_pad,
                       P * 576, P * 576, P * P_pad, n,
                       meta=dict(kind="gca_scores_gemm", flops=2 * n * P * P * 576, bytes=4 * n * (2 * P * 576 + P * P)))
            self._call("tcv_gca_softmax", Sm.data_ptr(), mm.data_ptr(), n, P, P_pad, None, 0)
            self._call("tcv_gemm_tn_f32", Sm.data_ptr(), Vt.data_ptr(), O.data_ptr(), P, 2048, P_pad, P_pad, P_pad, 2048,
                       P * P_pad, 2048 * P_pad, P * 2048, n,
                       meta=dict(kind="gca_pv_gemm", flops=2 * n * P * P * 2048,
                                 bytes=4 * n * (P * P + 2048 * P + P * 2048)))
        Y = self._act(n, h, w, 128)
        self._call("tcv_gca_fold", O.data_ptr(), n, h, w, Y.ptr)
        self.last_gca_scales = scales
        return self.conv(Y, _k(p, "W.0"), bn=_k(p, "W.1"), res1=feat)

    def _gca_shift_sum(self, p: str, g: Act, feat: Act, unknown: torch.Tensor, mm, scales) -> Act:
        """The aggregation half of GuidedCxtAtten.forward in its shift-sum form (csrc/gca.cu, include/tcvom_b200.h):
        fold(A.V)/4 == A2.F on the (hh+1) x (ww+1) grid -- [Pk x ld].[ld x 512] instead of [P x P].[P x 2048] + overlap-add
        (ops.py:112-118,204).  All three GEMM operands are split-bf16 (three MMAs per K step)."""
        n, h, w = feat.n, feat.h, feat.w
        hh, ww = h // 2, w // 2
        P, Pk = hh * ww, (hh + 1) * (ww + 1)
        ld = (Pk + 63) // 64 * 64
        Q = self._empty((2, n, P, 576), torch.bfloat16)
        Kn = self._empty((2, n, Pk, 576), torch.bfloat16)
        self._call("tcv_gca_prep_grid", g.ptr, unknown.data_ptr(), n, h, w, Q.data_ptr(), Kn.data_ptr(), mm.data_ptr(),
                   scales.data_ptr())
        Ft = self._empty((2, n, 512, ld), torch.bfloat16)
        self._call("tcv_gca_values_parity", feat.ptr, n, h, w, ld, Ft.data_ptr(),
                   meta=dict(kind="tcv_gca_values_parity", bytes=n * (4 * h * w * 128 + 4 * 512 * ld)))
        Sm = self._empty((n, P, ld))
        self._call("tcv_gemm_tn_tc", Q.data_ptr(), n * P * 576, Kn.data_ptr(), n * Pk * 576, Sm.data_ptr(), P, Pk,
                   576, ld, P * ld, n, 3, 0, 0,
                   meta=dict(kind="gca_scores_gemm_tc", flops=2 * n * P * P * 576,
                             bytes=n * (2 * 2 * 2 * P * 576 + 4 * P * P)))
        stats = self._empty((n, P, 2))
        A2 = self._empty((2, n, Pk, ld), torch.bfloat16)
        if self.gca_softmax_in_consumer:
            self._call("tcv_gca_rowstats", Sm.data_ptr(), mm.data_ptr(), n, h, w, ld, stats.data_ptr(), 0,
                       meta=dict(kind="tcv_gca_rowstats", bytes=4 * n * P * ld))
            self._call("tcv_gca_softmax_shift", Sm.data_ptr(), stats.data_ptr(), mm.data_ptr(), n, h, w, ld, A2.data_ptr(),
                       meta=dict(kind="tcv_gca_softmax_shift", bytes=n * (4 * P * ld + 4 * Pk * ld)))
        else:
            self._call("tcv_gca_rowstats", Sm.data_ptr(), mm.data_ptr(), n, h, w, ld, stats.data_ptr(), 1,
                       meta=dict(kind="tcv_gca_softmax", bytes=8 * n * P * ld))
            self._call("tcv_gca_shift_add", Sm.data_ptr(), n, h, w, ld, A2.data_ptr(),
                       meta=dict(kind="tcv_gca_shift_add", bytes=n * (4 * P * ld + 4 * Pk * ld)))
        O2 = self._empty((n, Pk, 512))
        # flops: the reference's count for this aggregation (2*P*P*2048 per image); executed: 2*Pk*ld*512 (x3 split)
        self._call("tcv_gemm_tn_tc", A2.data_ptr(), n * Pk * ld, Ft.data_ptr(), n * 512 * ld, O2.data_ptr(), Pk, 512, ld,
                   512, Pk * 512, n, 3, 0, 0,
                   meta=dict(kind="gca_pv_gemm_tc", flops=2 * n * P * P * 2048, flops_executed=2 * n * Pk * ld * 512,
                             bytes=n * (4 * (Pk * ld + 512 * ld) + 4 * Pk * 512)))
        Y = self._act(n, h, w, 128)
        self._call("tcv_gca_unfold_parity", O2.data_ptr(), n, h, w, Y.ptr,
                   meta=dict(kind="tcv_gca_unfold_parity", bytes=n * (4 * Pk * 512 + 4 * h * w * 128)))
        self.last_gca_scales = scales
        return self.conv(Y, _k(p, "W.0"), bn=_k(p, "W.1"), res1=feat)

    def tam(self, p: str, x: Act, xb: Act, xf: Act, mask_ptr: int, mask_stride: int, mh: int, mw: int,
            attb_ptr: int, attf_ptr: int, sm_ptr: int) -> Act:
        """FeatureAggregationModule.forward (VMN_model.py:18-68)."""
        q = self.conv(x, _k(p, "query_conv"), bias=True)
        v = self.conv(x, _k(p, "value_conv"), bias=True)
        kb = self.conv(xb, _k(p, "key_conv"), bias=True)
        kf = self.conv(xf, _k(p, "key_conv"), bias=True)
        out = self._act(x.n, x.h, x.w, x.c)
        self._call("tcv_tam_attend", q.ptr, v.ptr, kb.ptr, kf.ptr, mask_ptr, mask_stride, mh, mw, x.n, x.h, x.w, x.c,
                   self.window, out.ptr, attb_ptr, attf_ptr, sm_ptr)
        return out

    # ------------------------------------------------------------------ network program
    def _enc_block(self, x: Act, p: str, stride: int) -> Act:
        o = self.conv(x, p + ".conv1", stride=stride, bn=p + ".bn1", act=ACT_RELU)
        idt = x
        if stride != 1:
            idt = self.conv(self.avgpool2(x), p + ".downsample.1", bn=p + ".downsample.2")
        return self.conv(o, p + ".conv2", bn=p + ".bn2", res1=idt, act=ACT_RELU)

    def _shortcut(self, x: Act, p: str) -> Act:
        o = self.conv(x, p + ".0", act=ACT_RELU, bn2=p + ".2")
        return self.conv(o, p + ".3", act=ACT_RELU, bn2=p + ".5")

    def _dec_layer(self, x: Act, p: str, blocks: int, res2: Optional[Act]) -> Act:
        for i in range(blocks):
            bp = f"{p}.{i}"
            last = res2 if i == blocks - 1 else None
            if i == 0:
                o = self.deconv4x4s2(x, bp + ".conv1", bn=bp + ".bn1", act=ACT_LEAKY02)
                idt = self.conv(x, bp + ".upsample.1", bn=bp + ".upsample.2")    # 1x1 at low res; nearest-up commutes
                x = self.conv(o, bp + ".conv2", bn=bp + ".bn2", res1=idt, res1_shift=1, act=ACT_LEAKY02, res2=last)
            else:
                o = self.conv(x, bp + ".conv1", bn=bp + ".bn1", act=ACT_LEAKY02)
                x = self.conv(o, bp + ".conv2", bn=bp + ".bn2", res1=x, act=ACT_LEAKY02, res2=last)
        return x

    def per_frame(self, x8: Act, shortcuts: str = "all") -> dict:
        """encoder + decoder head for all frames at once (VMN_model.py:93-98).

        shortcuts="all": the five shortcut branches of every frame (res_gca_enc.py:84-88), as the reference computes them
        (FrameStream caches them per frame: every frame becomes a centre frame once).  shortcuts="head": only branches 3
        and 4, which the per-frame decoder head consumes (VMN_GCA.py:28-31); branches 0..2 are read by the decoder TAIL of
        CENTRE frames only (VMN_GCA.py:38-44), so `tail` computes them for exactly those frames from the returned sources
        -- the reference evaluates them for the end frames too and throws the result away (a 3-frame window: 2/3 of the two
        full-resolution and the half- / quarter-resolution branch convolutions are dead code; same values for the rest)."""
        e = "encoder"
        s2d = self.s2d_stride2 and self.use_tc_conv
        if s2d:
            c1 = self.conv_s2d(x8, e + ".conv1", bn=e + ".bn1", act=ACT_RELU)
        else:
            c1 = self.conv(x8, e + ".conv1", stride=2, bn=e + ".bn1", act=ACT_RELU)
        x1 = self.conv(c1, e + ".conv2", bn=e + ".bn2", act=ACT_RELU)
        c3 = self.conv(x1, e + ".conv3", stride=2, bn=e + ".bn3", act=ACT_RELU)
        g = x8
        for ci, bi in ((1, 3), (5, 7), (9, 11)):                                # guidance head (res_gca_enc.py:20-33)
            if s2d and self.s2d_guidance and ci in (1, 5):
                # reflect border materialised once, then the space-to-depth form (output channels of .1 padded 16 -> 32
                # with zero weights / affine, which .5 reads through zero weight rows)
                g = self.conv_s2d(self.pad_reflect1(g), f"{e}.guidance_head.{ci}", act=ACT_RELU,
                                  bn2=f"{e}.guidance_head.{bi}" + (self.S2D if ci == 1 else ""))
            elif self.use_tc_conv and g.c % 32 == 0:
                # 32 -> 128: reflect border materialised once, then the stride-2 tcgen05 path
                g = self.conv(self.pad_reflect1(g), f"{e}.guidance_head.{ci}", stride=2, prepadded=True, act=ACT_RELU,
                              bn2=f"{e}.guidance_head.{bi}")
            else:
                g = self.conv(g, f"{e}.guidance_head.{ci}", stride=2, pad=PAD_REFLECT, act=ACT_RELU,
                              bn2=f"{e}.guidance_head.{bi}")
        im_fea = g
        unknown = self._empty((x8.n, x8.h // 8, x8.w // 8))
        self._call("tcv_unknown_os8", x8.ptr, x8.n, x8.h, x8.w, unknown.data_ptr())
        feats = []
        cur = c3
        for name, planes, blocks, stride in ENC_LAYERS:
            if name == "layer3":
                cur = self.gca(e + ".gca", im_fea, cur, unknown)
                feats[-1] = cur
            for i in range(blocks):
                cur = self._enc_block(cur, f"{e}.{name}.{i}", stride if i == 0 else 1)
            feats.append(cur)
        x2, x3, x4, emb = feats
        srcs = (x8, x1, x2, x3, x4)
        assert shortcuts in ("all", "head")
        fea = [self._shortcut(t, f"{e}.shortcut.{i}") if (shortcuts == "all" or i >= 3) else None
               for i, t in enumerate(srcs)]
        d = self._dec_layer(emb, "decoder.layer1", DEC_LAYERS[0][2], fea[4])
        d = self._dec_layer(d, "decoder.layer2", DEC_LAYERS[1][2], fea[3])
        feat = self.gca("decoder.gca", im_fea, d, unknown)
        out = dict(fea=fea, feat=feat, im_fea=im_fea, unknown=unknown)
        if shortcuts == "head":
            out["shortcut_src"] = list(srcs[:3])
        return out

    def tail(self, pf: dict, n0: int, ncen: int, mask_ptr: int, mask_stride: int, H: int, W: int, pred_ptr: int,
             attb_ptr: int, attf_ptr: int, sm_ptr: int) -> None:
        """decoder tail for `ncen` consecutive centre frames starting at image n0+1 (VMN_GCA.py:35-49)."""
        feat: Act = pf["feat"]
        x = feat.slice(n0 + 1, n0 + 1 + ncen)
        xb = feat.slice(n0, n0 + ncen)
        xf = feat.slice(n0 + 2, n0 + 2 + ncen)
        fea = [f.slice(n0 + 1, n0 + 1 + ncen) if f is not None else
               (self._shortcut(pf["shortcut_src"][i].slice(n0 + 1, n0 + 1 + ncen), f"encoder.shortcut.{i}") if i < 3 else None)
               for i, f in enumerate(pf["fea"])]
        t = self.tam("decoder.fam", x, xb, xf, mask_ptr, mask_stride, H, W, attb_ptr, attf_ptr, sm_ptr)
        t = self._dec_layer(t, "decoder.layer3", DEC_LAYERS[2][2], fea[2])
        t = self._dec_layer(t, "decoder.layer4", DEC_LAYERS[3][2], fea[1])
        t = self.deconv4x4s2(t, "decoder.conv1", bn="decoder.bn1", act=ACT_LEAKY02, res2=fea[0])
        hk = "decoder.conv2" + self.HEAD32
        hw_ = self.w["decoder.conv2"]
        if hw_["cin"] == 32 and hw_["cout"] == 1 and hw_["k"] == 3 and os.environ.get("TCV_HEAD_DIRECT", "1") == "1":
            # one HBM-bound pass: 3x3 conv to the single alpha channel + (tanh+1)/2 (the padded 32-channel tensor-core form
            # below writes and re-reads a 32-channel full-resolution tensor for one real channel)
            self._call("tcv_head_conv_tanh01", t.ptr, t.plane, t.n, t.h, t.w, hw_["w"].data_ptr(),
                       self.bias["decoder.conv2"].data_ptr(), pred_ptr,
                       meta=dict(kind="tcv_head_conv_tanh01", bytes=t.n * t.h * t.w * (4 * 32 + 4),
                                 flops=2 * t.n * t.h * t.w * 9 * 32))
        elif self.use_tc_conv and hk in self.w and os.environ.get("TCV_HEAD32", "1") == "1":
            z = self.conv(t, hk, bias=True)
            self._call("tcv_head_tanh01", z.ptr, z.plane, z.n * z.h * z.w, z.c, pred_ptr,
                       meta=dict(kind="tcv_head_tanh01", bytes=z.n * z.h * z.w * (2 * 32 + 4)))
        else:
            self.conv(t, "decoder.conv2", bias=True, act=ACT_TANH01, f32_ptr=pred_ptr, want_split=False)

    def window_program(self, x8: Act, trimask: torch.Tensor, B: int, S: int, H: int, W: int) -> dict:
        """Runs (and records) the whole VMN forward on preprocessed input.  trimask fp32 [B*S,H,W]."""
        dev = self.device
        ncen = S - 2
        N8 = (H // 8) * (W // 8)
        w2 = self.window * self.window
        pred = self._empty((B, ncen, 1, H, W))
        attb = self._empty((B, ncen, w2, N8))
        attf = self._empty((B, ncen, w2, N8))
        sm = self._empty((B, ncen, 1, H // 8, W // 8), torch.uint8)
        pf = self.per_frame(x8, shortcuts="head" if self.centre_shortcuts else "all")
        for b in range(B):
            n0 = b * S
            self.tail(pf, n0, ncen, trimask.data_ptr() + 4 * (n0 + 1) * H * W, H * W, H, W,
                      pred[b].data_ptr(), attb[b].data_ptr(), attf[b].data_ptr(), sm[b].data_ptr())
        return dict(pred=pred, attb=attb, attf=attf, small_mask=sm, feat=pf["feat"], pf=pf)
