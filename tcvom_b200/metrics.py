"""Evaluation metrics of ``calc_metric.py`` on the GPU (second half).

The reference computes mSAD / MSE / SSDA / dtSSD with numpy on the host and MESSDdt with ``F.grid_sample`` on the CPU
(calc_metric.py:22-46, utils/utils.py:76-127), one frame per worker process.  Here one kernel pass
(``tcv_frame_metrics``, tcvom_b200/csrc/fba.cu) over the uint8 images as read from the PNGs produces the seven sums all of
them are made of; the host only divides and takes square roots.  No CPU path: without the CUDA library this module raises.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch

from . import _cabi

KEYS = ("mSAD", "MSE", "SSDA", "dtSSD", "MESSDdt_fix", "MESSDdt", "pixel_count", "flow_pixel_count")


def _u8(t: Optional[torch.Tensor], name: str, shape=None) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.uint8 or t.dim() != 2:
        raise TypeError(f"tcvom_b200.metrics: {name} must be a uint8 [H, W] tensor (the PNG's values)")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"tcvom_b200.metrics: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def frame_sums(alpha: torch.Tensor, gt: torch.Tensor, tri: torch.Tensor, next_alpha: Optional[torch.Tensor] = None,
               next_gt: Optional[torch.Tensor] = None, flow: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, _stream_ptr: Optional[int] = None) -> torch.Tensor:
    """The seven sums of one frame as a float64 [7] tensor on the images' device (no synchronisation):
    (pixel_count, sum |a-g|, sum (a-g)^2, sum ((a-ha)-(g-hg))^2, sum |(a-g)-(pa-pg)|, sum |(a-g)^2-(pa-pg)^2|,
    flow_pixel_count).  flow: float32 [H, W, 2] in pixels, NaN = invalid (calc_metric.py:64-70)."""
    alpha = _u8(alpha, "alpha")
    H, W = alpha.shape
    gt, tri = _u8(gt, "gt", (H, W)), _u8(tri, "tri", (H, W))
    next_alpha, next_gt = _u8(next_alpha, "next_alpha", (H, W)), _u8(next_gt, "next_gt", (H, W))
    if (next_alpha is None) != (next_gt is None):
        raise ValueError("tcvom_b200.metrics: next_alpha and next_gt come together")
    if flow is not None:
        if next_alpha is None:
            raise ValueError("tcvom_b200.metrics: flow needs the next frame")
        if flow.dtype != torch.float32 or tuple(flow.shape) != (H, W, 2):
            raise TypeError("tcvom_b200.metrics: flow must be float32 [H, W, 2]")
        flow = flow.contiguous()
    dev = alpha.device
    for t in (gt, tri, next_alpha, next_gt, flow):
        if t is not None and t.device != dev:
            raise ValueError("tcvom_b200.metrics: all inputs on one device")
    if _stream_ptr is None:
        if dev.type != "cuda":
            raise RuntimeError("tcvom_b200.metrics: CUDA tensors required (no CPU path)")
        _stream_ptr = torch.cuda.current_stream(dev).cuda_stream
    if out is None:
        out = torch.empty(7, dtype=torch.float64, device=dev)
    p = lambda t: 0 if t is None else t.data_ptr()
    _cabi.check(_cabi.lib().tcv_frame_metrics(p(alpha), p(gt), p(tri), p(next_alpha), p(next_gt), p(flow), H, W,
                                              out.data_ptr(), _stream_ptr), "frame_metrics")
    return out


def finish(sums, paired: bool) -> Dict[str, float]:
    """The reference's result dictionary (calc_metric.py:127-128) from the seven sums (a host sequence)."""
    s = [float(v) for v in sums]
    n = s[0]
    return {"mSAD": s[1] / n if n else float("nan"), "MSE": s[2] / n if n else float("nan"), "SSDA": math.sqrt(s[2]),
            "dtSSD": math.sqrt(s[3]) if paired else 0, "MESSDdt_fix": s[4] if paired else 0,
            "MESSDdt": s[5] if paired else 0, "pixel_count": int(n), "flow_pixel_count": int(s[6]) if paired else 0}


def frame_metrics(alpha, gt, tri, next_alpha=None, next_gt=None, flow=None) -> Dict[str, float]:
    """One frame's metrics, the dictionary ``calc_metric.calc_metric`` returns (one device->host read of 56 bytes)."""
    sums = frame_sums(alpha, gt, tri, next_alpha, next_gt, flow)
    return finish(sums.cpu().tolist(), next_alpha is not None)


def calc_metric(fn, args, print_fn=True, device=None) -> Dict[str, float]:
    """Drop-in for ``calc_metric.calc_metric(fn, args)`` (calc_metric.py:48-128 without the ``--vis`` branch): fn = (current
    frame, next frame or ''), args.pred / args.data the prediction and dataset folders.  Reads the PNGs with OpenCV like
    the reference, uploads the uint8 images and runs the metric kernel."""
    import cv2 as cv
    import numpy as np
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")

    def read(name):
        a = cv.imread(os.path.join(args.pred, name + "_pred.png"), cv.IMREAD_GRAYSCALE)
        t = cv.imread(os.path.join(args.pred, name + "_tri.png"), cv.IMREAD_GRAYSCALE)
        g = cv.imread(os.path.join(args.data, "FG_done", name + ".png"), cv.IMREAD_UNCHANGED)[..., -1]
        return [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (a, t, g)]

    if print_fn:
        print(fn[0])
    cf, nf = fn
    cfn = os.path.splitext(cf)[0]
    a, t, g = read(cfn)
    if nf == "":
        return frame_metrics(a, g, t)
    nfn = os.path.splitext(nf)[0]
    ha, _, hg = read(nfn)
    dirbase = os.path.dirname(cfn)
    assert dirbase == os.path.dirname(nfn), "{} | {}".format(cfn, nfn)
    x = cv.imread(os.path.join(args.data, "flow_png", dirbase, "flow_{}_{}.png".format(
        os.path.basename(cfn), os.path.basename(nfn))), cv.IMREAD_UNCHANGED)
    flow = np.float32(np.int16(x[..., :-1]))
    flow[x[..., -1] == 0] = np.nan
    return frame_metrics(a, g, t, ha, hg, torch.from_numpy(flow).to(dev) / 100.0)
