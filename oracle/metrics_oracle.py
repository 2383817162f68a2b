"""CPU oracle for the evaluation metrics of ``calc_metric.py``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tcvom_b200/`` imports this module; ``tests/`` uses it as the checker.

A restatement (numpy + ``F.grid_sample``) of calc_metric.py:22-46 and utils/utils.py:72-127 at commit f5fa07a.  Parity
pinning: upstream has no fixtures for this path; the oracle is pinned against the reference's own functions executed in the
build container by ``tests/golden/make_golden_metrics.py`` (committed vectors ``tests/golden/metrics_*.npz``,
``tests/test_oracle_metrics.py``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def preprocess(alpha_u8, gt_u8, tri_u8):
    """calc_metric.py:59-63"""
    return np.float32(alpha_u8 / 255.0), np.float32(gt_u8 / 255.0), (tri_u8 > 0) * (tri_u8 < 255)


def warped_errors(a, g, m, ha, hg, flow):
    """MESSDdt -- calc_metric.py:36-46 with utils.flow_dt(metric=True) (utils/utils.py:92-127).  flow [H, W, 2] float32."""
    H, W = a.shape
    flow = torch.from_numpy(np.array(flow, copy=True))
    nan = torch.isnan(flow)
    flow[nan] = 0
    valid = (~nan[..., 0]) & torch.from_numpy(m)           # the x channel's mask only (utils/utils.py:113)
    if int(valid.sum()) == 0:
        return 0.0, 0.0, 0
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    cx, cy = xs + flow[..., 0], ys + flow[..., 1]          # utils/utils.py:72-75,111
    grid = torch.stack([2 * cx / (W - 1) - 1, 2 * cy / (H - 1) - 1], -1)[None]   # utils/utils.py:82-86

    def warp(img):
        return F.grid_sample(torch.from_numpy(img)[None, None], grid, mode="bilinear", align_corners=True)[0, 0]

    pa, pg = warp(ha), warp(hg)
    d = torch.from_numpy(a)[valid] - torch.from_numpy(g)[valid]
    e = pa[valid] - pg[valid]
    return float(torch.abs(d - e).sum()), float(torch.abs(d ** 2 - e ** 2).sum()), int(valid.sum())


def frame_metrics(alpha_u8, gt_u8, tri_u8, next_alpha_u8=None, next_gt_u8=None, flow=None):
    """The dictionary calc_metric.calc_metric returns (calc_metric.py:72-98,127-128)."""
    a, g, m = preprocess(alpha_u8, gt_u8, tri_u8)
    out = {"mSAD": float(np.mean(np.abs(a[m] - g[m]))), "MSE": float(np.mean((a[m] - g[m]) ** 2)),
           "SSDA": float(np.sqrt(np.sum((a[m] - g[m]) ** 2))), "dtSSD": 0, "MESSDdt_fix": 0, "MESSDdt": 0,
           "pixel_count": int(np.sum(m)), "flow_pixel_count": 0}
    if next_alpha_u8 is not None:
        ha, hg, _ = preprocess(next_alpha_u8, next_gt_u8, tri_u8)
        out["dtSSD"] = float(np.sqrt(np.sum(((a - ha)[m] - (g - hg)[m]) ** 2)))
        out["MESSDdt_fix"], out["MESSDdt"], out["flow_pixel_count"] = warped_errors(a, g, m, ha, hg, flow)
    return out
