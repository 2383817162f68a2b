"""GPU parity of the evaluation-metric kernel (tcv_frame_metrics) with the reference's golden vectors and the oracle."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import golden
from oracle import metrics_oracle as mo
from test_oracle_metrics import CASES, KEYS, close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_frame_metrics_match_reference_golden(name):
    from tcvom_b200 import metrics
    g = golden(f"metrics_{name}.npz")
    t = lambda k: torch.from_numpy(g[k]).cuda()
    close(metrics.frame_metrics(t("a0"), t("g0"), t("t0"), t("a1"), t("g1"), t("flow")), g["pair"])
    close(metrics.frame_metrics(t("a1"), t("g1"), t("t1")), g["single"])


def test_frame_metrics_1080p_match_oracle():
    from tcvom_b200 import metrics
    rng = np.random.default_rng(11)
    h, w = 1080, 1920
    g0 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    a0 = np.uint8(np.clip(g0.astype(np.int32) + rng.integers(-20, 21, (h, w)), 0, 255))
    g1 = np.roll(g0, 3, 1)
    a1 = np.uint8(np.clip(g1.astype(np.int32) + rng.integers(-20, 21, (h, w)), 0, 255))
    tri = rng.choice(np.array([0, 128, 255], np.uint8), (h, w), p=[0.4, 0.2, 0.4])
    flow = (rng.normal(0, 4, (h, w, 2)) + [3, 0]).astype(np.float32)
    flow[rng.random((h, w)) < 0.1] = np.nan
    want = mo.frame_metrics(a0, g0, tri, a1, g1, flow)
    got = metrics.frame_metrics(*(torch.from_numpy(v).cuda() for v in (a0, g0, tri, a1, g1, flow)))
    close(got, [want[k] for k in KEYS])
    # a frame with no unknown pixel: counts are zero, sums are zero
    z = metrics.frame_sums(*(torch.from_numpy(v).cuda() for v in (a0, g0, np.zeros_like(tri), a1, g1, flow)))
    assert z.cpu().tolist() == [0.0] * 7


def test_calc_metric_reads_the_reference_folder_layout(tmp_path):
    """Drop-in for calc_metric.calc_metric(fn, args): same files, same dictionary."""
    import cv2 as cv
    from tcvom_b200 import metrics
    g = golden("metrics_blob96x128.npz")
    pred, data = str(tmp_path / "pred"), str(tmp_path / "data")
    for d in (os.path.join(pred, "v"), os.path.join(data, "FG_done", "v"), os.path.join(data, "flow_png", "v")):
        os.makedirs(d)
    for t in range(2):
        cv.imwrite(os.path.join(pred, "v", f"{t:05d}_pred.png"), g[f"a{t}"])
        cv.imwrite(os.path.join(pred, "v", f"{t:05d}_tri.png"), g[f"t{t}"])
        fg = np.zeros(g["g0"].shape + (4,), np.uint8)
        fg[..., 3] = g[f"g{t}"]
        cv.imwrite(os.path.join(data, "FG_done", "v", f"{t:05d}.png"), fg)
    fl = g["flow"]
    png = np.zeros(fl.shape[:2] + (3,), np.uint16)
    png[..., :2] = np.int16(np.round(np.nan_to_num(fl) * 100)).view(np.uint16)
    png[..., 2] = ~np.isnan(fl[..., 0])
    cv.imwrite(os.path.join(data, "flow_png", "v", "flow_00000_00001.png"), png)
    args = types.SimpleNamespace(pred=pred, data=data, vis=False)
    close(metrics.calc_metric(("v/00000.png", "v/00001.png"), args, print_fn=False), g["pair"])
    close(metrics.calc_metric(("v/00001.png", ""), args, print_fn=False), g["single"])
