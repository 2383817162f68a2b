"""Golden vectors of the evaluation metrics: the UNMODIFIED reference's calc_metric.calc_metric (calc_metric.py:48-128) run in
the build container on a small synthetic dataset folder written to a temporary directory with OpenCV (prediction /
trimap / FG_done / flow_png PNGs in the layout the reference reads).  The decoded arrays and the reference's result
dictionaries are committed as tests/golden/metrics_*.npz.  Usage (build container only, needs /root/reference):
    python tests/golden/make_golden_metrics.py
"""
import os
import sys
import tempfile
import types

import cv2 as cv
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import calc_metric as ref  # noqa: E402

KEYS = ("mSAD", "MSE", "SSDA", "dtSSD", "MESSDdt_fix", "MESSDdt", "pixel_count", "flow_pixel_count")


def blob(rng, h, w, t):
    ys, xs = np.mgrid[0:h, 0:w]
    cx, cy = w * (0.45 + 0.02 * t), h * (0.5 + 0.01 * t)
    r = np.hypot((xs - cx) / (0.3 * w), (ys - cy) / (0.35 * h))
    return np.clip((1.15 - r) * 4.0, 0, 1)


def make_case(name, h, w, seed, flow_mag, invalid_frac, all_invalid=False):
    rng = np.random.default_rng(seed)
    root = tempfile.mkdtemp()
    pred, data = os.path.join(root, "pred"), os.path.join(root, "data")
    for d in (os.path.join(pred, "v"), os.path.join(data, "FG_done", "v"), os.path.join(data, "flow_png", "v")):
        os.makedirs(d)
    frames = {}
    for t in range(2):
        g = blob(rng, h, w, t)
        a = np.clip(g + rng.normal(0, 0.08, g.shape) * ((g > 0) & (g < 1)) + rng.normal(0, 0.01, g.shape), 0, 1)
        g8, a8 = np.uint8(np.round(g * 255)), np.uint8(np.round(a * 255))
        tri = np.where(g8 == 0, 0, np.where(g8 == 255, 255, 128)).astype(np.uint8)
        tri = np.where(cv.dilate(np.uint8(tri == 128), np.ones((5, 5), np.uint8)) > 0, 128, tri).astype(np.uint8)
        fg = np.zeros((h, w, 4), np.uint8)
        fg[..., :3] = rng.integers(0, 256, (h, w, 3))
        fg[..., 3] = g8
        cv.imwrite(os.path.join(pred, "v", f"{t:05d}_pred.png"), a8)
        cv.imwrite(os.path.join(pred, "v", f"{t:05d}_tri.png"), tri)
        cv.imwrite(os.path.join(data, "FG_done", "v", f"{t:05d}.png"), fg)
        frames[t] = (a8, g8, tri)
    # flow PNG: 16-bit, channels (fx*100, fy*100, valid) read back as int16 (calc_metric.py:64-70)
    fl = rng.normal(0, flow_mag, (h, w, 2)) + np.array([0.02 * w, 0.01 * h])
    fl16 = np.int16(np.round(fl * 100))
    valid = (rng.random((h, w)) >= invalid_frac) & (not all_invalid)
    png = np.zeros((h, w, 3), np.uint16)
    png[..., :2] = fl16.view(np.uint16)
    png[..., 2] = valid.astype(np.uint16)
    cv.imwrite(os.path.join(data, "flow_png", "v", "flow_00000_00001.png"), png)
    args = types.SimpleNamespace(pred=pred, data=data, vis=False)
    pair = ref.calc_metric(("v/00000.png", "v/00001.png"), args, print_fn=False)
    single = ref.calc_metric(("v/00001.png", ""), args, print_fn=False)
    flow = np.float32(fl16)
    flow[~valid] = np.nan
    flow = flow / np.float32(100.0)
    np.savez_compressed(os.path.join(HERE, f"metrics_{name}.npz"),
                        a0=frames[0][0], g0=frames[0][1], t0=frames[0][2], a1=frames[1][0], g1=frames[1][1],
                        t1=frames[1][2], flow=flow, pair=np.array([pair[k] for k in KEYS], np.float64),
                        single=np.array([single[k] for k in KEYS], np.float64))
    print(name, pair, single)


if __name__ == "__main__":
    make_case("blob96x128", 96, 128, 0, 1.5, 0.2)
    make_case("bigflow64x80", 64, 80, 1, 30.0, 0.05)          # samples that leave the image: zero padding
    make_case("noflow48x64", 48, 64, 2, 1.0, 1.0, all_invalid=True)
