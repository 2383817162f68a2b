"""Deterministic synthetic inputs and fixture weights (no dataset / checkpoint is available).

* ``make_window``   -- a 3-frame RGB+trimap window shaped like what ``pred_test.py:70-84``
  hands to ``EvalModel`` (BGR 0..255, trimap in {0,128,255}); SURVEY.md section 8(d) recipe:
  low-frequency texture + noise, disc foreground / ring unknown / background, moving centre.
* ``fixture_state_dict`` -- a full 584-key ``vmn_gca`` ``NET.state_dict()``.  Large conv
  weights are regenerated from a per-key seed (numpy PCG64, platform independent); the
  small data-dependent tensors (spectral-norm u/v after power iteration, BatchNorm running
  statistics after warm-up) come from ``tests/golden/fixture_vmn_gca_small.npz``, which
  ``tests/golden/make_golden.py`` produced by running the *reference* in train mode
  (SURVEY.md section 4 recipe: plain random init is degenerate -- zero-gamma residual
  branches, unconverged spectral norm).
"""
from __future__ import annotations

import os
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SMALL_STATE = os.path.join(GOLDEN_DIR, "fixture_vmn_gca_small.npz")


def _rng(tag: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(tag.encode())])


def _upsample_bilinear(a: np.ndarray, H: int, W: int) -> np.ndarray:
    h, w = a.shape
    ys = (np.arange(H) + 0.5) * h / H - 0.5
    xs = (np.arange(W) + 0.5) * w / W - 0.5
    y0 = np.clip(np.floor(ys).astype(np.int64), 0, h - 1)
    x0 = np.clip(np.floor(xs).astype(np.int64), 0, w - 1)
    y1 = np.clip(y0 + 1, 0, h - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    top = a[y0][:, x0] * (1 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (1 - fx) + a[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def make_window(H: int, W: int, seed: int = 7, frames: int = 3, batch: int = 1,
                trimap: str = "ring") -> Tuple[np.ndarray, np.ndarray]:
    """Returns (imgs uint8 [B,S,3,H,W] BGR, tris uint8 [B,S,1,H,W] in {0,128,255}).

    trimap: "ring" (disc fg, ring unknown of width ~H/8, bg), "all_unknown", "no_unknown".
    """
    imgs = np.zeros((batch, frames, 3, H, W), np.uint8)
    tris = np.zeros((batch, frames, 1, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(batch):
        rng = _rng(f"window{b}", seed)
        lh, lw = max(H // 8, 2) + 2, max(W // 8, 2) + 2
        base = rng.uniform(0, 255, size=(3, lh, lw))
        for s in range(frames):
            for c in range(3):
                tex = _upsample_bilinear(base[c], H + 16, W + 16)
                oy, ox = (2 * s) % 17, (3 * s) % 17         # texture drifts 2-3 px per frame (wraps inside the 16 px margin)
                t = tex[oy:oy + H, ox:ox + W] + rng.normal(0, 4.0, size=(H, W))
                imgs[b, s, c] = np.clip(np.floor(t), 0, 255).astype(np.uint8)
            if trimap == "all_unknown":
                tris[b, s, 0] = 128
            elif trimap == "no_unknown":
                cy, cx = H / 2 + 2 * s, W / 2 + 3 * s
                r = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
                tris[b, s, 0] = np.where(r < min(H, W) / 4, 255, 0)
            else:
                cy, cx = H / 2 + 2 * s - 1, W / 2 + 3 * s - 2
                r = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
                r_in = min(H, W) / 4.0
                r_out = r_in + max(H / 8.0, 2.0)
                tris[b, s, 0] = np.where(r < r_in, 255, np.where(r < r_out, 128, 0))
    return imgs, tris


def make_train_batch(B: int, S: int, H: int, W: int, seed: int = 21):
    """Synthetic training batch shaped like what dataset/VMD.py:300-301 hands to FullModel_VMD (SURVEY.md section 8d,
    config 3): alpha uint8 [B,S,1,H,W] with a soft-edged moving blob (every frame has unknown pixels), fg / bg uint8
    [B,S,3,H,W] low-frequency texture + noise (BGR 0..255)."""
    yy, xx = np.mgrid[0:H, 0:W]
    a = np.zeros((B, S, 1, H, W), np.uint8)
    fg = np.zeros((B, S, 3, H, W), np.uint8)
    bg = np.zeros((B, S, 3, H, W), np.uint8)
    for b in range(B):
        for s in range(S):
            cy, cx = H / 2 + 2 * s - 3 + 5 * b, W / 2 + 3 * s - 5 - 4 * b
            r = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
            a[b, s, 0] = np.round(np.clip((min(H, W) / 3.0 - r) / (min(H, W) / 6.0), 0, 1) * 255)
        fg[b] = make_window(H, W, seed=seed + 1 + 10 * b, frames=S)[0][0]
        bg[b] = make_window(H, W, seed=seed + 2 + 10 * b, frames=S)[0][0]
    return a, fg, bg


# ----------------------------------------------------------------------------- fixture weights
def _xavier(tag: str, shape, seed: int) -> np.ndarray:
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    bound = np.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
    return _rng(tag, seed).uniform(-bound, bound, size=shape).astype(np.float32)


def seeded_tensor(key: str, shape, seed: int = 0) -> np.ndarray:
    """Value of every state_dict entry that does NOT come from the calibrated npz."""
    if key.endswith("weight_bar") or (key.endswith(".weight") and len(shape) == 4):
        w = _xavier(key, shape, seed)
        if key == "encoder.conv1.module.weight_bar":
            pass                                            # keep trimap channels alive in the fixture
        return w
    if key.endswith(".bias") and ("conv" in key):          # conv biases (guidance_conv, fam.*, conv2)
        return _rng(key, seed).uniform(-0.05, 0.05, size=shape).astype(np.float32)
    if key.endswith(".weight") and len(shape) == 1:        # BN gamma
        g = _rng(key, seed).uniform(0.8, 1.2, size=shape).astype(np.float32)
        if ".bn2." in key or ".W.1." in key:
            g *= 0.5                                        # non-degenerate residual branches
        return g
    if key.endswith(".bias") and len(shape) == 1:          # BN beta
        return _rng(key, seed).uniform(-0.1, 0.1, size=shape).astype(np.float32)
    raise KeyError(key)


CALIBRATED_SUFFIXES = ("weight_u", "weight_v", "running_mean", "running_var", "num_batches_tracked")


def fixture_state_dict(shapes: Dict[str, tuple], seed: int = 0, small_npz: str = SMALL_STATE,
                       calibrated: bool = True) -> Dict[str, torch.Tensor]:
    """Assemble the fixture checkpoint.  ``shapes`` maps state_dict key -> shape (e.g. from
    ``tcvom_b200.get_VMN_models('vmn_gca', 7).state_dict()``)."""
    small = np.load(small_npz) if calibrated else None
    out: Dict[str, torch.Tensor] = {}
    for k, shp in shapes.items():
        if k.endswith(CALIBRATED_SUFFIXES):
            if small is None:
                continue
            out[k] = torch.from_numpy(np.asarray(small[k]))
        else:
            out[k] = torch.from_numpy(seeded_tensor(k, tuple(shp), seed))
    return out


# ----------------------------------------------------------------------------- FBA fixture weights
def seeded_tensor_fba(key: str, shape, shapes: Dict[str, tuple], seed: int = 0) -> np.ndarray:
    """Fixture value of one ``vmn_fba`` state_dict entry (203 keys, no calibrated part: GroupNorm has no running
    statistics).  Plain random init saturates alpha = clamp(out[:, 0], 0, 1) at 0 for 99.7 % of the pixels
    (measured on the reference), which would make an alpha parity test vacuous: the last 1x1 conv's alpha row is
    rescaled and its bias zeroed so that the unknown band is spread over (0, 1) (mean 0.55, std 0.27, 10 % saturated)."""
    if len(shape) == 4:
        w = _xavier(key, shape, seed)
        if key.endswith("conv_up4.4.weight"):
            w[0] *= 0.6
        return w
    if key.endswith(".weight"):                                # GroupNorm gamma
        g = _rng(key, seed).uniform(0.8, 1.2, size=shape).astype(np.float32)
        if ".bn3." in key:
            g *= 0.3        # damped residual branches: a random-weight ResNet-50 amplifies rounding noise ~500x otherwise
        return g
    sibling = shapes.get(key[: -len(".bias")] + ".weight")
    if sibling is not None and len(sibling) == 1:              # GroupNorm beta
        return _rng(key, seed).uniform(-0.1, 0.1, size=shape).astype(np.float32)
    b = _rng(key, seed).uniform(-0.05, 0.05, size=shape).astype(np.float32)   # conv bias
    if key.endswith("conv_up4.4.bias"):
        b[0] = 0.0
    return b


def fixture_state_dict_fba(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Fixture checkpoint of ``get_VMN_models('vmn_fba')`` (shapes: state_dict key -> shape)."""
    return {k: torch.from_numpy(seeded_tensor_fba(k, tuple(s), shapes, seed)) for k, s in shapes.items()}
