"""GPU tests of tcvom_b200.FrameStream (per-frame feature reuse across sliding windows), the
FBA path at S=5 / B=2 and the trimap_transform drop-in (first green on a B200 in round 2, gpurun call 1).

The streamed mattes must equal what EvalModel.forward returns for every 3-frame window of the clip (same kernels on
the same values; the FBA path's GroupNorm sums are grouped differently for 1 and 3 images per launch -- fp64 partials, but a
scale / shift can land on the neighbouring float -- hence a 5e-5 bound there instead of bit equality)."""
import pytest
import torch

from helpers import fixture_sd, fixture_sd_dim, fixture_sd_fba, fixture_sd_index

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _clip(H, W, frames, seed):
    from tcvom_b200 import synthetic
    imgs, tris = synthetic.make_window(H, W, seed=seed, frames=frames)
    return torch.from_numpy(imgs).to(DEV), torch.from_numpy(tris).to(DEV)


@pytest.mark.parametrize("arch,dilate", [("vmn_gca", None), ("vmn_gca", 3), ("vmn_fba", None), ("vmn_dim", 2),
                                         ("vmn_index", None)])
def test_stream_equals_windowed_forward(arch, dilate):
    import tcvom_b200
    m = tcvom_b200.EvalModel(model=arch, agg_window=7, dilate_kernel=dilate)
    sd = {"vmn_gca": fixture_sd, "vmn_fba": fixture_sd_fba, "vmn_dim": fixture_sd_dim, "vmn_index": fixture_sd_index}[arch]()
    m.NET.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    H, W, T = 64, 96, 6
    imgs, tris = _clip(H, W, T, seed=15)
    tol = 5e-5 if arch == "vmn_fba" else 1e-6
    with torch.no_grad():
        stream = tcvom_b200.FrameStream(m, H, W, u8=True)
        outs = [stream.push(imgs[0, t], tris[0, t]) for t in range(T)]
        assert outs[0] is None and outs[1] is None
        for t in range(1, T - 1):
            ref = m(imgs[:, t - 1:t + 2].contiguous(), tris[:, t - 1:t + 2].contiguous())
            got = outs[t + 1]
            if arch == "vmn_fba":
                for g, r in zip(got, ref):
                    assert float((g - r[0, 1]).abs().max()) < tol
            else:
                assert float((got - ref[0, 1]).abs().max()) < tol
        # a new clip after reset(): the first two pushes return nothing again, the third equals the window
        stream.reset()
        assert stream.push(imgs[0, 2], tris[0, 2]) is None and stream.push(imgs[0, 3], tris[0, 3]) is None
        got = stream.push(imgs[0, 4], tris[0, 4])
        ref = m(imgs[:, 2:5].contiguous(), tris[:, 2:5].contiguous())
        a = got[0] if arch == "vmn_fba" else got
        r = ref[0] if arch == "vmn_fba" else ref
        assert float((a - r[0, 1]).abs().max()) < tol


def test_stream_rejects_cpu_frames_and_train_mode():
    import tcvom_b200
    m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7)
    m.NET.load_state_dict(fixture_sd(), strict=True)
    m = m.to(DEV).eval()
    stream = tcvom_b200.FrameStream(m, 64, 64)
    with pytest.raises(RuntimeError):
        stream.push(torch.zeros(3, 64, 64, dtype=torch.uint8), torch.zeros(1, 64, 64, dtype=torch.uint8))
    with pytest.raises(ValueError):
        tcvom_b200.FrameStream(m, 60, 64)
    m.train()
    with pytest.raises(NotImplementedError):
        tcvom_b200.FrameStream(m, 64, 64)


def test_fba_five_frame_samples_match_oracle():
    """EvalModel('vmn_fba') with S = 5 and B = 2 (tail launched with three centre frames per sample) vs the CPU oracle."""
    import tcvom_b200
    from oracle import vmn_fba_oracle as O
    from tcvom_b200 import synthetic
    sd = fixture_sd_fba()
    m = tcvom_b200.EvalModel(model="vmn_fba", agg_window=7)
    m.NET.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    imgs, tris = synthetic.make_window(64, 64, seed=21, frames=5, batch=2)
    ti, tt = torch.from_numpy(imgs), torch.from_numpy(tris)
    with torch.no_grad():
        a, Fg, Bg = m(ti.to(DEV), tt.to(DEV))
        ra, rF, rB = O.eval_forward(sd, ti.float(), tt.float())
    assert float((a.cpu() - ra).abs().max()) < 1e-3
    assert float((Fg.cpu() - rF).abs().max()) < 1e-3 and float((Bg.cpu() - rB).abs().max()) < 1e-3


def test_trimap_transform_operator():
    """tcvom_b200.trimap_transform (drop-in for utils/utils.py:25-39) vs the oracle on a 1088x1920 trimap pair."""
    import numpy as np
    import tcvom_b200
    from oracle import vmn_fba_oracle as O
    rng = np.random.default_rng(3)
    u = rng.uniform(size=(1, 2, 1088, 1920))
    t = np.zeros((1, 2, 2, 1088, 1920), np.float32)
    t[:, :, 0] = u < 0.0005
    t[:, :, 1] = u > 0.9999
    trimap = torch.from_numpy(t)
    got = tcvom_b200.trimap_transform(trimap.to(DEV)).cpu()
    assert float((got - O.trimap_transform(trimap)).abs().max()) < 2e-5
