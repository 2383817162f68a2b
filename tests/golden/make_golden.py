"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

Two non-invasive shims are needed to import/run the reference on a CPU-only host
(SURVEY.md section 8c): stub ``matplotlib`` (imported by models/Index/hldecoder.py:36) and
redirect ``torch.cuda.current_device`` (VMN_model.py:47,54 hard-code CUDA).

Outputs (all small, committed):
  fixture_vmn_gca_small.npz   spectral-norm u/v + BatchNorm running stats after warm-up
  vmn_gca_keys.json           ordered state_dict keys/shapes + ordered trainable parameter names
  eval_<case>.npz             EvalModel.forward inputs (uint8) and outputs
  op_gca.npz / op_tam.npz     GuidedCxtAtten / FeatureAggregationModule outputs on seeded inputs
  train_s5.npz                FullModel_VMD.forward (S=5) losses and alphas
"""
import json
import os
import sys
import types

REF = os.environ.get("TCVOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.cuda.current_device = lambda: torch.device("cpu")
torch.set_num_threads(8)

from models.model import EvalModel, FullModel_VMD  # noqa: E402  (reference)

from tcvom_b200 import synthetic  # noqa: E402

SEED = 0
WARMUP_PASSES = 24


def op_inputs(tag, shape, seed=11):
    return synthetic._rng(tag, seed).standard_normal(size=shape).astype(np.float32)


def train_inputs(H, W, S=5, seed=21):
    """alpha GT with a soft-edged moving blob (every frame has unknown pixels); fg/bg low-frequency texture
    + noise like the eval windows.  (White-noise fg/bg make the random-weight fixture chaotic: two fp32 CPU
    implementations already differ by 1e-4 on alpha, so it cannot serve as a 1e-3 parity vector.)"""
    rng = synthetic._rng("train", seed)
    yy, xx = np.mgrid[0:H, 0:W]
    a = np.zeros((1, S, 1, H, W), np.uint8)
    for s in range(S):
        cy, cx = H / 2 + 2 * s - 3, W / 2 + 3 * s - 5
        r = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
        soft = np.clip((min(H, W) / 3.0 - r) / (min(H, W) / 6.0), 0, 1)
        a[0, s, 0] = np.round(soft * 255)
    fg = synthetic.make_window(H, W, seed=seed + 1, frames=S)[0]
    bg = synthetic.make_window(H, W, seed=seed + 2, frames=S)[0]
    return a, fg, bg


def main(only_train=False):
    torch.manual_seed(0)
    if only_train:
        from helpers_golden import load_full
        full = load_full()
        tm = FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
        tm.NET.load_state_dict(full, strict=True)
        tm.eval()
        a, fg, bg = train_inputs(64, 64)
        with torch.no_grad():
            out = tm(torch.from_numpy(a).float(), torch.from_numpy(fg).float(), torch.from_numpy(bg).float())
        np.savez_compressed(os.path.join(HERE, "train_s5.npz"), a=a, fg=fg, bg=bg,
                            losses=np.array([float(o) for o in out[:5]], np.float64),
                            alphas=out[7].numpy(), comps=out[8].numpy(), tris_vis=out[6].numpy(),
                            scaled_imgs=out[5].numpy())
        print("train losses", [float(o) for o in out[:5]])
        return
    model = EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None)
    net = model.NET
    sd0 = net.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd0.items()}
    with open(os.path.join(HERE, "vmn_gca_keys.json"), "w") as f:
        json.dump({"state_dict": [[k, list(s)] for k, s in shapes.items()],
                   "trainable": [n for n, p in net.named_parameters() if p.requires_grad]}, f)

    seeded = synthetic.fixture_state_dict(shapes, SEED, calibrated=False)
    missing = net.load_state_dict(seeded, strict=False)
    assert all(k.endswith(synthetic.CALIBRATED_SUFFIXES) for k in missing.missing_keys), missing

    # warm-up: converge power iteration, make BN running stats real
    model.train()
    with torch.no_grad():
        for i in range(WARMUP_PASSES):
            imgs, tris = synthetic.make_window(128, 128, seed=100 + i)
            model(torch.from_numpy(imgs).float(), torch.from_numpy(tris).float())
            print("warmup", i, flush=True)
    model.eval()
    sd = net.state_dict()
    small = {k: v.numpy() for k, v in sd.items() if k.endswith(synthetic.CALIBRATED_SUFFIXES)}
    np.savez_compressed(os.path.join(HERE, "fixture_vmn_gca_small.npz"), **small)
    # round trip: what tests will load must equal what the reference holds now
    full = synthetic.fixture_state_dict(shapes, SEED)
    for k in sd:
        assert torch.equal(full[k], sd[k]), k

    cases = {
        "ring64": dict(H=64, W=64, seed=7, trimap="ring", dilate=None, batch=1),
        "ring96x128": dict(H=96, W=128, seed=8, trimap="ring", dilate=None, batch=1),
        "allunk64": dict(H=64, W=64, seed=9, trimap="all_unknown", dilate=None, batch=1),
        "nounk64": dict(H=64, W=64, seed=10, trimap="no_unknown", dilate=None, batch=1),
        "dil64x96": dict(H=64, W=96, seed=12, trimap="ring", dilate=5, batch=1),
        "batch2_64": dict(H=64, W=64, seed=13, trimap="ring", dilate=None, batch=2),
    }
    with torch.no_grad():
        for name, c in cases.items():
            imgs, tris = synthetic.make_window(c["H"], c["W"], seed=c["seed"], trimap=c["trimap"],
                                               batch=c["batch"])
            model.DILATION_KERNEL = c["dilate"]
            ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
            alphas = model(ti, tt)
            scaled_imgs, scaled_tris, trimasks, nimgs = model.preprocess(ti, tt)
            inputs = list(torch.cat([nimgs, scaled_tris], dim=2).split(1, dim=1))
            preds, attb, attf, small_mask = net(inputs, trimasks.split(1, dim=1))
            np.savez_compressed(
                os.path.join(HERE, f"eval_{name}.npz"), imgs=imgs, tris=tris,
                dilate=np.array(-1 if c["dilate"] is None else c["dilate"]),
                alphas=alphas.numpy(), pred1=preds[1].numpy(), attb1=attb[1].numpy(),
                attf1=attf[1].numpy(), small_mask1=small_mask[1].numpy(),
                trimask=trimasks.numpy().astype(np.uint8))
            print(name, "alpha mean", float(alphas[:, 1].mean()),
                  "unknown frac", float(trimasks.mean()), flush=True)
            m = trimasks[:, 1] > 0
            if m.any():
                p = preds[1][m]
                print("   unknown-band pred mean/std", float(p.mean()), float(p.std()),
                      "saturated", float(((p < 1e-3) | (p > 1 - 1e-3)).float().mean()))
        model.DILATION_KERNEL = None

        # ---- operator-level goldens
        gca = net.decoder.gca
        f = torch.from_numpy(op_inputs("gca_f", (2, 128, 16, 24)))
        al = torch.from_numpy(op_inputs("gca_alpha", (2, 128, 16, 24)))
        unk = torch.from_numpy((op_inputs("gca_unk", (2, 1, 16, 24)) > 0.3).astype(np.float32))
        y, (offsets, scale) = gca(f, al, unk)
        np.savez_compressed(os.path.join(HERE, "op_gca.npz"), y=y.numpy(), scale=scale.numpy())

        fam = net.decoder.fam
        x = torch.from_numpy(op_inputs("tam_x", (2, 128, 12, 16)))
        b = torch.from_numpy(op_inputs("tam_b", (2, 128, 12, 16)))
        fw = torch.from_numpy(op_inputs("tam_f", (2, 128, 12, 16)))
        mask = torch.from_numpy((op_inputs("tam_m", (2, 1, 96, 128)) > 0.5).astype(np.float32))
        feat, attb, attf, sm = fam(x, b, fw, mask)
        np.savez_compressed(os.path.join(HERE, "op_tam.npz"), feat=feat.numpy(), attb=attb.numpy(),
                            attf=attf.numpy(), small_mask=sm.numpy())

        # ---- training-side wrapper (losses), S=5 so that L_tc is non-zero
        tm = FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3)
        tm.NET.load_state_dict(full, strict=True)
        tm.eval()
        a, fg, bg = train_inputs(64, 64)
        out = tm(torch.from_numpy(a).float(), torch.from_numpy(fg).float(), torch.from_numpy(bg).float())
        np.savez_compressed(os.path.join(HERE, "train_s5.npz"), a=a, fg=fg, bg=bg,
                            losses=np.array([float(o) for o in out[:5]], np.float64),
                            alphas=out[7].numpy(), comps=out[8].numpy(), tris_vis=out[6].numpy(),
                            scaled_imgs=out[5].numpy())
        print("train losses", [float(o) for o in out[:5]])


LOSS_WEIGHTS = (1.0, 1.0, 1.0, 0.5, 0.25)     # train_ddp.py:61
GRAD_STRIDE = 257                               # every 257th element of each flattened gradient is stored


def grad_sample(g):
    """whole tensor when small, else every GRAD_STRIDE-th element"""
    f = g.detach().flatten()
    return f if f.numel() <= 8192 else f[::GRAD_STRIDE]


DAMP = 0.04      # residual-branch BatchNorm gains of the well-conditioned fixture = DAMP x the standard fixture's


def damp_state(sd, damp=DAMP):
    """The standard fixture with every residual-branch BatchNorm gain (``*.bn2.weight``, ``gca.W.1.weight``) scaled by
    `damp`: the network stays the same program, but a perturbation is no longer amplified ~1000x through the 29 residual
    blocks, so a whole-step gradient comparison resolves 1e-3 instead of sitting at the 1e-2 noise floor of the standard
    fixture (measured on the CPU oracle, fp32 against fp64: median gradient rel-L2 2.5e-5 instead of 2.7e-3)."""
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith(".bn2.weight") or k.endswith("W.1.weight"):
            out[k] = out[k] * damp
    return out


def train_step_golden(B=2, S=5, H=64, W=64, damp=None, freeze=False):
    """One reference training step (train_ddp.py:52-65 without the optimizer): FullModel_VMD in .train() mode,
    loss = L_alpha + L_comp + L_grad + 0.5 L_dt + 0.25 L_att, backward.  Stores the losses, every gradient's
    L2 norm / sum and a strided sample, and the state the forward mutates (spectral-norm u/v, BatchNorm
    running statistics).  dilate_kernel is fixed so that no host RNG is involved."""
    from helpers_golden import load_full
    full = load_full()
    if damp is not None:
        full = damp_state(full, damp)
    # freeze=True: the TAM pre-training mode (train_single_ddp.py:184-185, VMN_model.py:77-81,99-103, VMN_GCA.py:18-24):
    # encoder + decoder.layer1 / layer2 / gca in eval mode under no_grad, the rest of the decoder trains
    tm = FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=3, **(dict(freeze_backbone=True) if freeze else {}))
    tm.NET.load_state_dict(full, strict=True)
    tm.train()
    aa, ff, bb = [], [], []
    for b in range(B):
        a, fg, bg = train_inputs(H, W, S=S, seed=31 + 10 * b)
        aa.append(a); ff.append(fg); bb.append(bg)
    a, fg, bg = np.concatenate(aa), np.concatenate(ff), np.concatenate(bb)
    out = tm(torch.from_numpy(a).float(), torch.from_numpy(fg).float(), torch.from_numpy(bg).float())
    loss = sum(wt * o.mean() for wt, o in zip(LOSS_WEIGHTS, out[:5]))
    tm.zero_grad()
    loss.backward()
    res = dict(a=a, fg=fg, bg=bg, losses=np.array([float(o) for o in out[:5]], np.float64),
               alphas=out[7].detach().numpy())
    names = []
    res["nograd"] = np.array([n for n, p in tm.NET.named_parameters() if p.requires_grad and p.grad is None])
    for n, p in tm.NET.named_parameters():
        if not p.requires_grad:
            continue
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(n)
        res["gn:" + n] = np.array([float(g.double().norm()), float(g.double().sum())])
        res["gs:" + n] = grad_sample(g).numpy().copy()
    for k, v in tm.NET.state_dict().items():
        if k.endswith(("weight_u", "weight_v", "running_mean", "running_var", "num_batches_tracked")):
            res["st:" + k] = v.numpy().copy()
    fname = "train_step_s5_freeze.npz" if freeze else ("train_step_s5.npz" if damp is None else "train_step_s5_damped.npz")
    np.savez_compressed(os.path.join(HERE, fname), **res)
    print("train step losses", res["losses"], "params", len(names), "damp", damp)


if __name__ == "__main__":
    if "--train-step-freeze" in sys.argv:
        train_step_golden(damp=DAMP, freeze=True)
    elif "--train-step-damped" in sys.argv:
        train_step_golden(damp=DAMP)
    elif "--train-step" in sys.argv:
        train_step_golden()
    else:
        main(only_train="--only-train" in sys.argv)
