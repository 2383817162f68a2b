"""Generates the FBA+TAM golden vectors (tests/golden/fba_*.npz, vmn_fba_keys.json) by running the UNMODIFIED
reference ``EvalModel('vmn_fba')`` on the CPU of the build container (needs /root/reference):

    python tests/golden/make_golden_fba.py

Shims as in make_golden.py (stub matplotlib; torch.cuda.current_device -> cpu for VMN_model.py:47,54).
The fixture checkpoint is regenerated from seeds (tcvom_b200.synthetic.fixture_state_dict_fba): GroupNorm
carries no running statistics, so nothing but the key/shape table has to be committed.

Per case: inputs (uint8), alphas / Fs / Bs of EvalModel.forward, the centre frame's raw 7-channel
prediction, TAM logits and mask, and (first case only) the preprocessed 11-channel input and the TAM input
feature, which pin the distance-transform encoding and the encoder + pyramid-pooling head separately.
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

from models.model import EvalModel  # noqa: E402  (reference)

from tcvom_b200 import synthetic  # noqa: E402

CASES = {
    "ring64": dict(H=64, W=64, seed=7, trimap="ring", dilate=None, batch=1),
    "ring96x128": dict(H=96, W=128, seed=8, trimap="ring", dilate=None, batch=1),
    "allunk64": dict(H=64, W=64, seed=9, trimap="all_unknown", dilate=None, batch=1),
    "nounk64": dict(H=64, W=64, seed=10, trimap="no_unknown", dilate=None, batch=1),
    "dil64x96": dict(H=64, W=96, seed=12, trimap="ring", dilate=5, batch=1),
    "batch2_64": dict(H=64, W=64, seed=13, trimap="ring", dilate=None, batch=2),
}


def main():
    torch.manual_seed(0)
    model = EvalModel(model="vmn_fba", agg_window=7, dilate_kernel=None)
    net = model.NET
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(HERE, "vmn_fba_keys.json"), "w") as f:
        json.dump({"state_dict": [[k, list(s)] for k, s in shapes.items()],
                   "trainable": [n for n, p in net.named_parameters() if p.requires_grad]}, f)
    net.load_state_dict(synthetic.fixture_state_dict_fba(shapes, 0), strict=True)
    model.eval()
    with torch.no_grad():
        for name, c in CASES.items():
            imgs, tris = synthetic.make_window(c["H"], c["W"], seed=c["seed"], trimap=c["trimap"], batch=c["batch"])
            model.DILATION_KERNEL = c["dilate"]
            ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
            alphas, Fs, Bs = model(ti, tt)
            scaled_imgs, scaled_tris, trimasks, nimgs = model.preprocess(ti, tt)
            S = ti.shape[1]
            inputs = list(torch.cat([nimgs, scaled_tris], dim=2).split(1, dim=1))
            extras = [[scaled_imgs[:, i], scaled_tris[:, i, -2:]] for i in range(S)]
            preds, attb, attf, small_mask = net(inputs, trimasks.split(1, dim=1), extras=extras)
            out = dict(imgs=imgs, tris=tris, dilate=np.array(-1 if c["dilate"] is None else c["dilate"]),
                       alphas=alphas.numpy(), Fs=Fs[:, 1].numpy().astype(np.float16), Bs=Bs[:, 1].numpy().astype(np.float16),
                       pred1=preds[1].numpy(), attb1=attb[1].numpy().astype(np.float16),
                       attf1=attf[1].numpy().astype(np.float16), small_mask1=small_mask[1].numpy(),
                       trimask=trimasks.numpy().astype(np.uint8))
            if name == "ring64":
                x11 = torch.cat([nimgs, scaled_tris], dim=2)
                co, indices = net.encoder(x11[:, 1])
                feat = net.decoder([co, indices, None, None], extract_feature=True)
                out.update(x11=x11.numpy(), feat1=feat.numpy(), conv5_sample=co[-1].numpy()[:, ::64])
            np.savez_compressed(os.path.join(HERE, f"fba_{name}.npz"), **out)
            m = trimasks[:, 1] > 0
            print(name, "alpha mean", float(alphas[:, 1].mean()), "unknown frac", float(trimasks.mean()), flush=True)
            if m.any():
                p = preds[1][:, :1][m]
                print("   unknown-band alpha mean/std", float(p.mean()), float(p.std()),
                      "saturated", float(((p < 1e-3) | (p > 1 - 1e-3)).float().mean()))
        model.DILATION_KERNEL = None


if __name__ == "__main__":
    main()
