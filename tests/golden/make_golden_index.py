"""Generates the IndexNet+TAM golden vectors (tests/golden/index_*.npz, vmn_index_keys.json) by running the UNMODIFIED reference
``EvalModel('vmn_index')`` on the CPU of the build container (needs /root/reference):

    python tests/golden/make_golden_index.py

Shims as in make_golden.py (stub matplotlib; torch.cuda.current_device -> cpu for VMN_model.py:47,54).  The fixture
checkpoint is regenerated from a seed (oracle.vmn_index_oracle.fixture_sd_index), so only the key / shape table is committed.

Per case: inputs (uint8), alphas of EvalModel.forward, the centre frame's raw prediction, TAM logits and mask and (first
case only) the 4-channel network input and the OS8 feature the TAM reads.
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

from oracle.vmn_index_oracle import fixture_sd_index  # noqa: E402
from tcvom_b200 import synthetic  # noqa: E402

CASES = {
    "ring64x96": dict(H=64, W=96, seed=3, trimap="ring", dilate=None, batch=1),
    "dil96x64": dict(H=96, W=64, seed=12, trimap="ring", dilate=3, batch=1),
    "batch2_64": dict(H=64, W=64, seed=13, trimap="ring", dilate=None, batch=2),
}


def main():
    model = EvalModel(model="vmn_index", agg_window=7, dilate_kernel=None)
    net = model.NET
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(HERE, "vmn_index_keys.json"), "w") as f:
        json.dump({"state_dict": [[k, list(s)] for k, s in shapes.items()]}, f)
    sd = fixture_sd_index(shapes)
    assert list(sd.keys()) == list(shapes.keys())
    net.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        for name, c in CASES.items():
            imgs, tris = synthetic.make_window(c["H"], c["W"], seed=c["seed"], trimap=c["trimap"], batch=c["batch"])
            model.DILATION_KERNEL = c["dilate"]
            ti, tt = torch.from_numpy(imgs).float(), torch.from_numpy(tris).float()
            alphas = model(ti, tt)
            scaled_imgs, scaled_tris, trimasks, nimgs = model.preprocess(ti, tt)
            x4 = torch.cat([nimgs, scaled_tris], dim=2)
            preds, attb, attf, small_mask = net(list(x4.split(1, dim=1)), trimasks.split(1, dim=1))
            out = dict(imgs=imgs, tris=tris, dilate=np.array(-1 if c["dilate"] is None else c["dilate"]),
                       alphas=alphas.numpy(), pred1=preds[1].numpy(), attb1=attb[1].numpy().astype(np.float16),
                       attf1=attf[1].numpy().astype(np.float16), small_mask1=small_mask[1].numpy(),
                       trimask=trimasks.numpy().astype(np.uint8))
            if name == "ring64x96":
                dec_in = list(net.encoder(x4[:, 1]))
                feat = net.decoder(dec_in, extract_feature=True)
                out.update(x4=x4.numpy(), feat1=feat.numpy())
            np.savez_compressed(os.path.join(HERE, f"index_{name}.npz"), **out)
            p = preds[1]
            print(name, "alpha mean", float(alphas[:, 1].mean()), "pred range", float(p.min()), float(p.max()),
                  "saturated", float(((p < 1e-3) | (p > 1 - 1e-3)).float().mean()), flush=True)
        model.DILATION_KERNEL = None


if __name__ == "__main__":
    main()
