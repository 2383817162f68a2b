"""Times the native training step (FullModel_VMD fwd + bwd, train_ddp.py:52-65 without the optimizer) on one GPU:
BASELINE.json configs[2] (512x512 crop, batch 4, S=5) by default.  Prints step time, centre windows/s,
peak memory and a per-kernel-family breakdown from CUDA events around every C-ABI call."""
import argparse
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tcvom_b200  # noqa: E402
from helpers import fixture_sd  # noqa: E402
from tcvom_b200 import _cabi, synthetic  # noqa: E402

LOSS_WEIGHTS = (1.0, 1.0, 1.0, 0.5, 0.25)


def train_inputs(B, S, H, W, seed=21):
    return synthetic.make_train_batch(B, S, H, W, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--breakdown", type=int, default=1)
    args = ap.parse_args()
    dev = "cuda:0"
    if os.environ.get("TCV_PRE_FORWARD") == "1":
        # what bench.py does before its training sections: one recorded 1088x1920 forward window, then everything released
        import gc
        from tcvom_b200.engine import release_idle_pools
        m = tcvom_b200.EvalModel(model="vmn_gca", agg_window=7, dilate_kernel=None)
        m.NET.load_state_dict(fixture_sd(), strict=True)
        m = m.to(dev).eval()
        imgs, tris = synthetic.make_window(1088, 1920, seed=7)
        with torch.no_grad():
            for _ in range(3):
                m(torch.from_numpy(imgs).to(dev), torch.from_numpy(tris).to(dev))
        torch.cuda.synchronize()
        m.NET.engine().plans.clear()
        del m
        gc.collect(); release_idle_pools(); torch.cuda.empty_cache()
    model = tcvom_b200.FullModel_VMD(model="vmn_gca", agg_window=7, dilate_kernel=None)
    model.NET.load_state_dict(fixture_sd(), strict=True)
    model = model.to(dev).train()
    B, S, H, W = args.batch, args.frames, args.height, args.width
    a, fg, bg = (torch.from_numpy(t).float().to(dev) for t in train_inputs(B, S, H, W))
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-4)

    def step():
        out = model(a, fg, bg)
        loss = sum(w * o.mean() for w, o in zip(LOSS_WEIGHTS, out[:5]))
        model.zero_grad()
        loss.backward()
        opt.step()
        return loss

    torch.manual_seed(0)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    n0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    import time
    g0 = [g["collections"] for g in gc.get_stats()]
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    cpu_ms = (time.perf_counter() - t0) * 1e3 / args.steps       # time to ISSUE a step (no synchronisation inside)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    g1 = [g["collections"] for g in gc.get_stats()]
    launches = (_cabi.launch_count() - n0) // args.steps
    res = dict(workload=f"GCA+TAM train step (L_im+L_tc+L_af, fwd+bwd+Adam) {H}x{W} crop, batch {B}, S={S}",
               ms_per_step=ms, samples_per_s=B / (ms / 1e3), centre_windows_per_s=B * (S - 2) / (ms / 1e3),
               cpu_issue_ms_per_step=cpu_ms, gc_collections=[b - a for a, b in zip(g0, g1)], gc_objects=len(gc.get_objects()),
               launches_per_step=launches, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30, loss=float(loss))
    if args.breakdown:
        eng = model.NET.__dict__["_train_engines"][0]
        eng._prof = []
        step()
        torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0.0, 0])
        for name, tag, x0, x1 in eng._prof:
            agg[name][0] += x0.elapsed_time(x1)
            agg[name][1] += 1
        top = collections.defaultdict(lambda: [0.0, 0])
        for name, tag, x0, x1 in eng._prof:
            if tag:
                top[name + " " + tag][0] += x0.elapsed_time(x1)
                top[name + " " + tag][1] += 1
        res["top_tagged_ms"] = {k: [round(v[0], 3), v[1]] for k, v in sorted(top.items(), key=lambda kv: -kv[1][0])[:24]}
        eng._prof = None
        res["breakdown_ms"] = {k: [round(v[0], 3), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
