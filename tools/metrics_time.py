"""Times tcv_frame_metrics on 1080p frames (CUDA events, inputs rotated through more than L2) beside the CPU restatement of
calc_metric.py's per-frame function (oracle/metrics_oracle.py; test infrastructure, here as the timed CPU baseline only)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tcvom_b200 import metrics  # noqa: E402
from oracle import metrics_oracle as mo  # noqa: E402

h, w, sets, reps = 1080, 1920, 12, 20
rng = np.random.default_rng(0)
host = []
for _ in range(sets):
    g0 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    a0 = np.uint8(np.clip(g0.astype(np.int32) + rng.integers(-20, 21, (h, w)), 0, 255))
    g1 = np.roll(g0, 3, 1)
    a1 = np.uint8(np.clip(g1.astype(np.int32) + rng.integers(-20, 21, (h, w)), 0, 255))
    # a matting trimap: an unknown band (about 15 % of the frame) around an elliptic foreground
    ys, xs = np.mgrid[0:h, 0:w]
    r = np.hypot((xs - 0.5 * w) / (0.3 * w), (ys - 0.5 * h) / (0.4 * h))
    tri = np.where(r < 0.85, 255, np.where(r < 1.1, 128, 0)).astype(np.uint8)
    flow = (rng.normal(0, 4, (h, w, 2)) + [3, 0]).astype(np.float32)
    flow[rng.random((h, w)) < 0.1] = np.nan
    host.append((a0, g0, tri, a1, g1, flow))
dev = [tuple(torch.from_numpy(v).cuda() for v in s) for s in host]
out = torch.empty(sets, 7, dtype=torch.float64, device="cuda")
for i, s in enumerate(dev):
    metrics.frame_sums(*s, out=out[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    for i, s in enumerate(dev):
        metrics.frame_sums(*s, out=out[i])
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * sets)
bytes_alg = h * w * (5 + 8)          # five uint8 images + the flow, each read once
t0 = time.perf_counter()
want = mo.frame_metrics(*host[0])
cpu_ms = (time.perf_counter() - t0) * 1e3
got = metrics.finish(out[0].cpu().tolist(), True)
# end to end from host arrays: upload + kernel + 56-byte read
pinned = [tuple(torch.from_numpy(v).pin_memory() for v in s) for s in host[:4]]
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in pinned:
    metrics.frame_metrics(*(v.cuda(non_blocking=True) for v in s))
e2e_ms = (time.perf_counter() - t0) * 1e3 / len(pinned)
print(json.dumps({"kernel_us": round(us, 2), "algorithmic_GB_s": round(bytes_alg / us / 1e3, 1), "e2e_ms_host_arrays": round(e2e_ms, 3),
                  "cpu_reference_port_ms": round(cpu_ms, 1), "max_rel_diff": max(
                      abs(got[k] - want[k]) / max(1.0, abs(want[k])) for k in want)}))
