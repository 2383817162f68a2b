"""Training-step timing on one GPU: python tools/train_time.py [512|1080p] [steps]   (bench.py's run_train_section alone)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "512"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
shape = None if which == "512" else (1, 5, bench.H, bench.W, steps, bench.TRAIN_1080_GFLOP_PER_SAMPLE, "configs[3]")
args = argparse.Namespace()
from tcvom_b200 import _cabi
_cabi.lib().tcv_set_debug_flags(int(os.environ.get("TCV_DEBUG_FLAGS", "0")))   # measurement switches (csrc/tc_common.cuh)
calls = [0]


def barrier():
    """run_train_section calls this right before and right after its timed steps: with TCV_PROFILE_STEP=1 those are the
    cudaProfilerStart/Stop marks for `ncu --profile-from-start off`"""
    torch.cuda.synchronize()
    if os.environ.get("TCV_PROFILE_STEP") == "1":
        (torch.cuda.profiler.start if calls[0] == 0 else torch.cuda.profiler.stop)()
    calls[0] += 1


out = bench.run_train_section(args, 0, 1, dev, barrier, lambda ms: ms, shape=shape)
print(json.dumps({k: out[k] for k in ("ms_per_step", "gpu_launches_per_step", "loss", "peak_mem_gb")}))
