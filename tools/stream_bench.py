"""Frames/s of tcvom_b200.FrameStream (per-frame feature reuse) next to the windowed EvalModel.forward on
the same clip:  python tools/stream_bench.py [vmn_gca|vmn_fba] [H W [frames]]   -> one JSON line"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tcvom_b200
from tcvom_b200 import synthetic
from helpers import fixture_sd, fixture_sd_fba

arch = sys.argv[1] if len(sys.argv) > 1 else "vmn_gca"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1088, 1920)
T = int(sys.argv[4]) if len(sys.argv) > 4 else 12
m = tcvom_b200.EvalModel(model=arch, agg_window=7)
m.NET.load_state_dict(fixture_sd() if arch == "vmn_gca" else fixture_sd_fba(), strict=True)
m = m.cuda().eval()
imgs, tris = synthetic.make_window(H, W, seed=7, frames=T)
imgs, tris = torch.from_numpy(imgs).cuda(), torch.from_numpy(tris).cuda()
with torch.no_grad():
    stream = tcvom_b200.FrameStream(m, H, W, u8=True)
    for t in range(3):
        stream.push(imgs[0, t], tris[0, t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for t in range(3, T):
        out = stream.push(imgs[0, t], tris[0, t])
    e1.record(); torch.cuda.synchronize()
    ms_stream = e0.elapsed_time(e1) / (T - 3)
    win = [(imgs[:, t - 1:t + 2].contiguous(), tris[:, t - 1:t + 2].contiguous()) for t in range(1, T - 1)]
    m(*win[0]); m(*win[1])
    torch.cuda.synchronize()
    e0.record()
    for a, b in win[2:]:
        ref = m(a, b)
    e1.record(); torch.cuda.synchronize()
    ms_win = e0.elapsed_time(e1) / (len(win) - 2)
    a = out[0] if arch == "vmn_fba" else out
    r = ref[0] if arch == "vmn_fba" else ref
    err = float((a - r[0, 1]).abs().max())
print(json.dumps(dict(arch=arch, size=[H, W], ms_per_frame_stream=ms_stream, frames_per_s_stream=1e3 / ms_stream,
                      ms_per_window_windowed=ms_win, frames_per_s_windowed=1e3 / ms_win, speedup=ms_win / ms_stream,
                      last_frame_max_abs_diff=err, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))
