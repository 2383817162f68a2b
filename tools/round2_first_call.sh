#!/bin/bash
# Everything that was written after round 1's GPU budget was spent, in ONE gpurun call (~2 GPU-min):
#   gpurun --timeout 400 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
  echo "== pending GPU checks (FrameStream == windowed forward, FBA S=5/B=2, trimap_transform)"
  timeout 120 python -m pytest tests/test_gpu_z_stream.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5
  echo "== FrameStream frames/s vs windowed"
  timeout 90 python tools/stream_bench.py vmn_gca 1088 1920 12 2>&1 | tail -1
  timeout 90 python tools/stream_bench.py vmn_fba 1088 1920 8 2>&1 | tail -1
  echo "== headline window: default / N=64 tile heuristic (flag 256) / space-to-depth stride-2 layers"
  timeout 60 python tools/time_window.py 1088 1920 10 2>&1 | grep "ms/window"
  TCV_DEBUG_FLAGS=256 timeout 60 python tools/time_window.py 1088 1920 10 2>&1 | grep "ms/window"
  TCV_S2D_STRIDE2=1 timeout 60 python tools/time_window.py 1088 1920 10 2>&1 | grep "ms/window"
  echo "== parity with the space-to-depth layers on (golden windows + 1080p vs oracle)"
  TCV_S2D_STRIDE2=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
  echo "== FBA window"
  timeout 60 python tools/fba_bench.py 2>&1 | tail -1 | cut -c1-400
} > gpurun_out/round2_first_call.log 2>&1
cat gpurun_out/round2_first_call.log
