#!/bin/bash
# ncu captures of the kernels added late in round 2 (one gpurun call, ~5 GPU-min): training-step kernels and the DIM path.
mkdir -p gpurun_out/ncu
cd "$(dirname "$0")/.."
cap() {   # name, kernel regex, launches to skip, count, command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s "$skip" -c "$cnt" -f \
      -o "gpurun_out/ncu/r02_$name" "$@" > "gpurun_out/ncu/$name.log" 2>&1
}
# one 1088x1920 S=5 training step (tools/train_time.py: 2 warm-up steps + 1 timed step; skip the launches of the warm-up)
cap wgrad_rows "conv_wgrad_tc_kernel" 188 6 python tools/train_time.py 1080p 1
cap gca_train "gca_softmax_bwd_grid_kernel|gca_shift_add_kernel|gca_rowstats_kernel|transpose_planes_kernel" 12 6 python tools/train_time.py 1080p 1
cap dim "maxpool2_idx_kernel|maxunpool2_kernel|head_conv5_clamp01_kernel" 0 6 python -m pytest tests/test_gpu_dim.py -q -m gpu -p no:cacheprovider -k full_hd
ls -la gpurun_out/ncu
