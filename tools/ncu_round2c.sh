#!/bin/bash
# ncu captures of the kernels added at the end of round 2: the evaluation-metric kernel and the 32-channel TAM backward.
mkdir -p gpurun_out/ncu
cd "$(dirname "$0")/.."
cap() {   # name, kernel regex, launches to skip, count, command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -s "$skip" -c "$cnt" -f \
      -o "gpurun_out/ncu/r02_$name" "$@" > "gpurun_out/ncu/$name.log" 2>&1
}
cap metrics "frame_metrics_kernel" 12 2 python tools/metrics_time.py
cap tam_bwd "tam_attend_bwd_kernel|tam_attend_kernel" 0 6 python -m pytest tests/test_gpu_operator_train.py -q -m gpu -p no:cacheprovider -k "tam_operator_train_mode"
ls -la gpurun_out/ncu
