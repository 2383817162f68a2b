#!/bin/bash
# ncu captures of round 2 (one gpurun call):  launch list of one window + --set full of one launch per kernel class.
mkdir -p gpurun_out/ncu
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/ncu/r02_launches.csv python tools/profile_targets.py window > gpurun_out/ncu/launches.log 2>&1
cap() {   # name, kernel regex, launches to skip, count, workload
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/ncu/r02_$1" \
      python tools/profile_targets.py "$5" > "gpurun_out/ncu/$1.log" 2>&1
}
# replayed plan: skip the recording call's launches of the same kernel (s = launches per window of that kernel)
cap conv_tc2p "conv_tc2p_kernel" 58 6 window
cap gemm_tc2 "gemm_tc2_kernel" 4 4 window
cap gca_elementwise "gca_rowstats_kernel|gca_shift_add_kernel|gca_prep_grid|gca_values_parity|gca_unfold_parity" 10 10 window
cap tam_head "tam_attend_kernel|head_conv_tanh01_kernel" 2 2 window
cap conv_tc3 "conv_tc3_kernel" 11 4 window
cap losses "loss" 0 6 losses
ls -la gpurun_out/ncu
