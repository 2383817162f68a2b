#!/bin/bash
# ncu captures of round 2 (one gpurun call, ~6 GPU-min):  launch list of one window + --set full of one launch per kernel class.
# tools/profile_targets.py runs the window twice (call 1 records the plan, call 2 replays it): the captures skip the recording pass.
mkdir -p gpurun_out/ncu
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/ncu/r02_launches.csv python tools/profile_targets.py window > gpurun_out/ncu/launches.log 2>&1
cap() {   # name, kernel regex, launches of that kernel to skip, count, workload
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f \
      -o "gpurun_out/ncu/r02_$1" python tools/profile_targets.py "$5" > "gpurun_out/ncu/$1.log" 2>&1
}
# conv_tc2p launches of one window in plan order (profiles/r02_bench_calls.jsonl): #0 64->64 3x3 n3, #7 128->128 3x3 n3,
# #16 256->256 3x3 n3, #24 512->512 3x3 n3; 70 per window
cap conv_tc2p_64 "conv_tc2p_kernel" 70 1 window
cap conv_tc2p_128 "conv_tc2p_kernel" 77 1 window
cap conv_tc2p_256 "conv_tc2p_kernel" 86 1 window
cap conv_tc2p_512 "conv_tc2p_kernel" 94 1 window
cap gemm_tc2 "gemm_tc2_kernel" 4 2 window
cap gca_elementwise "gca_rowstats_kernel|gca_shift_add_kernel" 4 2 window
cap tam_head "tam_attend_kernel|head_conv_tanh01_kernel" 2 2 window
cap conv_tc3 "conv_tc3_kernel" 15 3 window
cap losses "loss" 0 6 losses
ls -la gpurun_out/ncu
