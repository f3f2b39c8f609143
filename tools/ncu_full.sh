#!/bin/bash
# One `ncu --set full` capture per hot kernel family (GPU box).  Reports land in gpurun_out/*.ncu-rep.
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
run() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "$1: rc=$? $(ls -la gpurun_out/$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
run r01_lstm_rec lstm_rec_tc2 9 1
run r01_bottleneck tgemm_bnrelu 200 2
run r01_fcm_conv fcm_conv_kernel 109 2
run r01_dftmel dftmel 6 2
