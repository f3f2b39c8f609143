#!/bin/bash
# Profiles of the final build (GPU box): launch list of the bench + `ncu --set full` of the top kernel and the prompt-mel kernel.
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log > gpurun_out/plain_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
ncu --set full --clock-control none --import-source on -k regex:tgemm_bnrelu_kernel -s 200 -c 2 -f -o gpurun_out/r01b_bottleneck $CMD > gpurun_out/ncu_bottleneck.log 2>&1
echo "bottleneck rc=$? $(ls -la gpurun_out/r01b_bottleneck.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
python tools/pm_time.py > gpurun_out/pm_plain.log 2>&1 || { echo "pm_time failed"; tail -3 gpurun_out/pm_plain.log; }
tail -1 gpurun_out/pm_plain.log
ncu --set full --clock-control none --import-source on -k regex:promptmel_kernel -s 5 -c 1 -f -o gpurun_out/r01b_promptmel python tools/pm_time.py > gpurun_out/ncu_promptmel.log 2>&1
echo "promptmel rc=$? $(ls -la gpurun_out/r01b_promptmel.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
