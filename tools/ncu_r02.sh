#!/bin/bash
# Round-2 profiles (GPU box): sustained bench line with a clock trace, ncu launch list, one `ncu --set full` capture per kernel
# family.  Reports land in gpurun_out/r02/; tools/profile_summary_r02.py turns them into the tracked summaries under profiles/.
#   bash tools/ncu_r02.sh [tag] [families...]      families default: all
TAG=${1:-r02}; shift
FAMS=${@:-"bottleneck transit lstm_rec fcm_conv fcm_block dftmel pgemm tdnn stats_pool trim local_conv cam_gate conv1"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain 0"
# 1. sustained line (>= 300 steps, ~6 s timed) with an nvidia-smi trace beside it
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > $OUT/clock_trace.csv &
SMI=$!
python bench.py --steps 300 --warmup 5 --no-cpu-baseline --sustain 0 > $OUT/bench_sustained.json 2> $OUT/bench_sustained.err; echo "sustained rc=$?"
kill $SMI
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
tail -1 $OUT/plain.log > $OUT/plain_bench.json
# 2. launch list of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches.csv)"
# 3. --set full per family: name regex, launches to skip (past the warm-up steps), count
run() {
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/$1 $CMD > $OUT/ncu_$1.log 2>&1
  echo "$1: rc=$? $(ls -la $OUT/$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
for f in $FAMS; do
  case $f in
    # tgemm_bnrelu_kernel launches per step: block 1 = 12 bottlenecks + transit 1, block 2 = 24 + transit 2, block 3 = 16 + transit 3 (55);
    # the 4th step starts at launch 165: 200, 201 = block-2 layers with cin = 960 / 992; 202 = transit 2 (K = 1024, N = 512)
    bottleneck) run bottleneck tgemm_bnrelu_kernel 200 2 ;;
    transit)    run transit tgemm_bnrelu_kernel 202 1 ;;
    lstm_rec)   run lstm_rec lstm_rec_tc2 10 1 ;;
    fcm_conv)   run fcm_conv fcm_conv_kernel 85 2 ;;
    fcm_block)  run fcm_block fcm_block_kernel 24 2 ;;
    dftmel)     run dftmel dftmel 6 2 ;;
    cmn)        run cmn cmn_mean_kernel 3 1 ;;
    pgemm)      run pgemm pgemm_kernel 10 2 ;;
    tdnn)       run tdnn ^tgemm_kernel 3 1 ;;
    stats_pool) run stats_pool stats_pool_kernel 3 1 ;;
    trim)       run trim trim_plan_kernel 3 1 ;;
    local_conv) run local_conv local_conv_kernel 200 1 ;;
    cam_gate)   run cam_gate cam_gate_clip_kernel 200 1 ;;
    conv1)      run conv1 fcm_conv1_rows_kernel 13 1 ;;
  esac
done
