#!/bin/bash
# Sweep libcbx chunking options through bench.py (GPU box).  Usage: tools/sweep_opts.sh out.log
out=${1:-gpurun_out/sweep.log}
: > $out
for xv in 36000 300000; do for fcm in 4096 16384 65536; do for lstm in 2048 4096; do
  echo "== xv_chunk_rows=$xv fcm_chunk_rows=$fcm lstm_chunk_partials=$lstm" >> $out
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt xv_chunk_rows=$xv --opt fcm_chunk_rows=$fcm --opt lstm_chunk_partials=$lstm > gpurun_out/_s.json 2>> $out && python tools/show_bench.py gpurun_out/_s.json >> $out
done; done; done
