#!/bin/bash
# Sweep libcbx chunking options through bench.py (GPU box).  Usage: tools/sweep_opts.sh out.log "fcm_chunk_rows=2048 fcm_chunk_rows=4096 ..."
out=${1:-gpurun_out/sweep.log}
: > $out
for kv in $2; do
  echo "== $kv" >> $out
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --opt $kv > gpurun_out/_s.json 2>> $out && python tools/show_bench.py gpurun_out/_s.json 2>/dev/null | grep -E "^value|fcm_conv" >> $out
done
cat $out
