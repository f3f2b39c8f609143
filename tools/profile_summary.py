"""Turn gpurun_out/launches.csv + plain_bench.json + *.ncu-rep (tools/ncu_round.sh) into the tracked summaries under profiles/."""
import csv, collections, json, re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
with open(f'{root}/gpurun_out/launches.csv') as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
for x in r: rows.append(dict(zip(hdr, x)))
starts = [i for i, x in enumerate(rows) if 'trim_plan' in x['Kernel Name']]
per_step = starts[1] - starts[0]
w = 7
seg = rows[starts[w] - 1:starts[w] - 1 + per_step]
def short(n):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('cbx::', '')
    return re.sub(r'<.*', '', n)
agg = collections.OrderedDict(); tot = 0
for x in seg:
    k = short(x['Kernel Name']); v = float(x['Metric Value'].replace(',', '')); ms = v / 1e6
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
bench = json.loads(open(f'{root}/gpurun_out/plain_bench.json').read())
out = ["# Round 1, final build (tcgen05 TF32 mode): ncu launch list of one bench step (256 x 10 s clips, both encoders)\n",
       "Command (tools/ncu_round.sh): `python bench.py --steps 1 --warmup 3 --no-cpu-baseline > plain.log && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline`; the table is the 8th window of %d launches (= one step), summarised by tools/profile_summary.py." % per_step,
       "Per-launch times under ncu are cold-cache and serialised (no two-stream overlap, no programmatic dependent launch): compare SHARES.\n",
       f"Sum of kernel durations under ncu: {tot:.2f} ms over {len(seg)} launches; the same command without the profiler (one timed step): {bench['ms_per_step']:.2f} ms/step = {bench['value']:.0f} clips/s, e2e {bench['e2e']['value']:.0f} clips/s (5-step runs: profiles/r01_final_bench_1gpu.json).\n",
       "| kernel | launches | total ms | share |\n|---|---:|---:|---:|"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
out.append("\nbench.py's own per-kernel CUDA-event shares (same build, profiled pass: one stream, library tags):\n")
out.append("| tag | ms/step | share |\n|---|---:|---:|")
kt = sum(v['ms_per_step'] for v in bench['kernels'].values())
for k, v in bench['kernels'].items():
    out.append(f"| {k} | {v['ms_per_step']:.3f} | {100 * v['ms_per_step'] / kt:.1f}% |")
open(f'{root}/profiles/r01_final_launches.md', 'w').write("\n".join(out) + "\n")
print("\n".join(out[4:16]))
keep = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']
for n in ('r01b_bottleneck', 'r01b_promptmel'):
    raw = subprocess.run(['ncu', '-i', f'{root}/gpurun_out/{n}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    cols = [i for i, c in enumerate(h) if c in keep]
    csv.writer(open(f'{root}/profiles/{n}_ncu_full_summary.csv', 'w')).writerows([[h[i] for i in cols], [u[i] for i in cols]] + [[r_[i] for i in cols] for r_ in rr[2:]])
    for r_ in rr[2:]:
        d = dict(zip(h, r_))
        print(n, 'time', d['gpu__time_duration.sum'], 'dram MB', d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], 'tensor%', d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], 'regs', d['launch__registers_per_thread'], 'grid', d['Grid Size'], 'block', d['Block Size'])
