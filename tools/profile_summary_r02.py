"""gpurun_out/<tag>/ (tools/ncu_r02.sh) -> tracked summaries under profiles/: launch-list shares, per-family `ncu --set full` CSV
summaries, profiles/traffic.json (DRAM bytes per launch of every captured family), the sustained line and its clock trace.
    python tools/profile_summary_r02.py [tag]"""
import collections, csv, json, os, re, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = f"{root}/gpurun_out/{tag}"
prof = f"{root}/profiles"
def short(n):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('cbx::', '')
    return re.sub(r'<.*', '', n)
out = []
if os.path.exists(f"{src}/launches.csv"):
    rows = []
    with open(f"{src}/launches.csv") as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines); hdr = next(r)
    for x in r:
        rows.append(dict(zip(hdr, x)))
    starts = [i for i, x in enumerate(rows) if 'trim_plan' in x['Kernel Name']]
    per_step = starts[1] - starts[0]
    seg = rows[starts[3] - 1:starts[3] - 1 + per_step]
    agg = collections.OrderedDict(); tot = 0
    for x in seg:
        k = short(x['Kernel Name']); ms = float(x['Metric Value'].replace(',', '')) / 1e6
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
    bench = json.loads(open(f"{src}/plain_bench.json").read())
    out += [f"# {tag}: ncu launch list of one bench step (256 x 10 s clips, both encoders, tcgen05 TF32 mode)\n",
            f"Command (tools/ncu_r02.sh): `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` run plain, then under `ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv`; the table is the 4th window of {per_step} launches (= one step).",
            "Per-launch times under ncu are cold-cache and serialised (no two-stream overlap, no programmatic dependent launch): compare SHARES.\n",
            f"Sum of kernel durations under ncu: {tot:.2f} ms over {len(seg)} launches; the same command without the profiler: {bench['ms_per_step']:.2f} ms/step = {bench['value']:.0f} clips/s, e2e {bench['e2e']['value']:.0f} clips/s.\n",
            "| kernel | launches | total ms | share |\n|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    out.append("\nbench.py's own per-kernel CUDA-event shares (same build, profiled pass: one stream, library tags):\n")
    out.append("| tag | ms/step | share | GB/s (algorithmic) | TFLOP/s (algorithmic) |\n|---|---:|---:|---:|---:|")
    kt = sum(v['ms_per_step'] for v in bench['kernels'].values())
    for k, v in bench['kernels'].items():
        out.append(f"| {k} | {v['ms_per_step']:.3f} | {100 * v['ms_per_step'] / kt:.1f}% | {v['gbs'] and round(v['gbs'])} | {v['tflops'] and round(v['tflops'], 1)} |")
    open(f"{prof}/{tag}_launches.md", "w").write("\n".join(out) + "\n")
keep = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_xu.sum']
traffic = {}
for f in sorted(os.listdir(src)):
    if not f.endswith(".ncu-rep"):
        continue
    n = f[:-8]
    raw = subprocess.run(['ncu', '-i', f"{src}/{f}", '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        print(n, "empty report"); continue
    h, u = rr[0], rr[1]
    cols = [i for i, c in enumerate(h) if c in keep]
    csv.writer(open(f"{prof}/{tag}_{n}_ncu_full_summary.csv", "w")).writerows([[h[i] for i in cols], [u[i] for i in cols]] + [[r_[i] for i in cols] for r_ in rr[2:]])
    for r_ in rr[2:]:
        d = dict(zip(h, r_)); un = dict(zip(h, u))
        def num(k):
            return float(d[k].replace(',', ''))
        def to_bytes(k):
            v = num(k); unit = un[k].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        t_us = num('gpu__time_duration.sum') * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(un['gpu__time_duration.sum'], 1)
        dram = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
        print(f"{n}: {short(d['Kernel Name'])} time {t_us:.1f} us  dram {dram / 1e6:.1f} MB ({dram / t_us / 1e3:.0f} GB/s)  tensor {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']}%  "
              f"dram% {d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']}  regs {d['launch__registers_per_thread']} grid {d['Grid Size']} block {d['Block Size']}")
        traffic.setdefault(n, []).append({"kernel": short(d['Kernel Name']), "dram_bytes_per_launch": dram, "gpu_time_us": t_us, "grid": d['Grid Size']})
if traffic:
    json.dump(traffic, open(f"{prof}/{tag}_traffic.json", "w"), indent=1)
for f in ("bench_sustained.json", "bench.json", "clock_trace.csv"):
    if os.path.exists(f"{src}/{f}"):
        shutil.copy(f"{src}/{f}", f"{prof}/{tag}_{f}")
