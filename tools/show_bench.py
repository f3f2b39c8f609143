import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} clips/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  clocks {d['clocks']}")
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items()})
tot = sum(v["ms_per_step"] for v in d["kernels"].values())
print(f"sum of kernel ms/step {tot:.2f}")
for k, v in d["kernels"].items():
    tf = v["tflops"]
    gb = v.get("gbs")
    print(f"  {k:28s} {v['ms_per_step']:9.3f} ms {100*v['ms_per_step']/tot:5.1f}%  n={v['launches_per_step']:6.0f}  {'' if tf is None else f'{tf:7.1f} TF/s'}  {'' if gb is None else f'{gb:7.0f} GB/s'}")
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
