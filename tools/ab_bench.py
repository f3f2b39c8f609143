"""Same-box A/B of two builds of libcbx.so (boxes of the pool differ by +-3 %, a box drifts as it warms up: only alternating runs on
ONE box compare two kernels).

    python tools/ab_bench.py ab/libcbx_base.so ab/libcbx_new.so [--rounds 3] [--steps 20] [--tags dense_bottleneck_gemm,...] [-- bench args]

An arm may also be `lib@key=value,key=value` (library options passed as `--opt`), or just `@key=value` for the in-tree library:
    python tools/ab_bench.py @bn_ctas=2 @bn_ctas=3

Runs `bench.py --no-cpu-baseline --sustain 0` with CBX_LIB pointing at each library in turn, `rounds` times, and prints per run the
step time, clips/s and the per-step time of the named kernel tags (the library's own CUDA-event profile)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        i = argv.index("--")
        argv, extra = argv[:i], argv[i + 1:]
    libs, rounds, steps, tags = [], 3, 20, ["dense_bottleneck_gemm"]
    i = 0
    while i < len(argv):
        if argv[i] == "--rounds":
            rounds = int(argv[i + 1]); i += 2
        elif argv[i] == "--steps":
            steps = int(argv[i + 1]); i += 2
        elif argv[i] == "--tags":
            tags = argv[i + 1].split(","); i += 2
        else:
            libs.append(argv[i]); i += 1
    rows = {l: [] for l in libs}
    for r in range(rounds):
        for l in libs:
            path, _, opts = l.partition("@")
            env = dict(os.environ, CBX_LIB=os.path.abspath(path)) if path else dict(os.environ)
            optargs = [a for kv in opts.split(",") if kv for a in ("--opt", kv)]
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--sustain", "0", "--steps", str(steps), *optargs, *extra],
                                 env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print(f"{os.path.basename(l)} round {r}: FAILED\n{out.stderr[-2000:]}", flush=True)
                continue
            d = json.loads(out.stdout.strip().splitlines()[-1])
            k = d.get("kernels", {})
            fam = {}
            for name, v in k.items():
                fam[name.split(":")[0]] = fam.get(name.split(":")[0], 0.0) + v["ms_per_step"]
            fam.update({name: v["ms_per_step"] for name, v in k.items()})
            row = dict(ms=d["ms_per_step"], value=d["value"], e2e=d["e2e"]["value"], xv=d["parity"].get("max_abs_xv"), ve=d["parity"].get("max_abs_ve"),
                       **{t: fam.get(t) for t in tags})
            rows[l].append(row)
            print(f"{os.path.basename(l):24s} round {r}: {row['ms']:.3f} ms/step  {row['value']:.0f} clips/s  e2e {row['e2e']:.0f}  "
                  + "  ".join(f"{t} {row[t]:.3f}" if row[t] is not None else f"{t} -" for t in tags)
                  + f"  parity ve {row['ve']:.2e} xv {row['xv']:.2e}", flush=True)
    for l in libs:
        if rows[l]:
            n = len(rows[l])
            print(f"MEAN {os.path.basename(l):24s}: {sum(x['ms'] for x in rows[l]) / n:.3f} ms/step  "
                  + "  ".join(f"{t} {sum(x[t] for x in rows[l]) / n:.3f}" for t in tags if rows[l][0][t] is not None))


if __name__ == "__main__":
    main()
