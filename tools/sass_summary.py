"""Per-kernel SASS evidence of the Blackwell-native instructions in libcbx.so (runs anywhere cuobjdump is: no GPU needed).
    python tools/sass_summary.py [round tag]  ->  profiles/<tag>_sass_summary.md
Counts, per kernel of the product library: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTMALDG / UTMASTG (TMA tensor
load / store), UTMAPF (TMA L2 prefetch) / UBLKCP (bulk copies), SYNCS (mbarrier), HMMA / IMMA (legacy mma.sync: must be 0), plus registers."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(root, "chatterbox_embed_b200", "libcbx.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", ln)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                                ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("HMMA/IMMA", r"\b[HI]MMA"),
                                ("MUFU", r"\bMUFU"), ("SHFL", r"\bSHFL"), ("REDUX", r"\bREDUX")])
counts, order = {}, []
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); order.append(cur); continue
    if cur is None or "/*" not in ln:
        continue
    counts[cur]["total"] += 1
    for k, p in pats.items():
        if re.search(p, ln):
            counts[cur][k] += 1
dem = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
def short(n):
    n = re.sub(r"^void ", "", n)
    n = n.replace("cbx::", "")
    n = re.sub(r"\(.*$", "", n)
    return n if len(n) < 110 else n[:107] + "..."
out = [f"# {tag}: SASS summary of chatterbox_embed_b200/libcbx.so (product build, sm_100a)\n",
       "`cuobjdump -sass libcbx.so`, instruction mnemonics counted per kernel (tools/sass_summary.py).  UTC*MMA = tcgen05.mma, LDTM / STTM = "
       "tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops.  HMMA/IMMA (legacy mma.sync) must be 0.\n",
       "| kernel | SASS instr | regs | " + " | ".join(pats) + " |", "|---|---:|---:|" + "---:|" * len(pats)]
rows = []
for mangled, d in zip(order, dem):
    c = counts[mangled]
    rows.append((-(c["UTC*MMA"] > 0), -c["total"], f"| `{short(d)}` | {c['total']} | {regs.get(mangled, ('?',))[0]} | " + " | ".join(str(c[k]) for k in pats) + " |"))
out += [r[2] for r in sorted(rows)]
n_tc = sum(1 for m in order if counts[m]["UTC*MMA"])
out.append(f"\n{len(order)} kernels, {n_tc} with tcgen05.mma; legacy mma.sync instructions in the library: {sum(counts[m]['HMMA/IMMA'] for m in order)}.")
path = os.path.join(root, "profiles", f"{tag}_sass_summary.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:3 + 2 + n_tc + 3]))
