"""Per-function stall-reason breakdown (CUDA source view):  python tests/ncu_stalls.py report.ncu-rep source.cu [stall ...]"""
import csv, re, subprocess, sys, collections
rep, src = sys.argv[1], sys.argv[2]
want = sys.argv[3:] or ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_barrier", "stall_branch_resolving", "stall_mio", "stall_lg"]
lines = open(src).read().split("\n")
func_at = {}; cur = "?"
for i, l in enumerate(lines, 1):
    mm = re.match(r"^(?:template.*)?(?:extern \"C\" )?__(?:device|global)__.*?\b(\w+)\s*\(", l)
    if mm and not l.startswith(" "): cur = mm.group(1)
    func_at[i] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
hdr = rows[hi]
cols = {w: hdr.index(w) for w in want}
cs = hdr.index("# Samples"); ci = hdr.index("Instructions Executed")
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[hi + 1:]:
    if r and r[0]:
        try: f = func_at.get(int(r[0]), "?")
        except Exception: continue
        a = agg[f]
        try:
            a["samples"] += float(r[cs]); a["inst"] += float(r[ci])
            for w, c in cols.items(): a[w] += float(r[c] or 0)
        except Exception: pass
ts = sum(a["samples"] for a in agg.values())
print("function".ljust(22), "samples%", " ".join(w.replace("stall_", "")[:9].rjust(9) for w in want), "(% of all samples)")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:28]:
    print(f.ljust(22), f"{a['samples']/ts*100:7.1f} ", " ".join(f"{a[w]/ts*100:9.2f}" for w in want))
tot = collections.Counter()
for a in agg.values(): tot.update(a)
print("TOTAL".ljust(22), f"{100:7.1f} ", " ".join(f"{tot[w]/ts*100:9.2f}" for w in want))
