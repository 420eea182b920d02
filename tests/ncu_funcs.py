"""Per-function share of warp instructions / stall samples of one kernel of an ncu report:
    python tests/ncu_funcs.py report.ncu-rep source.cu [kernel-name]"""
import csv, re, subprocess, sys, collections
rep, src = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else None
lines = open(src).read().split("\n")
func_at = {}
cur = "?"
for i, l in enumerate(lines, 1):
    mm = re.match(r"^(?:template.*\n)?(?:extern \"C\" )?__(?:device|global)__.*?\b(\w+)\s*\(", l)
    if mm and not l.startswith(" "):
        cur = mm.group(1)
    func_at[i] = cur
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kern:
    cmd += ["--kernel-name", kern]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if r and r[0] and len(r) == len(hdr):
        try:
            f = func_at.get(int(r[0]), "?")
        except Exception:
            continue
        a = agg[f]
        a["inst"] += float(r[ci] or 0); a["samples"] += float(r[cs] or 0)
        for s_ in stalls:
            try: a[s_] += float(r[hdr.index(s_)] or 0)
            except Exception: pass
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samples"] for a in agg.values())
print(f"total warp-instructions {ti:.3e}, samples {ts:.0f}")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((a[s_], s_) for s_ in stalls), reverse=True)[:3]
    print(f"{a['samples']/ts*100:5.1f}% samples {a['inst']/ti*100:5.1f}% inst  {f:24s} " + " ".join(f"{n.replace('stall_','')}={v/ts*100:.1f}%" for v, n in top))
