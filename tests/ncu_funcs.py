"""Per-function share of warp instructions / stall samples:  python tests/ncu_funcs.py report.ncu-rep source.cu"""
import csv, re, subprocess, sys
rep, src = sys.argv[1], sys.argv[2]
lines = open(src).read().split("\n")
func_at = {}
cur = "?"
for i, l in enumerate(lines, 1):
    mm = re.match(r"^(?:template.*\n)?(?:extern \"C\" )?__(?:device|global)__.*?\b(\w+)\s*\(", l)
    if mm and not l.startswith(" "):
        cur = mm.group(1)
    func_at[i] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r in rows[3:]:
    if r and r[0]:
        try:
            f = func_at.get(int(r[0]), "?")
            a = agg.setdefault(f, [0.0, 0.0])
            a[0] += float(r[ci]); a[1] += float(r[cs])
        except Exception:
            pass
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-instructions {ti:.3e}")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]/ts*100:5.1f}% samples {a[0]/ti*100:5.1f}% inst  {f}")
