"""Summarise an ncu report per CUDA source line:  python tests/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[3:]:
    if r and r[0]:
        try:
            data.append((float(r[ci]), float(r[cs]), r[0], r[1].strip()[:120]))
        except Exception:
            pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print(f"total warp-instructions {tot:.3e}, samples {tots:.0f}")
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{d[1]/tots*100:5.1f}% smp {d[0]/tot*100:5.1f}% inst  L{d[2]}: {d[3]}")
