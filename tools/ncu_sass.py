"""Dump the SASS of the hottest region from an ncu report: per instruction, executed count and stall samples.
usage: ncu_sass.py report.ncu-rep [threshold 0..1] [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if len(sys.argv) > 3: cmd += ["-k", "regex:" + sys.argv[3], "-c", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ci = hdr.index("Instructions Executed"); si = hdr.index("Source"); wi = hdr.index("Warp Stall Sampling (All Samples)")
body = rows[2:]
tot = sum(int(r[ci] or 0) for r in body); tots = sum(int(r[wi] or 0) for r in body)
print("total warp instr", tot, "stall samples", tots)
mx = max(int(r[ci] or 0) for r in body)
for i, r in enumerate(body):
    n = int(r[ci] or 0)
    if n >= thr * mx:
        print(f"{i:5d} {n/tot*100:5.2f}% {int(r[wi] or 0)/max(1,tots)*100:5.2f}%  {r[si].strip()[:100]}")
