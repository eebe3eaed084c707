"""Per CUDA source line: share of executed warp instructions (ncu report captured with --import-source on).
usage: ncu_lines.py report.ncu-rep [topn] [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kre = sys.argv[3] if len(sys.argv) > 3 else None
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre: cmd += ["-k", "regex:" + kre, "-c", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = ""; ci = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": ci = r.index("Instructions Executed"); continue
    if ci is None or len(r) <= ci: continue
    if r[0].strip().isdigit() and r[ci].strip().isdigit():
        lines.append((int(r[ci]), cur_file, int(r[0]), r[1].strip()))
tot = sum(l[0] for l in lines)
print("total warp instr", tot)
if "--by-line" in sys.argv:
    lines.sort(key=lambda l: (l[1], l[2]))
    for n, f, ln, src in lines:
        if n: print(f"{n / tot * 100:6.2f}%  {f}:{ln:<4d} {src[:110]}")
else:
    lines.sort(key=lambda l: -l[0])
    for n, f, ln, src in lines[:topn]:
        print(f"{n / tot * 100:6.2f}%  {f}:{ln:<4d} {src[:110]}")
