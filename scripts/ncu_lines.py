"""Stall samples per CUDA source line of one ncu report: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fp, agg, tot = None, [], 0
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]
        continue
    if r[0].isdigit():
        try:
            s, inst = int(r[4]), int(r[7])
        except (ValueError, IndexError):
            continue
        agg.append((s, inst, fp, int(r[0]), r[1].strip()[:100]))
        tot += s
print("total samples", tot, "instructions", sum(a[1] for a in agg))
for s, inst, f, l, src in sorted(agg, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}% inst={inst:9d} {f}:{l}  {src}")
