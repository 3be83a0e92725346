#!/usr/bin/env python
"""Per-source-line hot spots of an `ncu --set full --import-source on` capture (needs -lineinfo):
usage: python scripts/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
lines, cur = [], None
fname = ""
for r in rows:
  if len(r) >= 2 and r[0] == "File Path":
    fname = r[1].split("/")[-1]
  if len(r) < 8 or r[0] == "Line No":
    continue
  if r[0] != "":
    try:
      cur = dict(file=fname, line=int(r[0]), src=r[1].strip(), samples=int(r[4] or 0), inst=int(r[7] or 0))
      lines.append(cur)
    except ValueError:
      cur = None
tot_s = sum(l["samples"] for l in lines) or 1
tot_i = sum(l["inst"] for l in lines) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
print("---- by stall samples")
for l in sorted(lines, key=lambda l: -l["samples"])[:top]:
  print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * l["samples"] / tot_s, 100.0 * l["inst"] / tot_i, l["file"], l["line"], l["src"][:110]))
print("---- by instructions")
for l in sorted(lines, key=lambda l: -l["inst"])[:top]:
  print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * l["inst"] / tot_i, 100.0 * l["samples"] / tot_s, l["file"], l["line"], l["src"][:110]))
