#!/usr/bin/env python
"""Counts the SASS mnemonics that prove which hardware path a kernel takes (tcgen05 = UTCHMMA + LDTM/STTM + UTCBAR, TMA tensor
maps = UTMALDG, bulk copies = UBLKCP, mbarriers = SYNCS, warp-level tensor cores = HMMA) per kernel of the built objects.
usage: python scripts/sass_counts.py > profiles/r5/sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "cartpoleplusplus_b200", "csrc", "build")
MN = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "FFMA"]
print("| object | kernel | " + " | ".join(MN) + " |")
print("|---|---|" + "---:|" * len(MN))
for obj in sorted(os.listdir(BUILD)):
  if not obj.endswith(".o"):
    continue
  txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
  fn, counts = None, collections.OrderedDict()
  for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      fn = m.group(1); counts[fn] = collections.Counter(); continue
    if fn is None:
      continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
      op = m.group(1)
      for k in MN:
        if op.startswith(k):
          counts[fn][k] += 1
  for fn, c in counts.items():
    if not any(c[k] for k in MN[:8]):
      continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("cpp::", "")
    print("| %s | `%s` | " % (obj, name) + " | ".join(str(c[k]) if c[k] else "" for k in MN) + " |")
