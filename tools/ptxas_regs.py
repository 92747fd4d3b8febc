#!/usr/bin/env python
"""tools/ptxas_regs.py LOG [filter]: registers / spills / smem per kernel from a `-Xptxas -v` log."""
import re, subprocess, sys
log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
names = re.findall(r"Compiling entry function '(\S+)'", log)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"ptxas info\s+: Compiling entry function ", log)[1:]
for name, blk in zip(dem, blocks):
    if flt and not re.search(flt, name):
        continue
    regs = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    smem = re.search(r"(\d+) bytes smem", blk)
    print(f"{int(regs.group(1)):4d} regs  spill {spill.group(1):>4s}/{spill.group(2):<4s} smem {smem.group(1) if smem else 0:>6}  {name[:150]}")
