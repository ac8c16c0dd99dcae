#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS view) with nvdisasm -g line info: instructions executed and stall samples per CUDA source line.
usage: ncu_lines.py <ncu-rep> <kernel regex> <cubin> <mangled-name substring> [top]"""
import csv, re, subprocess, sys
from collections import defaultdict

rep, kre, cubin, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []  # (offset, file, line) for the chosen function
cur = None; infn = False
for l in sass:
    if l.startswith("\t.section\t.text."):
        infn = sub in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: lines.append((int(m.group(1), 16), cur))
assert len(lines) == len(data), (len(lines), len(data))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
inst = defaultdict(float); samp = defaultdict(float); ops = defaultdict(lambda: defaultdict(float))
for (off, loc), r in zip(lines, data):
    inst[loc] += f(r, "Instructions Executed"); samp[loc] += f(r, "# Samples")
    ops[loc][r[ix["Source"]].replace("@", " ").split()[0 if not r[ix["Source"]].startswith("@") else 1].split(".")[0]] += f(r, "Instructions Executed")
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
for loc in sorted(inst, key=lambda k: -samp[k])[:top]:
    o = ", ".join(f"{k}:{v/ti*100:.1f}" for k, v in sorted(ops[loc].items(), key=lambda kv: -kv[1])[:4])
    print(f"{loc[0]}:{loc[1]:<5d} inst {inst[loc]/ti*100:5.1f}%  samples {samp[loc]/ts*100:5.1f}%   {o}")
