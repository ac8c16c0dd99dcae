#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA), UTCBAR (tcgen05.commit), plus registers and static
resources.  Runs here (no GPU needed):  python scripts/sass_summary.py > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pmt_learning_for_semantic_segmentation_and_disparity_b200", "libpmt_ops.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "FFMA",
            "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    counts[cur][p] += 1
                    break
    demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    cols = [p for p in PATTERNS if any(c[p] for c in counts.values())]
    print(f"# SASS summary of `{os.path.relpath(lib, ROOT)}` (cuobjdump -sass, sm_100a)\n")
    print("| kernel | instr | regs | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    for (mangled, c), name in zip(counts.items(), demangle):
        name = name.replace("pmt::(anonymous namespace)::", "").replace("pmt::", "")
        name = re.sub(r"^void ", "", re.sub(r"\((?!anonymous).*", "", name))
        r = regs.get(mangled, ("?", "?"))[0]
        print(f"| `{name}` | {c['_total']} | {r} | " + " | ".join(str(c[p]) if c[p] else "" for p in cols) + " |")


if __name__ == "__main__":
    main()
