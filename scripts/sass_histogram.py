#!/usr/bin/env python
"""Opcode histogram per kernel of libamira_b200.so (cuobjdump -sass), written to profiles/sass_opcodes_r2.txt.
The mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), DFMA / DADD / DMUL (the fp64 FFT), LDGSTS (cp.async), REDG / ATOMG."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "amira-rust-asr-server_b200", "libamira_b200.so")
KEY = ["UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "DFMA", "DADD", "DMUL", "F2F", "LDGSTS", "SHFL", "LDS", "STS", "LDG",
       "STG", "ATOMG", "REDG", "MEMBAR", "FFMA", "MUFU", "LDL", "STL", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("amira::", "")
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
    lines = ["# SASS opcode histogram per kernel of libamira_b200.so (sm_100a), scripts/sass_histogram.py", "# " + " ".join(f"{k:>8s}" for k in ["total"] + KEY)]
    for name, c in kernels.items():
        lines.append(f"{name}")
        lines.append("  " + " ".join(f"{v:8d}" for v in [sum(c.values())] + [c.get(k, 0) for k in KEY]))
    text = "\n".join(lines) + "\n"
    path = os.path.join(ROOT, "profiles", "sass_opcodes_r2.txt")
    open(path, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
