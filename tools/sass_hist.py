#!/usr/bin/env python
"""Opcode histogram of one kernel in libsdpcutsel.so (static SASS, cuobjdump).  python tools/sass_hist.py [lib] [kernel-substr]"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sdpcutsel-via-nn_b200/libsdpcutsel.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "k_mlp_i8ILi4"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ops = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
PIPE = {"fp64": ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"), "fma": ("IMAD", "FFMA", "FMUL", "FADD"),
        "alu": ("LOP3", "PRMT", "IADD3", "SHF", "FSEL", "SEL", "LEA", "ISETP", "VIMNMX", "VIADD", "IMNMX", "PLOP3", "MOV", "FMNMX")}
tot = sum(ops.values())
print("total", tot)
for p, pre in PIPE.items():
    print("%-5s %d" % (p, sum(v for k, v in ops.items() if k.startswith(pre))))
print(" ".join("%s:%d" % kv for kv in ops.most_common(45)))
