#!/usr/bin/env python
"""Developer tool: opcode histogram of a SASS line range (cuobjdump -sass output), grouped by issue pipe.
usage: sass_hist.py file.sass first_line last_line"""
import re, sys, collections
ALU = ("PRMT", "LOP3", "SHF", "IADD3", "IADD", "VIADD", "ISETP", "SEL", "HMNMX2", "VIMNMX", "FMNMX", "LEA", "MOV", "VABSDIFF", "POPC", "SGXT", "BMSK", "FSEL", "FSETP", "PLOP3", "VHMNMX", "IABS", "SHL", "SHR", "HSETP2", "I2IP")
FMA = ("HFMA2", "HADD2", "HMUL2", "HSET2", "IMAD", "FFMA", "FADD", "FMUL")
LSU = ("LDS", "STS", "LDG", "STG", "CCTL", "LDL", "STL", "ATOM", "RED")
def pipe(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in ALU: return "alu"
    if b in LSU: return "lsu"
    if b in ("LDC", "LDCU", "S2R", "S2UR"): return "adu"
    if b.startswith("U") or b.startswith("R2UR"): return "uni"
    return "other"
lines = open(sys.argv[1]).read().split("\n")
a, b = int(sys.argv[2]), int(sys.argv[3])
ops = collections.Counter(); pipes = collections.Counter()
for ln in lines[a - 1:b]:
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if not m: continue
    op = m.group(2); ops[op] += 1; pipes[pipe(op)] += 1
tot = sum(ops.values())
print("total", tot, dict(pipes))
for k, v in ops.most_common(): print(f"{v:5d} {pipe(k):5s} {k}")
