"""Static loop report of a kernel's SASS (`cuobjdump -sass -fun <mangled> lib.so > f.sass`):
every backward branch closes a loop; prints each loop's size and opcode-class histogram."""
import re
import sys
from collections import Counter

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr_index = {a: k for k, (a, _) in enumerate(ins)}


def cls(op):
    op = re.sub(r"^@!?U?P\d+\s+", "", op)
    name = op.split()[0].split(".")[0]
    if name in ("DFMA", "DMUL", "DADD", "DSETP"):
        return "FP64:" + name
    if name in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDSM", "ATOMS", "ATOMG", "RED", "LDC", "LDCU"):
        return "MEM:" + name
    if name.startswith("MUFU"):
        return "MUFU"
    if name in ("SHFL", "BAR", "WARPSYNC", "BRA", "BSSY", "BSYNC", "EXIT", "NOP", "VOTE", "VOTEU", "CALL", "RET"):
        return "CTL:" + name
    return "INT/MOV:" + name


loops = []
for k, (a, op) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+.*?(0x[0-9a-f]+)", op)
    if m:
        t = int(m.group(1), 16)
        if t <= a and t in addr_index:
            loops.append((addr_index[t], k))
print(f"{len(ins)} instructions, {len(loops)} loops")
for lo, hi in loops:
    body = ins[lo:hi + 1]
    c = Counter(cls(op) for _, op in body)
    fp64 = sum(v for k2, v in c.items() if k2.startswith("FP64"))
    mem = sum(v for k2, v in c.items() if k2.startswith("MEM"))
    print(f"\nloop {ins[lo][0]:#06x}..{ins[hi][0]:#06x}: {len(body)} instr, FP64 {fp64} ({100*fp64/len(body):.0f}%), MEM {mem}")
    if len(body) >= int(sys.argv[2]) if len(sys.argv) > 2 else 40:
        print("   ", ", ".join(f"{k2}={v}" for k2, v in sorted(c.items(), key=lambda kv: -kv[1])))
