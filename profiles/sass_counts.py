"""Instruction mix of the shipped wavefront kernels, from `cuobjdump -sass libsdtree.so` on stdin:
per kernel the instruction count, the opcode histogram, and the mnemonics the round's findings turn on
(ATOMS.CAST.SPIN = float shared-memory add as a compare-and-swap loop, ATOMS.POPC.INC / ATOMS.ADD = native integer add,
MATCH.ANY = warp aggregation, LDG.E.ENL2.256 = one 32 B record per level, PRMT = byte-coded child rank,
RED = fire-and-forget global add)."""
import collections
import re
import sys

cur, funcs = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
watch = ["ATOMS.CAST.SPIN", "ATOMS.POPC.INC", "ATOMS.ADD", "MATCH.ANY", "LDG.E.ENL2.256.CONSTANT", "PRMT", "RED.E.ADD.F32.FTZ.RN.STRONG.GPU", "REDG", "ATOMG", "BAR.SYNC.DEFER_BLOCKING"]
for name, ins in funcs.items():
    if "k_wavefront" not in name and "k_scan_fused" not in name:
        continue
    h = collections.Counter(i.split(".")[0] for i in ins)
    full = collections.Counter(ins)
    print(f"{name}: {len(ins)} instructions")
    print("   opcodes: " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
    print("   watched: " + ", ".join(f"{w} {sum(v for k, v in full.items() if k.startswith(w))}" for w in watch))
