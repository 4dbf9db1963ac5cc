"""Attribute the warp-stall samples / executed instructions of an `ncu --page source --csv` export of se3mpc_solve_kernel to
the device functions inside it (symbol offsets from `cuobjdump -elf libse3mpc.so`).  Usage: attrib.py prof_src.csv [libse3mpc.so]"""
import collections
import csv
import re
import subprocess
import sys

src_csv = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "learningagileflight_se3_b200/csrc/libse3mpc.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = [(0, "solve_kernel body (ipm loop, reward, init)")]
# the solver kernel is instantiated twice (se3mpc_solve_kernel<2>, <3>): pick the instance the capture is of
first = open(src_csv).readline()
inst = re.search(r"solve_kernel<(?:\(int\))?(\d)", first)
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from divergence_audit import kernel_tag  # noqa: E402
tag = kernel_tag(elf, inst.group(1)) if inst else "solve_kernel"      # <CTAS, FULL, REDO = false>
for line in elf.split("\n"):
    if tag not in line:
        continue
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$\S*solve_kernel[A-Za-z0-9_]*\$(\S+)", line)
    if m:
        name = m.group(3)
        mm = re.match(r"_ZN3se3(\d+)", name)
        if mm:
            n = int(mm.group(1)); name = name[len(mm.group(0)):len(mm.group(0)) + n]
        funcs.append((int(m.group(1), 16), name))
funcs = sorted(set(funcs))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
S = collections.Counter(); I = collections.Counter(); T = collections.Counter(); st = collections.defaultdict(collections.Counter)
ops = collections.Counter()
stcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    off = int(r[0], 16) - base
    f = [n for o, n in funcs if off >= o][-1]
    S[f] += int(r[ix["# Samples"]]); n = int(r[ix["Instructions Executed"]]); I[f] += n
    T[f] += int(r[ix["Thread Instructions Executed"]])
    for h in stcols: st[f][h] += int(r[ix[h]] or 0)
    op = [o for o in r[1].split() if not o.startswith("@")][0].split(".")[0]
    ops[op] += n
tot = sum(S.values()); toti = sum(I.values())
print(f"instructions executed: {toti}   samples: {tot}   SASS instructions: {len(rows) - 2}")
for f, _ in S.most_common():
    top = ", ".join(f"{h[6:]}:{v * 100 // max(1, S[f])}" for h, v in st[f].most_common(5))
    print(f"{f:44s} samples {S[f] * 100 / tot:5.1f}%  inst {I[f] * 100 / toti:5.1f}%  lanes {T[f] / max(1, I[f]):4.1f}   {top}")
print("opcode mix:", ", ".join(f"{o} {n * 100 / toti:.1f}%" for o, n in ops.most_common(14)))
