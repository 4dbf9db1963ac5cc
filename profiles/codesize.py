"""SASS bytes of se3mpc_solve_kernel<CTAS> and of the device functions inside it (cuobjdump -elf symbol table).
Usage: codesize.py [libse3mpc.so] [instance: 2 | 3]"""
import re, subprocess, sys
args = sys.argv[1:]
lib = next((a for a in args if a.endswith(".so")), "learningagileflight_se3_b200/csrc/libse3mpc.so")
inst = next((a for a in args if a.isdigit()), "2")
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from divergence_audit import kernel_tag  # noqa: E402
tag = kernel_tag(elf, inst)      # the throughput instance <CTAS, FULL, REDO = false>
seen = {}
for line in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+(\S*" + tag + r"\S*)", line)
    if m and int(m.group(2), 16) > 0:
        name = m.group(3).split("$")[-1] if "$" in m.group(3) else "(whole kernel)"
        mm = re.match(r"_ZN3se3(\d+)", name)
        if mm: name = name[len(mm.group(0)):len(mm.group(0)) + int(mm.group(1))]
        seen[name] = int(m.group(2), 16)
tot = seen.pop("(whole kernel)", 0)
if not tot:      # the kernel's own symbol has no '$': take the size of its .text section
    m = re.search(r"\s([0-9a-f]+)\s+0\s+80\s+PROGBITS\s+\S+\s+\S+\s+\S+\s+\.text\.\S*" + tag, elf)
    tot = int(m.group(1), 16) if m else 0
for k, v in sorted(seen.items(), key=lambda kv: -kv[1]): print(f"{v:8d} B  {k}")
print(f"{tot - sum(seen.values()):8d} B  kernel body (inlined ipm_solve, reward, init)")
print(f"{tot:8d} B  se3mpc_solve_kernel<{inst}> total")
