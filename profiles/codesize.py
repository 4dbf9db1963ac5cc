"""SASS bytes of se3mpc_solve_kernel and of the device functions inside it (cuobjdump -elf symbol table)."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "learningagileflight_se3_b200/csrc/libse3mpc.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
seen = {}
for line in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+(\S*solve_kernel\S*)", line)
    if m and int(m.group(2), 16) > 0:
        name = m.group(3).split("$")[-1] if "$" in m.group(3) else "se3mpc_solve_kernel (whole)"
        mm = re.match(r"_ZN3se3(\d+)", name)
        if mm: name = name[len(mm.group(0)):len(mm.group(0)) + int(mm.group(1))]
        seen[name] = int(m.group(2), 16)
tot = seen.get("se3mpc_solve_kernel (whole)", 0)
for k, v in sorted(seen.items(), key=lambda kv: -kv[1]): print(f"{v:8d} B  {k}")
print(f"{tot - sum(v for k, v in seen.items() if 'whole' not in k):8d} B  kernel body (inlined ipm_solve, reward, init)")
