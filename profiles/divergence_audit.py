"""Static audit of the solver kernel's SASS: per device function, how many instructions exist only because ptxas could not prove
the control flow warp-uniform (R2UR descriptor moves instead of LDCU, BRA.DIV guards and WARPSYNC.COLLECTIVE wrappers around
shuffles).  Usage: divergence_audit.py [libse3mpc.so] [instance 2|3]; tests/test_codegen.py asserts the invariants."""
import collections
import re
import subprocess
import sys

KEYS = ("total", "R2UR", "BRA.DIV", "WARPSYNC", "LDCU", "BSSY", "UMOV", "NOP")


def kernel_tag(elf_text, inst, redo=0):
    """mangled-name fragment of se3mpc_solve_kernel<inst, FULL, REDO = redo> as built (FULL depends on -DSE3_TWO_PASS)"""
    import re as _re
    names = sorted(set(_re.findall(r"solve_kernelILi%sELb[01]ELb%dE" % (inst, redo), elf_text)))
    return names[0] if names else "solve_kernelILi%sE" % inst


def audit(lib="learningagileflight_se3_b200/csrc/libse3mpc.so", inst="2"):
    """{function name: Counter of KEYS} for se3mpc_solve_kernel<inst> (the kernel body is 'kernel body')."""
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True, check=True).stdout
    tag = kernel_tag(elf, str(inst).rstrip("r"), 1 if str(inst).endswith("r") else 0)      # "2r": the second-pass instance of a -DSE3_TWO_PASS build
    funcs = [(0, "kernel body")]
    for line in elf.split("\n"):
        if tag not in line:
            continue
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$\S*solve_kernel[A-Za-z0-9_]*\$(\S+)", line)
        if m:
            name = m.group(3)
            mm = re.match(r"_ZN3se3(\d+)", name)
            if mm:
                name = name[len(mm.group(0)):len(mm.group(0)) + int(mm.group(1))]
            funcs.append((int(m.group(1), 16), name))
    funcs = sorted(set(funcs))
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.split("\n")
    start = [i for i, l in enumerate(sass) if "Function :" in l and tag in l][0]
    end = ([i for i, l in enumerate(sass) if "Function :" in l and i > start] + [len(sass)])[0]
    cnt = collections.defaultdict(collections.Counter)
    for l in sass[start:end]:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if not m:
            continue
        off, ins = int(m.group(1), 16), m.group(2)
        f = [n for o, n in funcs if off >= o][-1]
        op = [o for o in ins.split() if not o.startswith("@")][0]
        cnt[f]["total"] += 1
        for key in KEYS[1:]:
            if op.startswith(key):
                cnt[f][key] += 1
    return dict(cnt)


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else "learningagileflight_se3_b200/csrc/libse3mpc.so"
    inst = sys.argv[2] if len(sys.argv) > 2 else "2"
    tot = collections.Counter()
    for f, c in audit(lib, inst).items():
        tot.update(c)
        print(f[:28].ljust(28), " ".join(f"{k} {c[k]:5d}" for k in KEYS))
    print("ALL".ljust(28), " ".join(f"{k} {tot[k]:5d}" for k in KEYS))
