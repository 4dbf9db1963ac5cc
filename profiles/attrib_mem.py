"""Per device function: L1 data-pipe load of an `ncu --page source --csv` export of se3mpc_solve_kernel -- shared-memory wavefronts
(actual / ideal), L1 tag requests and theoretical L2 sectors of global accesses, local-memory sectors.  Usage: attrib_mem.py src.csv lib.so"""
import collections, csv, re, subprocess, sys
src_csv, lib = sys.argv[1], sys.argv[2]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
first = open(src_csv).readline()
inst = re.search(r"solve_kernel<(?:\(int\))?(\d)", first)
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from divergence_audit import kernel_tag  # noqa: E402
tag = kernel_tag(elf, inst.group(1))
funcs = [(0, "solve_kernel body")]
for line in elf.split("\n"):
    if tag not in line: continue
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$\S*solve_kernel[A-Za-z0-9_]*\$(\S+)", line)
    if m:
        name = m.group(3); mm = re.match(r"_ZN3se3(\d+)", name)
        if mm: n = int(mm.group(1)); name = name[len(mm.group(0)):len(mm.group(0)) + n]
        funcs.append((int(m.group(1), 16), name))
funcs = sorted(set(funcs))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
cols = ["L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global", "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Ideal", "L2 Theoretical Sectors Local", "# Samples"]
acc = collections.defaultdict(collections.Counter)
top = []
for r in rows[2:]:
    off = int(r[0], 16) - base
    f = [n for o, n in funcs if off >= o][-1]
    for c in cols: acc[f][c] += int(r[ix[c]] or 0)
    top.append((int(r[ix["L1 Wavefronts Shared"]] or 0) + int(r[ix["L1 Tag Requests Global"]] or 0), f, r[1].strip(), r[ix["L1 Wavefronts Shared"]], r[ix["L1 Wavefronts Shared Ideal"]], r[ix["L1 Tag Requests Global"]], r[ix["Instructions Executed"]]))
tot = collections.Counter()
for f in acc:
    for c in cols: tot[c] += acc[f][c]
print("totals:", dict(tot))
print(f"{'function':28s} {'samples%':>8s} {'smem wf%':>9s} {'wf/ideal':>8s} {'gl tags%':>9s} {'L2 sect%':>9s} {'sect/ideal':>10s} {'local%':>7s}")
for f in sorted(acc, key=lambda f: -acc[f]["# Samples"]):
    a = acc[f]
    print(f"{f:28s} {100*a['# Samples']/tot['# Samples']:8.1f} {100*a[cols[0]]/max(1,tot[cols[0]]):9.1f} {a[cols[0]]/max(1,a[cols[1]]):8.2f} {100*a[cols[2]]/max(1,tot[cols[2]]):9.1f} "
          f"{100*a[cols[3]]/max(1,tot[cols[3]]):9.1f} {a[cols[3]]/max(1,a[cols[4]]):10.2f} {100*a[cols[5]]/max(1,tot[cols[5]]):7.1f}")
print("top instructions by L1 load (smem wavefronts + global tags):")
for t in sorted(top, reverse=True)[:40]:
    print(f"  {t[0]:>12d} {t[1]:22s} {t[2][:60]:60s} smem wf {t[3]:>11s} ideal {t[4]:>11s} gl tags {t[5]:>11s} exec {t[6]}")
