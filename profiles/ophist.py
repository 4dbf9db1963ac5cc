"""Static opcode histogram of the device functions inside se3mpc_solve_kernel (nvdisasm of the embedded cubin)."""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else "learningagileflight_se3_b200/csrc/libse3mpc.so")
pat = sys.argv[-1] if len(sys.argv) > 1 and not sys.argv[-1].endswith(".so") else ""
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
i0 = re.search(r"\.text\.\S*se3mpc_solve_kernel", txt).start()
sec = txt[i0:]
nxt = re.search(r"\n//-+ \.text\.", sec[10:])
if nxt: sec = sec[:nxt.start() + 10]
parts = re.split(r"\n\$\S*solve_kernel\S*?\$(\S+):\n", sec)
names = ["kernel body"] + parts[1::2]; bodies = parts[0::2]
for name, body in zip(names, bodies):
    if pat and pat not in name: continue
    ops = collections.Counter()
    for l in body.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
        if m: ops[m.group(1)] += 1
    tot = sum(ops.values())
    if tot > 50: print(name[:48], tot, " ".join(f"{o}:{n}" for o, n in ops.most_common(24)))
