"""Opcode stream of riccati_factor's stage loop inside se3mpc_solve_kernel<2> (nvdisasm of the embedded cubin): shows what ptxas managed
to interleave with the DMMAs.  Usage: factor_sass.py [lib.so] [function substring]"""
import os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "learningagileflight_se3_b200/csrc/libse3mpc.so")
fn = sys.argv[2] if len(sys.argv) > 2 else "riccati_factorENS"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec = txt[re.search(r"\.text\.\S*se3mpc_solve_kernelILi2", txt).start():]
parts = re.split(r"\n\$\S*solve_kernel\S*?\$(\S+):\n", sec)
for n, b in zip(["body"] + parts[1::2], parts[0::2]):
    if fn in n:
        ops = [re.match(r"\s+/\*([0-9a-f]+)\*/\s+((?:@!?U?P\d\s+)?[A-Z0-9_.]+)", l).group(2) for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]+\*/", l)]
        dm = [i for i, o in enumerate(ops) if o.startswith("DMMA")]
        print(n, len(ops), "instructions,", len(dm), "DMMA")
        print(" ".join(ops[max(dm[0] - 30, 0):]) if dm else " ".join(ops))
