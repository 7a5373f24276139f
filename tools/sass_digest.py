"""SASS digest of the built library for profiles/: per kernel registers / shared memory (cuobjdump -res-usage) and
the instruction histogram of its SASS -- the evidence that the kernels are DPX code built for sm_100a only.
usage: python tools/sass_digest.py [lib.so] > profiles/rNN_sass_digest.md"""
import collections, re, subprocess, sys
from pathlib import Path
lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "tsta_b200/_build/libtsta_cuda.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m: cur = m.group(1); continue
    if cur and "REG:" in line: usage[cur] = line.strip(); cur = None
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("# SASS digest of %s\n" % Path(lib).name)
print("architectures in the fat binary: %s\n" % ", ".join(archs))
kern = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m: cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if cur and m: kern[cur][m.group(2)] += 1
tot = collections.Counter()
print("| kernel | instructions | registers / shared | most frequent SASS opcodes |\n|---|---|---|---|")
for name, ops in kern.items():
    n = sum(ops.values())
    if n < 40: continue
    tot.update(ops)
    top = ", ".join("%s %d" % kv for kv in ops.most_common(7))
    u = usage.get(name, "")
    u = " ".join(w for w in u.split() if w.startswith(("REG", "SHARED")))
    print("| `%s` | %d | %s | %s |" % (dem(name)[:110].replace("|", "/"), n, u, top))
print("\nDPX / packed integer opcodes over the whole library: " + ", ".join("%s %d" % (k, v) for k, v in tot.most_common() if k.startswith(("VIADDMNMX", "VIMNMX", "VIADD", "PRMT", "REDUX"))))
