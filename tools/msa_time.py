"""Time TSTA_msa (ours) and the reference on a synthetic MSA input.
usage: python tools/msa_time.py NREADS LEN [ref]"""
import os, subprocess, sys, time, hashlib
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import binding as ob
from tsta_b200 import synth

n, ln = int(sys.argv[1]), int(sys.argv[2])
anc = synth.random_read(ln, 11)
reads = [anc] + [synth.mutate(anc, 0.05, 200 + k)[:ln] for k in range(n - 1)]
d = Path("/tmp/msa_time"); d.mkdir(exist_ok=True)
fa = d / "in.fa"; ob.write_fasta(fa, reads)
t0 = time.perf_counter()
p = subprocess.run([str(ROOT / "tsta_b200/_build/TSTA_msa"), "-i", str(fa), "-o", str(d / "ours.txt")], capture_output=True)
t1 = time.perf_counter() - t0
print("ours   %.3f s rc=%d %s" % (t1, p.returncode, p.stderr.decode().strip()[-400:]))
print("  md5", hashlib.md5((d / "ours.txt").read_bytes()).hexdigest())
if len(sys.argv) > 3:
    for W in (10, 30):
        t0 = time.perf_counter()
        p = ob.run_ref("TSTA_msa", ["-T", os.cpu_count(), "-W", W, "-i", fa, "-o", d / "ref.txt"], timeout=3000)
        t2 = time.perf_counter() - t0
        print("ref W=%d  %.3f s  md5 %s" % (W, t2, hashlib.md5((d / "ref.txt").read_bytes()).hexdigest()))
