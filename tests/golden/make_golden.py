"""Generates tests/golden/*.json by running the REAL reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) -- run in the build container only:

    python tests/golden/make_golden.py

psa_example.json : the reference's example pair (example/psa/seq/*.fa) with the outputs the
                   reference produces for default and example-script flags (README.md:16-22).
psa_golden.json  : fuzz pairs (lengths straddling the tile edges 160/640, three divergence
                   modes, six parameter sets, several -W/-T) with maxsorce + md5 of the -o file.
"""
import hashlib
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import binding as ob  # noqa: E402
from tsta_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def fasta_first(path):
    lines = open(path, "rb").read().split(b"\n")
    return b"".join(l.strip() for l in lines[1:] if not l.startswith(b">")).decode()


def main():
    ref = Path("/root/reference/example/psa/seq")
    a, b = fasta_first(ref / "seqa1.fa"), fasta_first(ref / "seqb1.fa")
    cases = []
    for prm in [(2, -5, -2, -4), (2, -3, -2, -4)]:
        ms, out = ob.ref_psa(a.encode(), b.encode(), *prm, T=4, W=10)
        corner = ob.psa(a.encode(), b.encode(), *prm, trace=False).corner
        cases.append({"params": prm, "maxsorce": ms, "corner": corner, "out_len": len(out),
                      "out_md5": hashlib.md5(out).hexdigest()})
    json.dump({"seqa": a, "seqb": b, "cases": cases}, open(OUT / "psa_example.json", "w"))

    rnd = random.Random(20261018)
    params = [(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (5, -4, -1, -10),
              (1, -3, -1, -5), (2, -5, 0, 0), (10, -10, -5, -20)]
    lens = [1, 2, 5, 17, 100, 127, 128, 129, 159, 160, 161, 250, 255, 256, 257, 640, 641, 1000, 2049]
    gold = []
    for k in range(64):
        la, lb = rnd.choice(lens), rnd.choice(lens)
        mode = rnd.choice([None, 0.1, 0.02])
        x, y = synth.read_pair(la, lb, 1000 + k, mode)
        y = y or b"G"
        prm = rnd.choice(params)
        W, T = rnd.choice([1, 2, 3, 10, 16, 30]), rnd.choice([1, 3, 8])
        ms, out = ob.ref_psa(x, y, *prm, T=T, W=W)
        gold.append({"name": f"fuzz{k}_{la}x{lb}_{mode}", "x": x.decode(), "y": y.decode(), "params": prm,
                     "W": W, "T": T, "maxsorce": ms, "out_md5": hashlib.md5(out).hexdigest(),
                     "out_len": len(out)})
    json.dump(gold, open(OUT / "psa_golden.json", "w"))
    print("wrote", len(gold), "fuzz cases")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


# ------------------------------------------------------------------------------------------
# MSA goldens (appended): run with `python tests/golden/make_golden.py msa`
# ------------------------------------------------------------------------------------------
def msa_cases():
    """Deterministic MSA fuzz inputs: (name, reads, params, W)."""
    rnd = random.Random(424242)
    out = []
    params = [(2, -5, -2, -4), (2, -3, -2, -4), (1, -1, -1, 0), (2, -5, -2, 0), (1, -3, -1, -5)]
    for k in range(36):
        nreads = rnd.choice([2, 3, 5, 8])
        L0 = rnd.choice([50, 130, 300, 700, 1500])
        div = rnd.choice([0.02, 0.05, 0.15])
        anc = synth.random_read(L0, 3000 + k)
        reads = [anc]
        for r in range(1, nreads):
            x = synth.mutate(anc, div, 91 * k + r)
            mode = rnd.random()
            if mode < 0.3:      # extra prefix -> new head nodes (late heads exercise quirk Q1 at -W 1)
                x = synth.random_read(rnd.choice([1, 3, 10]), 7 * k + r) + x
            elif mode < 0.4:
                x = x[rnd.choice([1, 5, 20]):]
            elif mode < 0.5:
                x = x + synth.random_read(rnd.choice([1, 4]), 11 * k + r)
            x = x[:len(anc)] or b"A"      # the first read stays the longest (quirk Q2 is out of domain)
            reads.append(x)
        out.append((f"msa{k}", reads, rnd.choice(params), rnd.choice([1, 1, 2, 3, 10, 30])))
    return out


def run_ref_msa(reads, prm, W, isa=None, T=3):
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        fa, fo = os.path.join(d, "r.fa"), os.path.join(d, "o.txt")
        ob.write_fasta(fa, reads)
        p = ob.run_ref("TSTA_msa", ["-M", prm[0], "-X", prm[1], "-E", prm[2], "-O", prm[3], "-W", W, "-T", T,
                                    "-i", fa, "-o", fo], isa=isa)
        return open(fo, "rb").read(), p.stdout, p.stderr


def main_msa():
    ex = Path("/root/reference/example/msa/seq/seq1.fa").read_bytes().split(b"\n")
    reads = [l.strip() for l in ex if l.strip() and not l.startswith(b">")]
    gold = {"example_reads": [r.decode() for r in reads], "example": [], "fuzz": []}
    for prm, W in [((2, -5, -2, -4), 10), ((2, -3, -2, -4), 30)]:
        out, so, se = run_ref_msa(reads, prm, W, T=4)
        gold["example"].append({"params": prm, "W": W, "out_md5": hashlib.md5(out).hexdigest(), "out_len": len(out),
                                "stdout": so.decode(), "stderr": se.decode()})
    kept = 0
    for name, rds, prm, W in msa_cases():
        out, so, se = run_ref_msa(rds, prm, W, isa="avx512")
        # keep only cases inside the reference's exact domain: the same tile edge on another ISA agrees
        out2, _, _ = run_ref_msa(rds, prm, 2 * W, isa="avx2")
        if out != out2:
            print("skip (reference ISA builds disagree):", name)
            continue
        gold["fuzz"].append({"name": name, "reads": [r.decode() for r in rds], "params": prm, "W": W,
                             "out_md5": hashlib.md5(out).hexdigest(), "out_len": len(out)})
        kept += 1
    json.dump(gold, open(OUT / "msa_golden.json", "w"))
    print("wrote", kept, "msa fuzz cases")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "msa":
    main_msa()
