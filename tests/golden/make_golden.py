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


if __name__ == "__main__":
    main()
