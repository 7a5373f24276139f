"""The three executables are the reference's public interface: flags, stdout
and -o bytes must match.  CPU tests cover argument handling and the FASTA front-end; GPU tests
run the drivers end to end."""
import gzip
import hashlib
import json
import subprocess
from pathlib import Path

import pytest

from oracle import binding as ob
from tsta_b200 import synth
from conftest import GOLDEN

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tsta_b200" / "_build"


@pytest.fixture(scope="module")
def bins():
    from tsta_b200 import build
    build.build_all()
    return BIN


def run(exe, *args, timeout=300):
    return subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, timeout=timeout)


def fnv(b: bytes) -> str:
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.mark.parametrize("name", ["TSTA_psa", "TSTA_psa_notrace"])
def test_usage_and_missing_input_text(bins, name):
    ours = run(bins / name)
    assert ours.returncode == 0
    assert ours.stdout.startswith(b"input file is not specified\n-M ")
    bad = run(bins / name, "-Q")
    assert bad.returncode == 0 and b"-M                      the sorce of match [default: 2]" in bad.stdout
    if ob.ref_bin(name) is not None:      # byte-identical to the real reference when it is built
        ref = run(ob.ref_bin(name))
        assert ours.stdout == ref.stdout
        assert run(ob.ref_bin(name), "-Q").stdout == bad.stdout


def test_fasta_front_end(bins, tmp_path):
    recs = [synth.random_read(1000, 1), b"acgtNNNNacgt" * 7, synth.random_read(77, 3)]
    plain = tmp_path / "a.fa"
    ob.write_fasta(plain, recs, width=60)
    crlf = tmp_path / "b.fa"
    crlf.write_bytes(plain.read_bytes().replace(b"\n", b"\r\n"))
    gz = tmp_path / "c.fa.gz"
    gz.write_bytes(gzip.compress(plain.read_bytes()))
    blank = tmp_path / "d.fa"
    blank.write_bytes(b"\n\n>r1 some comment\n" + recs[0][:500] + b"\n\n" + recs[0][500:] + b"\n>r2\n" + recs[1] + b"\n>r3\n" + recs[2])
    want = "".join("%d %d %s\n" % (k, len(r), fnv(r)) for k, r in enumerate(recs))
    for f in (plain, crlf, gz, blank):
        out = run(bins / "tsta_fasta_dump", f)
        assert out.returncode == 0 and out.stdout.decode() == want, f.name
    first = run(bins / "tsta_fasta_dump", plain, 1).stdout.decode()
    assert first == "0 %d %s\n" % (len(recs[0]), fnv(recs[0]))
    missing = run(bins / "tsta_fasta_dump", tmp_path / "nope.fa")
    assert missing.returncode == 1 and b"does not exist" in missing.stderr
    fq = tmp_path / "e.fq"
    fq.write_bytes(b"@r\nACGT\n+\nIIII\n")
    assert run(bins / "tsta_fasta_dump", fq).returncode == 1


@pytest.mark.gpu
def test_psa_drivers_on_example_pair(bins, tmp_path, gpu):
    g = json.load(open(GOLDEN / "psa_example.json"))
    ob.write_fasta(tmp_path / "a.fa", [g["seqa"].encode()], width=0)
    ob.write_fasta(tmp_path / "b.fa", [g["seqb"].encode()], width=70)
    for case in g["cases"]:
        M, X, E, O = case["params"]
        flags = ["-M", M, "-X", X, "-E", E, "-O", O, "-T", 30, "-W", 30, "-1", tmp_path / "a.fa", "-2", tmp_path / "b.fa"]
        out = tmp_path / "o.txt"
        r = run(bins / "TSTA_psa", *flags, "-o", out)
        assert r.returncode == 0, r.stderr
        assert r.stdout == b"maxsorce=%d\n" % case["maxsorce"]
        assert hashlib.md5(out.read_bytes()).hexdigest() == case["out_md5"]
        r = run(bins / "TSTA_psa_notrace", *flags)
        assert r.stdout == b"maxsorce=%d\n" % case["maxsorce"]


@pytest.mark.gpu
def test_psa_driver_swaps_and_default_output(bins, tmp_path, gpu):
    x, y = synth.read_pair(300, 450, 8, None)      # second file longer -> becomes ">1"
    ob.write_fasta(tmp_path / "a.fa", [x, b"IGNORED"])
    ob.write_fasta(tmp_path / "b.fa", [y])
    r = subprocess.run([str(bins / "TSTA_psa"), "-1", "a.fa", "-2", "b.fa"], capture_output=True, cwd=tmp_path)
    o = ob.psa(x, y)
    assert r.stdout == b"maxsorce=%d\n" % o.maxsorce
    assert (tmp_path / "output.txt").read_bytes() == o.output_bytes()


def test_batch_driver_argument_errors(bins, tmp_path):
    r = run(bins / "TSTA_psa_batch")
    assert r.returncode == 0 and r.stdout.startswith(b"input file is not specified\n")
    ob.write_fasta(tmp_path / "a.fa", [b"ACGT", b"ACG"])
    ob.write_fasta(tmp_path / "b.fa", [b"ACGT"])
    r = run(bins / "TSTA_psa_batch", "-1", tmp_path / "a.fa", "-2", tmp_path / "b.fa")
    assert r.returncode == 1 and b"2 records" in r.stderr and b"1 in" in r.stderr


@pytest.mark.gpu
def test_batch_driver_equals_one_process_per_pair(bins, tmp_path, gpu):
    """TSTA_psa_batch on two multi-record files == TSTA_psa run once per pair."""
    import random
    rnd = random.Random(4)
    xs, ys = [], []
    for k in range(40):
        la, lb = rnd.choice([1, 30, 250, 257, 700]), rnd.choice([1, 30, 250, 257, 700])
        x, y = synth.read_pair(la, lb, 300 + k, rnd.choice([None, 0.05]))
        xs.append(x)
        ys.append(y or b"C")
    ob.write_fasta(tmp_path / "a.fa", xs, width=60)
    with gzip.open(tmp_path / "b.fa.gz", "wb") as f:
        for k, y in enumerate(ys):
            f.write(b">r%d\n" % k + y + b"\n")
    out = tmp_path / "batch.txt"
    r = run(bins / "TSTA_psa_batch", "-1", tmp_path / "a.fa", "-2", tmp_path / "b.fa.gz", "-o", out)
    assert r.returncode == 0, r.stderr
    want_stdout, want_file = b"", b""
    for x, y in zip(xs, ys):
        o = ob.psa(x, y)
        want_stdout += b"maxsorce=%d\n" % o.maxsorce
        want_file += o.output_bytes() + b"\n"
    assert r.stdout == want_stdout
    assert out.read_bytes() == want_file
    # score-only, linear gaps, and the multi-GPU flag (clamped to the GPUs present)
    r = run(bins / "TSTA_psa_batch", "-n", "-O", "0", "-G", "8", "-1", tmp_path / "a.fa", "-2", tmp_path / "b.fa.gz")
    assert r.returncode == 0, r.stderr
    assert r.stdout == b"".join(b"maxsorce=%d\n" % ob.psa(x, y, O=0, trace=False).maxsorce for x, y in zip(xs, ys))
