"""Parity at the sizes BASELINE.json states: the CUDA path
against the score-only oracle (maxsorce, corner, the WHOLE last row) and against the REAL
reference binaries of oracle/_ref (-o bytes of traced 100 kbp / 200 kbp-wide pairs, TSTA_msa on
16 x 50 kbp).  What only happens at scale is pinned here: hundreds of bands behind the ticket
order, epoch tags over thousands of chunks, re-centring drift over 1e5 rows, direction
addressing past 2^31.

The reference keeps 3 bytes per cell in host RAM (30 GB for the 100 kbp pair): those tests skip
when the host is too small or oracle/_ref has not been built.
"""
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tsta_b200" / "_build"
AFFINE = (2, -5, -2, -4)


def host_ram_bytes() -> int:
    try:
        return os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        return 0


def need_reference(cells: int):
    if ob.ref_isa() is None:
        pytest.skip("oracle/_ref not built (oracle/build_ref.sh needs /root/reference)")
    if 3 * cells > 0.6 * host_ram_bytes():
        pytest.skip("the reference keeps 3 bytes per cell: %d GB do not fit this host" % (3 * cells >> 30))


def make_reads(spec):
    """spec = (name, n, m, seed_x, seed_y | None, divergence): the reads of one job."""
    _, n, m, sx, sy, div = spec
    x = synth.random_read(n, sx)
    y = synth.random_read(m, sy) if div is None else synth.mutate(x, div, sy)
    return x, y


JOBS = {
    "random100k": ("random100k", 100_000, 100_000, 1, 2, None),       # the pair bench.py times
    "related100k": ("related100k", 100_000, 100_000, 3, 4, 0.05),
    "random200k": ("random200k", 200_000, 200_000, 1, 2, None),
}


def _oracle_job(args):
    spec, prm = args
    x, y = make_reads(spec)
    o = ob.psa_scores(x, y, *prm)
    return o.maxsorce, o.corner, o.lastrow


@pytest.fixture(scope="module")
def oracle_scores():
    """The score-only oracle needs ~30 s of one host core per 1e10 cells: all jobs of this module are
    started at once in worker processes and collected by the tests that need them."""
    from concurrent.futures import ProcessPoolExecutor
    ex = ProcessPoolExecutor(max_workers=4)
    fut = {(k, prm): ex.submit(_oracle_job, (JOBS[k], prm))
           for k, prm in [("random100k", AFFINE), ("related100k", AFFINE), ("related100k", (2, -5, -2, 0)),
                          ("random200k", AFFINE)]}
    yield fut
    ex.shutdown(wait=False, cancel_futures=True)


def check_scores(gpu, oracle_scores, key, prm):
    x, y = make_reads(JOBS[key])
    plan = gpu.PsaPlan(False)
    plan.upload(x, y, *prm)
    plan.run()
    r = plan.fetch()
    got_lastrow = plan.lastrow()
    plan.close()
    maxsorce, corner, lastrow = oracle_scores[(key, prm)].result(timeout=1200)
    assert (r.maxsorce, r.corner) == (maxsorce, corner)
    np.testing.assert_array_equal(got_lastrow, lastrow)
    return maxsorce, corner


def test_pair100k_random_scores_vs_oracle(gpu, oracle_scores):
    """BASELINE config 2, the pair bench.py times (seeds 1 and 2): maxsorce, corner, last row."""
    check_scores(gpu, oracle_scores, "random100k", AFFINE)


def test_pair100k_related_scores_vs_oracle(gpu, oracle_scores):
    """A 5 %-divergent pair: maxsorce is the corner region's score (~1.6e5), far from the random pair's -5."""
    maxsorce, _ = check_scores(gpu, oracle_scores, "related100k", AFFINE)
    assert maxsorce > 100_000
    check_scores(gpu, oracle_scores, "related100k", (2, -5, -2, 0))      # linear gaps (-O 0)


def test_pair100k_traced_vs_reference(gpu):
    """BASELINE config 3 at 1e10 cells: -o bytes and maxsorce against the real reference (AVX-512 build
    when the host has it), for the unrelated pair of the benchmark and a related one."""
    need_reference(100_000 * 100_000)
    for key in ("random100k", "related100k"):
        x, y = make_reads(JOBS[key])
        ms, want = ob.ref_psa(x, y, *AFFINE, T=os.cpu_count() or 4, W=30, trace=True, timeout=900)
        r = gpu.psa(x, y, *AFFINE, trace=True)
        assert r.maxsorce == ms
        assert hashlib.md5(r.output_bytes()).hexdigest() == hashlib.md5(want).hexdigest()
        assert r.output_bytes() == want


def test_pair200k_wide_traced_vs_reference(gpu):
    """200 000 columns x 60 000 rows (1.2e10 cells, 36 GB of reference matrices): the column geometry of
    BASELINE config 3 (1042 bands at C = 3, or 521 at C = 6) with as many rows as the reference can hold."""
    n, m = 200_000, 60_000
    need_reference(n * m)
    x, y = synth.random_read(n, 21), synth.random_read(m, 22)
    ms, want = ob.ref_psa(x, y, *AFFINE, T=os.cpu_count() or 4, W=30, trace=True, timeout=1200)
    r = gpu.psa(x, y, *AFFINE, trace=True)
    assert r.maxsorce == ms
    assert r.output_bytes() == want


def test_pair200k_traced_scores_and_alignment(gpu, oracle_scores):
    """BASELINE config 3 at its stated size (4e10 cells).  The reference would need 120 GB; pinned by the
    score-only oracle (maxsorce, corner, last row: ~2 min of one host core, started with the module) plus the alignment's own
    consistency (it spells both reads and re-scores to the oracle's corner)."""
    from test_psa_gpu import rescore
    x, y = make_reads(JOBS["random200k"])
    r = gpu.psa(x, y, *AFFINE, trace=True)
    assert r.a.replace(b"-", b"") == x and r.b.replace(b"-", b"") == y
    assert rescore(r.a, r.b, *AFFINE) == r.corner
    maxsorce, corner = check_scores(gpu, oracle_scores, "random200k", AFFINE)
    assert (r.maxsorce, r.corner) == (maxsorce, corner)


def test_msa16x50k_vs_reference(gpu, tmp_path):
    """BASELINE config 5: TSTA_msa on 16 synthetic 50 kbp reads, -o bytes against the real reference."""
    if ob.ref_isa() is None:
        pytest.skip("oracle/_ref not built")
    import bench
    reads = bench.make_msa_reads(16, 50_000, 0)
    fa = tmp_path / "in.fa"
    bench.write_fasta(fa, reads)
    ours = tmp_path / "ours.txt"
    p = subprocess.run([str(BIN / "TSTA_msa"), "-i", str(fa), "-o", str(ours)], capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr
    ref = tmp_path / "ref.txt"
    q = ob.run_ref("TSTA_msa", ["-T", os.cpu_count() or 4, "-W", 10, "-i", fa, "-o", ref], timeout=3000)
    assert q.returncode == 0, q.stderr
    assert p.stderr == q.stderr                      # seq_num: 16
    assert p.stdout == q.stdout                      # progress line
    assert hashlib.md5(ours.read_bytes()).hexdigest() == hashlib.md5(ref.read_bytes()).hexdigest()
    # the graph stays on the device between reads: against a full export per read, at least five
    # times fewer bytes cross PCIe -- and the file is the same
    nbytes = {}
    for mode in ("resident", "full"):
        env = dict(os.environ, TSTA_MSA_TIMING="1")
        if mode == "full":
            env["TSTA_MSA_FULL_EXPORT"] = "1"
        out = tmp_path / (mode + ".txt")
        r = subprocess.run([str(BIN / "TSTA_msa"), "-i", str(fa), "-o", str(out)], capture_output=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr
        assert out.read_bytes() == ours.read_bytes()
        nbytes[mode] = int([l for l in r.stderr.decode().splitlines() if l.startswith("export:")][0].split()[1])
    assert nbytes["full"] >= 5 * nbytes["resident"], nbytes
