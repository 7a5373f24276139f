#!/usr/bin/env python
"""bench.py -- GCUPS of the global-alignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline of the
line (value / e2e / roofline / cpu_baseline) is BASELINE.json configs[1] -- one synthetic
100 kbp x 100 kbp DNA pair per GPU, affine gaps, score-only (TSTA_psa_notrace); under torchrun
every rank aligns its own independent pair (round-robin sharding of independent units, no
data-path collective): weak scaling.

Without --workload the line also carries a `workloads` block, every entry measured in the same
process with the same rules (warm-up, L2 flush between timed steps, CUDA events, max over ranks):
  1 GPU   pair100k_trace (the north-star gate), pair200k_trace, batch250 / batch10k with affine
          and linear (-O 0) gaps, batch250_trace, batch10k_trace, msa16x50k
  N GPUs  strong scaling of FIXED jobs: batch10k_strong / batch250_strong (one fixed batch sharded
          round-robin, results gathered on rank 0 inside the timed region) and pair100k_bands /
          pair400k_bands (ONE pair as column slices, boundary column over NVLink)
Each entry reports value, e2e, roofline_frac, kernel_ms and `parity`: the result compared with the
unmodified reference binary (oracle/_ref) or, for the sliced pair, with the one-GPU run.

Workloads (--workload NAME makes NAME the headline):
  pair100k          configs[1]: one 100 kbp pair per GPU, score-only            (default)
  pair100k_bands    configs[1], multi-GPU form: ONE pair as column slices over the N GPUs
  pair400k_bands    the same for a 400 kbp x 100 kbp pair
  pair100k_trace    one 100 kbp pair per GPU with traceback
  pair200k_trace    configs[2]: 200 kbp pair with traceback
  batch250[_linear] configs[3]: 1,000,000 pairs of 250 bp per GPU, score-only (tsta_psa_batch)
  batch10k[_linear] configs[3]: 10,000 pairs of 10 kbp per GPU, score-only
  batch250_trace / batch10k_trace   the same with traceback (batch10k_trace: 2,000 pairs)
  batch10k_strong / batch250_strong a FIXED batch (40,000 x 10 kbp / 4,000,000 x 250 bp) over all GPUs
  msa16x50k         configs[4]: TSTA_msa graph DP, 16 reads x 50 kbp (one job per GPU)

  value     whole-job GCUPS with the inputs already resident in HBM where the ABI allows it
            (plan API: kernel time from CUDA events on the library's stream, max over ranks;
            batch / msa: the kernels' CUDA-event time reported by the library)
  e2e       same metric through the host-buffer C entry point (tsta_psa_align / tsta_psa_batch /
            the TSTA_msa driver loop): H2D of the reads, kernels, D2H of the results, every step
  roofline  fill kernel against the MEASURED int32 DPX issue rate (tsta_measure_dpx_peak):
            7 algorithmic lane-ops per cell
  cpu_baseline / --impl reference
            the reference's own AVX-512/AVX2/SSE + pthread implementation (oracle/_ref, built
            from the unmodified reference sources) on this box's host cores, same workload
            (bounded sample for the batch / msa shapes)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

AFFINE = dict(M=2, X=-5, E=-2, O=-4)
LINEAR = dict(M=2, X=-5, E=-2, O=0)
PARAMS = AFFINE
LANE_OPS_PER_CELL = 7  # 1 select, 1 max(e,f), 1 viaddmax, 1 add, 2 viaddmax, 1 running max

# name: kind, shape, trace, params, description
WORKLOADS = {
    "pair100k": ("pair", (100_000, 100_000), False, AFFINE,
                 "TSTA_psa_notrace: one synthetic 100 kbp x 100 kbp pair per GPU, affine (M2 X-5 E-2 O-4), score-only"),
    "pair100k_bands": ("bands", (100_000, 100_000), False, AFFINE,
                       "TSTA_psa_notrace: ONE synthetic 100 kbp x 100 kbp pair as column slices over all GPUs (NVLink boundary hand-off)"),
    "pair400k_bands": ("bands", (400_000, 100_000), False, AFFINE,
                       "TSTA_psa_notrace: ONE synthetic 400 kbp x 100 kbp pair as column slices over all GPUs (more bands than one GPU has sub-partitions)"),
    "pair100k_trace": ("pair", (100_000, 100_000), True, AFFINE,
                       "TSTA_psa: one synthetic 100 kbp x 100 kbp pair per GPU, affine, with traceback"),
    "pair200k_trace": ("pair", (200_000, 200_000), True, AFFINE,
                       "TSTA_psa: one synthetic 200 kbp x 200 kbp pair per GPU, affine, with traceback"),
    "pair20k": ("pair", (20_000, 20_000), False, AFFINE, "small single pair (debug)"),
    "pair20k_trace": ("pair", (20_000, 20_000), True, AFFINE, "small single pair with traceback (debug)"),
    "batch250": ("batch", (1_000_000, 250), False, AFFINE,
                 "batched PSA: 1,000,000 synthetic 250 bp pairs per GPU (half unrelated, half 5 % divergent), affine, score-only"),
    "batch250_linear": ("batch", (1_000_000, 250), False, LINEAR,
                        "batched PSA: 1,000,000 synthetic 250 bp pairs per GPU, linear gaps (-O 0), score-only"),
    "batch10k": ("batch", (10_000, 10_000), False, AFFINE,
                 "batched PSA: 10,000 synthetic 10 kbp pairs per GPU, affine, score-only"),
    "batch10k_linear": ("batch", (10_000, 10_000), False, LINEAR,
                        "batched PSA: 10,000 synthetic 10 kbp pairs per GPU, linear gaps (-O 0), score-only"),
    "batch250_trace": ("batch", (1_000_000, 250), True, AFFINE, "batched PSA with traceback: 1,000,000 pairs of 250 bp per GPU"),
    "batch250_2bit": ("batch", (1_000_000, 250), False, AFFINE,
                      "batched PSA: 1,000,000 synthetic 250 bp pairs per GPU, reads in the 2-bit wire format (tsta_psa_batch_2bit), score-only"),
    "batch250_trace_2bit": ("batch", (1_000_000, 250), True, AFFINE,
                            "batched PSA with traceback: 1,000,000 pairs of 250 bp per GPU, reads in the 2-bit wire format"),
    "batch10k_trace": ("batch", (2_000, 10_000), True, AFFINE, "batched PSA with traceback: 2,000 pairs of 10 kbp per GPU"),
    "batch10k_strong": ("batch_strong", (40_000, 10_000), False, AFFINE,
                        "batched PSA, FIXED job: 40,000 synthetic 10 kbp pairs sharded round-robin over all GPUs, results gathered on rank 0"),
    "batch250_strong": ("batch_strong", (4_000_000, 250), False, AFFINE,
                        "batched PSA, FIXED job: 4,000,000 synthetic 250 bp pairs sharded round-robin over all GPUs, results gathered on rank 0"),
    "msa16x50k": ("msa", (16, 50_000), True, AFFINE,
                  "TSTA_msa: partial-order alignment of 16 synthetic 50 kbp reads (5 % divergence), one job per GPU"),
    "msa16x10k": ("msa", (16, 10_000), True, AFFINE, "TSTA_msa: 16 x 10 kbp (debug)"),
}
SUITE_NGPU = ["batch10k_strong", "batch250_strong", "pair100k_bands", "pair400k_bands"]
# (the fixed jobs of the N-GPU suite also run on one GPU: the first point of their strong-scaling curves)
SUITE_1GPU = ["pair100k_trace", "pair200k_trace", "batch250", "batch250_linear", "batch10k", "batch10k_linear",
              "batch250_trace", "batch10k_trace", "batch250_2bit", "batch250_trace_2bit", "msa16x50k"] + SUITE_NGPU


def peaks():
    try:
        return json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int, interval_ms: int = 100):
        self.device = device
        self.interval_ms = interval_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", str(self.interval_ms)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                seen = set()
                for r in rows:
                    for k, nm in enumerate(names):
                        if r[5 + k].strip().lower().startswith("active"):
                            seen.add(nm)
                out["reasons"] = sorted(seen)
                out["samples"] = len(sm)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------
def make_pair(shape, rank: int):
    from tsta_b200 import synth
    n, m = shape
    return synth.random_read(n, 2 * rank + 1), synth.random_read(m, 2 * rank + 2)


_BATCH_CACHE = {}


def make_batch(npairs: int, ln: int, seed: int):
    """Rectangular batch: even pairs unrelated, odd pairs 5 %-substituted copies.  Returns
    (a, b) uint8 [npairs, ln]; cached, because several workloads share one shape."""
    key = (npairs, ln, seed)
    if key not in _BATCH_CACHE:
        from tsta_b200 import synth
        _BATCH_CACHE.clear()
        a = synth.random_bulk(npairs * ln, 2 * seed + 11).reshape(npairs, ln)
        b = synth.random_bulk(npairs * ln, 2 * seed + 12).reshape(npairs, ln)
        rnd = np.random.default_rng(seed)
        odd = np.arange(npairs) % 2 == 1
        block = 1 << 18                      # (bounded temporaries: the 250 bp batches have 1e9 bases)
        for o in range(0, npairs, block):
            sl = slice(o, min(npairs, o + block))
            od = odd[sl]
            keep = rnd.random((int(od.sum()), ln), dtype=np.float32) >= 0.05
            bb = b[sl]
            bb[od] = np.where(keep, a[sl][od], bb[od])
        _BATCH_CACHE[key] = (a, b)
    return _BATCH_CACHE[key]


def batch_arrays(a: np.ndarray, b: np.ndarray):
    npairs, ln = a.shape
    off = np.arange(npairs, dtype=np.int64) * ln
    lens = np.full(npairs, ln, dtype=np.int32)
    return a.reshape(-1), off, lens, b.reshape(-1), off.copy(), lens.copy()


def make_msa_reads(nreads: int, ln: int, seed: int):
    from tsta_b200 import synth
    anc = synth.random_read(ln, 11 + seed)
    return [anc] + [synth.mutate(anc, 0.05, 200 + 16 * seed + k)[:ln] for k in range(nreads - 1)]


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for k, r in enumerate(reads):
            f.write(b">%d\n" % (k + 1) + r + b"\n")


def host_ram_bytes() -> int:
    try:
        return os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except Exception:
        return 32 << 30


# ------------------------------------------------------------------------------------------
# reference arm: the unmodified reference binaries on this box's host cores
# ------------------------------------------------------------------------------------------
def _ref_psa_run(x: bytes, y: bytes, trace: bool, W: int, threads: int, prm=AFFINE):
    """One run of the reference binary: (wall seconds, maxsorce, md5 of the -o file | None)."""
    from oracle import binding as ob
    with tempfile.TemporaryDirectory() as d:
        f1, f2, fo = os.path.join(d, "a.fa"), os.path.join(d, "b.fa"), os.path.join(d, "o.txt")
        ob.write_fasta(f1, [x])
        ob.write_fasta(f2, [y])
        args = ["-M", prm["M"], "-X", prm["X"], "-E", prm["E"], "-O", prm["O"],
                "-T", threads, "-W", W, "-1", f1, "-2", f2]
        if trace:
            args += ["-o", fo]
        t0 = time.perf_counter()
        p = ob.run_ref("TSTA_psa" if trace else "TSTA_psa_notrace", args, timeout=900)
        dt = time.perf_counter() - t0
        if b"maxsorce=" not in p.stdout:
            raise RuntimeError("reference run failed: " + p.stderr.decode()[-200:])
        ms = int(p.stdout.decode().strip().split("maxsorce=")[1])
        md5 = hashlib.md5(open(fo, "rb").read()).hexdigest() if trace else None
    return dt, ms, md5


def cpu_reference(name: str, steps: int):
    """Times the reference (oracle/_ref) on a bounded sample of the workload; falls back to the
    scalar C port of the oracle when the reference binaries are not present.  Besides the timing it
    returns what the reference computed (`maxsorce`, `out_md5`, the sample's shape) so that the
    caller can check its own result against it."""
    from oracle import binding as ob
    from tsta_b200 import synth
    kind, shape, trace, prm, _ = WORKLOADS[name]
    threads = os.cpu_count() or 1
    isa = ob.ref_isa()
    steps = max(1, min(steps, 3))
    if isa is None:
        side = 12_000
        x, y = synth.random_read(side, 1), synth.random_read(side, 2)
        ob.psa_scores(x[:500], y[:500])
        t0 = time.perf_counter()
        o = ob.psa_scores(x, y, **prm)
        t = time.perf_counter() - t0
        return {"value": side * side / t / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port", "runs": 1,
                "sample": f"{side} x {side} pair through the scalar C oracle (oracle/_ref not built)", "seconds": t,
                "sample_shape": (side, side), "maxsorce": o.maxsorce, "out_md5": None}
    if kind in ("pair", "bands"):
        n, m = shape
        note = "the full workload (1 pair, %d x %d)" % (n, m)
        if trace:   # the reference keeps 3 bytes per cell in host RAM
            avail = host_ram_bytes()
            while 3 * n * m > 0.5 * avail:
                n //= 2
                m //= 2
            note = "a %d x %d pair (the reference keeps 3 bytes per cell in host RAM)" % (n, m)
        x, y = synth.random_read(n, 1), synth.random_read(m, 2)
        best = None
        runs = 0
        for W in (30, 10):
            _ref_psa_run(x[:2000], y[:2000], trace, W, threads, prm)      # page in the binary
            rs = [_ref_psa_run(x, y, trace, W, threads, prm) for _ in range(steps)]
            runs += len(rs)
            t = sum(r[0] for r in rs) / len(rs)
            if best is None or t < best[0]:
                best = (t, W, rs[-1][1], rs[-1][2])
        t, W, ms, md5 = best
        return {"value": n * m / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference", "runs": runs,
                "sample": f"{note}; reference {isa} build, -T {threads} -W {W}, whole-process wall time, "
                          f"mean of {steps} runs (best of -W 30 / -W 10)",
                "seconds": t, "sample_shape": (n, m), "maxsorce": ms, "out_md5": md5}
    if kind in ("batch", "batch_strong"):
        # the reference aligns one pair per process: one process per core at a time over a
        # time-bounded sample (a run that hits the reference's shutdown hang is killed and retried)
        _, ln = shape
        W = 4 if ln <= 1000 else 10
        budget_s = 15.0
        one_timeout = 0.5 if ln <= 1000 else 1.5      # ~10x a normal run; the reference sometimes hangs at exit
        procs = max(1, threads // 2)      # every reference process is a worker plus a busy-waiting main thread
        from concurrent.futures import ThreadPoolExecutor
        pairs = [(synth.random_read(ln, 9000 + k), synth.random_read(ln, 19000 + k)) for k in range(8)]
        with tempfile.TemporaryDirectory() as d:
            files = []
            for k, (x, y) in enumerate(pairs):
                f1, f2 = os.path.join(d, f"a{k}.fa"), os.path.join(d, f"b{k}.fa")
                ob.write_fasta(f1, [x])
                ob.write_fasta(f2, [y])
                files.append((f1, f2))
            exe = "TSTA_psa" if trace else "TSTA_psa_notrace"

            def one(k, out):
                f1, f2 = files[k % len(files)]
                args = ["-M", prm["M"], "-X", prm["X"], "-E", prm["E"], "-O", prm["O"],
                        "-T", 1, "-W", W, "-1", f1, "-2", f2] + (["-o", out] if trace else [])
                p = ob.run_ref(exe, args, timeout=one_timeout, retries=8)
                return b"maxsorce=" in p.stdout
            one(0, os.path.join(d, "warm.txt"))
            t0 = time.perf_counter()
            deadline = t0 + budget_s

            def worker(w):
                done = 0
                while time.perf_counter() < deadline:
                    if one(w + done, os.path.join(d, f"o{w}.txt")):
                        done += 1
                return done
            with ThreadPoolExecutor(procs) as ex:
                counts = list(ex.map(worker, range(procs)))
            t = time.perf_counter() - t0
        total = sum(counts)
        cells = float(total) * ln * ln
        return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference", "runs": total,
                "sample": f"{total} pairs of {ln} bp in {t:.1f} s, one reference {isa} process (-T 1 -W {W}) per pair, "
                          f"{procs} processes at a time = {threads} busy threads (the reference reads one pair per process and "
                          f"its main thread spins on the tile barrier; process start-up included)",
                "seconds": t, "sample_shape": (total, ln), "maxsorce": None, "out_md5": None}
    # msa: a smaller job of the same kind (node-cells counted by our own driver's rule)
    nreads, ln = shape
    ln_s = min(ln, 10_000)
    reads = make_msa_reads(nreads, ln_s, 0)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa")
        write_fasta(fa, reads)
        best = None
        md5 = None
        for W in (10, 30):
            t0 = time.perf_counter()
            ob.run_ref("TSTA_msa", ["-T", threads, "-W", W, "-i", fa, "-o", os.path.join(d, "ref.txt")], timeout=3000)
            t = time.perf_counter() - t0
            if W == 10:     # (-W changes the tile geometry and with it quirk Q1: our driver matches the -W 10 geometry)
                md5 = hashlib.md5(open(os.path.join(d, "ref.txt"), "rb").read()).hexdigest()
            if best is None or t < best[0]:
                best = (t, W)
        t, W = best
        cells = msa_node_cells(fa)
    return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference", "runs": 2,
            "sample": f"{nreads} reads x {ln_s} bp (same generator), reference {isa} build, -T {threads} -W {W}, "
                      "whole-process wall time; node-cells counted as sum over reads of graph rows x read length",
            "seconds": t, "sample_shape": (nreads, ln_s), "maxsorce": None, "out_md5": md5}


def msa_node_cells(fasta_path: str, device: int | None = None):
    """Runs OUR TSTA_msa driver with timing on and returns the node-cell count it reports."""
    env = dict(os.environ, TSTA_MSA_TIMING="1")
    if device is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(device)
    with tempfile.TemporaryDirectory() as d:
        p = subprocess.run([str(ROOT / "tsta_b200/_build/TSTA_msa"), "-i", fasta_path, "-o", os.path.join(d, "o.txt")],
                           capture_output=True, env=env)
    for line in p.stderr.decode().splitlines():
        if line.startswith("kernels:"):
            return float(line.split(",")[1].split()[0])
    raise RuntimeError("TSTA_msa timing line missing: " + p.stderr.decode()[-300:])


MSA_LAST = {}


def run_msa_once(fasta_path: str, out_path: str, device: int):
    """One TSTA_msa process on `device`: (wall s, kernel s, node-cells, launches)."""
    env = dict(os.environ, TSTA_MSA_TIMING="1")
    if "CUDA_VISIBLE_DEVICES" in env:
        vis = env["CUDA_VISIBLE_DEVICES"].split(",")
        env["CUDA_VISIBLE_DEVICES"] = vis[device] if device < len(vis) else vis[0]
    else:
        env["CUDA_VISIBLE_DEVICES"] = str(device)
    t0 = time.perf_counter()
    p = subprocess.run([str(ROOT / "tsta_b200/_build/TSTA_msa"), "-i", fasta_path, "-o", out_path],
                       capture_output=True, env=env)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("TSTA_msa failed: " + p.stderr.decode()[-300:])
    fill_ms = walk_ms = cells = 0.0
    loop_s = None
    global MSA_LAST
    for line in p.stderr.decode().splitlines():
        if line.startswith("kernels:"):
            parts = line.replace(",", " ").split()
            fill_ms, walk_ms, cells = float(parts[2]), float(parts[5]), float(parts[7])
            MSA_LAST = {"fill_ms": fill_ms, "walk_ms": walk_ms, "node_cells": cells,
                        "edge_cells": float(parts[-2]) if parts[-1] == "edge-cells" else None}
        if line.startswith("timing:"):
            w = line.split()
            loop_s = sum(float(w[k]) for k in (2, 4, 6, 8, 10))
        if line.startswith("export:"):
            MSA_LAST["export"] = line
    return wall, (fill_ms + walk_ms) * 1e-3, cells, loop_s


def rescore(a: bytes, b: bytes, M, X, E, O) -> int:
    """Affine score of an emitted alignment (a gap of length k costs O + k*E)."""
    A = np.frombuffer(a, dtype=np.uint8)
    B = np.frombuffer(b, dtype=np.uint8)
    ga, gb = A == ord("-"), B == ord("-")
    both = ~ga & ~gb
    s = int(np.sum(np.where(A[both] == B[both], M, X)))
    for g in (ga, gb):
        s += int(g.sum()) * E
        s += (int(np.sum(g[1:] & ~g[:-1])) + int(g[0])) * O
    return s


# ------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench invocation (one rank)."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        from tsta_b200 import api
        self.torch, self.dist, self.api = torch, dist, api
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.warmup = max(3, a.warmup)
        torch.cuda.set_device(self.local_rank)
        api.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        self.dpx_peak, _ = api.measure_dpx_peak()
        self.check_parity = not a.no_cpu_baseline

    def flush_l2(self):
        if self.flush is None:
            self.flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device="cuda")
        self.flush.zero_()
        self.torch.cuda.synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return [float(v) for v in t]

    def allmax(self, vals):
        return self._reduce(vals, self.dist.ReduceOp.MAX)

    def allsum(self, vals):
        return self._reduce(vals, self.dist.ReduceOp.SUM)


def issue_roofline(ctx, fill_rate, packed, kernel_name, kernel_ms):
    """Fill kernel against the measured DPX issue rate.  The packed kernels advance two cells per
    lane-op (int16 x 2), so their roofline is twice the measured int32 VIADDMNMX rate."""
    packf = 2 if packed else 1
    peak = packf * ctx.dpx_peak / 1e12
    achieved = fill_rate * LANE_OPS_PER_CELL / 1e12
    return {"bound": "int_alu", "bound_note": "integer-ALU / DPX issue (neither hbm nor tensor: SURVEY.md 8d)",
            "achieved": achieved, "peak": peak, "unit": "Tlaneop/s", "frac": achieved / peak,
            "peak_source": "measured live: dependency-free VIADDMNMX loop (tsta_measure_dpx_peak) = %.2f Tlaneop/s, x%d "
                           "cells per lane-op for the %s kernels" % (ctx.dpx_peak / 1e12, packf, "int16x2" if packf == 2 else "int32"),
            "algorithmic_ops_per_cell": LANE_OPS_PER_CELL, "kernel": kernel_name, "kernel_ms": kernel_ms, "traffic": None}


# ------------------------------------------------------------------------------------------
# workloads.  Each returns a dict: value, ms_per_step, steps, scaling, e2e{...}, gpu_launches,
# roofline, clocks, config (same keys on both arms), impl (how OUR arm ran it), result, parity
# ------------------------------------------------------------------------------------------
def wl_pair(ctx, name, steps):
    api = ctx.api
    kind, shape, trace, prm, desc = WORKLOADS[name]
    n, m = shape
    x, y = make_pair(shape, ctx.rank)
    plan = api.PsaPlan(trace)
    plan.upload(x, y, **prm)
    for _ in range(ctx.warmup):
        plan.run()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank, 100)
    sampler.start()
    fill_ms, trace_ms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        ctx.flush_l2()
        f, t = plan.run()          # CUDA events on the library's stream around the kernels
        fill_ms.append(f)
        trace_ms.append(t)
    ctx.barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    res = plan.fetch()
    info = plan.info()
    plan.close()
    step_ms, fill_only = ctx.allmax([(sum(fill_ms) + sum(trace_ms)) / steps, sum(fill_ms) / steps])
    cells_rank = float(n) * m
    cells_total = ctx.world * cells_rank
    kernel_name = "psa_pair16_kernel" if info["packed"] else "psa_pair_kernel"
    # end to end through the one-shot C entry point (host buffers in, host result out)
    for _ in range(2):
        api.psa(x, y, trace=trace, **prm)
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(steps, 10))
    for _ in range(e2e_steps):
        r2 = api.psa(x, y, trace=trace, **prm)
    ctx.barrier()
    (e2e_s,) = ctx.allmax([(time.perf_counter() - t0) / e2e_steps])
    assert (r2.maxsorce, r2.corner) == (res.maxsorce, res.corner)
    roof = issue_roofline(ctx, cells_rank / (fill_only * 1e-3), info["packed"], kernel_name, fill_only)
    roof["note"] = ("achieved = cells/s x 7 algorithmic lane-ops (SURVEY.md 8d) over the fill kernel(s) alone; "
                    "ncu summaries under profiles/")
    try:   # DRAM traffic of the dominant kernel, from the committed ncu capture of this workload
        tfile = next(f for f in (ROOT / "profiles" / "r02_traffic.json", ROOT / "profiles" / "r01_traffic.json") if f.exists())
        tr = json.load(open(tfile)).get(name)
        if tr and ctx.world == 1:
            roof["traffic"] = tr["bytes"]
            roof["traffic_source"] = tr["source"]
    except Exception:
        pass
    result = {"maxsorce": res.maxsorce, "corner": res.corner}
    parity = None
    if trace:
        result["alignment_length"] = len(res.a)
        result["out_md5"] = hashlib.md5(res.output_bytes()).hexdigest()
        s0, s1 = api.order_pair(x, y)
        parity = {"alignment_spells_both_reads": res.a.replace(b"-", b"") == s0 and res.b.replace(b"-", b"") == s1,
                  "rescored_alignment_equals_corner": rescore(res.a, res.b, **prm) == res.corner}
    return {"value": cells_total / (step_ms * 1e-3) / 1e9, "ms_per_step": step_ms, "steps": steps, "scaling": "weak",
            "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": len(x) + len(y),
                    "d2h_bytes_per_step": 8 + (2 * len(r2.a) if trace and r2.a else 0), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": steps * info["launches"], "roofline": roof, "clocks": clocks,
            "config": {"cells_per_step_per_gpu": int(cells_rank)},
            "impl": {"cols_per_lane": info["cols_per_lane"], "bands": info["bands"], "packed_int16x2": info["packed"],
                     "launches_per_step": info["launches"],
                     "parallelism": f"independent pairs, 1 per GPU x {ctx.world}" if ctx.world > 1 else "1 GPU"},
            "result": result, "parity": parity, "dtype": "int16x2" if info["packed"] else "int32",
            "extra": {"fill_ms": fill_only, "trace_ms": step_ms - fill_only, "wall_s": wall}}


def wl_bands(ctx, name, steps):
    api, dist, world, rank = ctx.api, ctx.dist, ctx.world, ctx.rank
    kind, shape, trace, prm, desc = WORKLOADS[name]
    n, m = shape
    x, y = make_pair(shape, 0)          # the SAME pair on every rank
    s0, s1 = api.order_pair(x, y)
    slices = api.slice_columns(len(s0), world)
    mine = slices[rank] if rank < len(slices) else None
    plan = api.PsaPlan(False)
    if mine is not None:
        plan.upload_slice(s0, s1, mine[0], mine[1], **prm)
    if world > 1:
        from tsta_b200 import shard
        h = plan.inbox_handle() if (mine is not None and mine[0] > 0) else b""
        handles = shard.exchange_inbox_handles(dist, h, world)
        if mine is not None and rank + 1 < len(slices):
            plan.connect_right(handles[rank + 1])

    def run_all():
        if mine is None:
            return 0.0
        plan.launch()
        return plan.wait()[0]
    for _ in range(ctx.warmup):
        ctx.barrier()
        run_all()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank, 100)
    sampler.start()
    tsum = 0.0
    walls = []
    for _ in range(steps):
        ctx.flush_l2()
        ctx.barrier()                       # all slices start together: slice g only ever waits for g-1
        t0 = time.perf_counter()
        f = run_all()
        ctx.torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        tsum += f
    ctx.barrier()
    clocks = sampler.stop()
    # a slice's kernel time includes waiting for its left neighbours: the job time is the max over ranks
    step_ms, = ctx.allmax([tsum / steps])
    cells_total = float(n) * m
    info = plan.info() if mine is not None else {"launches": 0, "cols_per_lane": 0, "bands": 0, "packed": True}
    res = plan.fetch() if mine is not None else None
    ms_all = ctx.allmax([res.maxsorce if res else -10**9])[0]
    corner = ctx.allmax([res.corner if (res and rank == len(slices) - 1) else -10**9])[0]
    result = {"maxsorce": int(ms_all), "corner": int(corner)}
    # e2e: upload + run + fetch of the slices with host buffers
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(steps, 10))
    for _ in range(e2e_steps):
        if mine is not None:
            plan.upload_slice(s0, s1, mine[0], mine[1], **prm)
        ctx.barrier()
        run_all()
        if mine is not None:
            plan.fetch()
    ctx.barrier()
    (e2e_s,) = ctx.allmax([(time.perf_counter() - t0) / e2e_steps])
    plan.close()
    # parity: the same pair on ONE GPU (rank 0's own device), whole
    parity = None
    if rank == 0:
        whole = api.PsaPlan(False)
        whole.upload(x, y, **prm)
        one_ms = whole.run()[0]
        w = whole.fetch()
        whole.close()
        parity = {"equals_one_gpu_run": (w.maxsorce, w.corner) == (result["maxsorce"], result["corner"]),
                  "one_gpu_ms": one_ms, "speedup_vs_one_gpu": one_ms / step_ms}
    ctx.barrier()
    roof = issue_roofline(ctx, cells_total / (step_ms * 1e-3) / world, info["packed"], "psa_pair16_kernel", step_ms)
    return {"value": cells_total / (step_ms * 1e-3) / 1e9, "ms_per_step": step_ms, "steps": steps, "scaling": "strong",
            "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": (mine[1] if mine else 0) + len(s1),
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": steps * info["launches"], "roofline": roof, "clocks": clocks,
            "config": {"cells_per_step": int(cells_total)},
            "impl": {"cols_per_lane": info["cols_per_lane"], "bands_per_gpu": info["bands"], "slices": slices,
                     "packed_int16x2": info["packed"],
                     "parallelism": f"column slices over {world} GPU(s), boundary column over NVLink"},
            "result": result, "parity": parity, "dtype": "int16x2",
            "extra": {"fill_ms": step_ms, "wall_ms_per_step": 1e3 * sum(walls) / len(walls)}}


def wl_batch(ctx, name, steps):
    api, world, rank = ctx.api, ctx.world, ctx.rank
    kind, shape, trace, prm, desc = WORKLOADS[name]
    npairs_total, ln = shape
    strong = kind == "batch_strong"
    if strong:
        from tsta_b200 import shard
        a, b = make_batch(npairs_total, ln, 77)           # the SAME fixed batch on every rank ...
        idx = shard.shard_indices(npairs_total, rank, world)   # ... of which this rank takes pairs r, r+W, ...
        a, b = a[idx], b[idx]
    else:
        a, b = make_batch(npairs_total, ln, rank)
        idx = None
    npairs = a.shape[0]
    # inputs and outputs live in page-locked host memory (tsta_host_alloc), as a caller that
    # streams batches would keep them
    bits2 = name.endswith("_2bit")
    arrs = list(batch_arrays(a, b))
    if bits2:            # the wire format: four bases to a byte (packed once, outside the timed region, like a 2-bit read store)
        arrs[0], arrs[3] = api.pack_2bit(arrs[0]), api.pack_2bit(arrs[3])
    pins = [api.pinned_copy(v) for v in arrs]
    s0, o0, l0, s1, o1, l1 = [p.array for p in pins]
    outs = [api.PinnedArray(npairs, np.int32), api.PinnedArray(npairs, np.int32)]
    if trace:
        # moves come back packed four to a byte (tsta_psa_batch_packed): 2 bits are all a move needs
        slot = (2 * ln + 3) & ~3
        ops_off = np.arange(npairs, dtype=np.int64) * slot
        outs += [api.PinnedArray(npairs * slot // 4, np.uint8), api.pinned_copy(ops_off), api.PinnedArray(npairs, np.int32)]
    out = tuple(p.array for p in outs)
    torch = ctx.torch
    gathered = None

    def one_step():
        nonlocal gathered
        r = api.psa_batch_arrays(s0, o0, l0, s1, o1, l1, trace=trace, out=out, packed=trace, bits2=bits2, **prm)
        if strong and world > 1:
            # the job ends when rank 0 holds every pair's result: ONE all_gather of (maxsorce, corner) per step
            mine = torch.from_numpy(np.stack([r.maxsorce, r.corner])).cuda()
            pad = (npairs_total + world - 1) // world
            if mine.shape[1] < pad:
                mine = torch.nn.functional.pad(mine, (0, pad - mine.shape[1]))
            parts = [torch.empty_like(mine) for _ in range(world)]
            ctx.dist.all_gather(parts, mine)
            if rank == 0:
                gathered = torch.stack(parts).cpu().numpy()       # [world, 2, pad]
        return r
    for _ in range(ctx.warmup):
        r = one_step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank, 100)
    sampler.start()
    kms, walls = [], []
    for _ in range(steps):
        ctx.flush_l2()
        if strong:
            ctx.barrier()
        t0 = time.perf_counter()
        r = one_step()
        walls.append(time.perf_counter() - t0)
        kms.append(r.kernel_ms)
    ctx.barrier()
    clocks = sampler.stop()
    step_ms, e2e_s = ctx.allmax([sum(kms) / steps, sum(walls) / steps])
    cells_rank = float(npairs) * ln * ln
    cells_total = float(npairs_total) * ln * ln if strong else world * cells_rank
    sums = ctx.allsum([float(r.maxsorce.astype(np.int64).sum()), float(r.corner.astype(np.int64).sum())])
    result = {"sum_maxsorce": int(sums[0]), "sum_corner": int(sums[1])}
    if strong and world > 1 and rank == 0:
        full = np.zeros((2, npairs_total), dtype=np.int64)
        for w in range(world):
            k = len(range(w, npairs_total, world))
            full[:, w::world] = gathered[w][:, :k]
        result["gathered_on_rank0"] = {"pairs": int(npairs_total), "sum_maxsorce": int(full[0].sum()), "sum_corner": int(full[1].sum())}
    h2d = int(s0.nbytes + s1.nbytes + o0.nbytes + o1.nbytes + l0.nbytes + l1.nbytes)
    d2h = int(r.maxsorce.nbytes + r.corner.nbytes + (r._ops.nbytes + r._alen.nbytes if trace else 0))
    # parity: a few pairs of this batch through the unmodified reference binary (one process per pair)
    parity = None
    if rank == 0 and ctx.check_parity:
        parity = batch_parity(a, b, r, prm, trace, ln)
    roof = issue_roofline(ctx, cells_rank / (step_ms * 1e-3), True, "psa_batch16_dual_kernel" if ln <= 256 else "psa_batch16_kernel", step_ms)
    launches = steps * (2 if trace else 1)
    del pins, outs
    return {"value": cells_total / (step_ms * 1e-3) / 1e9, "ms_per_step": step_ms, "steps": steps,
            "scaling": "strong" if strong else "weak",
            "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3,
                    "includes": "H2D of this rank's reads, kernels, D2H of the results" + (", all_gather of every pair's result" if strong and world > 1 else "")},
            "gpu_launches": launches, "roofline": roof, "clocks": clocks,
            "config": {"pairs": int(npairs_total) if strong else int(npairs), "pairs_are": "the whole job" if strong else "per GPU",
                       "read_length": ln, "cells_per_step": int(cells_total)},
            "impl": {"pairs_this_rank": int(npairs),
                     "parallelism": f"independent pairs sharded round-robin over {world} GPU(s)" + (", one all_gather of the results per step" if strong else ", no collective"),
                     "value_is": "CUDA-event time of the fill (+ traceback) kernels inside tsta_psa_batch, max over ranks"},
            "result": result, "parity": parity, "dtype": "int16x2", "extra": {"kernel_ms": step_ms}}


def ref_one_pair(x, y, prm, trace, W, timeout):
    """(maxsorce, -o bytes | None) of one pair from the reference binary."""
    from oracle import binding as ob
    with tempfile.TemporaryDirectory() as d:
        f1, f2, fo = os.path.join(d, "a.fa"), os.path.join(d, "b.fa"), os.path.join(d, "o.txt")
        ob.write_fasta(f1, [x])
        ob.write_fasta(f2, [y])
        args = ["-M", prm["M"], "-X", prm["X"], "-E", prm["E"], "-O", prm["O"], "-T", 1, "-W", W, "-1", f1, "-2", f2]
        p = ob.run_ref("TSTA_psa" if trace else "TSTA_psa_notrace", args + (["-o", fo] if trace else []), timeout=timeout, retries=12)
        ms = int(p.stdout.decode().strip().split("maxsorce=")[1])
        return ms, (open(fo, "rb").read() if trace else None)


def batch_parity(a, b, r, prm, trace, ln, samples=6):
    from oracle import binding as ob
    if ob.ref_isa() is None:
        return {"checked": False, "why": "oracle/_ref not built"}
    npairs = a.shape[0]
    ks = sorted(set(int(v) for v in np.linspace(0, npairs - 1, samples)) | {1})
    ok = True
    for k in ks:
        x, y = a[k].tobytes(), b[k].tobytes()
        # (short timeout, many retries: the reference sometimes hangs in its thread-pool teardown)
        ms, out = ref_one_pair(x, y, prm, trace, 4 if ln <= 1000 else 10, 2.0 if ln <= 1000 else 6.0)
        ok = ok and ms == int(r.maxsorce[k])
        if trace:
            nal = int(r._alen[k])
            ops = r.moves(k, 2 * ln)
            A = np.full(nal, ord("-"), dtype=np.uint8)
            B = np.full(nal, ord("-"), dtype=np.uint8)
            A[ops != 2] = a[k]
            B[ops != 0] = b[k]
            ok = ok and out == b">1\n" + A.tobytes() + b"\n>2\n" + B.tobytes()
    return {"checked": True, "ok": bool(ok), "pairs_compared_with_reference_binary": ks,
            "compared": "maxsorce" + (" and -o bytes" if trace else "")}


def wl_msa(ctx, name, steps):
    kind, shape, trace, prm, desc = WORKLOADS[name]
    nreads, ln = shape
    reads = make_msa_reads(nreads, ln, ctx.rank)
    d = tempfile.mkdtemp()
    fa = os.path.join(d, "in.fa")
    write_fasta(fa, reads)
    ctx.flush = None
    # TSTA_msa is its own process: this process gives back what it keeps cached on the GPU and in page-locked host
    # memory first (plans and batch buffers of the earlier workloads, the pinned batch arrays, torch's allocator)
    _BATCH_CACHE.clear()
    ctx.api.load().tsta_psa_release_cache()
    import gc
    gc.collect()
    ctx.torch.cuda.empty_cache()
    steps = max(1, min(steps, 3))
    run_msa_once(fa, os.path.join(d, "o.txt"), ctx.local_rank)          # warm-up (one is enough: a fresh process each time)
    ctx.barrier()
    # (nvidia-smi polling takes driver locks; the MSA driver makes many small CUDA calls per read, so it is sampled at 1 Hz)
    sampler = ClockSampler(ctx.local_rank, 1000)
    sampler.start()
    ks, ws, ls = [], [], []
    for _ in range(steps):
        wall_s, kern_s, cells_rank, loop_s = run_msa_once(fa, os.path.join(d, "o.txt"), ctx.local_rank)
        ks.append(kern_s)
        ws.append(wall_s)
        ls.append(loop_s)
    ctx.barrier()
    clocks = sampler.stop()
    step_ms, e2e_s, wall_s = ctx.allmax([1e3 * sum(ks) / steps, sum(ls) / steps, sum(ws) / steps])
    cells_total = ctx.world * cells_rank
    result = {"out_md5": hashlib.md5(open(os.path.join(d, "o.txt"), "rb").read()).hexdigest()}
    # algorithmic work of the graph DP: 2*in + 5 int32 lane-ops per cell, i.e.
    # 2 per edge-cell + 5 per node-cell; int32 kernel -> held against the measured VIADDMNMX rate
    ec, nc, fm = MSA_LAST.get("edge_cells"), MSA_LAST.get("node_cells"), MSA_LAST.get("fill_ms")
    ach = (2.0 * ec + 5.0 * nc) / (fm * 1e-3) / 1e12 if ec and nc and fm else None
    roof = {"bound": "int_alu", "bound_note": "int32 ALU issue; the kernel is bound by one warp's issue rate and the per-row scan latency (one warp per band, DESIGN.md 6)",
            "achieved": ach, "peak": ctx.dpx_peak / 1e12, "unit": "Tlaneop/s", "frac": (ach / (ctx.dpx_peak / 1e12)) if ach else None,
            "peak_source": "measured live: dependency-free VIADDMNMX loop (tsta_measure_dpx_peak)",
            "algorithmic_ops": "2 per edge-cell + 5 per node-cell (SURVEY.md 8d)", "edge_cells": ec, "node_cells": nc,
            "kernel": "poa_rows_kernel", "kernel_ms": fm, "fill_gcups": cells_rank / (fm * 1e-3) / 1e9 if fm else None, "traffic": None}
    # parity: OUR driver on the sample the reference arm is timed on (same generator, 10 kbp reads), -o md5 against
    # the reference's (the full 16 x 50 kbp comparison is tests/test_full_size_gpu.py)
    last = dict(MSA_LAST)            # (the parity run below is another TSTA_msa job: it overwrites MSA_LAST)
    parity = None
    if ctx.rank == 0 and ctx.check_parity:
        parity = msa_parity(nreads, min(ln, 10_000), ctx.local_rank)
    return {"value": cells_total / (step_ms * 1e-3) / 1e9, "ms_per_step": step_ms, "steps": steps, "scaling": "weak",
            "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS",
                    "h2d_bytes_per_step": export_bytes(last.get("export"), int(sum(len(r) for r in reads[1:]) + 30 * cells_rank / ln)),   # counted by the library (reads + changed node records + the permutation per read)
                    "d2h_bytes_per_step": int(12 * cells_rank / ln),                                      # lastsorce + path steps per round
                    "ms_per_step": e2e_s * 1e3,
                    "is": "driver loop per read: graph export + H2D + fill + walk + D2H path + graph fusion + re-sort "
                          "(process start-up / CUDA context creation excluded; whole-process wall in wall_s)"},
            "gpu_launches": steps * 6 * (nreads - 1), "roofline": roof, "clocks": clocks,
            "config": {"reads": nreads, "read_length": ln, "node_cells_per_gpu": int(cells_rank)},
            "impl": {"parallelism": f"replicas only: one MSA job per GPU x {ctx.world}",
                     "value_is": "CUDA-event time of the POA fill + walk kernels (TSTA_msa driver, TSTA_MSA_TIMING=1)"},
            "result": result, "parity": parity, "dtype": "int32",
            "extra": {"kernel_ms": step_ms, "fill_ms": last.get("fill_ms"), "walk_ms": last.get("walk_ms"),
                      "wall_s": wall_s, "driver_loop_s_per_run": ls, "process_wall_s_per_run": ws, "export": last.get("export")}}


def msa_parity(nreads, ln_s, device):
    from oracle import binding as ob
    if ob.ref_isa() is None:
        return {"checked": False, "why": "oracle/_ref not built"}
    reads = make_msa_reads(nreads, ln_s, 0)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa")
        write_fasta(fa, reads)
        run_msa_once(fa, os.path.join(d, "ours.txt"), device)
        ob.run_ref("TSTA_msa", ["-T", os.cpu_count() or 4, "-W", 10, "-i", fa, "-o", os.path.join(d, "ref.txt")], timeout=3000)
        ours = hashlib.md5(open(os.path.join(d, "ours.txt"), "rb").read()).hexdigest()
        ref = hashlib.md5(open(os.path.join(d, "ref.txt"), "rb").read()).hexdigest()
    return {"checked": True, "ok": ours == ref, "compared": f"-o md5 of TSTA_msa on {nreads} x {ln_s} bp against the reference binary",
            "out_md5": ours, "reference_out_md5": ref}


def export_bytes(line, fallback):
    """`export: N bytes ...` line of TSTA_msa (TSTA_MSA_TIMING=1): what tsta_poa_* copied to the device."""
    try:
        return int(line.split()[1])
    except Exception:
        return fallback


RUNNERS = {"pair": wl_pair, "bands": wl_bands, "batch": wl_batch, "batch_strong": wl_batch, "msa": wl_msa}


def pair_parity(name, res):
    """maxsorce (and, when the reference can hold the matrices, the -o file) of a single-pair workload
    against the unmodified reference binary on the same pair."""
    from oracle import binding as ob
    from tsta_b200 import synth
    kind, shape, trace, prm, _ = WORKLOADS[name]
    if ob.ref_isa() is None:
        return {"checked": False, "why": "oracle/_ref not built"}
    n, m = shape
    x, y = synth.random_read(n, 1), synth.random_read(m, 2)
    threads = os.cpu_count() or 1
    out = dict(res.get("parity") or {})
    full_trace = trace and 3 * n * m <= 0.5 * host_ram_bytes()
    _, ms, md5 = _ref_psa_run(x, y, full_trace, 30, threads, prm)
    out["checked"] = True
    out["maxsorce_equals_reference"] = ms == res["result"]["maxsorce"]
    out["reference_maxsorce"] = ms
    ok = out["maxsorce_equals_reference"]
    if trace:
        if full_trace:
            out["out_md5_equals_reference"] = md5 == res["result"]["out_md5"]
            ok = ok and out["out_md5_equals_reference"]
        else:
            out["out_md5_equals_reference"] = None
            out["why_no_md5"] = "the reference needs %d GB for this pair; its alignment is pinned by the self-consistency checks" % (3 * n * m >> 30)
        ok = ok and out.get("alignment_spells_both_reads", True) and out.get("rescored_alignment_equals_corner", True)
    out["ok"] = bool(ok)
    return out


def summarize(res):
    rf = res["roofline"]
    step_frac = None
    if rf.get("frac") and rf.get("kernel_ms") and res["ms_per_step"]:
        step_frac = rf["frac"] * rf["kernel_ms"] / res["ms_per_step"]        # all kernels of the step (fill + traceback) against the same peak
    return {"value": res["value"], "unit": "GCUPS", "ms_per_step": res["ms_per_step"], "steps": res["steps"],
            "scaling": res["scaling"], "kernel_ms": res["roofline"].get("kernel_ms"),
            "e2e": res["e2e"]["value"], "e2e_ms_per_step": res["e2e"]["ms_per_step"],
            "roofline_frac": res["roofline"]["frac"], "roofline_frac_whole_step": step_frac, "gpu_launches": res["gpu_launches"],
            "result": res["result"], "parity": res["parity"],
            "parity_checked": bool(res["parity"] and res["parity"].get("checked", "equals_one_gpu_run" in res["parity"])),
            "config": res["config"], "impl": res["impl"], "extra": res["extra"]}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference runs (cpu_baseline and parity checks)")
    ap.add_argument("--no-suite", action="store_true", help="headline only (no `workloads` block)")
    a = ap.parse_args()

    # Libraries (NCCL's version banner, ...) may write to stdout; the contract is ONE JSON line
    # there, so everything else goes to stderr until the line is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    headline = a.workload or "pair100k"
    kind, shape, trace, prm, desc = WORKLOADS[headline]
    base_config = {"workload": desc, "params": prm, "l2": "flushed between timed steps (256 MiB memset)"}

    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        cb = cpu_reference(headline, a.steps)
        n_s = cb["sample_shape"]
        cfg = dict(base_config)
        if kind in ("pair", "bands"):
            cfg["cells_per_step_per_gpu" if kind == "pair" else "cells_per_step"] = int(shape[0]) * int(shape[1])
        line = {"impl": "reference", "metric": "GCUPS", "value": cb["value"], "unit": "GCUPS",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "steps_run": cb["runs"],
                "steps_note": "the reference arm ran %d whole-process run(s) of its sample (see cpu_baseline.sample); "
                              "ms_per_step is their mean" % cb["runs"],
                "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": cfg,
                "result": {"maxsorce": cb["maxsorce"], "out_md5": cb["out_md5"], "sample_shape": list(n_s)},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    ctx = Ctx(a)
    res = RUNNERS[kind](ctx, headline, a.steps)
    line = None
    if ctx.rank == 0:
        cfg = dict(base_config)
        cfg.update(res["config"])
        line = {"metric": "GCUPS", "value": res["value"], "unit": "GCUPS", "n_gpus": ctx.world, "steps": res["steps"],
                "warmup": ctx.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": res["scaling"], "vs_baseline": None, "dtype": res["dtype"],
                "data": "synthetic", "config": cfg, "impl_config": res["impl"], "result": res["result"],
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "roofline": res["roofline"]}
        line.update(res["extra"])
        if ctx.world == 1 and not a.no_cpu_baseline:
            try:
                cb = cpu_reference(headline, 2)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                par = dict(res["parity"] or {})
                if kind == "pair" and cb["kind"] == "reference" and tuple(cb["sample_shape"]) == tuple(shape):
                    par.update(checked=True, reference_maxsorce=cb["maxsorce"],
                               maxsorce_equals_reference=cb["maxsorce"] == res["result"]["maxsorce"])
                    ok = par["maxsorce_equals_reference"]
                    if trace:
                        par["out_md5_equals_reference"] = cb["out_md5"] == res["result"]["out_md5"]
                        ok = ok and par["out_md5_equals_reference"]
                    par["ok"] = bool(ok)
                elif kind == "pair":
                    par = pair_parity(headline, res)
                line["parity"] = par or None
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "failed: " + str(e)[:200]}
        elif res["parity"]:
            line["parity"] = res["parity"]
        line["parity_checked"] = bool(line.get("parity") and line["parity"].get("checked", "equals_one_gpu_run" in line["parity"]))

    # ---- the other workloads, measured in the same process -------------------------------------
    if a.workload is None and not a.no_suite:
        suite = SUITE_1GPU if ctx.world == 1 else SUITE_NGPU
        block = {}
        t_suite0 = time.perf_counter()
        for name in suite:
            k2 = WORKLOADS[name][0]
            st = max(2, min(a.steps, 3 if name in ("pair200k_trace", "batch10k_strong", "batch250_strong") else 5))
            t0 = time.perf_counter()
            try:
                r = RUNNERS[k2](ctx, name, st)
                if ctx.rank == 0:
                    if k2 == "pair" and ctx.check_parity:
                        r["parity"] = pair_parity(name, r)
                    block[name] = summarize(r)
                    block[name]["description"] = WORKLOADS[name][4]
                    block[name]["params"] = WORKLOADS[name][3]
                    block[name]["bench_wall_s"] = time.perf_counter() - t0
            except Exception as e:      # one workload must not take the headline down with it
                if ctx.rank == 0:
                    block[name] = {"error": repr(e)[:300]}
                ctx.torch.cuda.synchronize()
        if ctx.rank == 0:
            line["workloads"] = block
            line["workloads_wall_s"] = time.perf_counter() - t_suite0
            line["workloads_note"] = ("same process, same timing rules as the headline; value = device-timed kernels, e2e = host buffers through "
                                      "the C ABI; parity = comparison with the unmodified reference binary (oracle/_ref) or, for the sliced pair, "
                                      "with the one-GPU run of the same pair")
    if ctx.rank == 0:
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
