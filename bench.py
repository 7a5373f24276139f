#!/usr/bin/env python
"""bench.py -- GCUPS of the global-alignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload
(N=1): BASELINE.json configs[1] -- one synthetic 100 kbp x 100 kbp DNA pair, affine gaps,
score-only (TSTA_psa_notrace).  Under torchrun every rank aligns its own independent pair
(round-robin sharding of independent units, no data-path collective): weak scaling.

Workloads (--workload):
  pair100k          configs[1]: one 100 kbp pair per GPU, score-only            (default)
  pair100k_bands    configs[1], multi-GPU form: ONE pair as column slices over the N GPUs,
                    boundary column by peer stores over NVLink (strong scaling)
  pair400k_bands    the same for a 400 kbp x 100 kbp pair (too many bands for one GPU's sub-partitions)
  pair100k_trace    one 100 kbp pair per GPU with 4-bit directions in HBM + traceback
  pair200k_trace    configs[2]: 200 kbp pair, 20 GB of packed directions + traceback
  batch250          configs[3]: 1,000,000 pairs of 250 bp per GPU, score-only (tsta_psa_batch)
  batch10k          configs[3]: 10,000 pairs of 10 kbp per GPU, score-only
  batch250_trace / batch10k_trace   the same with traceback (batch10k_trace: 2,000 pairs)
  msa16x50k         configs[4]: TSTA_msa graph DP, 16 reads x 50 kbp (one job per GPU)

  value     whole-job GCUPS with the inputs already resident in HBM where the ABI allows it
            (plan API: kernel time from CUDA events on the library's stream, max over ranks;
            batch / msa: the kernels' CUDA-event time reported by the library)
  e2e       same metric through the host-buffer C entry point (tsta_psa_align / tsta_psa_batch /
            the TSTA_msa driver loop): H2D of the reads, kernels, D2H of the results, every step
  roofline  fill kernel against the MEASURED int32 DPX issue rate (tsta_measure_dpx_peak):
            7 algorithmic lane-ops per cell; traceback workloads also report
            the packed-direction write rate against measured HBM bandwidth
  cpu_baseline / --impl reference
            the reference's own AVX-512/AVX2/SSE + pthread implementation (oracle/_ref, built
            from the unmodified reference sources) on this box's host cores, same workload
            (bounded sample for the batch / msa shapes)
"""
from __future__ import annotations

import argparse
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

PARAMS = dict(M=2, X=-5, E=-2, O=-4)
LANE_OPS_PER_CELL = 7  # 1 select, 1 max(e,f), 1 viaddmax, 1 add, 2 viaddmax, 1 running max

# name: kind, shape, trace, description
WORKLOADS = {
    "pair100k": ("pair", (100_000, 100_000), False,
                 "TSTA_psa_notrace: one synthetic 100 kbp x 100 kbp pair per GPU, affine (M2 X-5 E-2 O-4), score-only"),
    "pair100k_bands": ("bands", (100_000, 100_000), False,
                       "TSTA_psa_notrace: ONE synthetic 100 kbp x 100 kbp pair as column slices over all GPUs (NVLink boundary hand-off)"),
    "pair400k_bands": ("bands", (400_000, 100_000), False,
                       "TSTA_psa_notrace: ONE synthetic 400 kbp x 100 kbp pair as column slices over all GPUs (more bands than one GPU has sub-partitions)"),
    "pair100k_trace": ("pair", (100_000, 100_000), True,
                       "TSTA_psa: one synthetic 100 kbp x 100 kbp pair per GPU, affine, 4-bit directions in HBM + traceback"),
    "pair200k_trace": ("pair", (200_000, 200_000), True,
                       "TSTA_psa: one synthetic 200 kbp x 200 kbp pair per GPU, affine, 20 GB packed directions + traceback"),
    "pair20k": ("pair", (20_000, 20_000), False, "small single pair (debug)"),
    "batch250": ("batch", (1_000_000, 250), False,
                 "batched PSA: 1,000,000 synthetic 250 bp pairs per GPU (half unrelated, half 5 % divergent), affine, score-only"),
    "batch10k": ("batch", (10_000, 10_000), False,
                 "batched PSA: 10,000 synthetic 10 kbp pairs per GPU, affine, score-only"),
    "batch250_trace": ("batch", (1_000_000, 250), True, "batched PSA with traceback: 1,000,000 pairs of 250 bp per GPU"),
    "batch10k_trace": ("batch", (2_000, 10_000), True, "batched PSA with traceback: 2,000 pairs of 10 kbp per GPU"),
    "msa16x50k": ("msa", (16, 50_000), True,
                  "TSTA_msa: partial-order alignment of 16 synthetic 50 kbp reads (5 % divergence), one job per GPU"),
    "msa16x10k": ("msa", (16, 10_000), True, "TSTA_msa: 16 x 10 kbp (debug)"),
}


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


def make_batch(npairs: int, ln: int, seed: int):
    """Rectangular batch: even pairs unrelated, odd pairs 5 %-substituted copies."""
    from tsta_b200 import synth
    a = np.frombuffer(synth.random_read(npairs * ln, 2 * seed + 11), dtype=np.uint8).reshape(npairs, ln)
    b = np.frombuffer(synth.random_read(npairs * ln, 2 * seed + 12), dtype=np.uint8).reshape(npairs, ln).copy()
    rnd = np.random.default_rng(seed)
    odd = np.arange(npairs) % 2 == 1
    keep = rnd.random((int(odd.sum()), ln)) >= 0.05
    b[odd] = np.where(keep, a[odd], b[odd])
    off = np.arange(npairs, dtype=np.int64) * ln
    lens = np.full(npairs, ln, dtype=np.int32)
    return a.reshape(-1).copy(), off, lens, b.reshape(-1).copy(), off.copy(), lens.copy()


def make_msa_reads(nreads: int, ln: int, seed: int):
    from tsta_b200 import synth
    anc = synth.random_read(ln, 11 + seed)
    return [anc] + [synth.mutate(anc, 0.05, 200 + 16 * seed + k)[:ln] for k in range(nreads - 1)]


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for k, r in enumerate(reads):
            f.write(b">%d\n" % (k + 1) + r + b"\n")


# ------------------------------------------------------------------------------------------
# reference arm: the unmodified reference binaries on this box's host cores
# ------------------------------------------------------------------------------------------
def _ref_psa_seconds(x: bytes, y: bytes, trace: bool, W: int, threads: int) -> float:
    from oracle import binding as ob
    with tempfile.TemporaryDirectory() as d:
        f1, f2 = os.path.join(d, "a.fa"), os.path.join(d, "b.fa")
        ob.write_fasta(f1, [x])
        ob.write_fasta(f2, [y])
        args = ["-M", PARAMS["M"], "-X", PARAMS["X"], "-E", PARAMS["E"], "-O", PARAMS["O"],
                "-T", threads, "-W", W, "-1", f1, "-2", f2]
        if trace:
            args += ["-o", os.path.join(d, "o.txt")]
        t0 = time.perf_counter()
        p = ob.run_ref("TSTA_psa" if trace else "TSTA_psa_notrace", args, timeout=900)
        dt = time.perf_counter() - t0
        if b"maxsorce=" not in p.stdout:
            raise RuntimeError("reference run failed: " + p.stderr.decode()[-200:])
    return dt


def cpu_reference(name: str, steps: int):
    """Times the reference (oracle/_ref) on a bounded sample of the workload; falls back to the
    scalar C port of the oracle when the reference binaries are not present."""
    from oracle import binding as ob
    from tsta_b200 import synth
    kind, shape, trace, _ = WORKLOADS[name]
    threads = os.cpu_count() or 1
    isa = ob.ref_isa()
    steps = max(1, min(steps, 3))
    if isa is None:
        side = 12_000
        x, y = synth.random_read(side, 1), synth.random_read(side, 2)
        ob.psa(x[:500], y[:500], trace=False)
        t0 = time.perf_counter()
        ob.psa(x, y, trace=False, **PARAMS)
        t = time.perf_counter() - t0
        return {"value": side * side / t / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
                "sample": f"{side} x {side} pair through the scalar C oracle (oracle/_ref not built)", "seconds": t}
    if kind in ("pair", "bands"):
        n, m = shape
        note = "the full workload (1 pair, %d x %d)" % (n, m)
        if trace:   # the reference keeps 3 bytes per cell in host RAM
            try:
                avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
            except Exception:
                avail = 32 << 30
            while 3 * n * m > 0.5 * avail:
                n //= 2
                m //= 2
            note = "a %d x %d pair (the reference keeps 3 bytes per cell in host RAM)" % (n, m)
        x, y = synth.random_read(n, 1), synth.random_read(m, 2)
        best = None
        for W in (30, 10):
            _ref_psa_seconds(x[:2000], y[:2000], trace, W, threads)      # page in the binary
            ts = [_ref_psa_seconds(x, y, trace, W, threads) for _ in range(steps)]
            t = sum(ts) / len(ts)
            if best is None or t < best[0]:
                best = (t, W)
        t, W = best
        return {"value": n * m / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
                "sample": f"{note}; reference {isa} build, -T {threads} -W {W}, whole-process wall time", "seconds": t}
    if kind == "batch":
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
                args = ["-M", PARAMS["M"], "-X", PARAMS["X"], "-E", PARAMS["E"], "-O", PARAMS["O"],
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
        return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
                "sample": f"{total} pairs of {ln} bp in {t:.1f} s, one reference {isa} process (-T 1 -W {W}) per pair, "
                          f"{procs} processes at a time = {threads} busy threads (the reference reads one pair per process and "
                          f"its main thread spins on the tile barrier; process start-up included)",
                "seconds": t}
    # msa: a smaller job of the same kind (node-cells counted by our own driver's rule)
    nreads, ln = shape
    ln_s = min(ln, 10_000)
    reads = make_msa_reads(nreads, ln_s, 0)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "in.fa")
        write_fasta(fa, reads)
        best = None
        for W in (10, 30):
            t0 = time.perf_counter()
            ob.run_ref("TSTA_msa", ["-T", threads, "-W", W, "-i", fa, "-o", os.path.join(d, "ref.txt")], timeout=3000)
            t = time.perf_counter() - t0
            if best is None or t < best[0]:
                best = (t, W)
        t, W = best
        cells = msa_node_cells(fa)
    return {"value": cells / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
            "sample": f"{nreads} reads x {ln_s} bp (same generator), reference {isa} build, -T {threads} -W {W}, "
                      "whole-process wall time; node-cells counted as sum over reads of graph rows x read length",
            "seconds": t}


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
    return wall, (fill_ms + walk_ms) * 1e-3, cells, loop_s


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="pair100k", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, shape, trace, desc = WORKLOADS[a.workload]
    warmup = max(3, a.warmup)
    config = {"workload": desc, "params": PARAMS,
              "l2": "flushed between timed steps (256 MiB memset)"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(a.workload, a.steps)
        line = {"impl": "reference", "metric": "GCUPS", "value": cb["value"], "unit": "GCUPS",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from tsta_b200 import api

    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dpx_peak, _ = api.measure_dpx_peak()
    # (nvidia-smi polling takes driver locks; the MSA driver makes many small CUDA calls per read, so it is sampled at 1 Hz)
    sampler = ClockSampler(local_rank, 1000 if kind == "msa" else 100)
    extra = {}
    scaling = "weak"
    kernel_name = "psa_pair16_kernel"
    result = {}

    # --------------------------------------------------------------------------------------
    if kind == "pair":
        n, m = shape
        x, y = make_pair(shape, rank)
        plan = api.PsaPlan(trace)
        plan.upload(x, y, **PARAMS)
        for _ in range(warmup):
            plan.run()
        barrier()
        sampler.start()
        fill_ms, trace_ms = [], []
        t_wall0 = time.perf_counter()
        for _ in range(a.steps):
            flush.zero_()
            torch.cuda.synchronize()
            f, t = plan.run()          # CUDA events on the library's stream around the kernels
            fill_ms.append(f)
            trace_ms.append(t)
        barrier()
        wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
        res = plan.fetch()
        info = plan.info()
        step_ms, fill_only = allmax([(sum(fill_ms) + sum(trace_ms)) / a.steps, sum(fill_ms) / a.steps])
        cells_rank = float(n) * m
        cells_total = world * cells_rank
        launches = a.steps * info["launches"]
        kernel_name = ("psa_pair16_kernel" if info["packed"] else "psa_pair_kernel")
        config.update(cells_per_step_per_gpu=int(cells_rank), cols_per_lane=info["cols_per_lane"], bands=info["bands"],
                      packed_int16x2=info["packed"],
                      parallelism=f"independent pairs, 1 per GPU x {world}" if world > 1 else "1 GPU")
        result = {"maxsorce": res.maxsorce, "corner": res.corner}
        # end to end through the one-shot C entry point (host buffers in, host result out)
        for _ in range(2):
            api.psa(x, y, trace=trace, **PARAMS)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(a.steps, 10))
        for _ in range(e2e_steps):
            r2 = api.psa(x, y, trace=trace, **PARAMS)
        barrier()
        (e2e_s,) = allmax([(time.perf_counter() - t0) / e2e_steps])
        assert (r2.maxsorce, r2.corner) == (res.maxsorce, res.corner)
        h2d = len(x) + len(y)
        d2h = 8 + (2 * len(r2.a) if trace and r2.a else 0)
        extra = {"fill_ms": fill_only, "trace_ms": step_ms - fill_only, "wall_s": wall}
        fill_rate = cells_rank / (fill_only * 1e-3)

    # --------------------------------------------------------------------------------------
    elif kind == "bands":
        scaling = "strong"
        n, m = shape
        x, y = make_pair(shape, 0)          # the SAME pair on every rank
        s0, s1 = api.order_pair(x, y)
        slices = api.slice_columns(len(s0), world)
        mine = slices[rank] if rank < len(slices) else None
        plan = api.PsaPlan(False)
        if mine is not None:
            plan.upload_slice(s0, s1, mine[0], mine[1], **PARAMS)
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
        for _ in range(warmup):
            barrier()
            run_all()
        barrier()
        sampler.start()
        tsum = 0.0
        walls = []
        for _ in range(a.steps):
            flush.zero_()
            barrier()                       # all slices start together: slice g only ever waits for g-1
            t0 = time.perf_counter()
            f = run_all()
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
            tsum += f
        barrier()
        clocks = sampler.stop()
        # a slice's kernel time includes waiting for its left neighbours: the job time is the max over ranks
        step_ms, = allmax([tsum / a.steps])
        fill_only = step_ms
        cells_total = float(n) * m
        cells_rank = cells_total / world
        info = plan.info() if mine is not None else {"launches": 0, "cols_per_lane": 0, "bands": 0, "packed": True}
        launches = a.steps * info["launches"]
        res = plan.fetch() if mine is not None else None
        ms_all = allmax([res.maxsorce if res else -10**9])[0]
        corner = allmax([res.corner if (res and rank == len(slices) - 1) else -10**9])[0]
        result = {"maxsorce": int(ms_all), "corner": int(corner)}
        config.update(cells_per_step=int(cells_total), cols_per_lane=info["cols_per_lane"], bands_per_gpu=info["bands"],
                      slices=slices, parallelism=f"column slices over {world} GPU(s), boundary column over NVLink")
        # e2e: upload + run + fetch of the slices with host buffers
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(a.steps, 10))
        for _ in range(e2e_steps):
            if mine is not None:
                plan.upload_slice(s0, s1, mine[0], mine[1], **PARAMS)
            barrier()
            run_all()
            if mine is not None:
                plan.fetch()
        barrier()
        (e2e_s,) = allmax([(time.perf_counter() - t0) / e2e_steps])
        h2d = (mine[1] if mine else 0) + len(s1)
        d2h = 8
        extra = {"fill_ms": fill_only, "wall_ms_per_step": 1e3 * sum(walls) / len(walls)}
        fill_rate = cells_total / (fill_only * 1e-3) / world

    # --------------------------------------------------------------------------------------
    elif kind == "batch":
        npairs, ln = shape
        # inputs and outputs live in page-locked host memory (tsta_host_alloc), as a caller that
        # streams batches would keep them
        pins = [api.pinned_copy(x) for x in make_batch(npairs, ln, rank)]
        s0, o0, l0, s1, o1, l1 = [p.array for p in pins]
        outs = [api.PinnedArray(npairs, np.int32), api.PinnedArray(npairs, np.int32)]
        if trace:
            ops_off = np.arange(npairs, dtype=np.int64) * (2 * ln)
            outs += [api.PinnedArray(npairs * 2 * ln, np.uint8), api.pinned_copy(ops_off), api.PinnedArray(npairs, np.int32)]
        out = tuple(p.array for p in outs)
        for _ in range(warmup):
            r = api.psa_batch_arrays(s0, o0, l0, s1, o1, l1, trace=trace, out=out, **PARAMS)
        barrier()
        sampler.start()
        kms, walls = [], []
        for _ in range(a.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = api.psa_batch_arrays(s0, o0, l0, s1, o1, l1, trace=trace, out=out, **PARAMS)
            walls.append(time.perf_counter() - t0)
            kms.append(r.kernel_ms)
        barrier()
        clocks = sampler.stop()
        step_ms, e2e_s = allmax([sum(kms) / a.steps, sum(walls) / a.steps])
        fill_only = step_ms
        cells_rank = float(npairs) * ln * ln
        cells_total = world * cells_rank
        launches = a.steps * (2 if trace else 1)
        kernel_name = "psa_batch16_kernel"
        result = {"sum_maxsorce": int(r.maxsorce.astype(np.int64).sum()), "sum_corner": int(r.corner.astype(np.int64).sum())}
        config.update(pairs_per_gpu=npairs, read_length=ln, cells_per_step_per_gpu=int(cells_rank),
                      parallelism=f"independent pairs sharded over {world} GPU(s), no collective",
                      value_is="CUDA-event time of the fill (+ traceback) kernels inside tsta_psa_batch")
        h2d = int(s0.nbytes + s1.nbytes + o0.nbytes + o1.nbytes + l0.nbytes + l1.nbytes)
        d2h = int(r.maxsorce.nbytes + r.corner.nbytes + (r._ops.nbytes + r._alen.nbytes if trace else 0))
        extra = {"kernel_ms": step_ms}
        fill_rate = cells_rank / (fill_only * 1e-3)

    # --------------------------------------------------------------------------------------
    else:   # msa
        nreads, ln = shape
        reads = make_msa_reads(nreads, ln, rank)
        d = tempfile.mkdtemp()
        fa = os.path.join(d, "in.fa")
        write_fasta(fa, reads)
        del flush
        torch.cuda.empty_cache()
        steps = max(1, min(a.steps, 3))
        run_msa_once(fa, os.path.join(d, "o.txt"), local_rank)          # warm-up (one is enough: a fresh process each time)
        barrier()
        sampler.start()
        ks, ws, ls = [], [], []
        for _ in range(steps):
            wall_s, kern_s, cells_rank, loop_s = run_msa_once(fa, os.path.join(d, "o.txt"), local_rank)
            ks.append(kern_s)
            ws.append(wall_s)
            ls.append(loop_s)
        barrier()
        clocks = sampler.stop()
        a.steps = steps
        step_ms, e2e_s, wall_s = allmax([1e3 * sum(ks) / steps, sum(ls) / steps, sum(ws) / steps])
        fill_only = step_ms
        cells_total = world * cells_rank
        launches = steps * 2 * (nreads - 1)
        kernel_name = "poa_fill_kernel"
        import hashlib
        result = {"out_md5": hashlib.md5(open(os.path.join(d, "o.txt"), "rb").read()).hexdigest()}
        config.update(reads=nreads, read_length=ln, node_cells_per_gpu=int(cells_rank),
                      parallelism=f"replicas only: one MSA job per GPU x {world}",
                      value_is="CUDA-event time of the POA fill + walk kernels (TSTA_msa driver, TSTA_MSA_TIMING=1)",
                      e2e_is="driver loop per read: graph export + H2D + fill + walk + D2H path + graph fusion + re-sort "
                             "(process start-up / CUDA context creation excluded; whole-process wall in wall_s)")
        h2d = int(sum(len(r) for r in reads[1:]) + 30 * cells_rank / ln)     # reads + CSR graph + descriptors per round
        d2h = int(12 * cells_rank / ln)                                      # lastsorce + path steps per round
        extra = {"kernel_ms": step_ms, "wall_s": wall_s, "driver_loop_s_per_run": ls, "process_wall_s_per_run": ws}
        fill_rate = cells_rank / (fill_only * 1e-3)

    value = cells_total / (step_ms * 1e-3) / 1e9
    e2e_value = cells_total / e2e_s / 1e9

    if rank == 0:
        pk = peaks()
        achieved = fill_rate * LANE_OPS_PER_CELL / 1e12
        if kind == "msa":
            # algorithmic work of the graph DP: 2*in + 5 int32 lane-ops per cell, i.e.
            # 2 per edge-cell + 5 per node-cell; int32 kernel -> held against the measured VIADDMNMX rate
            ec, nc, fm = MSA_LAST.get("edge_cells"), MSA_LAST.get("node_cells"), MSA_LAST.get("fill_ms")
            ach = (2.0 * ec + 5.0 * nc) / (fm * 1e-3) / 1e12 if ec and nc and fm else None
            roof = {"bound": "int_alu", "bound_note": "int32 ALU issue; the kernel is latency bound (one warp per SM sub-partition, DESIGN.md 6)",
                    "achieved": ach, "peak": dpx_peak / 1e12, "unit": "Tlaneop/s", "frac": (ach / (dpx_peak / 1e12)) if ach else None,
                    "peak_source": "measured live: dependency-free VIADDMNMX loop (tsta_measure_dpx_peak)",
                    "algorithmic_ops": "2 per edge-cell + 5 per node-cell (SURVEY.md 8d)", "edge_cells": ec, "node_cells": nc,
                    "kernel": kernel_name, "kernel_ms": fm, "fill_gcups": fill_rate / 1e9, "traffic": None,
                    "note": "profiles/r01_poa_fill_16x20k_bytekeys.md: 636 cycles per row-step of 128 cells for 193 instructions"}
        else:
            # The packed kernels advance two cells per lane-op (int16 x 2), so their issue roofline is
            # twice the measured int32 VIADDMNMX rate; int32 kernels (fallback) are held against the rate itself.
            packf = 2 if config.get("packed_int16x2", True) else 1
            peak = packf * dpx_peak / 1e12
            roof = {"bound": "int_alu", "bound_note": "integer-ALU / DPX issue (neither hbm nor tensor: SURVEY.md 8d)", "achieved": achieved,
                    "peak": peak, "unit": "Tlaneop/s", "frac": achieved / peak,
                    "peak_source": "measured live: dependency-free VIADDMNMX loop (tsta_measure_dpx_peak) = %.2f Tlaneop/s, x%d "
                                   "cells per lane-op for the %s kernels" % (dpx_peak / 1e12, packf, "int16x2" if packf == 2 else "int32"),
                    "algorithmic_ops_per_cell": LANE_OPS_PER_CELL, "kernel": kernel_name,
                    "kernel_ms": fill_only, "traffic": None,
                    "note": "achieved = cells/s x 7 algorithmic lane-ops (SURVEY.md 8d). ncu (profiles/): ALU pipe 38 % active for the "
                            "single 100 kbp pair (latency/ramp bound, DESIGN.md 5), 90-92 % for the batched kernels"}
            try:   # DRAM traffic of the dominant kernel, from the committed ncu capture of this workload
                tr = json.load(open(ROOT / "profiles" / "r01_traffic.json")).get(a.workload)
                if tr and world == 1:
                    roof["traffic"] = tr["bytes"]
                    roof["traffic_source"] = tr["source"]
            except Exception:
                pass
            if trace:
                hbm = pk.get("hbm_gbs", 6650.0)
                roof["hbm_write"] = {"achieved": 0.5 * fill_rate / 1e9, "peak": hbm, "unit": "GB/s",
                                     "frac": 0.5 * fill_rate / 1e9 / hbm, "bytes_per_cell": 0.5,
                                     "peak_source": "MEASURED_PEAKS.json" if pk else "fallback 6650"}
        line = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": a.steps,
                "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "int16x2" if config.get("packed_int16x2", True) and kind != "msa" else "int32",
                "data": "synthetic", "config": config, "result": result,
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof}
        line.update(extra)
        if world == 1 and not a.no_cpu_baseline:
            try:
                cb = cpu_reference(a.workload, 2)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "failed: " + str(e)[:200]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
