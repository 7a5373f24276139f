#!/usr/bin/env python
"""bench.py -- GCUPS of the global-alignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload
(N=1): BASELINE.json configs[1] -- one synthetic 100 kbp x 100 kbp DNA pair, affine gaps,
score-only (TSTA_psa_notrace).  Under torchrun every rank aligns its own independent pair
(round-robin batch sharding, no data-path collective): weak scaling.

  value     whole-job GCUPS with the reads already resident in HBM (kernel time from CUDA events
            on the library's stream, max over ranks)
  e2e       same metric through the one-shot C entry point tsta_psa_align with HOST buffers:
            H2D of the reads, kernels, D2H of the result, every step
  roofline  fill kernel against the MEASURED int32 DPX issue rate (tsta_measure_dpx_peak):
            7 algorithmic lane-ops per cell; traceback workloads also report
            the packed-direction write rate against measured HBM bandwidth
  cpu_baseline / --impl reference
            the reference's own AVX-512/AVX2/SSE + pthread implementation (oracle/_ref, built
            from the unmodified reference sources) on this box's host cores, same workload
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

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (n, m, trace, divergence, description)
    "pair100k": (100_000, 100_000, False, None,
                 "TSTA_psa_notrace: one synthetic 100 kbp x 100 kbp pair per GPU, affine (M2 X-5 E-2 O-4), score-only"),
    "pair100k_trace": (100_000, 100_000, True, None,
                       "TSTA_psa: one synthetic 100 kbp x 100 kbp pair per GPU, affine, 4-bit directions in HBM + traceback"),
    "pair200k_trace": (200_000, 200_000, True, None,
                       "TSTA_psa: one synthetic 200 kbp x 200 kbp pair per GPU, affine, 20 GB packed directions + traceback"),
    "pair20k": (20_000, 20_000, False, None, "small single pair (debug)"),
}
PARAMS = dict(M=2, X=-5, E=-2, O=-4)
LANE_OPS_PER_CELL = 7  # 1 select, 1 max(e,f), 1 viaddmax, 1 add, 2 viaddmax, 1 running max


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

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
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


def make_pair(name: str, rank: int):
    from tsta_b200 import synth
    n, m, trace, div, _ = WORKLOADS[name]
    if div is None:
        x, y = synth.random_read(n, 2 * rank + 1), synth.random_read(m, 2 * rank + 2)
    else:
        x = synth.random_read(n, 2 * rank + 1)
        y = synth.mutate(x, div, rank)
    return x, y


# ------------------------------------------------------------------------------------------
# reference arm: the unmodified reference binary on this box's host cores
# ------------------------------------------------------------------------------------------
def run_reference_once(name: str, x: bytes, y: bytes, W: int, threads: int) -> float:
    """Wall seconds of one reference process on the pair (includes its FASTA read + alloc)."""
    from oracle import binding as ob
    _, _, trace, _, _ = WORKLOADS[name]
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


def reference_sample(name: str):
    """(sample workload name, note): the reference needs 3 bytes per cell for traceback, so
    traced workloads are sampled at a size that fits host RAM."""
    n, m, trace, _, _ = WORKLOADS[name]
    if not trace:
        return n, m, "the full workload (1 pair, %d x %d)" % (n, m)
    try:
        avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except Exception:
        avail = 32 << 30
    side = n
    while 3 * side * side > 0.5 * avail:
        side //= 2
    return side, side, "a %d x %d pair (the reference keeps 3 bytes per cell in host RAM)" % (side, side)


def cpu_reference(name: str, steps: int, warmup: int):
    from oracle import binding as ob
    from tsta_b200 import synth
    threads = os.cpu_count() or 1
    isa = ob.ref_isa()
    n, m, note = reference_sample(name)
    x, y = synth.random_read(n, 1), synth.random_read(m, 2)
    if isa is not None:
        try:
            best = None
            for W in (30, 10):
                run_reference_once(name, x[:2000], y[:2000], W, threads)  # page in the binary
                times = [run_reference_once(name, x, y, W, threads) for _ in range(max(1, warmup and 1) + steps)][1:] \
                    or [run_reference_once(name, x, y, W, threads)]
                t = sum(times) / len(times)
                if best is None or t < best[0]:
                    best = (t, W)
            t, W = best
            return {"value": n * m / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "reference",
                    "sample": f"{note}; reference {isa} build, -T {threads} -W {W}, whole-process wall time",
                    "seconds": t}
        except Exception as e:  # fall through to the port
            err = str(e)[:120]
    else:
        err = "oracle/_ref not built"
    # port: the scalar C oracle on a bounded sample
    side = 12_000
    x, y = synth.random_read(side, 1), synth.random_read(side, 2)
    ob.psa(x[:500], y[:500], trace=False)
    t0 = time.perf_counter()
    ob.psa(x, y, trace=False, **PARAMS)
    t = time.perf_counter() - t0
    return {"value": side * side / t / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
            "sample": f"{side} x {side} pair through the scalar C oracle ({err})", "seconds": t}


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

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n, m, trace, _, desc = WORKLOADS[a.workload]
    config = {"workload": desc, "cells_per_step_per_gpu": n * m, "params": PARAMS,
              "parallelism": f"independent pairs, 1 per GPU x {world}" if world > 1 else "1 GPU",
              "l2": "flushed between timed steps (256 MiB memset)"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(a.workload, max(1, a.steps), a.warmup)
        line = {"impl": "reference", "metric": "GCUPS", "value": cb["value"], "unit": "GCUPS",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
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

    x, y = make_pair(a.workload, rank)
    plan = api.PsaPlan(trace)
    plan.upload(x, y, **PARAMS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    for _ in range(max(3, a.warmup)):
        plan.run()
    dpx_peak, _ = api.measure_dpx_peak()

    sampler = ClockSampler(local_rank)
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

    step_ms = (sum(fill_ms) + sum(trace_ms)) / a.steps
    # max over ranks of the device time
    tt = torch.tensor([step_ms, sum(fill_ms) / a.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms_max, fill_ms_max = float(tt[0]), float(tt[1])
    cells = n * m
    value = world * cells / (step_ms_max * 1e-3) / 1e9

    # end to end through the one-shot C entry point (host buffers in, host result out)
    for _ in range(2):
        api.psa(x, y, trace=trace, **PARAMS)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(a.steps, 10))
    for _ in range(e2e_steps):
        r2 = api.psa(x, y, trace=trace, **PARAMS)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * cells / float(te[0]) / 1e9
    assert (r2.maxsorce, r2.corner) == (res.maxsorce, res.corner)
    d2h = 8 + (2 * len(r2.a) if trace and r2.a else 0)

    if rank == 0:
        pk = peaks()
        fill_rate = cells / (fill_ms_max * 1e-3)
        achieved = fill_rate * LANE_OPS_PER_CELL / 1e12
        roof = {"bound": "int32 DPX/ALU issue (not hbm, not tensor: SURVEY.md 8d)", "achieved": achieved,
                "peak": dpx_peak / 1e12, "unit": "Tlaneop/s", "frac": achieved / (dpx_peak / 1e12),
                "peak_source": "measured live: dependency-free VIADDMNMX loop (tsta_measure_dpx_peak)",
                "algorithmic_ops_per_cell": LANE_OPS_PER_CELL, "kernel": "psa_pair_kernel",
                "kernel_ms": fill_ms_max, "traffic": None}
        if trace:
            hbm = pk.get("hbm_gbs", 6650.0)
            roof["hbm_write"] = {"achieved": 0.5 * fill_rate / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": 0.5 * fill_rate / 1e9 / hbm, "bytes_per_cell": 0.5,
                                 "peak_source": "MEASURED_PEAKS.json" if pk else "fallback 6650"}
        line = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": a.steps,
                "warmup": max(3, a.warmup), "ms_per_step": step_ms_max, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": dict(config, cols_per_lane=info["cols_per_lane"], bands=info["bands"]),
                "result": {"maxsorce": res.maxsorce, "corner": res.corner},
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": len(x) + len(y),
                        "d2h_bytes_per_step": d2h, "ms_per_step": float(te[0]) * 1e3},
                "gpu_launches": a.steps * info["launches"],
                "clocks": clocks, "roofline": roof,
                "fill_ms": fill_ms_max, "trace_ms": step_ms_max - fill_ms_max, "wall_s": wall}
        if world == 1 and not a.no_cpu_baseline:
            cb = cpu_reference(a.workload, 2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
