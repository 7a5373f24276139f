"""The packed int16x2 fill (csrc/psa_fill16.cuh) against the oracle: every strip width, single
and multi-band pipelines, ramp-only shapes, the int16 range limit and the fall-back to int32."""
import os
import random

import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture
def env():
    saved = {k: os.environ.get(k) for k in ("TSTA_PSA_C16", "TSTA_PSA_PACKED", "TSTA_PSA_CH")}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("c16", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("ch", [16, 32])
def test_every_strip_width_vs_oracle(gpu, env, c16, ch):
    env["TSTA_PSA_C16"] = str(c16)
    env["TSTA_PSA_CH"] = str(ch)
    rnd = random.Random(100 * c16 + ch)
    band = 64 * c16
    shapes = [(1, 1), (5, 3), (band - 1, 40), (band, band), (band + 1, 70), (3 * band + 17, 100),
              (2 * band, 2 * band - 5), (1500, 1400), (4000, 64), (700, 699)]
    plan = gpu.PsaPlan(False)
    for k, (la, lb) in enumerate(shapes):
        div = rnd.choice([None, 0.05, 0.2])
        x, y = synth.read_pair(la, lb, 31000 + 50 * c16 + k, div)
        y = y or b"G"
        for prm in [(2, -5, -2, -4), (2, -5, -2, 0), (1, -1, -1, 0)]:
            o = ob.psa(x, y, *prm, trace=False, want_lastrow=True)
            plan.upload(x, y, *prm)
            plan.run()
            info = plan.info()
            assert info["packed"] and info["cols_per_lane"] == c16
            r = plan.fetch()
            assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner), (la, lb, prm, div)
            np.testing.assert_array_equal(plan.lastrow(), o.lastrow)
    plan.close()


def test_range_limit_falls_back_to_int32(gpu, env):
    x, y = synth.read_pair(900, 850, 5, 0.1)
    plan = gpu.PsaPlan(False)
    for prm, want_packed in [((2, -5, -2, -4), True), ((10, -10, -5, -20), True), ((60, -5, -2, -4), True),
                             ((2, -60, -2, -4), True), ((30, -30, -10, -40), False), ((100, -100, -20, -100), False)]:
        env["TSTA_PSA_C16"] = "4"
        o = ob.psa(x, y, *prm, trace=False, want_lastrow=True)
        plan.upload(x, y, *prm)
        plan.run()
        assert plan.info()["packed"] == want_packed, prm
        r = plan.fetch()
        assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner), prm
        np.testing.assert_array_equal(plan.lastrow(), o.lastrow)
    # five symbols: no packed table
    plan.upload(x + b"N", y)
    plan.run()
    assert not plan.info()["packed"]
    o = ob.psa(x + b"N", y, trace=False)
    assert plan.fetch().maxsorce == o.maxsorce
    plan.close()


def test_steep_scores_stay_inside_int16(gpu, env):
    """Long runs of matches / gaps drive the scores far from zero; the warp base follows them."""
    plan = gpu.PsaPlan(False)
    a = synth.random_read(6000, 3)
    cases = [(a, a), (a, a[:3000]), (a, a[2500:]), (b"A" * 5000, b"A" * 4000 + b"C" * 900), (a, synth.mutate(a, 0.3, 1))]
    for prm in [(2, -5, -2, -4), (60, -5, -2, -4), (10, -10, -5, -20)]:
        for x, y in cases:
            env["TSTA_PSA_C16"] = "3"
            o = ob.psa(x, y, *prm, trace=False, want_lastrow=True)
            plan.upload(x, y, *prm)
            plan.run()
            assert plan.info()["packed"]
            r = plan.fetch()
            assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner), prm
            np.testing.assert_array_equal(plan.lastrow(), o.lastrow)
    plan.close()


def test_packed_and_int32_kernels_agree_at_100k(gpu, env):
    x, y = synth.read_pair(100_000, 100_000, 1)
    plan = gpu.PsaPlan(False)
    out = {}
    for mode in ("1", "0"):
        env["TSTA_PSA_PACKED"] = mode
        plan.upload(x, y)
        plan.run()
        assert plan.info()["packed"] == (mode == "1")
        r = plan.fetch()
        out[mode] = (r.maxsorce, r.corner, plan.lastrow())
    assert out["1"][:2] == out["0"][:2]
    np.testing.assert_array_equal(out["1"][2], out["0"][2])
    plan.close()


def test_packed_batch_vs_oracle(gpu, env):
    rnd = random.Random(17)
    for lens in ([1, 2, 30, 64, 100, 127, 128], [129, 200, 250, 255, 256], [257, 300, 600, 1100, 2300]):
        pairs = []
        for k in range(150 if lens[-1] < 1000 else 40):
            la, lb = rnd.choice(lens), rnd.choice(lens)
            x, y = synth.read_pair(la, lb, 52000 + k, rnd.choice([None, 0.05, 0.15]))
            pairs.append((x, y or b"A"))
        for prm in [(2, -5, -2, -4), (2, -5, -2, 0)]:
            r = gpu.psa_batch(pairs, *prm, trace=False)
            for k, (x, y) in enumerate(pairs):
                o = ob.psa(x, y, *prm, trace=False)
                assert (r.maxsorce[k], r.corner[k]) == (o.maxsorce, o.corner), (k, len(x), len(y), prm)


@pytest.mark.parametrize("lens", [[1, 2, 30, 64, 100, 127, 128], [129, 200, 250, 255, 256]])
def test_dual_pair_mode_with_traceback(gpu, env, lens):
    """Short reads: two pairs per warp (one per int16 half).  Odd pair count, unequal lengths in
    the two halves, traceback, against the oracle and against the one-pair-per-warp path."""
    rnd = random.Random(len(lens))
    pairs = []
    for k in range(101):
        la, lb = rnd.choice(lens), rnd.choice(lens)
        x, y = synth.read_pair(la, lb, 64000 + k, rnd.choice([None, 0.05, 0.15]))
        pairs.append((x, y or b"A"))
    for prm in [(2, -5, -2, -4), (2, -5, -2, 0), (1, -1, -1, 0)]:
        env.pop("TSTA_PSA_DUAL", None)
        r = gpu.psa_batch(pairs, *prm, trace=True)
        env["TSTA_PSA_DUAL"] = "0"
        r1 = gpu.psa_batch(pairs, *prm, trace=True)
        np.testing.assert_array_equal(r.maxsorce, r1.maxsorce)
        np.testing.assert_array_equal(r.corner, r1.corner)
        for k, (x, y) in enumerate(pairs):
            o = ob.psa(x, y, *prm)
            assert (r.maxsorce[k], r.corner[k]) == (o.maxsorce, o.corner), (k, len(x), len(y), prm)
            assert r.output_bytes(k) == o.output_bytes(), (k, len(x), len(y), prm)
            assert r1.output_bytes(k) == o.output_bytes()


@pytest.fixture
def trace_env():
    keys = ("TSTA_PSA_TRACE_MODE", "TSTA_PSA_CORRIDOR", "TSTA_PSA_TILE_BUDGET_MB", "TSTA_PSA_C16", "TSTA_PSA_WALK")
    saved = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _traced(gpu, x, y, prm=(2, -5, -2, -4)):
    plan = gpu.PsaPlan(True)
    plan.upload(x, y, *prm)
    plan.run()
    r = plan.fetch()
    st = plan.trace_stats()
    plan.close()
    return r, st


@pytest.mark.parametrize("walk", ["chase", "serial"])
def test_checkpointed_traceback_leaves_and_reenters_corridor(gpu, trace_env, walk):
    """A 3000-base deletion moves the path 3000 columns sideways: the walk leaves the first corridor
    (margin 1024) and is resumed in new ones; the alignment is the oracle's byte for byte."""
    trace_env["TSTA_PSA_CORRIDOR"] = "1024"
    trace_env["TSTA_PSA_WALK"] = walk
    x = synth.random_read(20_000, 71)
    y = x[:8_000] + x[11_000:]
    o = ob.psa(x, y)
    r, st = _traced(gpu, x, y)
    assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner)
    assert r.output_bytes() == o.output_bytes()
    assert st["corridors"] >= 2 and st["tiles"] < st["tiles_of_matrix"]
    # an insertion instead (the path moves up), and both in one pair
    y2 = x[:5_000] + synth.random_read(2_500, 72) + x[5_000:14_000] + x[16_500:]
    o2 = ob.psa(x, y2)
    r2, st2 = _traced(gpu, x, y2)
    assert r2.output_bytes() == o2.output_bytes() and st2["corridors"] >= 2


def test_checkpointed_traceback_tile_budget(gpu, trace_env):
    """A tile budget far below one corridor: the walk is resumed stage after stage."""
    trace_env["TSTA_PSA_TILE_BUDGET_MB"] = "16"
    for c16 in ("2", "3", "6"):
        trace_env["TSTA_PSA_C16"] = c16
        x, y = synth.read_pair(30_000, 29_000, 81, 0.1)
        o = ob.psa(x, y)
        r, st = _traced(gpu, x, y)
        assert r.output_bytes() == o.output_bytes() and r.maxsorce == o.maxsorce
        assert st["corridors"] >= 2


@pytest.mark.parametrize("mode", ["full", "ckpt", "ckpt_serial"])
def test_trace_modes_agree_with_oracle(gpu, trace_env, mode):
    """The traceback modes of the packed kernels (direction words of the whole matrix / checkpoints +
    corridor with the parallel tile chase / the same with the serial window walker) against the oracle,
    every strip width, linear and affine gaps, ragged shapes."""
    if mode == "full":
        trace_env["TSTA_PSA_TRACE_MODE"] = "full"
    if mode == "ckpt_serial":
        trace_env["TSTA_PSA_WALK"] = "serial"
    rnd = random.Random(7)
    for c16 in (2, 3, 4, 6, 8):
        trace_env["TSTA_PSA_C16"] = str(c16)
        band = 64 * c16
        for k, (la, lb) in enumerate([(1, 1), (band + 1, 70), (3 * band + 17, 300), (1500, 1400), (700, 699), (5000, 4100)]):
            x, y = synth.read_pair(la, lb, 500 + 10 * c16 + k, rnd.choice([None, 0.05, 0.2]))
            y = y or b"G"
            for prm in [(2, -5, -2, -4), (2, -5, -2, 0)]:
                o = ob.psa(x, y, *prm)
                r, st = _traced(gpu, x, y, prm)
                assert (r.maxsorce, r.corner) == (o.maxsorce, o.corner), (mode, c16, la, lb, prm)
                assert r.output_bytes() == o.output_bytes(), (mode, c16, la, lb, prm)
                assert (st["corridors"] >= 1) == (mode != "full")
