"""One long pair as column slices over several GPUs (tsta_psa_plan_upload_slice): the boundary
column crosses NVLink as tagged elements; results must equal the single-GPU run and the oracle."""
import numpy as np
import pytest

from oracle import binding as ob
from tsta_b200 import synth


def test_slice_columns_host_logic():
    from tsta_b200 import api
    for n in (1, 1535, 1536, 1537, 50_000, 100_000, 303_200):
        for parts in (1, 2, 3, 4, 8):
            sl = api.slice_columns(n, parts)
            assert 1 <= len(sl) <= parts
            assert sl[0][0] == 0 and sum(w for _, w in sl) == n
            for (c0, w), (c1, _) in zip(sl, sl[1:]):
                assert c0 + w == c1 and w % 1536 == 0      # every slice but the last: whole bands
            assert all(w > 0 for _, w in sl)
    assert api.slice_columns(0, 4) == []


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,parts", [(20_000, 3000, 2), (100_000, 2500, 2), (60_000, 4000, 4)])
def test_sliced_pair_equals_oracle(gpu, n, m, parts):
    if gpu.device_count() < parts:
        pytest.skip(f"needs {parts} GPUs")
    x, y = synth.read_pair(n, m, 77, None)
    s0, s1 = ob.order_pair(x, y)
    o = ob.psa(x, y, trace=False, want_lastrow=True)
    slices = gpu.slice_columns(len(s0), parts)
    plans = []
    for g, (c0, w) in enumerate(slices):
        gpu.set_device(g)
        p = gpu.PsaPlan(False)
        p.upload_slice(s0, s1, c0, w)
        plans.append(p)
    for g in range(len(plans) - 1):
        plans[g].connect_right_local(plans[g + 1])
    for rep in range(3):                      # the run counter is the tag: all slices in lockstep
        for p in plans:
            p.launch()
        for p in plans:
            p.wait()
    res = [p.fetch() for p in plans]
    assert max(r.maxsorce for r in res) == o.maxsorce
    assert res[-1].corner == o.corner
    last = np.concatenate([p.lastrow() for p in plans])
    np.testing.assert_array_equal(last, o.lastrow)
    for p in plans:
        p.close()
    gpu.set_device(0)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,parts,div", [(6000, 5000, 2, 0.1), (40_000, 3000, 2, None), (20_000, 18_000, 2, 0.05),
                                           (30_000, 2500, 4, 0.1)])
def test_sliced_pair_traceback_equals_oracle(gpu, n, m, parts, div):
    """Traced slices: every GPU keeps the direction matrix of its columns; the walk goes right to
    left through them (tsta_psa_plan_trace_from), the -o bytes must equal the oracle's."""
    if gpu.device_count() < parts:
        pytest.skip(f"needs {parts} GPUs")
    x, y = synth.read_pair(n, m, 91, div)
    s0, s1 = ob.order_pair(x, y)
    o = ob.psa(x, y)
    slices = gpu.slice_columns(len(s0), parts)
    plans = []
    for g, (c0, w) in enumerate(slices):
        gpu.set_device(g)
        p = gpu.PsaPlan(True)
        p.upload_slice(s0, s1, c0, w)
        plans.append(p)
    for g in range(len(plans) - 1):
        plans[g].connect_right_local(plans[g + 1])
    for p in plans:
        p.launch()
    for p in plans:
        p.wait()
    res = [p.fetch() for p in plans]
    assert max(r.maxsorce for r in res) == o.maxsorce and res[-1].corner == o.corner
    moves, i, forced = [], len(s1) - 1, 0
    for p in reversed(plans):
        mv, i, forced, _ = p.trace_from(i, forced)
        moves.append(mv)
        if i < 0:
            break
    a, b = gpu.assemble_alignment(s0, s1, np.concatenate(moves))
    assert b">1\n" + a + b"\n>2\n" + b == o.output_bytes()
    for p in plans:
        p.close()
    gpu.set_device(0)


@pytest.mark.gpu
def test_trace_from_on_a_whole_pair(gpu):
    """trace_from / assemble_alignment on an unsliced traced plan reproduce fetch()."""
    x, y = synth.read_pair(3000, 2800, 5, 0.08)
    s0, s1 = ob.order_pair(x, y)
    p = gpu.PsaPlan(True)
    p.upload(x, y)
    p.run()
    r = p.fetch()
    mv, i, forced, _ = p.trace_from(len(s1) - 1, 0)
    a, b = gpu.assemble_alignment(s0, s1, mv)
    assert (a, b) == (r.a, r.b)
    p.close()
