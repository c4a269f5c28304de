"""CPU tests of the host-built execution plan of the fused graph-convolution kernel
(mgnns_b200.api.graph_util.fused_plan_arrays; consumed by mgnns_gcn_fused_tc, include/mgnns_b200.h)."""
import numpy as np
import pytest

from mgnns_b200 import synth
from mgnns_b200.api.graph_util import FUSED_SEG_EDGES, FUSED_TILE_ROWS, fused_plan_arrays


def _emulate(plan, rowptr, col, val, x):
    """Execute the plan the way the kernel does (tile -> segments -> rows) in float64."""
    n = rowptr.shape[0] - 1
    y = np.zeros((n, x.shape[1]))
    seen = np.zeros(n, dtype=np.int64)
    ldx4 = plan['ldx4']
    for t in range(plan['n_tiles']):
        z = np.zeros((FUSED_TILE_ROWS, x.shape[1]))
        written = np.zeros(FUSED_TILE_ROWS, dtype=np.int64)
        sole_flags = {}
        for beg, cnt, slot, sole in plan['segs'][plan['tile_seg_ptr'][t]:plan['tile_seg_ptr'][t + 1]]:
            assert 0 <= cnt <= FUSED_SEG_EDGES
            e = plan['edges'][beg:beg + cnt]
            cols = e[:, 0].view(np.uint32).astype(np.int64) // ldx4
            vals = e[:, 1].copy().view(np.float32).astype(np.float64)
            z[slot] += vals @ x[cols] if cnt else 0.0
            written[slot] += 1
            sole_flags.setdefault(int(slot), []).append(bool(sole))
        multi = set(plan['multi_rows'][plan['tile_multi_ptr'][t]:plan['tile_multi_ptr'][t + 1]].tolist())
        assert multi == set(np.nonzero(written > 1)[0].tolist())
        assert (written >= 1).all()                      # every slot is zeroed or written, padding slots too
        for slot, flags in sole_flags.items():           # 'sole' <=> the row has exactly one segment
            assert all(f == (len(flags) == 1) for f in flags)
        for slot in range(FUSED_TILE_ROWS):
            r = plan['tile_rows'][t * FUSED_TILE_ROWS + slot]
            if r >= 0:
                y[r] = z[slot]
                seen[r] += 1
    assert (seen == 1).all()
    return y


@pytest.mark.parametrize("n,mean_degree", [(64, 5), (300, 20), (1000, 64)])
def test_fused_plan_reproduces_spmm(n, mean_degree):
    rowptr, col, val = synth.cfg2_word_graph(n, mean_degree, seed=3)
    ldx = 12
    plan = fused_plan_arrays(rowptr, col, val, ldx)
    plan['ldx4'] = ldx * 4
    rs = np.random.RandomState(0)
    x = rs.randn(n, 7)
    dense = np.zeros((n, n))
    dense[np.repeat(np.arange(n), np.diff(rowptr)), col] = val
    y = _emulate(plan, rowptr, col, val, x)
    np.testing.assert_allclose(y, dense @ x, rtol=1e-12, atol=1e-12)


def test_fused_plan_hub_rows_empty_rows_and_balance():
    n = 700
    rs = np.random.RandomState(1)
    dense = (rs.rand(n, n) < 0.02) * rs.randn(n, n)
    dense[5] = rs.randn(n)               # 700 edges -> 6 segments
    dense[9] = 0                         # empty row
    rows, cols = np.nonzero(dense)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))])
    plan = fused_plan_arrays(rowptr, cols, dense[rows, cols].astype(np.float32), 300)
    plan['ldx4'] = 1200
    x = rs.randn(n, 3)
    y = _emulate(plan, rowptr, cols, dense[rows, cols], x)
    np.testing.assert_allclose(y, dense.astype(np.float32).astype(np.float64) @ x, rtol=1e-6, atol=1e-6)
    segs, ptr = plan['segs'], plan['tile_seg_ptr']
    assert plan['n_tiles'] == 6
    hub_slot = int(np.nonzero(plan['tile_rows'] == 5)[0][0])
    assert hub_slot % FUSED_TILE_ROWS == 0      # the heaviest row is rank 0 -> tile 0, slot 0
    for t in range(plan['n_tiles']):
        lens = segs[ptr[t]:ptr[t + 1], 1]
        assert (np.diff(lens) <= 0).all()       # longest first within a tile
    per_tile = np.array([segs[ptr[t]:ptr[t + 1], 1].sum() for t in range(plan['n_tiles'])])
    assert per_tile.max() - per_tile.min() <= 700   # rank dealing: tiles differ by at most the hub row


def test_cfg2_word_graph_has_the_specified_size():
    rowptr, col, val = synth.cfg2_word_graph(2000, 64, seed=0)
    assert col.shape[0] == 65 * 2000 and rowptr[-1] == col.shape[0]
    deg = np.diff(rowptr)
    rows = np.repeat(np.arange(2000), deg)
    assert ((col == rows).reshape(-1).sum()) == 2000                       # one self loop per row
    same_row = np.diff(rows) == 0
    assert (np.diff(col)[same_row] > 0).all()                              # sorted, no duplicates
    np.testing.assert_allclose(np.add.reduceat(val, rowptr[:-1]), 1.0, rtol=1e-5)
    assert deg.max() > 10 * np.median(deg)                                 # heavy tail


def test_hub_spmm_plan_reproduces_spmm_and_stages_the_hubs():
    """Emulate mgnns_spmm_hub_f32's walk over the plan (chunk -> segment -> hub edges from the table, other edges from
    X) in float64 and compare with the dense product."""
    from mgnns_b200.api.graph_util import hub_plan_arrays
    n, F, ldx = 900, 8, 12
    rowptr, col, val = synth.cfg2_word_graph(n, 40, seed=5)
    plan = hub_plan_arrays(rowptr, col, val, n, ldx, F, hub_capacity=25, n_chunks=7, seg_edges=64)
    rs = np.random.RandomState(2)
    x = rs.randn(n, ldx)
    table = x[plan['hub_cols'], :F]
    y = np.zeros((n, F))
    written = np.zeros(n, dtype=np.int64)
    for ch in range(plan['n_chunks']):
        seg = plan['segs'][plan['chunk_seg_ptr'][ch]:plan['chunk_seg_ptr'][ch + 1]]
        assert (np.diff(seg[:, 2]) <= 0).all()                      # longest first
        for beg, n_hub, cnt, roww in seg:
            row, sole = int(roww) & 0x7fffffff, int(roww) < 0
            assert 0 <= n_hub <= cnt <= 64
            e = plan['edges'][beg:beg + cnt]
            off = e[:, 0].view(np.uint32).astype(np.int64)
            w = e[:, 1].copy().view(np.float32).astype(np.float64)
            acc = w[:n_hub] @ table[off[:n_hub] // (F * 4)] if n_hub else 0.0
            acc = acc + (w[n_hub:] @ x[off[n_hub:] // (ldx * 4), :F] if cnt > n_hub else 0.0)
            y[row] += acc
            written[row] += 1
            assert sole == (np.diff(rowptr)[row] <= 64)
    assert set(np.nonzero(written > 1)[0].tolist()) == set(plan['multi_rows'].tolist())
    dense = np.zeros((n, n))
    dense[np.repeat(np.arange(n), np.diff(rowptr)), col] = val
    np.testing.assert_allclose(y, dense @ x[:, :F], rtol=1e-10, atol=1e-12)
    coldeg = np.bincount(col, minlength=n)
    assert coldeg[plan['hub_cols']].min() >= np.sort(coldeg)[-25]   # the most referenced columns are the staged ones
    assert 0.1 < plan['hub_edge_fraction'] < 1.0
    per_chunk = np.array([plan['segs'][plan['chunk_seg_ptr'][c]:plan['chunk_seg_ptr'][c + 1], 2].sum() for c in range(7)])
    assert per_chunk.max() < 1.3 * per_chunk.mean()
