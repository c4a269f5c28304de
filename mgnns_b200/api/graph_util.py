"""Label-graph adjacency: gen_A / gen_adj (ref: utils/util.py:382-398, :421-426) and the CSR form the
SpMM kernel consumes."""
import pickle

import numpy as np
import torch

from .. import ops


def gen_A(num_classes, t, adj_file, gama=0.2):
    """Binarised, re-weighted co-occurrence adjacency (ref: utils/util.py:382-398).

    The reference defines four positional arguments but calls it with three
    (models/Multi_GCN_Multihead_att.py:338,:344); its own comment (util.py:396) says 0.2 is the
    paper's p, hence the default.  `adj_file` may be a path to the pickle or an already loaded
    {'adj','nums'} mapping.  Returns (A float64 [N,N], nums float64 [N,1]) like the reference.
    """
    if isinstance(adj_file, (str, bytes)):
        with open(adj_file, "rb") as f:
            result = pickle.load(f)
    else:
        result = adj_file
    counts = np.asarray(result["adj"], dtype=np.float64)
    nums = np.asarray(result["nums"], dtype=np.float64)[:, np.newaxis]
    cond = counts / nums                      # P(j | i) = co-occurrence / occurrences of i
    binary = np.where(cond < t, 0.0, 1.0)     # NaN (0/0) compares False -> 1, exactly as the reference's two masks
    binary[np.isnan(cond)] = cond[np.isnan(cond)]
    A = binary * gama / (binary.sum(0, keepdims=True) + 1e-6)
    A = A + (1 - gama) * np.identity(num_classes, int)
    return A, nums


def gen_adj(A):
    """Symmetric-style normalisation Â = (A·D)ᵀ·D with D = diag(rowsum(A)^-1/2) (ref: utils/util.py:421-426).

    D is diagonal, so the two dense N³ products of the reference reduce to
    Â[i,j] = (A[j,i]·d[i])·d[j]; the multiplication order is kept so fp32 results are identical.
    """
    d = torch.pow(A.sum(1).float(), -0.5)
    return (A.float().t() * d.unsqueeze(1)) * d.unsqueeze(0)


class CSRAdjacency:
    """Â in CSR plus its transpose (for the backward SpMM), all on one CUDA device."""

    def __init__(self, rowptr, col, val, t_rowptr, t_col, t_val, n_rows, n_cols):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.t_rowptr, self.t_col, self.t_val = t_rowptr, t_col, t_val
        self.n_rows, self.n_cols = n_rows, n_cols

    @property
    def nnz(self):
        return int(self.col.numel())

    @classmethod
    def from_dense(cls, adj: torch.Tensor) -> "CSRAdjacency":
        """Dense [N,M] CUDA matrix -> CSR via the library's own conversion kernels."""
        adj = adj.detach()
        rowptr, col, val = ops.dense_to_csr(adj)
        t_rowptr, t_col, t_val = ops.dense_to_csr(adj.t().contiguous())
        return cls(rowptr, col, val, t_rowptr, t_col, t_val, adj.shape[0], adj.shape[1])

    @classmethod
    def from_scipy_like(cls, rowptr, col, val, n_cols, device) -> "CSRAdjacency":
        """Host CSR arrays (numpy) -> device CSR + transpose (host-side transpose, one-off)."""
        rowptr = np.asarray(rowptr, dtype=np.int64)
        col = np.asarray(col, dtype=np.int64)
        val = np.asarray(val, dtype=np.float32)
        n_rows = rowptr.shape[0] - 1
        rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(rowptr))
        order = np.lexsort((rows, col))
        t_col = rows[order]
        t_val = val[order]
        t_rowptr = np.zeros(n_cols + 1, dtype=np.int64)
        np.add.at(t_rowptr, col + 1, 1)
        t_rowptr = np.cumsum(t_rowptr)

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)

        return cls(dev(rowptr, torch.int32), dev(col, torch.int32), dev(val, torch.float32),
                   dev(t_rowptr, torch.int32), dev(t_col, torch.int32), dev(t_val, torch.float32), n_rows, n_cols)

    def spmm(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.mgnns.spmm_csr(self.rowptr, self.col, self.val, x,
                                        self.t_rowptr, self.t_col, self.t_val, self.n_rows)

    def fused_plan(self, ldx: int) -> "FusedGcnPlan":
        """Execution plan of this matrix for the fused layer kernel (mgnns_gcn_fused_tc), for node features whose
        rows are `ldx` floats apart.  Built once on the host and cached."""
        plans = self.__dict__.setdefault('_fused_plans', {})
        plan = plans.get(ldx)
        if plan is None:
            plan = plans[ldx] = FusedGcnPlan.build(self.rowptr.cpu().numpy(), self.col.cpu().numpy(),
                                                   self.val.cpu().numpy(), ldx, self.rowptr.device)
        return plan


FUSED_TILE_ROWS = 128
FUSED_SEG_EDGES = 128


def fused_plan_arrays(rowptr, col, val, ldx, tile_rows=FUSED_TILE_ROWS, seg_edges=FUSED_SEG_EDGES):
    """Host arrays of the fused-layer plan (see include/mgnns_b200.h, mgnns_gcn_fused_tc).

    Rows are ranked by degree (descending, stable) and dealt to the T = ceil(n/128) tiles round-robin — tile t holds
    the rows ranked t, t+T, t+2T, ... in slots 0, 1, 2, ... — so every tile carries the same share of the edges.
    Each row becomes ceil(degree/128) segments (at least one, so empty rows still zero their slot); a tile's
    segments are sorted by length (descending) because the kernel deals them to its gather warps round-robin."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.ascontiguousarray(np.asarray(val, dtype=np.float32))
    n = rowptr.shape[0] - 1
    if int(col.max(initial=0)) * ldx * 4 >= 2 ** 32:
        raise ValueError("fused GCN plan: neighbour-row byte offsets do not fit 32 bits")
    deg = np.diff(rowptr)
    n_tiles = max(1, (n + tile_rows - 1) // tile_rows)
    order = np.argsort(-deg, kind='stable')
    rank = np.arange(n)
    tile_of, slot_of = rank % n_tiles, rank // n_tiles
    rows_tbl = np.full(n_tiles * tile_rows, -1, dtype=np.int32)
    rows_tbl[tile_of * tile_rows + slot_of] = order
    # one descriptor per (tile, slot), padding slots included (they must be zeroed too)
    slot_row = rows_tbl.astype(np.int64)
    slot_deg = np.where(slot_row >= 0, deg[np.maximum(slot_row, 0)], 0)
    slot_beg = np.where(slot_row >= 0, rowptr[np.maximum(slot_row, 0)], 0)
    n_seg = np.maximum(1, (slot_deg + seg_edges - 1) // seg_edges)
    total = int(n_seg.sum())
    owner = np.repeat(np.arange(n_tiles * tile_rows), n_seg)              # slot index of every segment
    first = np.cumsum(n_seg) - n_seg
    j = np.arange(total) - np.repeat(first, n_seg)                        # index of the segment within its row
    seg_beg = slot_beg[owner] + j * seg_edges
    seg_len = np.minimum(seg_edges, slot_deg[owner] - j * seg_edges).clip(min=0)
    seg_tile = owner // tile_rows
    key = np.lexsort((-seg_len, seg_tile))                                # by tile, longest first
    segs = np.stack([seg_beg[key], seg_len[key], (owner % tile_rows)[key], (n_seg[owner] == 1)[key].astype(np.int64)],
                    axis=1).astype(np.int32)
    tile_seg_ptr = np.zeros(n_tiles + 1, dtype=np.int64)
    np.add.at(tile_seg_ptr, seg_tile + 1, 1)
    tile_seg_ptr = np.cumsum(tile_seg_ptr).astype(np.int32)
    multi_slot = np.nonzero(n_seg > 1)[0]
    tile_multi_ptr = np.zeros(n_tiles + 1, dtype=np.int64)
    np.add.at(tile_multi_ptr, multi_slot // tile_rows + 1, 1)
    tile_multi_ptr = np.cumsum(tile_multi_ptr).astype(np.int32)
    multi_rows = (multi_slot % tile_rows).astype(np.int32)
    edges = np.empty((col.shape[0], 2), dtype=np.int32)
    edges[:, 0] = (col * ldx * 4).astype(np.uint32).view(np.int32)
    edges[:, 1] = val.view(np.int32)
    return dict(tile_seg_ptr=tile_seg_ptr, segs=np.ascontiguousarray(segs), edges=edges, tile_rows=rows_tbl,
                tile_multi_ptr=tile_multi_ptr, multi_rows=multi_rows, n_tiles=n_tiles, n_rows=n)


def hub_plan_arrays(rowptr, col, val, n_cols, ldx, F, hub_capacity, n_chunks, seg_edges=FUSED_SEG_EDGES):
    """Host arrays of the hub-staged SpMM plan (see csrc/spmm_hub.cu, mgnns_spmm_hub_f32).

    The `hub_capacity` most referenced columns become the shared-memory hub table (slot s holds X[b, hub_cols[s], :]).
    Every row's edges are reordered hub-first and re-indexed (hub edge: byte offset of its slot in the table, other
    edge: byte offset of its column's row in X[b]); rows are cut into segments of at most `seg_edges` edges; segments
    are dealt, in row order, to `n_chunks` chunks of about equal edge count and sorted longest-first inside a chunk
    (the kernel's warps take them round-robin)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.ascontiguousarray(np.asarray(val, dtype=np.float32))
    n = rowptr.shape[0] - 1
    nnz = col.shape[0]
    if max(n_cols * ldx, hub_capacity * F) * 4 >= 2 ** 32:
        raise ValueError("hub SpMM plan: byte offsets do not fit 32 bits")
    deg = np.diff(rowptr)
    coldeg = np.bincount(col, minlength=n_cols)
    hub_cols = np.argsort(-coldeg, kind='stable')[:max(0, min(hub_capacity, n_cols))]
    hub_cols = hub_cols[coldeg[hub_cols] > 1]                          # a column used once gains nothing from staging
    slot = np.full(n_cols, -1, dtype=np.int64)
    slot[hub_cols] = np.arange(hub_cols.shape[0])
    is_hub = slot[col] >= 0
    rows = np.repeat(np.arange(n), deg)
    order = np.lexsort((np.arange(nnz), ~is_hub, rows))                # by row, hub edges first, original order kept
    col2, val2, hub2 = col[order], val[order], is_hub[order]
    edges = np.empty((nnz, 2), dtype=np.int32)
    edges[:, 0] = np.where(hub2, slot[col2] * F * 4, col2 * ldx * 4).astype(np.uint32).view(np.int32)
    edges[:, 1] = val2.view(np.int32)
    hub_per_row = np.bincount(rows, weights=is_hub, minlength=n).astype(np.int64)
    n_seg = np.maximum(1, (deg + seg_edges - 1) // seg_edges)
    total = int(n_seg.sum())
    owner = np.repeat(np.arange(n), n_seg)
    first = np.cumsum(n_seg) - n_seg
    j = np.arange(total) - np.repeat(first, n_seg)
    seg_len = np.minimum(seg_edges, deg[owner] - j * seg_edges).clip(min=0)
    seg_hub = np.clip(hub_per_row[owner] - j * seg_edges, 0, seg_len)
    seg_beg = rowptr[owner] + j * seg_edges
    sole = n_seg[owner] == 1
    # contiguous chunks of about equal edge count (every segment costs at least one unit so empty rows spread too)
    n_chunks = max(1, min(n_chunks, total))
    cost = np.cumsum(np.maximum(seg_len, 1))
    chunk_of = np.minimum(n_chunks - 1, (cost - 1) * n_chunks // cost[-1])
    key = np.lexsort((-seg_len, chunk_of))
    roww = (owner[key] | np.where(sole[key], 1 << 31, 0)).astype(np.uint32).view(np.int32)
    segs = np.stack([seg_beg[key].astype(np.int32), seg_hub[key].astype(np.int32), seg_len[key].astype(np.int32), roww], axis=1)
    chunk_seg_ptr = np.zeros(n_chunks + 1, dtype=np.int64)
    np.add.at(chunk_seg_ptr, chunk_of + 1, 1)
    chunk_seg_ptr = np.cumsum(chunk_seg_ptr).astype(np.int32)
    multi_rows = np.nonzero(n_seg > 1)[0].astype(np.int32)
    return dict(hub_cols=hub_cols.astype(np.int32), chunk_seg_ptr=chunk_seg_ptr, segs=np.ascontiguousarray(segs),
                edges=edges, multi_rows=multi_rows, n_chunks=n_chunks, n_rows=n,
                hub_edge_fraction=float(is_hub.mean()) if nnz else 0.0)


class HubSpmmPlan:
    """Device copy of hub_plan_arrays()."""

    def __init__(self, arrays, ldx, F, device):
        self.ldx, self.F, self.n_chunks, self.n_rows = ldx, F, arrays['n_chunks'], arrays['n_rows']
        self.hub_edge_fraction = arrays['hub_edge_fraction']
        self.n_hub, self.n_multi = int(arrays['hub_cols'].shape[0]), int(arrays['multi_rows'].shape[0])
        for k in ('hub_cols', 'chunk_seg_ptr', 'segs', 'edges', 'multi_rows'):
            a = arrays[k]
            if a.size == 0:
                a = np.zeros((1,) + a.shape[1:], dtype=np.int32)
            setattr(self, k, torch.from_numpy(np.ascontiguousarray(a)).to(device))


class FusedGcnPlan:
    """Device copy of fused_plan_arrays()."""

    def __init__(self, arrays, ldx, device):
        self.ldx, self.n_tiles, self.n_rows = ldx, arrays['n_tiles'], arrays['n_rows']
        for k in ('tile_seg_ptr', 'segs', 'edges', 'tile_rows', 'tile_multi_ptr', 'multi_rows'):
            a = arrays[k]
            if a.size == 0:
                a = np.zeros((1,) + a.shape[1:], dtype=np.int32)            # keep a valid pointer
            setattr(self, k, torch.from_numpy(np.ascontiguousarray(a)).to(device))

    @classmethod
    def build(cls, rowptr, col, val, ldx, device):
        return cls(fused_plan_arrays(rowptr, col, val, ldx), ldx, device)


_CSR_CACHE = {}


def as_csr(adj) -> CSRAdjacency:
    """Accept a CSRAdjacency or a dense tensor (the reference's calling convention).

    Dense inputs are converted on the device; conversions are memoised on (storage, version) so a
    constant Â (ref: gen_adj(self.object_A).detach(), models/Multi_GCN_Multihead_att.py:461) is
    converted once, not every forward.
    """
    if isinstance(adj, CSRAdjacency):
        return adj
    key = (adj.data_ptr(), adj._version, tuple(adj.shape), adj.device)
    hit = _CSR_CACHE.get(key)
    if hit is not None and hit[0]() is adj:
        return hit[1]
    import weakref
    csr = CSRAdjacency.from_dense(adj)
    if len(_CSR_CACHE) > 64:
        _CSR_CACHE.clear()
    _CSR_CACHE[key] = (weakref.ref(adj), csr)
    return csr
