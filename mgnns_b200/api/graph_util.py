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
