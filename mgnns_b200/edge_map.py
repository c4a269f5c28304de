"""CSR stand-in for the reference's dense `edges_mappings[V,V]` (ref: utils/pmi.py:89-105), plus its on-disk form.

Pure numpy on purpose: the CPU reference arm of bench.py and the synthetic-corpus helpers use it without loading
the CUDA library (mgnns_b200.ops).
"""
import numpy as np


class SparseEdgeMap:
    """CSR stand-in for the reference's dense int `edges_mappings[V,V]` (3.25 GB at V=20k).

    Supports the two things the reference does with the matrix: `m[i, j]` lookups
    (models/Text_GCN.py:160,:164) and `.shape`; `toarray()` densifies for small V.
    Edge ids are 1 + CSR position (row-major enumeration, ref: utils/pmi.py:92-97); 0 = no edge.
    """

    def __init__(self, rowptr, col, n, eid=None):
        self.rowptr = np.asarray(rowptr, dtype=np.int64)
        self.col = np.asarray(col, dtype=np.int64)
        self.eid = None if eid is None else np.asarray(eid, dtype=np.int64)
        self.shape = (n, n)

    @property
    def nnz(self):
        return int(self.col.shape[0])

    def __getitem__(self, ij):
        i, j = int(ij[0]), int(ij[1])
        lo, hi = self.rowptr[i], self.rowptr[i + 1]
        k = lo + np.searchsorted(self.col[lo:hi], j)
        if k < hi and self.col[k] == j:
            return int(k + 1) if self.eid is None else int(self.eid[k])
        return 0

    def toarray(self):
        out = np.zeros(self.shape, dtype=np.int64)
        rows = np.repeat(np.arange(self.shape[0]), np.diff(self.rowptr))
        out[rows, self.col] = (np.arange(self.nnz) + 1) if self.eid is None else self.eid
        return out

    @classmethod
    def from_dense(cls, m):
        m = np.asarray(m)
        rows, cols = np.nonzero(m)            # row-major order
        rowptr = np.zeros(m.shape[0] + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        return cls(np.cumsum(rowptr), cols, m.shape[0], eid=m[rows, cols])

    # ---- on-disk form: the reference pickles / re-computes a dense int [V,V] array (3.25 GB at
    # V=20k, 20 GB at V=50k); the CSR arrays go into one compressed .npz instead
    def save(self, path, weights=None):
        """Write rowptr/col(/eid) (+ optional edge weights [count,1]) to `path` (.npz)."""
        arrays = dict(rowptr=self.rowptr, col=self.col.astype(np.int32 if self.shape[0] < 2 ** 31 else np.int64),
                      n=np.int64(self.shape[0]))
        if self.eid is not None:
            arrays['eid'] = self.eid
        if weights is not None:
            arrays['weights'] = np.asarray(weights, dtype=np.float32)
        np.savez_compressed(path, **arrays)

    @classmethod
    def load(cls, path):
        """-> (SparseEdgeMap, weights or None); the inverse of save()."""
        z = np.load(path)
        m = cls(z['rowptr'], z['col'], int(z['n']), eid=z['eid'] if 'eid' in z.files else None)
        return m, (z['weights'] if 'weights' in z.files else None)
