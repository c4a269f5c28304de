"""SpMM-only timing on the cfg-2 word graph (CUDA events): python scripts/spmm_bench.py [B]."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mgnns_b200 import ops, synth
from mgnns_b200.api.graph_util import CSRAdjacency

dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N, F = 10000, 300
rowptr, cols, val = synth.cfg2_word_graph(N, seed=0)
csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)
x = torch.randn(B, N, F, device=dev)
A = torch.zeros(N, N, dtype=torch.float64)
A[torch.from_numpy(np.repeat(np.arange(N), np.diff(rowptr))), torch.from_numpy(cols)] = torch.from_numpy(val).double()
ref = A @ x[B - 1].double().cpu()
with torch.no_grad():
    for _ in range(2):
        y = csr.spmm(x)
    err = (y[B - 1].double().cpu() - ref).abs().max().item()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = csr.spmm(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('hub kernel' if os.environ.get('MGNNS_SPMM_HUB', '1') != '0' else 'plain kernel')
nnz = cols.shape[0]
print("B=%d: %.3f ms  gather %.1f TB/s  hbm-algorithmic %.0f GB/s  max|err| %.1e"
      % (B, ms, nnz * B * F * 4 / ms / 1e9, 8.0 * B * N * F / ms / 1e6, err))
