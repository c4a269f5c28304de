"""cfg 2 microbench driver: GraphConvolution(300->512)+ReLU on the 10k-node word graph, batch 256.
Prints per-kernel CUDA-event times for each precision mode and hot-column setting; checks one sample against
a float64 torch reference.  Usage: python scripts/cfg2_bench.py [B] [hot_cols ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mgnns_b200 import ops, synth
from mgnns_b200.api.graph_util import CSRAdjacency
from mgnns_b200.api.multi_gcn import GraphConvolution

dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hots = [0]
N, Fin, Fout = 10000, 300, 512
rowptr, cols, val = synth.cfg2_word_graph(N, seed=0)
nnz = cols.shape[0]
torch.manual_seed(0)
gc = GraphConvolution(Fin, Fout).to(dev)
x = torch.randn(B, N, Fin, device=dev)

# float64 reference for sample 0
A = torch.zeros(N, N, dtype=torch.float64)
rows = np.repeat(np.arange(N), np.diff(rowptr))
A[torch.from_numpy(rows), torch.from_numpy(cols)] = torch.from_numpy(val).double()
ref = torch.relu(A @ x[0].double().cpu() @ gc.weight.detach().double().cpu())

for hot in hots:
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)
    for mode in ('tf32x3', 'tf32', 'fp32'):
        if mode == 'fp32' and B > 64:
            continue
        ops.set_precision(mode)
        with torch.no_grad():
            for _ in range(2):
                y = gc(x, csr, ops.ACT_RELU)
            err = (y[0].double().cpu() - ref).abs().max().item()
            ops.KernelTimers.reset(['spmm_csr', 'linear_tc'])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                y = gc(x, csr, ops.ACT_RELU)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        sp, _ = ops.KernelTimers.mean_ms('spmm_csr')
        li, _ = ops.KernelTimers.mean_ms('linear_tc')
        ops.KernelTimers.reset([])
        print("B=%d mode=%s: total %.3f ms  spmm %s ms  linear_tc %s ms  max|err| %.2e (ref rms %.3f)"
              % (B, mode, ms, "%.3f" % sp if sp else "-", "%.3f" % li if li else "-", err, ref.pow(2).mean().sqrt().item()),
              flush=True)
        del y
