"""cfg 2 microbench driver: GraphConvolution(300->512)+ReLU on the 10k-node word graph (nnz 650,000).
Times the fused kernel (gather -> smem operand -> tcgen05) and the two-kernel path (SpMM + tcgen05 dense layer) with
CUDA events, per precision mode, and checks one sample against a float64 torch reference.
Usage: python scripts/cfg2_bench.py [B] [iters]"""
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
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, Fin, Fout = 10000, 300, 512
rowptr, cols, val = synth.cfg2_word_graph(N, seed=0)
nnz = cols.shape[0]
torch.manual_seed(0)
gc = GraphConvolution(Fin, Fout).to(dev)
x = torch.randn(B, N, Fin, device=dev)
csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)

A = torch.zeros(N, N, dtype=torch.float64)
rows = np.repeat(np.arange(N), np.diff(rowptr))
A[torch.from_numpy(rows), torch.from_numpy(cols)] = torch.from_numpy(val).double()
ref = torch.relu(A @ x[B - 1].double().cpu() @ gc.weight.detach().double().cpu())
alg_bytes = 4 * (B * N * Fin + B * N * Fout + Fin * Fout) + 8 * nnz + 4 * (N + 1)
print("cfg2: N=%d nnz=%d B=%d  algorithmic bytes %.3f GB, gather bytes %.1f GB" % (N, nnz, B, alg_bytes / 1e9, 4.0 * nnz * B * Fin / 1e9), flush=True)

for fused in ('0', '1'):
    os.environ['MGNNS_GCN_FUSED'] = fused
    for mode in ('tf32x3', 'tf32'):
        ops.set_precision(mode)
        names = ['gcn_fused'] if fused == '1' else ['spmm_hub', 'spmm_csr', 'linear_tc']
        with torch.no_grad():
            for _ in range(2):
                y = gc(x, csr, ops.ACT_RELU)
            err = (y[B - 1].double().cpu() - ref).abs().max().item()
            ops.KernelTimers.reset(names)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                y = gc(x, csr, ops.ACT_RELU)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        parts = "  ".join("%s %.3f ms" % (n, ops.KernelTimers.mean_ms(n)[0]) for n in names if ops.KernelTimers.mean_ms(n)[0])
        ops.KernelTimers.reset([])
        print("fused=%s mode=%s: total %.3f ms (%.0f GB/s algorithmic = %.3f of 6535.7)  %s  max|err| %.2e (ref rms %.3f)"
              % (fused, mode, ms, alg_bytes / ms / 1e6, alg_bytes / ms / 1e6 / 6535.7, parts, err, ref.pow(2).mean().sqrt().item()),
              flush=True)
        del y
