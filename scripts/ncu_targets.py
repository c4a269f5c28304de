"""Small driver that launches the hot kernels a few times (for ncu --set full captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgnns_b200 import ops
dev = torch.device('cuda', 0)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
B = int(os.environ.get('B', 512))
torch.manual_seed(0)
if which in ('imgbank', 'all'):
    f = torch.relu(torch.randn(B, 2048, 14, 14, device=dev)).requires_grad_(False)
    w = (torch.randn(300, 2048, device=dev) * 0.02).requires_grad_()
    b = torch.zeros(300, device=dev, requires_grad=True)
    for _ in range(3):
        bank, pooled, _ = torch.ops.mgnns.imgbank(f, w, b)
        (bank.sum() + pooled.sum()).backward()
    torch.cuda.synchronize()
if which in ('attn', 'all'):
    for L, masked in ((196, False), (100, True)):
        u = (torch.randn(B, 4, 300, device=dev) * 0.1).requires_grad_()
        bank = torch.randn(B, L, 300, device=dev).requires_grad_()
        mask = None
        if masked:
            lens = torch.randint(2, 40, (B,), device=dev)
            mask = (torch.arange(L, device=dev).unsqueeze(0) < lens.unsqueeze(1)).float()
        for _ in range(3):
            ctx, attn, psum, lse = torch.ops.mgnns.attn_q1(u, bank, mask, 0.088, 0.0, 0)
            ctx.sum().backward()
    torch.cuda.synchronize()
if which in ('spmm', 'all'):
    import numpy as np
    from mgnns_b200.api.graph_util import CSRAdjacency
    N, F, Bs = 10000, 300, int(os.environ.get('BS', 32))
    rs = np.random.RandomState(0)
    deg = np.clip((rs.pareto(1.3, N) + 1) * 20, 1, 3000).astype(np.int64)
    deg = np.maximum(1, (deg * (64.0 * N / deg.sum())).astype(np.int64))
    pop = (rs.pareto(1.1, N) + 1); pop /= pop.sum()
    rows, cols = [], []
    for i in range(N):
        c = np.unique(np.concatenate([rs.choice(N, deg[i], p=pop), [i]]))
        rows.append(np.full(c.shape, i)); cols.append(c)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=N))])
    val = (1.0 / np.diff(rowptr))[rows].astype(np.float32)
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)
    x = torch.randn(Bs, N, F, device=dev)
    for _ in range(3):
        y = csr.spmm(x)
    torch.cuda.synchronize()
print('done', which)
