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
if which in ('spmm', 'gcn', 'all'):
    from mgnns_b200 import synth
    from mgnns_b200.api.graph_util import CSRAdjacency
    from mgnns_b200.api.multi_gcn import GraphConvolution
    N, F, Bs = 10000, 300, int(os.environ.get('BS', 32))
    rowptr, cols, val = synth.cfg2_word_graph(N, seed=0)
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)
    gc = GraphConvolution(F, 512).to(dev)
    x = torch.randn(Bs, N, F, device=dev)
    with torch.no_grad():
        for _ in range(3):
            y = gc(x, csr, ops.ACT_RELU) if which != 'spmm' else csr.spmm(x)
    torch.cuda.synchronize()
if which in ('lstm', 'all'):
    sys.path.insert(0, ROOT)
    import bench
    lstm = torch.nn.LSTM(300, 150, num_layers=2, batch_first=True, bidirectional=True).to(dev)
    hb = bench.host_batch(B, 0)
    lens = hb['lens']
    plan = ops.LstmPlan(lens, 100, dev, None)
    x = torch.randn(plan.capacity, 300, device=dev, requires_grad=True)
    for _ in range(3):
        y = ops.packed_bilstm(lstm, x, plan, False)
        y.sum().backward()
    torch.cuda.synchronize()
print('done', which)
