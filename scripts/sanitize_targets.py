"""Small-shape launches of every kernel family, for `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_targets.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import mgnns_test_helpers as H
from mgnns_b200 import ops, synth
from mgnns_b200.api.graph_util import CSRAdjacency

dev = torch.device('cuda', 0)
torch.manual_seed(0)
# tensor-core dense layer / weight gradient / image bank (ragged tiles on every edge)
a = torch.randn(2100, 300, device=dev, requires_grad=True)
w = torch.randn(300, 516, device=dev, requires_grad=True)
y = torch.ops.mgnns.mm(a, w, None, False, False, ops.ACT_RELU, 0.0)
y.sum().backward()
g = torch.randn(2100, 1200, device=dev)
torch.ops.mgnns.mm(g[:, 600:], a.detach(), None, True, False, ops.ACT_NONE, 0.0)
f = torch.relu(torch.randn(5, 2048, 14, 14, device=dev))
wl = (torch.randn(300, 2048, device=dev) * 0.02).requires_grad_()
bl = torch.zeros(300, device=dev, requires_grad=True)
bank, pooled, _ = torch.ops.mgnns.imgbank(f, wl, bl)
(bank.sum() + pooled.sum()).backward()
# attention (masked + dropout, unmasked), SpMM, rowmax
for L, masked in ((196, False), (100, True)):
    u = (torch.randn(6, 4, 300, device=dev) * 0.1).requires_grad_()
    bk = torch.randn(6, L, 300, device=dev).requires_grad_()
    mask = None
    if masked:
        lens = torch.tensor([3, 100, 1, 17, 40, 9], device=dev)
        mask = (torch.arange(L, device=dev).unsqueeze(0) < lens.unsqueeze(1)).float()
    for p in (0.0, 0.1):
        torch.ops.mgnns.attn_q1(u, bk, mask, 0.088, p, 7)[0].sum().backward()
rowptr, cols, val = synth.cfg2_word_graph(500, mean_degree=12, seed=1)
csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, 500, dev)
csr.spmm(torch.randn(3, 500, 300, device=dev))
ops.rowmax(torch.randn(300, 196, device=dev))
# whole model, forked streams + deferred weight gradients
from test_gpu_parity import build_model
cfg = dict(H.MODEL_CFG, B=6, V=300, seed=5)
emap, count = synth.synthetic_edge_map(cfg['V'], seed=5, docs=300)
model = build_model(dev, cfg, emap, count).train()
model.branch_streams = True
text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
loss = torch.nn.functional.cross_entropy(model(text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev)), labels.to(dev))
prev = ops.defer_weight_grads(True)
loss.backward()
ops.defer_weight_grads(prev)
ops.join_deferred()
torch.cuda.synchronize()
print("sanitize targets done, loss", loss.item())
