"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the golden
vectors produced by the reference's own code.

Tolerances (fp32 mode, stated per BASELINE.json north_star): logits 1e-3 relative / 1e-4 absolute,
argmax identical, integer counts / edge ids bit-exact.  Kernel-level checks use tighter bounds.
"""
import math
import os

import numpy as np
import pytest
import torch

import mgnns_test_helpers as H
from mgnns_b200 import synth
from oracle import mgnns_oracle as O
from oracle import pmi_oracle as PO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device('cuda', 0)


@pytest.fixture(scope="module")
def ops():
    from mgnns_b200 import ops as _ops
    return _ops


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(a, b, rtol=1e-4, atol=1e-5, msg=''):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(37, 301, 77), (128, 128, 64), (1, 7, 300), (513, 300, 2048), (300, 100, 5)])
def test_gemm_all_layouts(ops, dev, ta, tb, M, N, K):
    a = rnd(K, M, seed=1) if ta else rnd(M, K, seed=1)
    b = rnd(N, K, seed=2) if tb else rnd(K, N, seed=2)
    bias = rnd(N, seed=3)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + bias.double()
    out = torch.ops.mgnns.mm(a.to(dev), b.to(dev), bias.to(dev), bool(ta), bool(tb), ops.ACT_NONE, 0.0)
    close(out, ref, rtol=2e-5, atol=2e-5 * math.sqrt(K))
    out = torch.ops.mgnns.mm(a.to(dev), b.to(dev), bias.to(dev), bool(ta), bool(tb), ops.ACT_LEAKY, 0.2)
    close(out, torch.nn.functional.leaky_relu(ref, 0.2), rtol=2e-5, atol=2e-5 * math.sqrt(K))


def test_gemm_strided_rows_and_autograd(ops, dev):
    big = rnd(64, 200, seed=4).to(dev)
    a = big[:, 3:80].detach().requires_grad_()        # row stride 200, unaligned start
    w = rnd(50, 77, seed=5).to(dev).requires_grad_()
    bias = rnd(50, seed=6).to(dev).requires_grad_()
    y = ops.linear(a, w, bias, ops.ACT_RELU)
    r = rnd(64, 50, seed=7).to(dev)
    (y * r).sum().backward()
    ac, wc, bc = a.detach().cpu().double().requires_grad_(), w.detach().cpu().double().requires_grad_(), \
        bias.detach().cpu().double().requires_grad_()
    yr = torch.relu(ac @ wc.t() + bc)
    (yr * r.cpu().double()).sum().backward()
    close(y, yr, 1e-5, 1e-5)
    close(a.grad, ac.grad, 1e-4, 1e-5)
    close(w.grad, wc.grad, 1e-4, 1e-5)
    close(bias.grad, bc.grad, 1e-4, 1e-5)


def test_gemm_batch_reduce_accumulate(ops, dev):
    B, M, N, K = 12, 33, 70, 19
    a, b = rnd(B, M, K, seed=8), rnd(B, K, N, seed=9)
    c = torch.zeros(3, M, N, device=dev)
    ops.gemm_raw(0, 0, M, N, K, a.to(dev), K, M * K, b.to(dev), N, K * N, c, N, M * N, batch=B, reduce=4)
    ref = torch.bmm(a.double(), b.double()).view(3, 4, M, N).sum(1)
    close(c, ref, 1e-5, 1e-4)
    c2 = torch.ones(M, N, device=dev)
    ops.gemm_raw(0, 0, M, N, K, a.to(dev), K, M * K, b.to(dev), N, K * N, c2, N, 0, batch=B, reduce=3, accumulate=1)
    close(c2, 1 + torch.bmm(a.double(), b.double()).sum(0), 1e-5, 1e-4)


def test_head_mm_fwd_bwd(ops, dev):
    B, Hh, dk, D = 9, 4, 128, 300
    x = rnd(B, Hh * dk, seed=10).to(dev).requires_grad_()
    w = rnd(Hh * dk, D, seed=11, scale=0.1).to(dev).requires_grad_()
    u = torch.ops.mgnns.head_mm(x, w, Hh, 0)
    r = rnd(B, Hh * D, seed=12).to(dev)
    (u * r).sum().backward()
    xc, wc = x.detach().cpu().double().requires_grad_(), w.detach().cpu().double().requires_grad_()
    ur = torch.einsum('bhk,hkd->bhd', xc.view(B, Hh, dk), wc.view(Hh, dk, D)).reshape(B, Hh * D)
    (ur * r.cpu().double()).sum().backward()
    close(u, ur, 1e-5, 1e-4)
    close(x.grad, xc.grad, 1e-4, 1e-4)
    close(w.grad, wc.grad, 1e-4, 1e-4)
    c = rnd(B, Hh * D, seed=13).to(dev).requires_grad_()
    w2 = rnd(Hh * dk, D, seed=14, scale=0.1).to(dev).requires_grad_()
    o = torch.ops.mgnns.head_mm(c, w2, Hh, 1)
    r2 = rnd(B, Hh * dk, seed=15).to(dev)
    (o * r2).sum().backward()
    cc, w2c = c.detach().cpu().double().requires_grad_(), w2.detach().cpu().double().requires_grad_()
    orf = torch.einsum('bhd,hkd->bhk', cc.view(B, Hh, D), w2c.view(Hh, dk, D)).reshape(B, Hh * dk)
    (orf * r2.cpu().double()).sum().backward()
    close(o, orf, 1e-5, 1e-4)
    close(c.grad, cc.grad, 1e-4, 1e-4)
    close(w2.grad, w2c.grad, 1e-4, 1e-4)


# ------------------------------------------------------------------------------------------- SpMM / CSR
def random_adj(n, m, density, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n, m, generator=g)
    v = torch.randn(n, m, generator=g)
    return torch.where(a < density, v, torch.zeros(()))


@pytest.mark.parametrize("n,F,batch", [(80, 300, 0), (365, 1024, 0), (200, 7, 3), (1500, 300, 5), (64, 2048, 2)])
def test_spmm_matches_dense(dev, n, F, batch):
    from mgnns_b200.api.graph_util import CSRAdjacency
    adj = random_adj(n, n, 0.05, seed=n)
    adj[3] = 0           # an empty row
    x = rnd(*( (batch, n, F) if batch else (n, F)), seed=F)
    csr = CSRAdjacency.from_dense(adj.to(dev))
    assert csr.nnz == int((adj != 0).sum())
    xg = x.to(dev).requires_grad_()
    y = csr.spmm(xg)
    ref = torch.matmul(adj.double(), x.double())
    close(y, ref, 1e-5, 1e-5)
    r = rnd(*y.shape, seed=1)
    (y * r.to(dev)).sum().backward()
    close(xg.grad, torch.matmul(adj.double().t(), r.double()), 1e-5, 1e-5)


@pytest.mark.parametrize("n,F,batch,density", [(3000, 300, 20, 0.04), (2500, 64, 17, 0.08), (2048, 132, 33, 0.06)])
def test_spmm_hub_staged_kernel_matches_dense_and_plain_kernel(ops, dev, n, F, batch, density):
    """The persistent SpMM with shared-memory hub rows (batched features, nnz >= 100k): hub columns, a 1,500-edge hub
    ROW (segments combined with vector atomics), an empty row; forward and the transpose product of the backward,
    against float64 and against the plain kernel (the default; the hub kernel is MGNNS_SPMM_HUB=1)."""
    from mgnns_b200.api.graph_util import CSRAdjacency
    adj = random_adj(n, n, density, seed=n)
    g = torch.Generator().manual_seed(n + 7)
    adj[:, 5] = torch.randn(n, generator=g)
    adj[:, 17] = torch.randn(n, generator=g)
    adj[9, :1500] = torch.randn(1500, generator=g)
    adj[3] = 0
    adj = adj / adj.abs().sum(1, keepdim=True).clamp(min=1.0)
    rowptr, cols, val = _csr_host(adj)
    assert cols.shape[0] >= 100000
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, n, dev)
    x = rnd(batch, n, F, seed=F)
    xg = x.to(dev).requires_grad_()
    ops.KernelTimers.reset(['spmm_hub', 'spmm_csr'])
    os.environ['MGNNS_SPMM_HUB'] = '1'
    try:
        y = csr.spmm(xg)
        r = rnd(*y.shape, seed=1)
        (y * r.to(dev)).sum().backward()
        torch.cuda.synchronize()
        with torch.no_grad():
            os.environ['MGNNS_SPMM_PAD'] = '1'
            y_pad = csr.spmm(x.to(dev))                  # rows copied to a 128-byte stride first
            del os.environ['MGNNS_SPMM_PAD']
    finally:
        del os.environ['MGNNS_SPMM_HUB']
    torch.cuda.synchronize()
    assert ops.KernelTimers.mean_ms('spmm_hub')[1] == 3 and ops.KernelTimers.mean_ms('spmm_csr')[1] == 0
    ops.KernelTimers.reset([])
    assert (y_pad - y).abs().max().item() < 1e-6
    close(y, torch.matmul(adj.double(), x.double()), 1e-5, 1e-5)
    close(xg.grad, torch.matmul(adj.double().t(), r.double()), 1e-5, 1e-5)
    assert (y[:, 3] == 0).all()
    with torch.no_grad():
        y0 = csr.spmm(x.to(dev))                         # the default (plain) kernel
    assert (y - y0).abs().max().item() < 1e-5


def test_dense_to_csr_ordering_and_scan(ops, dev):
    adj = random_adj(3000, 517, 0.02, seed=5)
    rowptr, col, val = ops.dense_to_csr(adj.to(dev))
    rows, cols = np.nonzero(adj.numpy())
    assert np.array_equal(col.cpu().numpy(), cols)
    assert np.array_equal(val.cpu().numpy(), adj.numpy()[rows, cols])
    ref_ptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=3000))])
    assert np.array_equal(rowptr.cpu().numpy(), ref_ptr)


def test_graph_convolution_matches_reference_golden(dev, golden):
    from mgnns_b200.api.graph_util import gen_A, gen_adj
    from mgnns_b200.api.multi_gcn import GraphConvolution
    z = golden('modules.npz')
    lg = synth.label_graphs()
    A, _ = gen_A(80, 0.4, synth.adj_dict('object'))
    adj = gen_adj(torch.from_numpy(A).float().to(dev))
    gc = GraphConvolution(300, 64)
    synth.fill_parameters(gc, seed=3)
    gc.to(dev)
    out = gc(torch.from_numpy(lg['object_glove']).float().to(dev), adj)
    close(out, z['gc_out'], 1e-4, 1e-5)
    gcb = GraphConvolution(300, 32, bias=True)
    synth.fill_parameters(gcb, seed=4)
    gcb.to(dev)
    out = gcb(torch.from_numpy(z['gcb_x']).to(dev), adj)
    close(out, z['gcb_out'], 1e-4, 1e-5)
    # out_features < in_features takes the other association order
    gcs = GraphConvolution(300, 16)
    synth.fill_parameters(gcs, seed=6)
    x = torch.from_numpy(lg['object_glove']).float()
    ref = O.graph_convolution(x, adj.cpu(), gcs.weight.detach())
    close(gcs.to(dev)(x.to(dev), adj), ref, 1e-4, 1e-5)


# ------------------------------------------------------------------------------------------- fused GCN layer
def _csr_host(adj):
    rows, cols = np.nonzero(adj.numpy())
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=adj.shape[0]))])
    return rowptr, cols, adj.numpy()[rows, cols]


@pytest.mark.parametrize("n,K,N,batch,bias,act", [
    (300, 300, 512, 9, False, 1),       # cfg-2 feature sizes on a small graph, ragged last K chunk (300 = 9*32 + 12)
    (130, 64, 256, 20, True, 2),        # single accumulator half, bias + LeakyReLU, two row tiles with padding slots
    (1000, 128, 320, 3, True, 0),       # N1 = 64 second half, no activation
    (97, 32, 32, 25, False, 1),         # one K chunk, smallest N
])
def test_gcn_fused_matches_float64(ops, dev, n, K, N, batch, bias, act):
    """mgnns_gcn_fused_tc against relu((Â·X)·W + b) in float64: hub rows longer than one 128-edge segment (combined
    through the shared-memory scratch tile), empty rows, rows of every tile slot."""
    from mgnns_b200.api.graph_util import CSRAdjacency
    from mgnns_b200.api.multi_gcn import GraphConvolution
    adj = random_adj(n, n, 0.04, seed=n)
    g = torch.Generator().manual_seed(n + 1)
    adj[5] = torch.randn(n, generator=g)                 # a dense hub row: ceil(n/128) segments
    adj[7, : min(n, 200)] = torch.randn(min(n, 200), generator=g)
    adj[3] = 0                                           # an empty row
    adj[:, 11] = torch.randn(n, generator=g)             # a hub column
    adj[3] = 0
    adj = adj / adj.abs().sum(1, keepdim=True).clamp(min=1.0)      # Â-like row scale: outputs stay O(1)
    rowptr, cols, val = _csr_host(adj)
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, n, dev)
    gc = GraphConvolution(K, N, bias=bias)
    synth.fill_parameters(gc, seed=K)
    gc.to(dev)
    x = rnd(batch, n, K, seed=N)
    slope = 0.2
    os.environ['MGNNS_GCN_FUSED'] = '1'
    try:
        assert ops.gcn_fused_ok(x.to(dev), gc.weight)
        l0 = ops._abi.launch_count()
        with torch.no_grad():
            y = gc(x.to(dev), csr, act, slope)
        assert ops._abi.launch_count() - l0 == 2         # weight split + the fused kernel, nothing else
    finally:
        del os.environ['MGNNS_GCN_FUSED']
    ref = adj.double() @ x.double() @ gc.weight.detach().cpu().double()
    if bias:
        ref = ref + gc.bias.detach().cpu().double().view(1, 1, -1)
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.leaky_relu(ref, slope)
    close(y, ref, 1e-4, 2e-5)
    # the unfused path (SpMM + dense layer) on the same inputs, and the autograd path, agree with it
    with torch.no_grad():
        y2 = gc(x.to(dev), csr, act, slope)
    close(y2, ref, 1e-4, 2e-5)
    prev = ops.set_precision('tf32')
    os.environ['MGNNS_GCN_FUSED'] = '1'
    try:
        with torch.no_grad():
            y3 = gc(x.to(dev), csr, act, slope)
    finally:
        ops.set_precision(prev)
        del os.environ['MGNNS_GCN_FUSED']
    close(y3, ref, 2e-2, 5e-3)                           # plain TF32 operands: looser, stated bound


def test_cfg2_bench_shape_matches_oracle_float64(ops, dev):
    """BASELINE.json configs[1] at the bench shape — N=10,000 word graph (nnz = 650,000), 300 -> 512, ReLU — on a
    64-sample chunk of the batch: the fused kernel AND the two-kernel path (2-D SpMM grid with rows_per_cta > 1,
    79 x 2 tensor-core tiles per sample) against oracle.graph_convolution in float64 on four samples."""
    from mgnns_b200.api.graph_util import CSRAdjacency
    from mgnns_b200.api.multi_gcn import GraphConvolution
    N, Fin, Fout, B = 10000, 300, 512, 64
    rowptr, cols, val = synth.cfg2_word_graph(N, seed=0)
    assert cols.shape[0] == 650000
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, dev)
    torch.manual_seed(0)
    gc = GraphConvolution(Fin, Fout).to(dev)
    x = torch.randn(B, N, Fin, generator=torch.Generator().manual_seed(5))
    xd = x.to(dev)
    A = torch.zeros(N, N, dtype=torch.float64)
    A[torch.from_numpy(np.repeat(np.arange(N), np.diff(rowptr))), torch.from_numpy(cols)] = torch.from_numpy(val).double()
    w64 = gc.weight.detach().cpu().double()
    picks = (0, 1, 37, B - 1)
    refs = {b: torch.relu(O.graph_convolution(x[b].double(), A, w64)) for b in picks}
    with torch.no_grad():
        y_two = gc(xd, csr, ops.ACT_RELU)
        os.environ['MGNNS_GCN_FUSED'] = '1'
        try:
            y_fused = gc(xd, csr, ops.ACT_RELU)
        finally:
            del os.environ['MGNNS_GCN_FUSED']
    for b in picks:
        close(y_fused[b], refs[b], 1e-4, 1e-5)
        close(y_two[b], refs[b], 1e-4, 1e-5)
    # every sample, not just the four: the two implementations agree with each other to fp32 rounding
    assert (y_fused - y_two).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------- text GCN
def _text_case(V, B, L, ngram, seed, dev, extra_docs=()):
    text, lens, _ = synth.make_texts(B, V, L, seed=seed)
    for i, d in enumerate(extra_docs):
        text[i] = torch.tensor(d + [0] * (L - len(d)))
    emap, count = synth.synthetic_edge_map(V, seed=seed, docs=400)
    h = rnd(V, 300, seed=seed + 1, scale=0.4)
    w = 1.0 + 0.3 * rnd(count, 1, seed=seed + 2)
    return text, emap, count, h, w


@pytest.mark.parametrize("ngram", [1, 4, 5])
def test_text_maxagg_forward_backward_vs_oracle(dev, ngram):
    V, B, L = 200, 12, 100
    extra = [[0] * 0,                                # all PAD
             [5, 0, 0, 7, 5, 0, 9],                  # interior PADs, duplicates
             [3],                                    # single token
             list(range(2, 102)),                    # full length, all distinct
             [4] * 100]                              # full length, one word
    text, emap, count, h, w = _text_case(V, B, L, ngram, 3, dev, extra)
    text[0] = 0
    hc, wc = h.clone().requires_grad_(), w.clone().requires_grad_()
    ref = O.text_gcn_forward(text, hc, wc, lambda u, v: emap[u, v], ngram)
    r = rnd(B, 300, seed=9)
    (ref * r).sum().backward()
    hg, wg = h.to(dev).requires_grad_(), w.to(dev).requires_grad_()
    rp = torch.from_numpy(emap.rowptr.astype(np.int32)).to(dev)
    cl = torch.from_numpy(emap.col.astype(np.int32)).to(dev)
    out = torch.ops.mgnns.text_maxagg(text.to(dev), hg, wg, rp, cl, None, ngram, 100, True)
    close(out, ref, 1e-5, 1e-5)
    (out * r.to(dev)).sum().backward()
    close(hg.grad, hc.grad, 1e-4, 1e-5)
    close(wg.grad, wc.grad, 1e-4, 1e-5)


def test_text_gcn_module_truncation_and_explicit_eids(dev):
    from mgnns_b200.api.pmi import SparseEdgeMap
    from mgnns_b200.api.text_gcn import Model
    V, B, L = 150, 8, 120            # longer than max_length -> truncated to 100
    text, _, _ = synth.make_texts(B, V, L, seed=21)
    text[0, :] = torch.arange(2, 122) % (V - 2) + 2
    emap, count = synth.synthetic_edge_map(V, seed=21, docs=300)
    dense = emap.toarray()
    perm = np.random.RandomState(0).permutation(count - 1) + 1       # arbitrary (non row-major) ids
    dense[dense > 0] = perm[dense[dense > 0] - 1]
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, V)]
    m = Model(7, 300, vocab, 4, 0.5, count, dense, pmi=torch.zeros(count, 1))
    synth.fill_parameters(m, seed=2)
    m.eval().to(dev)
    ref = O.text_gcn_forward(text, m.node_hidden.weight.detach().cpu(), m.seq_edge_w.weight.detach().cpu(),
                             lambda u, v: int(dense[u, v]), 4, max_length=100)
    close(m(text.to(dev)), ref, 1e-5, 1e-5)
    # train mode: dropout(0.5) then ReLU == ReLU then dropout; keep-rate and scaling
    m.train()
    torch.manual_seed(0)
    y = m(text.to(dev))
    nz = ref > 0
    kept = (y.cpu() != 0) & nz
    assert 0.4 < kept.sum().item() / nz.sum().item() < 0.6
    close(y.cpu()[kept], 2 * ref[kept], 1e-5, 1e-5)


# ------------------------------------------------------------------------------------------- attention layers
def test_layernorm_fwd_bwd(dev, golden):
    from mgnns_b200.api.layers import LayerNorm
    z = golden('modules.npz')
    ln = LayerNorm(300)
    synth.fill_parameters(ln, seed=1)
    ln.to(dev)
    close(ln(torch.from_numpy(z['ln_x']).to(dev)), z['ln_y'], 1e-5, 1e-5)
    for D, rows in ((300, 37), (64, 5), (1000, 3)):
        x, res = rnd(rows, D, seed=D), rnd(rows, D, seed=D + 1)
        gam, bet = 1 + 0.1 * rnd(D, seed=2), 0.1 * rnd(D, seed=3)
        r = rnd(rows, D, seed=4)
        leaves = [t.clone().double().requires_grad_() for t in (x, res, gam, bet)]
        (O.layer_norm(leaves[0] + leaves[1], leaves[2], leaves[3]) * r.double()).sum().backward()
        gl = [t.to(dev).requires_grad_() for t in (x, res, gam, bet)]
        y = torch.ops.mgnns.add_layernorm(gl[0], gl[1], gl[2], gl[3], 1e-6)
        close(y, O.layer_norm(x.double() + res.double(), gam.double(), bet.double()), 1e-5, 1e-5)
        (y * r.to(dev)).sum().backward()
        for a, b, nm in zip(gl, leaves, 'x res gamma beta'.split()):
            close(a.grad, b.grad, 2e-4, 2e-5, msg='%s D=%d' % (nm, D))


def _mha_layer(dev, seed=2, dropout=0.5):
    from mgnns_b200.api.layers import MyMultiHeadAttention
    layer = MyMultiHeadAttention(4, 300, 128, dropout=dropout, need_mask=False)
    synth.fill_parameters(layer, seed=seed)
    return layer.to(dev)


def test_mha_layer_matches_reference_golden(dev, golden):
    z = golden('modules.npz')
    layer = _mha_layer(dev).eval()
    q = torch.from_numpy(z['mha_q']).to(dev).requires_grad_()
    bank = torch.from_numpy(z['mha_bank']).to(dev).requires_grad_()
    mask = torch.from_numpy(z['mha_mask']).to(dev)
    y, attn = layer(q, bank, bank, mask)
    close(y, z['mha_out_masked'], 1e-4, 1e-4)
    assert attn.shape == z['mha_attn_masked'].shape
    close(attn, z['mha_attn_masked'], 1e-4, 1e-6)
    yu, attnu = layer(q, bank, bank, None)
    close(yu, z['mha_out_unmasked'], 1e-4, 1e-4)
    close(attnu, z['mha_attn_unmasked'], 1e-4, 1e-6)
    (y * torch.from_numpy(z['mha_r']).to(dev)).sum().backward()
    close(q.grad, z['mha_gq'], 1e-3, 1e-4)
    close(bank.grad, z['mha_gbank'], 1e-3, 1e-5)
    for n, p in layer.named_parameters():
        g = z['mha_g_' + n]
        if n == 'slf_attn.w_ks.bias':
            # the key bias shifts every score of a head equally; softmax cancels it (gradient == 0 up to
            # rounding in the reference, exactly absent here)
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6
            assert float(np.abs(g).max()) < 1e-5
            continue
        if g.ndim == 0:
            np.testing.assert_allclose(p.grad.norm().item(), float(g), rtol=1e-3, err_msg=n)
        else:
            close(p.grad, g, 1e-3, 1e-5, msg=n)


@pytest.fixture(params=['tc', 'scalar'])
def attn_impl(request):
    """Both attention implementations behind the one op: mma.sync fragments (default) and the CUDA-core kernels."""
    prev = os.environ.get('MGNNS_ATTN')
    os.environ['MGNNS_ATTN'] = request.param
    yield request.param
    if prev is None:
        del os.environ['MGNNS_ATTN']
    else:
        os.environ['MGNNS_ATTN'] = prev


@pytest.mark.parametrize("B,L,Hh,masked,D", [(7, 196, 4, False, 300), (5, 100, 4, True, 300), (3, 33, 16, True, 300),
                                               (2, 1, 4, False, 300), (4, 70, 5, True, 300), (3, 65, 3, False, 64),
                                               (2, 32, 8, True, 8), (3, 97, 12, False, 132)])
def test_attn_q1_core_vs_dense_formula(dev, ops, attn_impl, B, L, Hh, masked, D):
    assert ops.attn_uses_tensor_cores(Hh, L, D) == (attn_impl == 'tc')
    u, bank = rnd(B, Hh, D, seed=1, scale=0.2), rnd(B, L, D, seed=2)
    mask = None
    if masked:
        lens = torch.randint(1, L + 1, (B,), generator=torch.Generator().manual_seed(3))
        mask = (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)).float()
    scale = 1 / math.sqrt(128)
    ud, bd = u.double().requires_grad_(), bank.double().requires_grad_()
    s = torch.einsum('bhd,bld->bhl', ud, bd) * scale
    if mask is not None:
        s = s.masked_fill(mask.unsqueeze(1) == 0, float('-inf'))
    p = torch.softmax(s, -1)
    ctx_ref = torch.einsum('bhl,bld->bhd', p, bd)
    r, r2 = rnd(B, Hh, D, seed=4).double(), rnd(B, Hh, seed=5).double()
    ((ctx_ref * r).sum() + (p.sum(-1) * r2).sum()).backward()
    ug, bg = u.to(dev).requires_grad_(), bank.to(dev).requires_grad_()
    ctx, attn, psum, lse = torch.ops.mgnns.attn_q1(ug, bg, None if mask is None else mask.to(dev), scale, 0.0, 0)
    close(ctx, ctx_ref, 1e-4, 1e-5)
    close(attn.view(Hh, B, L).permute(1, 0, 2), p, 1e-4, 1e-6)
    close(psum, torch.ones(B, Hh), 1e-5, 1e-5)
    close(lse, torch.logsumexp(s, -1), 1e-5, 1e-5)
    ((ctx * r.float().to(dev)).sum() + (psum * r2.float().to(dev)).sum()).backward()
    close(ug.grad, ud.grad, 1e-3, 1e-5)
    close(bg.grad, bd.grad, 1e-3, 1e-5)


def test_attn_q1_masked_rows_inside_the_sequence_and_zero_gradient_rows(dev, attn_impl):
    """Interior masked rows (not only a padded tail), an all-but-one-masked sample, and rows past the last live chunk:
    their probabilities and bank gradients are exactly zero."""
    B, L, Hh, D = 4, 100, 4, 300
    u, bank = rnd(B, Hh, D, seed=11, scale=0.2), rnd(B, L, D, seed=12)
    mask = torch.ones(B, L)
    mask[0, 3:40] = 0
    mask[0, 77:] = 0
    mask[1, 1:] = 0
    mask[2, :50] = 0
    ud, bd = u.double().requires_grad_(), bank.double().requires_grad_()
    s = (torch.einsum('bhd,bld->bhl', ud, bd) * 0.1).masked_fill(mask.unsqueeze(1) == 0, float('-inf'))
    p = torch.softmax(s, -1)
    r = rnd(B, Hh, D, seed=13).double()
    (torch.einsum('bhl,bld->bhd', p, bd) * r).sum().backward()
    ug, bg = u.to(dev).requires_grad_(), bank.to(dev).requires_grad_()
    ctx, attn, psum, lse = torch.ops.mgnns.attn_q1(ug, bg, mask.to(dev), 0.1, 0.0, 0)
    close(ctx, torch.einsum('bhl,bld->bhd', p, bd), 1e-4, 1e-5)
    a = attn.view(Hh, B, L).permute(1, 0, 2).cpu()
    assert (a[mask.unsqueeze(1).expand(B, Hh, L) == 0] == 0).all()
    (ctx * r.float().to(dev)).sum().backward()
    close(ug.grad, ud.grad, 1e-3, 1e-5)
    close(bg.grad, bd.grad, 1e-3, 1e-5)
    assert (bg.grad.cpu()[mask == 0] == 0).all()


def test_attention_dropout_statistics_and_determinism(dev, attn_impl):
    B, L, Hh, D = 64, 100, 4, 300
    u, bank = rnd(B, Hh, D, seed=1, scale=0.05).to(dev), rnd(B, L, D, seed=2).to(dev)
    _, attn0, _, _ = torch.ops.mgnns.attn_q1(u, bank, None, 0.1, 0.0, 0)
    ctx1, attn1, psum1, _ = torch.ops.mgnns.attn_q1(u, bank, None, 0.1, 0.1, 1234)
    ctx2, attn2, _, _ = torch.ops.mgnns.attn_q1(u, bank, None, 0.1, 0.1, 1234)
    _, attn3, _, _ = torch.ops.mgnns.attn_q1(u, bank, None, 0.1, 0.1, 99)
    assert torch.equal(attn1, attn2) and torch.equal(ctx1, ctx2)
    assert not torch.equal(attn1, attn3)
    kept = attn1 != 0
    assert abs(kept.float().mean().item() - 0.9) < 0.01
    close(attn1[kept], attn0[kept] / 0.9, 1e-5, 1e-7)
    close(psum1, attn1.view(Hh, B, L).sum(-1).t(), 1e-5, 1e-6)
    close(ctx1, torch.einsum('hbl,bld->bhd', attn1.view(Hh, B, L), bank), 1e-4, 1e-5)
    # backward regenerates the same mask: finite-difference-free check against autograd on the dense formula
    ug, bg = u.clone().requires_grad_(), bank.clone().requires_grad_()
    ctx, _, psum, _ = torch.ops.mgnns.attn_q1(ug, bg, None, 0.1, 0.1, 1234)
    r = rnd(B, Hh, D, seed=7).to(dev)
    ((ctx * r).sum() + psum.sum()).backward()
    ud, bd = u.double().requires_grad_(), bank.double().requires_grad_()
    p = torch.softmax(torch.einsum('bhd,bld->bhl', ud, bd) * 0.1, -1)
    pt = p * (kept.view(Hh, B, L).permute(1, 0, 2).double() / 0.9)
    ((torch.einsum('bhl,bld->bhd', pt, bd) * r.double()).sum() + pt.sum()).backward()
    close(ug.grad, ud.grad, 1e-3, 1e-5)
    close(bg.grad, bd.grad, 1e-3, 1e-5)


def test_label_attention_matches_reference_golden_and_grads(dev, golden):
    from mgnns_b200.api.multi_gcn import Attention
    z = golden('modules.npz')
    lg = synth.label_graphs()
    att = Attention(hid_dim=300, image_dim=80, n_heads=5, dropout=0.5)
    synth.fill_parameters(att, seed=5)
    att.eval().to(dev)
    key = torch.from_numpy(z['latt_key']).to(dev).requires_grad_()
    query = torch.from_numpy(lg['label_glove'])
    out = att(query, key, key)
    close(out, z['latt_out'], 1e-4, 1e-5)
    r = rnd(*out.shape, seed=3)
    (out * r.to(dev)).sum().backward()
    P = {n: p.detach().cpu().double().requires_grad_() for n, p in att.named_parameters()}
    kc = torch.from_numpy(z['latt_key']).double().requires_grad_()
    (O.label_attention(P, '', query.double(), kc, kc, 5) * r.double()).sum().backward()
    close(key.grad, kc.grad, 1e-3, 1e-6)
    for n, p in att.named_parameters():
        close(p.grad, P[n].grad, 1e-3, 1e-5, msg=n)
    # dropout on the softmax probabilities (train mode): determinism + keep rate
    att.train()
    torch.manual_seed(5)
    o1 = att(query, key.detach(), key.detach())
    torch.manual_seed(5)
    o2 = att(query, key.detach(), key.detach())
    assert torch.equal(o1, o2)


# ------------------------------------------------------------------------------------------- packed bi-LSTM
@pytest.mark.parametrize("B,L,seed", [(5, 12, 0), (37, 100, 1), (8, 100, 2)])
def test_packed_bilstm_vs_torch_lstm(dev, ops, B, L, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    if seed == 2:
        lens[:] = torch.tensor([100, 100, 1, 1, 2, 50, 99, 3])
    x = torch.randn(B, L, 300, generator=g)
    ref = torch.nn.LSTM(300, 150, num_layers=2, bidirectional=True, batch_first=True, dropout=0.5).eval()
    synth.fill_parameters(ref, seed=seed)
    xr = x.clone().double().requires_grad_()
    refd = ref.double()
    packed = torch.nn.utils.rnn.pack_padded_sequence(xr, lens, batch_first=True, enforce_sorted=False)
    out, (hn, _) = refd(packed)
    bank_ref, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=L)
    r = torch.randn(B, L, 300, generator=g)
    (bank_ref * r.double()).sum().backward()

    mine = torch.nn.LSTM(300, 150, num_layers=2, bidirectional=True, batch_first=True, dropout=0.5).eval()
    synth.fill_parameters(mine, seed=seed)
    mine = mine.to(dev)
    plan = ops.LstmPlan(lens, L, dev)
    assert plan.N == int(lens.sum())
    xg = x.to(dev).requires_grad_()
    xc = xg.reshape(B * L, 300).index_select(0, plan.flat_idx)
    y = ops.packed_bilstm(mine, xc, plan, False)
    bank = y.new_zeros(B * L, 300).index_copy(0, plan.flat_idx, y).view(B, L, 300)
    close(bank, bank_ref, 1e-4, 1e-5)
    # final states: forward direction at the last token, reverse direction at the first token
    close(y.index_select(0, plan.last_idx)[:, :150], hn[-2], 1e-4, 1e-5)
    close(y.index_select(0, plan.first_idx)[:, 150:], hn[-1], 1e-4, 1e-5)
    (bank * r.to(dev)).sum().backward()
    close(xg.grad, xr.grad, 1e-3, 1e-5)
    for (n, p), (_, q) in zip(mine.named_parameters(), refd.named_parameters()):
        close(p.grad, q.grad, 1e-3, 2e-5, msg=n)


# ------------------------------------------------------------------------------------------- image bank
@pytest.mark.parametrize("mode,rtol,atol", [("fp32", 1e-4, 1e-4), ("tf32x3", 1e-4, 1e-4), ("tf32", 5e-3, 5e-2)])
@pytest.mark.parametrize("B", [5, 37])
def test_imgbank_fwd_bwd(dev, ops, mode, rtol, atol, B):
    prev = ops.set_precision(mode)
    try:
        _imgbank_case(dev, B, rtol, atol)
    finally:
        ops.set_precision(prev)


def test_imgbank_fused_maxpool_negative_maps_nan_and_ragged_positions(dev, ops):
    """The 3xTF32 forward fuses the global max pool into its operand pass (ref: nn.MaxPool2d(14,14), model:302):
    all-negative maps (zero-filled TMA padding must not win), NaN, -0.0, and position counts that end inside a
    32-position box, against torch."""
    prev = ops.set_precision("tf32x3")
    try:
        for B, C, P_ in ((3, 64, 196), (5, 96, 100), (2, 32, 32), (9, 64, 8), (4, 128, 260)):
            f = -torch.rand(B, C, P_, generator=torch.Generator().manual_seed(P_)) - 0.5
            f[0, 1] = rnd(P_, seed=1)
            f[0, 2, P_ - 1] = 7.0
            f[0, 3, 0] = float('nan')
            f[0, 4] = 0.0
            f[0, 4, P_ // 2] = -0.0
            f[B - 1, C - 1, P_ - 1] = float('inf')
            w, b = rnd(300, C, seed=2, scale=0.1), rnd(300, seed=3)
            bank, pooled, _ = torch.ops.mgnns.imgbank(f.to(dev), w.to(dev), b.to(dev))
            ref = f.max(dim=2)[0]
            got = pooled.cpu()
            assert torch.equal(torch.isnan(got), torch.isnan(ref)), (B, C, P_)
            assert torch.equal(got[~torch.isnan(ref)], ref[~torch.isnan(ref)]), (B, C, P_)
            ok = ~torch.isnan(f).any(2) & ~torch.isinf(f).any(2)
            bank_ref = torch.nn.functional.linear(f.double().permute(0, 2, 1), w.double(), b.double())
            sel = ok.all(1)
            close(bank[sel.to(dev)], bank_ref[sel], 1e-4, 1e-4)
    finally:
        ops.set_precision(prev)


def _imgbank_case(dev, B, rtol, atol):
    C, P_, Oo = 2048, 196, 300
    f = torch.relu(rnd(B, C, 14, 14, seed=1))
    f[0, 7] = 0                                  # an all-zero channel: arg-max must be position 0
    w, b = rnd(Oo, C, seed=2, scale=0.02), rnd(Oo, seed=3)
    fg, wg, bg = f.to(dev).requires_grad_(), w.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    bank, pooled, argmax = torch.ops.mgnns.imgbank(fg, wg, bg)
    fd, wd, bd = f.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    bank_ref = torch.nn.functional.linear(fd.view(B, C, -1).permute(0, 2, 1), wd, bd)
    pooled_ref = torch.nn.functional.max_pool2d(fd, 14, 14).view(B, C)
    close(bank, bank_ref, rtol, atol)
    assert torch.equal(pooled.cpu(), pooled_ref.float())
    if argmax.numel():                           # unfused path returns the arg-max; the fused one recomputes it in backward
        assert int(argmax[0, 7]) == 0
    r1, r2 = rnd(B, P_, Oo, seed=4), rnd(B, C, seed=5)
    ((bank * r1.to(dev)).sum() + (pooled * r2.to(dev)).sum()).backward()
    ((bank_ref * r1.double()).sum() + (pooled_ref * r2.double()).sum()).backward()
    close(wg.grad, wd.grad, rtol, atol * 10 * math.sqrt(B))
    close(bg.grad, bd.grad, 1e-4, 1e-3)
    close(fg.grad, fd.grad, 1e-4, 1e-4)


# ------------------------------------------------------------------------------------------- tensor-core dense layer
@pytest.mark.parametrize("mode,rtol,atol", [("tf32x3", 5e-5, 3e-6), ("tf32", 5e-3, 1e-3)])
@pytest.mark.parametrize("w_kn", [True, False])
@pytest.mark.parametrize("M,N,K", [(2048, 512, 300), (2500, 300, 128), (4099, 1200, 300), (2304, 44, 36), (3000, 520, 2048)])
def test_linear_tc_matches_fp64(dev, ops, mode, rtol, atol, w_kn, M, N, K):
    """mgnns::mm routes products with >= 2048 rows to the tcgen05 kernel (ref: the X.W of
    GraphConvolution.forward, model:53, fused with bias + activation)."""
    prev = ops.set_precision(mode)
    try:
        a = rnd(M, K, seed=1)
        w = rnd(K, N, seed=2) if w_kn else rnd(N, K, seed=2)
        bias = rnd(N, seed=3)
        ref = a.double() @ (w.double() if w_kn else w.double().t()) + bias.double()
        before = ops.KernelTimers.records.get("linear_tc")
        ops.KernelTimers.reset(["linear_tc"])
        out = torch.ops.mgnns.mm(a.to(dev), w.to(dev), bias.to(dev), False, not w_kn, ops.ACT_NONE, 0.0)
        assert len(ops.KernelTimers.records["linear_tc"]) == 1, "tensor-core path was not taken"
        ops.KernelTimers.reset([])
        # the TMEM accumulator truncates (round toward zero) on every add: the bound grows with K, not sqrt(K)
        close(out, ref, rtol=rtol, atol=atol * K)
        out = torch.ops.mgnns.mm(a.to(dev), w.to(dev), None, False, not w_kn, ops.ACT_RELU, 0.0)
        close(out, torch.relu(ref - bias.double()), rtol=rtol, atol=atol * K)
        out = torch.ops.mgnns.mm(a.to(dev), w.to(dev), bias.to(dev), False, not w_kn, ops.ACT_LEAKY, 0.2)
        close(out, torch.nn.functional.leaky_relu(ref, 0.2), rtol=rtol, atol=atol * K)
    finally:
        ops.set_precision(prev)


@pytest.mark.parametrize("mode,rtol,atol", [("tf32x3", 5e-5, 3e-6), ("tf32", 5e-3, 1e-3)])
@pytest.mark.parametrize("M,N,K", [(600, 300, 8192), (600, 152, 2051), (1200, 300, 4100), (36, 44, 2048), (132, 520, 3000)])
def test_wgrad_tc_matches_fp64(dev, ops, mode, rtol, atol, M, N, K):
    """C = A^T . B with a long reduction (the LSTM / many-row Linear weight gradients) on the tcgen05 kernel:
    mgnns::mm with trans_a and >= 2048 reduction rows, including strided column slices of a wider matrix."""
    prev = ops.set_precision(mode)
    try:
        wide = rnd(K, 2 * M, seed=1).to(dev)
        a = wide[:, M:]                               # column slice: row stride 2M, 16-byte aligned start
        b = rnd(K, N, seed=2).to(dev)
        ref = a.double().cpu().t() @ b.double().cpu()
        ops.KernelTimers.reset(["wgrad_tc"])
        out = torch.ops.mgnns.mm(a, b, None, True, False, ops.ACT_NONE, 0.0)
        assert len(ops.KernelTimers.records["wgrad_tc"]) == 1, "tensor-core weight-gradient path was not taken"
        ops.KernelTimers.reset([])
        close(out, ref, rtol=rtol, atol=atol * K)
    finally:
        ops.set_precision(prev)


def test_linear_tc_fp32_mode_keeps_cuda_core_path_and_autograd(dev, ops):
    a = rnd(2100, 300, seed=5).to(dev).requires_grad_()
    w = rnd(300, 512, seed=6, scale=0.05).to(dev).requires_grad_()
    prev = ops.set_precision("fp32")
    try:
        y32 = torch.ops.mgnns.mm(a, w, None, False, False, ops.ACT_RELU, 0.0)
    finally:
        ops.set_precision(prev)
    y = torch.ops.mgnns.mm(a, w, None, False, False, ops.ACT_RELU, 0.0)       # default mode: tf32x3 on tensor cores
    close(y, y32, 2e-5, 2e-5)
    r = rnd(2100, 512, seed=7).to(dev)
    (y * r).sum().backward()
    ad, wd = a.detach().double().cpu().requires_grad_(), w.detach().double().cpu().requires_grad_()
    (torch.relu(ad @ wd) * r.double().cpu()).sum().backward()
    close(a.grad, ad.grad, 1e-4, 1e-4)
    close(w.grad, wd.grad, 1e-4, 1e-3)


@pytest.mark.parametrize("P", [196, 4, 128, 260, 512, 197, 7])
def test_rowmax_ties_nan_inf_vs_torch(dev, ops, P):
    """Global spatial max + first-index arg-max (ref: nn.MaxPool2d(14,14), model:302): vectorised kernel for
    P % 4 == 0, scalar kernel otherwise; ties -> lowest index, NaN wins, all -inf rows -> index 0."""
    rows = 1000
    x = rnd(rows, P, seed=P)
    x[1] = x[1].round()                       # many exact ties
    x[2] = -3.5                               # constant row
    x[3] = float('-inf')
    x[4, P // 2] = float('nan')
    x[5, P - 1] = float('nan')
    x[5, 0] = float('nan')
    x[6] = -x[6].abs() - 1.0                  # all negative
    x[7, P - 1] = 100.0                       # max in the last slot
    x[8] = 0.0
    x[8, P // 3] = -0.0
    pooled, argmax = ops.rowmax(x.to(dev))
    ref_v, ref_i = x.max(dim=1)
    nan_rows = torch.isnan(x).any(1)
    assert torch.equal(torch.isnan(pooled.cpu()), nan_rows)
    assert torch.equal(pooled.cpu()[~nan_rows], ref_v[~nan_rows])
    first = (x == ref_v.unsqueeze(1)).float().argmax(1)                 # lowest index attaining the max
    assert torch.equal(argmax.cpu().long()[~nan_rows], first[~nan_rows])
    assert int(argmax[4]) == P // 2 and int(argmax[5]) == 0 and int(argmax[3]) == 0


# ------------------------------------------------------------------------------------------- PMI
KAT_VOCAB = ['PAD', 'UNK', 'a', 'b', 'c', 'd', 'e']
KAT_DOCS = ["a b c a d", "b c d e", "a a b zzz c", "e d c b a b c"]


def test_pmi_counts_bit_exact_kat(dev, golden):
    from mgnns_b200.api import pmi
    z = golden('pmi_kat.npz')
    for tag, mc in (('w2m1', 1), ('w2m2', 2)):
        w, emap, count = pmi.cal_PMI_from_texts(KAT_DOCS, KAT_VOCAB, 2, mc, device=dev)
        assert count == int(z['kat_%s_count' % tag])
        assert np.array_equal(emap.toarray(), z['kat_%s_map' % tag])
        np.testing.assert_allclose(w.numpy(), z['kat_%s_weights' % tag], rtol=1e-6)
    pair, wc = PO.counts_loops(KAT_DOCS, KAT_VOCAB, 2)
    _, emap, _ = pmi.cal_PMI_from_texts(KAT_DOCS, KAT_VOCAB, 2, 1, device=dev)
    rowptr, col, cnt = emap.pair_counts
    dense = np.zeros_like(pair)
    dense[np.repeat(np.arange(7), np.diff(rowptr)), col] = cnt
    assert np.array_equal(dense, pair) and np.array_equal(emap.word_count, wc)


def test_pmi_real_text_bit_exact_vs_reference(dev, golden):
    from mgnns_b200.api import pmi
    z = golden('pmi_val400.npz')
    texts, vocab = list(z['texts']), list(z['vocab'])
    w, emap, count = pmi.cal_PMI_from_texts(texts, vocab, int(z['window']), int(z['min_cooc']), device=dev)
    assert count == int(z['count'])
    rows = np.repeat(np.arange(len(vocab)), np.diff(emap.rowptr))
    assert np.array_equal(rows, z['rows']) and np.array_equal(emap.col, z['cols'])
    assert np.array_equal(np.arange(1, count), z['ids'])
    np.testing.assert_allclose(w.numpy(), z['weights'], rtol=1e-6)
    # integer counts against the numpy oracle, cell by cell
    pair, wc = PO.counts_numpy(*PO.encode(texts, vocab), len(vocab), int(z['window']))
    rowptr, col, cnt = emap.pair_counts
    dense = np.zeros_like(pair)
    dense[np.repeat(np.arange(len(vocab)), np.diff(rowptr)), col] = cnt
    pair[pair < int(z['min_cooc'])] = 0
    assert np.array_equal(dense, pair) and np.array_equal(emap.word_count, wc)


def test_pmi_counts_large_synthetic_checksums(dev):
    """Full-size property check: the sum of all pair counts equals the number of (centre, target)
    pairs the corpus contains, and sharding the corpus by document adds up exactly."""
    from mgnns_b200 import ops
    V, D, L, w = 20154, 20000, 100, 6
    ids, lens, _ = synth.make_texts(D, V, L, seed=3)
    tok = ids.to(torch.int32)
    tok[ids == 1] = -1                     # treat UNK as out-of-vocabulary for this test
    full = ops.pmi_count(tok.to(dev), V, w, 0, 1)
    a = ops.pmi_count(tok[:D // 2].to(dev), V, w, 0, 1)
    b = ops.pmi_count(tok[D // 2:].to(dev), V, w, 0, 1)
    assert torch.equal(full[3], a[3] + b[3])
    assert int(full[2].sum()) == int(a[2].sum()) + int(b[2].sum())
    t = tok.numpy()
    centre = (t > 0)
    expected = 0
    for off in range(-w, w):
        if off == 0:
            continue
        c = centre[:, max(0, -off):L - max(0, off)]
        tg = t[:, max(0, off):L + min(0, off)]
        expected += int((c & (tg >= 0)).sum())
    assert int(full[2].sum()) == expected
    assert int(full[3].sum()) == int(centre.sum())


def _pair_keys_numpy(t, V, w, pad_id=0):
    """Sorted (centre*V + target) keys with counts, straight from the definition (ref: utils/pmi.py:40-58)."""
    L = t.shape[1]
    centre = (t >= 0) & (t != pad_id)
    keys = []
    for off in range(-w, w):
        if off == 0:
            continue
        c = t[:, max(0, -off):L - max(0, off)]
        ok = centre[:, max(0, -off):L - max(0, off)]
        tg = t[:, max(0, off):L + min(0, off)]
        m = ok & (tg >= 0)
        keys.append(c[m].astype(np.int64) * V + tg[m])
    return np.unique(np.concatenate(keys), return_counts=True)


@pytest.mark.parametrize("V,D,w,mc", [(977, 300, 6, 1), (977, 300, 3, 2), (60013, 400, 6, 1), (131, 50, 1, 1)])
def test_pmi_sparse_count_bit_exact_vs_definition_and_dense_table(dev, V, D, w, mc):
    """The table-free count (row buckets + shared-memory column counters) against the definition and against the
    dense-table kernel; V=60,013 exceeds the 56k-word shared-memory chunk, so the column-chunk loop runs."""
    from mgnns_b200 import ops
    ids, lens, _ = synth.make_texts(D, V, 100, seed=V % 97, zipf_a=1.05)
    tok = ids.to(torch.int32)
    tok[ids == 1] = -1
    tok[0, :] = 0                                   # an all-PAD document
    tok[1, :] = 5                                   # one word repeated: a single heavy cell (5,5)
    rowptr, col, cnt, wc = ops.pmi_count(tok.to(dev), V, w, 0, mc)
    keys, kc = _pair_keys_numpy(tok.numpy(), V, w)
    keep = kc >= mc
    rows = np.repeat(np.arange(V), np.diff(rowptr.cpu().numpy().astype(np.int64)))
    got = rows * V + col.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, keys[keep])          # row-major order, bit-exact cell set
    assert np.array_equal(cnt.cpu().numpy(), kc[keep])
    t = tok.numpy()
    assert np.array_equal(wc.cpu().numpy(), np.bincount(t[t > 0], minlength=V))
    if V < 5000:
        d = ops.pmi_count_dense(tok.to(dev), V, w, 0, mc)
        for a, b in zip((rowptr, col, cnt, wc), d):
            assert torch.equal(a, b)


def test_pmi_count_v50k_checksums_and_row_sharding(dev):
    """cfg 5 vocabulary (V=50,000; the reference's dense table would be 20 GB): size-independent properties —
    total of all cells = number of pairs in the corpus, word counts = live tokens, disjoint row ranges partition the
    result exactly (how the count shards across ranks), and the CSR is row-major with strictly increasing columns."""
    from mgnns_b200 import ops
    V, D, L, w = 50000, 40000, 100, 6
    ids, lens, _ = synth.make_texts(D, V, L, seed=11)
    tok = ids.to(torch.int32).to(dev)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    rowptr, col, cnt, wc = ops.pmi_count(tok, V, w, 0, 1)
    assert torch.cuda.max_memory_allocated() - base < 1 << 30          # < 1 GB of working memory at V=50k
    keys, kc = _pair_keys_numpy(ids.numpy(), V, w)
    assert int(cnt.sum()) == int(kc.sum()) == ops.pmi_count.last_pairs
    assert col.numel() == keys.shape[0]
    rp = rowptr.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(V), np.diff(rp))
    assert np.array_equal(rows * V + col.cpu().numpy().astype(np.int64), keys)
    assert np.array_equal(cnt.cpu().numpy(), kc)
    assert int(wc.sum()) == int((ids > 0).sum())
    cut = 137
    lo = ops.pmi_count(tok, V, w, 0, 1, row_range=(0, cut))
    hi = ops.pmi_count(tok, V, w, 0, 1, row_range=(cut, V))
    assert torch.equal(torch.cat([lo[1], hi[1]]), col) and torch.equal(torch.cat([lo[2], hi[2]]), cnt)
    assert torch.equal(lo[0][:cut + 1], rowptr[:cut + 1]) and torch.equal(lo[3] + hi[3], wc)
    two = ops.pmi_count(tok, V, w, 0, 2)
    assert int(two[2].sum()) == int(kc[kc >= 2].sum()) and two[1].numel() == int((kc >= 2).sum())


# ------------------------------------------------------------------------------------------- whole model
def build_model(dev, cfg, edge_map, edge_count, dropout=0.5):
    from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
    from mgnns_b200.api.text_gcn import Model as TextModel
    opt = dict(emb_path='', bidirectional=True, hidden_size=cfg['hidden_size'], emb_size=300,
               num_layers=cfg['num_layers'], dropout=dropout, emb_type='random', vocab_size=cfg['V'],
               stack_num=cfg['stack_num'], n_head=cfg['n_head'], d_kv=cfg['d_kv'], is_regu=False)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, cfg['V'])]
    text_model = TextModel(cfg['num_labels'], 300, vocab, cfg['ngram'], 0.5, edge_count, edge_map,
                           pmi=torch.zeros(edge_count, 1))
    model = Multi_GCN_Multihead_Att(opt, cfg['num_labels'], text_model, IdentityTrunk(), IdentityTrunk(), 80, 365,
                                    object_t=cfg['object_t'], place_t=cfg['place_t'], in_channel=300,
                                    object_adj_file=synth.adj_dict('object'), place_adj_file=synth.adj_dict('place'))
    synth.fill_parameters(model, seed=cfg['seed'])
    return model.to(dev)


def test_full_model_logits_and_grads_vs_reference_golden(dev, golden):
    z, zp = golden('model.npz'), golden('pmi_synth300.npz')
    cfg = H.MODEL_CFG
    model = build_model(dev, cfg, H.edge_map_from_golden(zp, cfg['V']), int(zp['count'])).eval()
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    # eval-mode forward + gradients (the bi-LSTM is the hand-written recurrence kernel, so no cuDNN eval-mode caveat)
    logits = model(text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    close(model.text_features(text.to(dev)), z['text_feature'], 1e-4, 1e-5)
    # north-star tolerance, fp32 mode
    close(logits, z['logits'], 1e-3, 1e-4)
    assert np.array_equal(logits.argmax(1).cpu().numpy(), z['logits'].argmax(1))
    loss = torch.nn.functional.cross_entropy(logits, labels.to(dev))
    np.testing.assert_allclose(loss.item(), float(z['loss']), rtol=1e-4)
    loss.backward()
    norms = dict(zip(z['grad_names'], z['grad_norms']))
    params = dict(model.named_parameters())
    checked = 0
    for n, ref in norms.items():
        p = params[n]
        if n.endswith('w_ks.bias'):
            continue
        assert p.grad is not None, n
        np.testing.assert_allclose(p.grad.norm().item(), ref, rtol=5e-3, atol=1e-7, err_msg=n)
        checked += 1
    assert checked > 90
    for k in z.files:
        if k.startswith('grad::') and k[6:] in params:
            close(params[k[6:]].grad, z[k], 5e-3, 5e-6, msg=k)
    close(params['text_features.node_hidden.weight'].grad.sum(1), z['grad::node_hidden_rowsum'], 5e-3, 5e-6)
    # parameters the reference never gives a gradient stay gradient-free (DDP bucket contract)
    for n in ('rnn.weight_ih_l0', 'object_gate.weight', 'text_object_text_multi_head_att.slf_attn.fc.weight',
              'object_linear_1.weight', 'text_features.Linear.weight'):
        assert params[n].grad is None, n


def test_branch_streams_match_single_stream(dev, golden):
    """Forking the independent channels / attention stacks onto side CUDA streams (model.branch_streams) runs the
    same kernels in the same per-branch order: logits and every gradient must match the single-stream run."""
    zp = golden('pmi_synth300.npz')
    cfg = dict(H.MODEL_CFG, B=16)
    model = build_model(dev, cfg, H.edge_map_from_golden(zp, cfg['V']), int(zp['count'])).eval()
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    args = (text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    out = {}
    for mode in (False, True):
        model.branch_streams = mode
        model.zero_grad(set_to_none=True)
        for _ in range(3):                                   # repeat: allocator reuse across streams
            model.zero_grad(set_to_none=True)
            logits = model(*args)
            torch.nn.functional.cross_entropy(logits, labels.to(dev)).backward()
        torch.cuda.synchronize()
        out[mode] = (logits.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters()
                                               if p.grad is not None})
    assert torch.equal(out[False][0], out[True][0])
    assert out[False][1].keys() == out[True][1].keys()
    for n, g in out[False][1].items():
        # weight gradients accumulated with fp32 atomics (image bank, text GCN) differ in summation order only
        close(out[True][1][n], g, 1e-4, 1e-6, msg=n)


def test_deferred_lstm_weight_grads_match_inline(dev, ops, golden):
    """ops.defer_weight_grads: the LSTM weight-gradient products run on a side stream and are added to .grad by
    join_deferred(); every gradient must equal the inline autograd path (same kernels, same operands)."""
    zp = golden('pmi_synth300.npz')
    cfg = dict(H.MODEL_CFG, B=12)
    model = build_model(dev, cfg, H.edge_map_from_golden(zp, cfg['V']), int(zp['count'])).eval()
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    args = (text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    grads = {}
    for defer in (False, True):
        for streams in (False, True):
            model.branch_streams = streams
            for _ in range(2):
                model.zero_grad(set_to_none=True)
                loss = torch.nn.functional.cross_entropy(model(*args), labels.to(dev))
                prev = ops.defer_weight_grads(defer)
                try:
                    loss.backward()
                finally:
                    ops.defer_weight_grads(prev)
                    ops.join_deferred()
            torch.cuda.synchronize()
            grads[(defer, streams)] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    base = grads[(False, False)]
    lstm_names = [n for n in base if n.startswith('lstm.')]
    assert len(lstm_names) == 16
    for key, g in grads.items():
        assert g.keys() == base.keys(), key
        for n in lstm_names:
            close(g[n], base[n], 2e-5, 1e-6, msg=str((key, n)))     # split-K atomics: summation order only
        for n in base:
            close(g[n], base[n], 1e-4, 1e-6, msg=str((key, n)))


def test_full_model_bigger_batch_vs_oracle_and_mvsa_labels(dev):
    """B=24 against the CPU oracle on a fresh seed, 7 and 3 labels (MVSA-shaped)."""
    for num_labels, seed in ((7, 23), (3, 24)):
        cfg = dict(H.MODEL_CFG, B=24, V=400, seed=seed, num_labels=num_labels)
        emap, count = synth.synthetic_edge_map(cfg['V'], seed=seed, docs=800)
        model = build_model(dev, cfg, emap, count).eval()
        text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
        with torch.no_grad():
            logits = model(text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
        P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        query = torch.from_numpy(synth.label_graphs()['label_glove'])[:num_labels]
        with torch.no_grad():
            ref = O.model_forward(P, text, lens, mask, fo, fp, oinp[0], pinp[0], query, lambda u, v: emap[u, v], cfg)
        close(logits, ref, 1e-3, 1e-4)
        assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))


def test_cfg5_shaped_model_16_heads_large_label_graphs_vs_oracle(dev):
    """BASELINE cfg 5 at reduced size: n_head=16 (four head groups in the attention kernels), synthetic label graphs
    with 512 object / 640 scene nodes (the SpMM / label-attention paths beyond the shipped 80 / 365), V=500."""
    from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
    from mgnns_b200.api.text_gcn import Model as TextModel
    cfg = dict(H.MODEL_CFG, B=10, V=500, seed=41, n_head=16)
    n_obj, n_plc = 512, 640
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=41, docs=900)
    rs = np.random.RandomState(5)

    def cooc(n):
        nums = rs.randint(50, 500, size=n).astype(np.float64)
        adj = np.zeros((n, n))
        for i in range(n):
            js = rs.choice(n, 6, replace=False)
            adj[i, js] = rs.randint(1, 60, size=6)
        adj = np.minimum(adj + adj.T, nums[:, None])
        np.fill_diagonal(adj, 0)
        return {'adj': adj, 'nums': nums}
    adj_o, adj_p = cooc(n_obj), cooc(n_plc)
    opt = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5,
               emb_type='random', vocab_size=cfg['V'], stack_num=2, n_head=16, d_kv=128, is_regu=False)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, cfg['V'])]
    tm = TextModel(7, 300, vocab, cfg['ngram'], 0.5, count, emap, pmi=torch.zeros(count, 1))
    model = Multi_GCN_Multihead_Att(opt, 7, tm, IdentityTrunk(), IdentityTrunk(), n_obj, n_plc, object_t=0.1, place_t=0.1,
                                    in_channel=300, object_adj_file=adj_o, place_adj_file=adj_p)
    synth.fill_parameters(model, seed=41)
    model = model.to(dev).eval()
    text, lens, mask = synth.make_texts(cfg['B'], cfg['V'], cfg['L'], seed=41)
    fo, fp = synth.make_fmaps(cfg['B'], seed=41), synth.make_fmaps(cfg['B'], seed=42)
    oinp, pinp = synth.label_inputs(cfg['B'], n_obj, n_plc, seed=3)
    labels = synth.make_labels(cfg['B'], 7, seed=41)
    for streams in (False, True):
        model.branch_streams = streams
        model.zero_grad(set_to_none=True)
        logits = model(text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
        torch.nn.functional.cross_entropy(logits, labels.to(dev)).backward()
        P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        for k in list(P):
            if P[k].dtype.is_floating_point and k not in ('object_A', 'place_A'):
                P[k].requires_grad_()
        query = torch.from_numpy(synth.label_graphs()['label_glove'])
        ref = O.model_forward(P, text, lens, mask, fo, fp, oinp[0], pinp[0], query, lambda u, v: emap[u, v], cfg)
        close(logits, ref, 1e-3, 1e-4)
        assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
        torch.nn.functional.cross_entropy(ref, labels).backward()
        params = dict(model.named_parameters())
        for n in ('gc1.weight', 'gc2.weight', 'object_attention.w_k.weight', 'place_attention.w_v.weight',
                  'text_img_object_multi_head_att.0.slf_attn.w_qs.weight', 'img_place_text_multi_head_att.1.slf_attn.w_vs.weight',
                  'img_object_text_multi_head_att.1.slf_attn.fc.weight', 'lstm.weight_hh_l1_reverse', 'liner_img_place.weight',
                  'multi_linear_1.weight'):
            np.testing.assert_allclose(params[n].grad.norm().item(), P[n].grad.norm().item(), rtol=5e-3, err_msg=n)


def test_cfg5_full_size_32_sample_slice_vs_oracle(dev, ops):
    """BASELINE.json configs[4] at its full model size — V = 50,000 vocabulary, 4096 object + 4096 scene label nodes
    (A_hat 0.4 % full: the label GCN runs on the SpMM + tcgen05 path with 4096 rows), 16 attention heads (two head
    tiles in the tensor-core attention) — on a 32-sample slice of the batch against the CPU oracle: logits within
    1e-3 rel / 1e-4 abs, identical arg-max, gradient norms of one parameter per kernel family."""
    from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
    from mgnns_b200.api.text_gcn import Model as TextModel
    n_obj = n_plc = 4096
    cfg = dict(H.MODEL_CFG, B=32, V=50000, seed=51, n_head=16, n_obj=n_obj, n_plc=n_plc, object_t=0.04, place_t=0.04)
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=51, docs=3000)
    adj_o, adj_p = synth.synthetic_label_graph(n_obj, seed=80), synth.synthetic_label_graph(n_plc, seed=365)
    opt = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5,
               emb_type='random', vocab_size=cfg['V'], stack_num=2, n_head=16, d_kv=128, is_regu=False)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, cfg['V'])]
    tm = TextModel(7, 300, vocab, cfg['ngram'], 0.5, count, emap, pmi=torch.zeros(count, 1))
    model = Multi_GCN_Multihead_Att(opt, 7, tm, IdentityTrunk(), IdentityTrunk(), n_obj, n_plc, object_t=0.04, place_t=0.04,
                                    in_channel=300, object_adj_file=adj_o, place_adj_file=adj_p)
    synth.fill_parameters(model, seed=51)
    model = model.to(dev).eval()
    assert ops.attn_uses_tensor_cores(16, 196, 300) and ops.attn_uses_tensor_cores(16, 100, 300)
    text, lens, mask = synth.make_texts(cfg['B'], cfg['V'], cfg['L'], seed=51)
    fo, fp = synth.make_fmaps(cfg['B'], seed=51), synth.make_fmaps(cfg['B'], seed=52)
    oinp, pinp = synth.label_inputs(cfg['B'], n_obj, n_plc, seed=3)
    labels = synth.make_labels(cfg['B'], 7, seed=51)
    model.branch_streams = True
    logits = model(text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    torch.nn.functional.cross_entropy(logits, labels.to(dev)).backward()
    csr = model._adj_csr('object_A')
    assert 0.003 < csr.nnz / n_obj ** 2 < 0.006
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    watch = ('gc1.weight', 'gc2.weight', 'object_attention.w_k.weight', 'text_img_place_multi_head_att.0.slf_attn.w_qs.weight',
             'img_object_text_multi_head_att.1.slf_attn.w_vs.weight', 'lstm.weight_hh_l0', 'liner_img_object.weight')
    for k in watch:
        P[k].requires_grad_()
    query = torch.from_numpy(synth.label_graphs()['label_glove'])
    ref = O.model_forward(P, text, lens, mask, fo, fp, oinp[0], pinp[0], query, lambda u, v: emap[u, v], cfg)
    close(logits, ref, 1e-3, 1e-4)
    assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))
    torch.nn.functional.cross_entropy(ref, labels).backward()
    params = dict(model.named_parameters())
    for n in watch:
        np.testing.assert_allclose(params[n].grad.norm().item(), P[n].grad.norm().item(), rtol=5e-3, err_msg=n)


def test_flat_clip_adam_matches_clip_grad_norm_plus_torch_adam(dev):
    """mgnns_sqnorm_f32 + mgnns_clip_adam_f32 (two launches over flat buffers) against the reference's sequence —
    optimizer.zero_grad(); backward; clip_grad_norm_(model.parameters(), 10); torch.optim.Adam.step() with the twelve
    parameter groups of get_config_optim — over four steps with a clip that bites (max_norm 0.05) and one that does
    not: same parameters, same Adam moments, same (scaled, accumulating) gradients of the never-stepped parameters.

    Both optimizers are fed the SAME fresh gradients every step (torch.autograd.grad of model A, accumulated into both
    twins the way AccumulateGrad does: assigned where .grad is None, added in place otherwise).  Two separately
    differentiated twins would not do: their gradients differ by fp32 atomic-order noise, which Adam (update ~ lr *
    sign(g) wherever g is small) amplifies into O(lr) parameter differences that then compound through the model."""
    from mgnns_b200.optim import FlatClipAdam, FlatGradients
    cfg = dict(H.MODEL_CFG, B=8, seed=17)
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=17, docs=400)
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    args = (text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    # Conditioning: with a clip that bites AND weight decay, Adam sees coef*g + wd*p; the two implementations' clip
    # coefficients differ by ~1e-6 relative (double vs float accumulation of the norm), and wherever the two terms
    # nearly cancel that is amplified ~1000x into the normalised update (observed 1.6e-6 absolute on 1 of 307,200
    # elements).  So: tight bounds on (clip, no decay) — Adam is invariant to the gradient scale — and on (no clip,
    # decay) — coefficient exactly 1 in both — and 1e-3 of the largest possible move for the combination.
    for max_norm, wd, p_atol in ((0.05, 0.0, 1e-6), (1e6, 1e-2, 1e-6), (0.05, 1e-2, 5e-5)):
        ma = build_model(dev, cfg, emap, count).eval()
        mb = build_model(dev, cfg, emap, count).eval()
        oa = torch.optim.Adam(ma.get_config_optim(5e-3, 0.1), lr=5e-3, weight_decay=wd)
        ob = torch.optim.Adam(mb.get_config_optim(5e-3, 0.1), lr=5e-3, weight_decay=wd)
        pa_list, pb_list = list(ma.parameters()), list(mb.parameters())
        fg = flat = None
        for step in range(4):
            oa.zero_grad()
            if flat is None:
                ob.zero_grad()
            else:
                flat.zero_grad()
            loss = torch.nn.functional.cross_entropy(ma(*args), labels.to(dev))
            fresh = torch.autograd.grad(loss, pa_list, allow_unused=True)
            for a, b, g in zip(pa_list, pb_list, fresh):
                if g is None:
                    continue
                for p in (a, b):
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.add_(g)
            na = torch.nn.utils.clip_grad_norm_(ma.parameters(), max_norm=max_norm)
            oa.step()
            if fg is None:
                fg = FlatGradients(mb.parameters())
            fg.pack()
            if flat is None:
                flat = FlatClipAdam(ob, fg, max_norm)
            flat.step()
            np.testing.assert_allclose(flat.total_norm(), float(na), rtol=2e-5)
        pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
        moved = 0
        for n in pa:
            # parameters moved by up to 4 x lr x 10 = 0.2: 1e-6 absolute is 5e-6 of the update (fp32 rounding of m / sqrt(v))
            close(pb[n], pa[n], 2e-5, p_atol, msg=n)
            if pa[n].grad is not None:
                close(pb[n].grad, pa[n].grad, 2e-5, 1e-9, msg='grad ' + n)
        for p in flat.owned_params:
            st = oa.state[[q for n, q in pa.items() if pb[n] is p][0]]
            o = flat.p_flat.data_ptr()
            off = (p.data_ptr() - o) // 4
            close(flat.m_flat[off:off + p.numel()].view_as(p), st['exp_avg'], 2e-5, 1e-9 if p_atol <= 1e-6 else 1e-8)
            close(flat.v_flat[off:off + p.numel()].view_as(p), st['exp_avg_sq'], 2e-5, 1e-12)
            moved += 1
        assert moved > 80 and int(flat.step_count) == 4
        assert pb['multi_linear_1.weight'].grad is not None and pb['multi_linear_1.weight'] not in set(flat.owned_params)
        # the twin whose parameters now live in the flat buffer still runs, and agrees with the torch-stepped model
        with torch.no_grad():
            close(mb(*args), ma(*args), 1e-3, 1e-4)


@pytest.mark.parametrize("flat", [False, True])
def test_graphed_train_step_matches_eager_steps(dev, ops, flat):
    """GraphedTrainStep (whole step captured as one CUDA graph: four-stream forward/backward, deferred weight
    gradients, clip, Adam — torch's fused Adam or, flat=True, the two-kernel FlatClipAdam) against the same three
    steps run eagerly on one stream with clip_grad_norm_ + torch.optim.Adam — eval mode, so no dropout randomness;
    parameters must agree to summation-order noise.  Also exercises update_lengths() on a second batch."""
    from mgnns_b200.graph_step import GraphedTrainStep
    cfg = dict(H.MODEL_CFG, B=16, V=300, seed=51)
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=51, docs=500)
    batches = []
    for sd in (51, 52):
        c = dict(cfg, seed=sd)
        text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(c)
        batches.append(dict(text=text.to(dev), lens=lens, mask=mask.to(dev), fo=fo.to(dev), fp=fp.to(dev),
                            oinp=oinp.to(dev), pinp=pinp.to(dev), labels=labels.to(dev)))
    crit = torch.nn.CrossEntropyLoss()

    def fresh():
        m = build_model(dev, cfg, emap, count).eval()
        # a large eps keeps Adam's normalised update proportional to the gradient for near-zero gradients, so that
        # summation-order noise (fp32 atomics) cannot flip an update of size lr
        o = torch.optim.Adam(m.get_config_optim(1e-3, 0.1), lr=1e-3, weight_decay=1e-5, eps=1e-2, capturable=True,
                             fused=True)
        return m, o

    # eager reference: batch 0, batch 1, batch 0
    m_e, o_e = fresh()
    losses_e = []
    for b in (batches[0], batches[1], batches[0]):
        o_e.zero_grad(set_to_none=True)
        loss = crit(m_e(b['text'], b['lens'], b['mask'], b['fo'], b['fp'], b['oinp'], b['pinp']), b['labels'])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), 10.0)
        o_e.step()
        losses_e.append(loss.item())

    # graph: static buffers start as batch 0; warm-up steps inside the constructor would move the parameters, so the
    # model/optimizer state is restored after capture
    m_g, o_g = fresh()
    m_g.branch_streams = True
    static = {k: (v.clone() if torch.is_tensor(v) and k != 'lens' else v) for k, v in batches[0].items()}
    p0 = {n: p.detach().clone() for n, p in m_g.named_parameters()}
    g = GraphedTrainStep(m_g, o_g, crit, static, clip_norm=10.0, world_size=1, warmup=1, plan_capacity=16 * 100,
                         flat_optimizer=True if flat else None)
    with torch.no_grad():
        for n, p in m_g.named_parameters():
            p.copy_(p0[n])
    for st in o_g.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                v.zero_()
    if flat:
        assert len(o_g.state) == 0 and g.flat_opt.p_flat.numel() > 1e6
        g.flat_opt.m_flat.zero_()
        g.flat_opt.v_flat.zero_()
        g.flat_opt.step_count.zero_()
    # Like the reference engine (optimizer.zero_grad(), engine:841), the step only clears the gradients of the
    # parameters the optimizer owns: the never-stepped ones (classifier tail, image-bank Linears, ... SURVEY §0.4)
    # keep ACCUMULATING across steps and count toward the clip norm.  The warm-up step left one such contribution in
    # the graph's static .grad buffers; clear it so that both runs start from the same state.
    for p in m_g.parameters():
        if p.grad is not None:
            p.grad.zero_()
    losses_g = []
    for b in (batches[0], batches[1], batches[0]):
        for k in ('text', 'mask', 'fo', 'fp', 'labels'):
            static[k].copy_(b[k])
        g.update_lengths(b['lens'])
        losses_g.append(g.replay().item())
    torch.cuda.synchronize()
    np.testing.assert_allclose(losses_g, losses_e, rtol=2e-4)
    pe = dict(m_e.named_parameters())
    worst = 0.0
    for n, p in m_g.named_parameters():
        d = (p.detach() - pe[n].detach()).abs().max().item()
        worst = max(worst, d)
        assert d < 2e-5, (n, d)
    moved = max((p.detach() - p0[n]).abs().max().item() for n, p in m_g.named_parameters())
    assert moved > 1e-4, "the captured step did not update the parameters"


@pytest.mark.parametrize("num_labels", [7, 3])
def test_cfg3_full_size_inference_batch_independence(dev, num_labels):
    """BASELINE cfg 3 (inference, batch 1024, TumEmo 7 / MVSA 3 classes) at full size through size-independent
    properties: every sample's logits are independent of what else is in the batch (the first 32 samples alone, and a
    permuted batch, give the same rows), and a 32-sample slice matches the CPU oracle."""
    B = 1024
    cfg = dict(H.MODEL_CFG, B=B, V=2000, seed=61 + num_labels, num_labels=num_labels)
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=61, docs=3000)
    model = build_model(dev, cfg, emap, count).eval()
    model.branch_streams = True
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    dt = lambda t: t.to(dev)          # noqa: E731
    with torch.no_grad():
        full = model(dt(text), lens, dt(mask), dt(fo), dt(fp), dt(oinp), dt(pinp))
        head = model(dt(text[:32]), lens[:32], dt(mask[:32]), dt(fo[:32]), dt(fp[:32]), dt(oinp[:32]), dt(pinp[:32]))
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
        shuf = model(dt(text[perm]), lens[perm], dt(mask[perm]), dt(fo[perm]), dt(fp[perm]), dt(oinp[perm]), dt(pinp[perm]))
    assert full.shape == (B, num_labels) and torch.isfinite(full).all()
    # the image-bank tiles straddle sample boundaries and the LSTM tiles depend on the length order: fp32 rounding only
    close(full[:32], head, 1e-4, 1e-5)
    close(shuf, full[perm.to(dev)], 1e-4, 1e-5)
    assert torch.equal(shuf.argmax(1), full[perm.to(dev)].argmax(1))
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    query = torch.from_numpy(synth.label_graphs()['label_glove'])[:num_labels]
    with torch.no_grad():
        ref = O.model_forward(P, text[:32], lens[:32], mask[:32], fo[:32], fp[:32], oinp[0], pinp[0], query,
                              lambda u, v: emap[u, v], cfg)
    close(full[:32], ref, 1e-3, 1e-4)
    assert torch.equal(full[:32].argmax(1).cpu(), ref.argmax(1))


def test_train_mode_step_is_finite_and_deterministic(dev):
    cfg = dict(H.MODEL_CFG, B=8, V=300, seed=31)
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=31, docs=500)
    model = build_model(dev, cfg, emap, count).train()
    text, lens, mask, fo, fp, oinp, pinp, labels = H.model_inputs(cfg)
    args = (text.to(dev), lens, mask.to(dev), fo.to(dev), fp.to(dev), oinp.to(dev), pinp.to(dev))
    outs = []
    for _ in range(2):
        torch.manual_seed(123)
        torch.cuda.manual_seed(123)
        model.zero_grad(set_to_none=True)
        logits = model(*args)
        loss = torch.nn.functional.cross_entropy(logits, labels.to(dev))
        loss.backward()
        outs.append((logits.detach().clone(), model.gc1.weight.grad.detach().clone()))
        assert torch.isfinite(logits).all()
        for n, p in model.named_parameters():
            if p.grad is not None:
                assert torch.isfinite(p.grad).all(), n
    assert torch.equal(outs[0][0], outs[1][0])
    close(outs[0][1], outs[1][1], 1e-4, 1e-6)   # float atomics may reorder sums


def test_cpu_tensors_are_rejected(dev):
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.mgnns.mm(torch.randn(3, 3), torch.randn(3, 3), None, False, False, 0, 0.0)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.mgnns.add_layernorm(torch.randn(3, 8), None, torch.ones(8), torch.zeros(8), 1e-6)


def test_text_bank_glue_kernels_match_torch(dev, ops):
    """mgnns_pad_rows_fwd/bwd and mgnns_embedding_bwd against the torch formulation they replace (new_zeros + index_copy
    into the padded bank; nn.Embedding's dense backward with padding_idx): values bit-equal (pure data movement),
    table gradient to summation order; lengths include 1, the maximum, and a text longer than L (clamped)."""
    torch.manual_seed(3)
    B, L, F, V, E = 9, 100, 300, 57, 300
    lens = torch.tensor([1, 100, 7, 130, 2, 55, 100, 3, 18])
    plan = ops.LstmPlan(lens, L, dev, capacity=512)
    assert int(plan.N) == int(lens.clamp(max=L).sum())
    y = torch.randn(plan.capacity, F, device=dev)
    y[plan.N:] = 0
    y1 = y.clone().requires_grad_()
    bank = ops.pad_text_bank(y1, plan, B, L)
    y2 = y.clone().requires_grad_()
    ref = y2.new_zeros(B * L + 1, F).index_copy(0, plan.flat_idx, y2)[:B * L].view(B, L, F)
    assert torch.equal(bank, ref)
    for b in range(B):
        assert bank[b, min(int(lens[b]), L):].abs().max().item() == 0 if int(lens[b]) < L else True
    g = torch.randn_like(bank)
    bank.backward(g)
    ref.backward(g)
    assert torch.equal(y1.grad, y2.grad)
    # embedding rows: tokens with repeats and the padding index
    emb = torch.nn.Embedding(V, E, padding_idx=0).to(dev)
    tokens = torch.randint(0, V, (700,), device=dev)
    tokens[::13] = 0
    w1 = emb.weight.detach().clone().requires_grad_()
    out = ops.embedding_rows(w1, tokens, 0)
    ref = emb(tokens)
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g)
    close(w1.grad, emb.weight.grad, 1e-5, 1e-6)
    assert w1.grad[0].abs().max().item() == 0
