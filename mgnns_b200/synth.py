"""Synthetic TumEmo-shaped inputs and deterministic weights.

Used by tests, bench.py and the golden-vector generator so that every consumer sees bit-identical
inputs: all randomness comes from CPU torch.Generators seeded from explicit integers.
"""
import os
import zlib

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'label_graphs.npz')
TUMEMO_LABEL_P = (.075, .142, .084, .109, .298, .203, .088)   # angry, bored, calm, fear, happy, love, sad


def label_graphs():
    """Shipped label-graph fixtures: raw co-occurrence counts, occurrence counts, label-node GloVe-300."""
    z = np.load(_DATA)
    return {k: z[k] for k in z.files}


def adj_dict(kind):
    z = label_graphs()
    return {'adj': z[kind + '_adj'].astype(np.float64), 'nums': z[kind + '_nums'].astype(np.float64)}


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def make_texts(B, V, L=100, seed=0, unk_rate=0.05, zipf_a=1.1):
    """ids int64 [B,L] (0 = PAD tail, 1 = UNK), lens int64 [B], mask f32 [B,L].

    Lengths ~ clip(round(LogNormal(2.3, 1.0)), 2, L) (median 10, mean 16, p95 52 — the TumEmo val
    split); ids Zipf(1.1) over [2, V) with 5 % UNK.
    """
    g = _gen(1234 + seed)
    lens = torch.exp(torch.randn(B, generator=g) * 1.0 + 2.3).round().clamp(2, L).to(torch.int64)
    ranks = torch.arange(1, V - 1, dtype=torch.float64)
    prob = ranks.pow(-zipf_a)
    prob = prob / prob.sum()
    ids = torch.multinomial(prob, B * L, replacement=True, generator=g).view(B, L) + 2
    unk = torch.rand(B, L, generator=g) < unk_rate
    ids[unk] = 1
    pos = torch.arange(L).unsqueeze(0)
    mask = (pos < lens.unsqueeze(1))
    ids = ids * mask
    return ids.to(torch.int64), lens, mask.to(torch.float32)


def make_fmaps(B, seed=0, C=2048, hw=14):
    """relu(N(0,1)) trunk outputs [B,C,hw,hw] (head-only runs)."""
    g = _gen(4321 + seed)
    return torch.relu(torch.randn(B, C, hw, hw, generator=g))


def make_labels(B, num_labels=7, seed=0):
    g = _gen(999 + seed)
    if num_labels == 7:
        p = torch.tensor(TUMEMO_LABEL_P, dtype=torch.float64)
    else:
        p = torch.full((num_labels,), 1.0 / num_labels, dtype=torch.float64)
    return torch.multinomial(p, B, replacement=True, generator=g)


def label_inputs(B, n_obj=80, n_place=365, seed=0):
    """object_inp [B,n_obj,300], place_inp [B,n_place,300]: the shipped GloVe matrices broadcast
    (ref dataset:265), or N(0,0.5^2) for non-default sizes."""
    z = label_graphs()
    g = _gen(77 + seed)

    def one(n, key):
        if z[key].shape[0] == n:
            m = torch.from_numpy(z[key]).float()
        else:
            m = torch.randn(n, 300, generator=g) * 0.5
        return m.unsqueeze(0).expand(B, n, 300)
    return one(n_obj, 'object_glove'), one(n_place, 'place_glove')


def synthetic_edge_map(V, seed=0, docs=2000, window=6, min_cooc=2, L=100):
    """PMI edge-id map from a synthetic corpus drawn by the same text generator, computed on the host
    with numpy (setup only; the product's counting path is mgnns_b200.api.pmi)."""
    ids, _, _ = make_texts(docs, V, L, seed=seed + 101)
    ids = ids.numpy()
    tok = np.where(ids == 0, 0, ids)            # PAD is vocabulary index 0, like the reference vocab
    centre_ok = tok != 0
    keys = []
    for off in range(-window, window):
        if off == 0:
            continue
        if off > 0:
            c, t, ok = tok[:, :L - off], tok[:, off:], centre_ok[:, :L - off]
        else:
            c, t, ok = tok[:, -off:], tok[:, :L + off], centre_ok[:, -off:]
        keys.append((c[ok].astype(np.int64) * V + t[ok]))
    keys = np.concatenate(keys)
    uniq, cnt = np.unique(keys, return_counts=True)
    wc = np.bincount(tok[centre_ok], minlength=V).astype(np.int64)
    uniq, cnt = uniq[cnt >= min_cooc], cnt[cnt >= min_cooc]
    rows, cols = uniq // V, uniq % V
    total = wc.sum()
    pw = wc / total
    with np.errstate(divide='ignore', invalid='ignore'):
        pmi = np.log((cnt / total) / (pw[rows] * pw[cols]))
    keep = np.nan_to_num(pmi) > 0
    keep &= (pw[rows] * pw[cols]) != 0
    rows, cols = rows[keep], cols[keep]
    rowptr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    from .edge_map import SparseEdgeMap
    return SparseEdgeMap(np.cumsum(rowptr), cols, V), int(keep.sum()) + 1


def fill_parameters(module_or_dict, seed=0, scale=None):
    """Deterministic, construction-order independent weights: each tensor is drawn from its own
    generator seeded by crc32(name) ^ seed.  Shapes decide the scale (fan-in style) unless given.
    Adjacency parameters (object_A / place_A) and LayerNorm gamma/beta are left untouched."""
    items = module_or_dict.items() if isinstance(module_or_dict, dict) else module_or_dict.state_dict().items()
    for name, t in items:
        if not t.dtype.is_floating_point or name.endswith('_A') or name in ('object_A', 'place_A'):
            continue
        g = _gen((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if name.endswith('gamma'):
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif name.endswith('beta') or name.endswith('bias'):
            v = 0.05 * torch.randn(t.shape, generator=g)
        elif name.endswith('seq_edge_w.weight'):
            v = 1.0 + 0.2 * torch.randn(t.shape, generator=g)
        elif name.endswith('node_hidden.weight') or name == 'embedding.weight':
            v = 0.4 * torch.randn(t.shape, generator=g)
        else:
            fan_in = t.shape[1] if t.dim() >= 2 else t.shape[0]
            if name in ('gc1.weight', 'gc2.weight'):
                fan_in = t.shape[0]
            s = scale if scale is not None else 1.0 / (fan_in ** 0.5)
            v = s * torch.randn(t.shape, generator=g)
        with torch.no_grad():
            t.copy_(v.to(t.dtype))
    if not isinstance(module_or_dict, dict) and hasattr(module_or_dict, 'embedding'):
        with torch.no_grad():
            module_or_dict.embedding.weight[0].zero_()


def cfg2_word_graph(N=10000, mean_degree=64, seed=0):
    """SURVEY §8d cfg 2 adjacency: PMI-like word graph in CSR — power-law row degrees (exactly
    `mean_degree`·N distinct off-diagonal neighbours in total, each row's drawn WITHOUT replacement from a
    heavy-tailed popularity law by the Gumbel top-k trick) plus a self loop on every row, so
    nnz = (mean_degree + 1)·N (650,000 at the defaults); rows normalised to sum 1.
    Returns host arrays (rowptr int64 [N+1], col int64 [nnz] sorted within a row, val float32 [nnz])."""
    rs = np.random.RandomState(seed)
    raw = np.clip((rs.pareto(1.3, N) + 1) * 20, 1, min(3000, N - 1))
    deg = np.maximum(1, np.floor(raw * (float(mean_degree) * N / raw.sum()))).astype(np.int64)
    deg = np.minimum(deg, min(3000, N - 1))
    # fix the total to mean_degree*N exactly (largest rows absorb the rounding remainder)
    order = np.argsort(-raw, kind='stable')
    rem, k = int(mean_degree) * N - int(deg.sum()), 0
    while rem != 0:
        i = order[k % N]
        step = 1 if rem > 0 else -1
        if 1 <= deg[i] + step <= min(3000, N - 1):
            deg[i] += step
            rem -= step
        k += 1
    pop = (rs.pareto(1.1, N) + 1)
    logp = np.log(pop / pop.sum())
    cols = []
    for r0 in range(0, N, 500):
        r1 = min(N, r0 + 500)
        key = logp[None, :] + rs.gumbel(size=(r1 - r0, N))
        key[np.arange(r1 - r0), np.arange(r0, r1)] = -np.inf          # the self loop is added explicitly
        kmax = int(deg[r0:r1].max())
        top = np.argpartition(-key, kmax - 1, axis=1)[:, :kmax]
        topk = np.take_along_axis(key, top, axis=1)
        srt = np.argsort(-topk, axis=1)
        top = np.take_along_axis(top, srt, axis=1)
        for i in range(r0, r1):
            cols.append(np.sort(np.concatenate([top[i - r0, :deg[i]], [i]])))
    rowptr = np.concatenate([[0], np.cumsum(deg + 1)]).astype(np.int64)
    cols = np.concatenate(cols).astype(np.int64)
    rows = np.repeat(np.arange(N), deg + 1)
    val = (1.0 / (deg + 1))[rows].astype(np.float32)
    return rowptr, cols, val


def synthetic_label_graph(n, neighbours=8, seed=0):
    """Co-occurrence statistics of an n-node label graph in the shape of the shipped pickles ({'adj','nums'},
    ref: utils/util.py:359-380): every node co-occurs with `neighbours` random others (symmetrised, so ~2x that
    per row).  With t = 0.04 every listed edge survives gen_A's threshold: Â has ~(2*neighbours+1)*n non-zeros
    (0.4 % at n = 4096, the BASELINE cfg-5 label graphs)."""
    rs = np.random.RandomState(seed)
    nums = np.full(n, 1000.0)
    adj = np.zeros((n, n))
    rows = np.repeat(np.arange(n), neighbours)
    cols = rs.randint(0, n, size=n * neighbours)
    adj[rows, cols] = rs.randint(50, 500, size=n * neighbours)
    adj = np.maximum(adj, adj.T)
    np.fill_diagonal(adj, 0)
    return {'adj': adj, 'nums': nums}
