"""CPU oracle for the MGNNS hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch (CPU, fp32 or fp64) / numpy restatement of the reference's algorithm, each function
citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; nothing under mgnns_b200/ does.

Pinning (see oracle/make_golden.py and tests/test_oracle_golden.py): every function below is
checked against outputs of the reference's own code executed in the authoring container
(models/submodules.py, models/moudles.py, GraphConvolution, Attention, Multi_GCN_Multihead_Att,
utils/util.py gen_A/gen_adj, utils/pmi.py cal_PMI), frozen under tests/golden/.
PARITY UNPINNED at one boundary: DGL (dgl.DGLGraph / update_all(src_mul_edge, max) / sum_nodes,
models/Text_GCN.py:181-268) is an un-vendored, un-versioned dependency; its semantics are restated
from its published behaviour (builtin max over in-edges, zero fill for zero-in-degree nodes).

All functions are functional: parameters come from a dict keyed by the reference's state_dict
names, so the same state_dict drives the oracle and the CUDA modules, and autograd through the
oracle yields reference gradients.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------ label graph
def gen_A(num_classes, t, adj, nums, gama=0.2):
    """utils/util.py:382-398 — conditional probabilities, binarise at t, re-weight, add self loops."""
    adj = np.asarray(adj, dtype=np.float64)
    nums = np.asarray(nums, dtype=np.float64).reshape(-1, 1)
    p = adj / nums
    b = np.zeros_like(p)
    b[p >= t] = 1.0
    b[np.isnan(p)] = np.nan
    a = b * gama / (b.sum(0, keepdims=True) + 1e-6)
    return a + (1 - gama) * np.identity(num_classes)


def gen_adj(A):
    """utils/util.py:421-426 — (A·D)^T·D with D = diag(rowsum^-1/2), as two dense products."""
    D = torch.diag(torch.pow(A.sum(1), -0.5))
    return torch.matmul(torch.matmul(A, D).t(), D)


def graph_convolution(x, adj, weight, bias=None):
    """models/Multi_GCN_Multihead_att.py:52-58 — support = X·W; out = Â·support (+bias)."""
    out = torch.matmul(adj, torch.matmul(x, weight))
    return out if bias is None else out + bias


# ------------------------------------------------------------------------------ label attention
def label_attention(P, prefix, query, key, value, n_heads, dropout_mask=None):
    """models/Multi_GCN_Multihead_att.py:88-133 in closed form.

    energy[b,c,h,d] = Q[c,h,d]*K[b,h,d]/sqrt(dh); softmax over d; (dropout); * V[b,h,d]; fc.
    The reference builds energy with an O(B^2) cat loop (:111-115) — same values.
    """
    hid = P[prefix + 'w_q.weight'].shape[0]
    dh = hid // n_heads
    Q = F.linear(query.to(P[prefix + 'w_q.weight'].dtype), P[prefix + 'w_q.weight'], P[prefix + 'w_q.bias'])
    K = F.linear(key, P[prefix + 'w_k.weight'], P[prefix + 'w_k.bias'])
    V = F.linear(value, P[prefix + 'w_v.weight'], P[prefix + 'w_v.bias'])
    C, B = Q.shape[0], K.shape[0]
    Q = Q.view(1, C, n_heads, dh)
    K = K.view(B, 1, n_heads, dh)
    V = V.view(B, 1, n_heads, dh)
    energy = (Q * K) / math.sqrt(dh)
    att = torch.softmax(energy, dim=-1)
    if dropout_mask is not None:
        att = att * dropout_mask
    x = (att * V).reshape(B, C, hid)
    return F.linear(x, P[prefix + 'fc.weight'], P[prefix + 'fc.bias'])


# ------------------------------------------------------------------------------ transformer pieces
def layer_norm(x, gamma, beta, eps=1e-6):
    """models/submodules.py:153-156 — unbiased std, eps added to std."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return gamma * (x - mean) / (std + eps) + beta


def multi_head_attention(P, prefix, q, k, v, mask, n_head, d_k, d_v):
    """models/submodules.py:55-94 (+ :106-119), eval mode (no dropout).

    q [B,lq,d], k/v [B,lk,d], mask [B,lq,lk] float (0 = masked) or None.
    Returns (out [B,lq,d], attn [n_head*B, lq, lk]) with the reference's head-major layout.
    """
    B, lq, _ = q.shape
    lk = k.shape[1]
    residual = q
    qh = F.linear(q, P[prefix + 'w_qs.weight'], P[prefix + 'w_qs.bias']).view(B, lq, n_head, d_k)
    kh = F.linear(k, P[prefix + 'w_ks.weight'], P[prefix + 'w_ks.bias']).view(B, lk, n_head, d_k)
    vh = F.linear(v, P[prefix + 'w_vs.weight'], P[prefix + 'w_vs.bias']).view(B, lk, n_head, d_v)
    qh = qh.permute(2, 0, 1, 3).reshape(-1, lq, d_k)
    kh = kh.permute(2, 0, 1, 3).reshape(-1, lk, d_k)
    vh = vh.permute(2, 0, 1, 3).reshape(-1, lk, d_v)
    attn = torch.bmm(qh, kh.transpose(1, 2)) / (d_k ** 0.5)
    if mask is not None:
        attn = attn.masked_fill(mask.repeat(n_head, 1, 1) == 0.0, float('-inf'))
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, vh).view(n_head, B, lq, d_v).permute(1, 2, 0, 3).reshape(B, lq, -1)
    out = F.linear(out, P[prefix + 'fc.weight'], P[prefix + 'fc.bias'])
    out = layer_norm(out + residual, P[prefix + 'layer_norm.gamma'], P[prefix + 'layer_norm.beta'])
    return out, attn


def positionwise_ffn(P, prefix, x):
    """models/submodules.py:132-139 — Conv1d(k=1) pair == two Linears over the feature dim."""
    w1 = P[prefix + 'w_1.weight'].squeeze(-1)
    w2 = P[prefix + 'w_2.weight'].squeeze(-1)
    o = F.linear(F.relu(F.linear(x, w1, P[prefix + 'w_1.bias'])), w2, P[prefix + 'w_2.bias'])
    return layer_norm(o + x, P[prefix + 'layer_norm.gamma'], P[prefix + 'layer_norm.beta'])


def my_multi_head_attention(P, prefix, q, k, v, mask, n_head, d_kv):
    """models/moudles.py:207-230 — unsqueeze q/mask, slf_attn, pos_ffn, squeeze."""
    if q.dim() == 2:
        q = q.unsqueeze(1)
    if mask is not None:
        mask = mask.unsqueeze(1)
    out, attn = multi_head_attention(P, prefix + 'slf_attn.', q, k, v, mask, n_head, d_kv, d_kv)
    out = positionwise_ffn(P, prefix + 'pos_ffn.', out)
    return out.squeeze(1), attn


# ------------------------------------------------------------------------------ TextLevelGCN
def text_doc_edges(doc_ids, ngram, max_length=100):
    """models/Text_GCN.py:168-174 (node set incl. PAD) and :142-166 (edges, PAD stripped, explicit
    self loop).  Returns (sorted node list, [(src_word, dst_word), ...])."""
    ids = list(doc_ids)[:max_length]
    nodes = sorted(set(ids))
    seq = [t for t in ids if t != 0]
    edges = []
    for p, s in enumerate(seq):
        for q in range(max(0, p - ngram), min(p + ngram + 1, len(seq))):
            edges.append((s, seq[q]))
        edges.append((s, s))
    return nodes, edges


def text_gcn_forward(doc_ids, node_hidden, seq_edge_w, edge_id, ngram, max_length=100, relu=True):
    """models/Text_GCN.py:213-275 without DGL, eval mode.

    message(u->v) = h[u] * w[edges_matrix[u,v]] (:242-245, dgl.function.src_mul_edge); reduce = max
    over in-edges per feature (:247); nodes with no in-edge (the PAD node) get 0 — DGL's documented
    zero fill for builtin reducers [PARITY UNPINNED: DGL unversioned]; eta = 0 so the old state is
    discarded (:258-262); readout = sum over the document's nodes (:268); ReLU (:271).
    edge_id(u, v) -> int is the edges_matrix lookup.
    """
    outs = []
    w_flat = seq_edge_w.reshape(-1)
    for doc in doc_ids.tolist():
        nodes, edges = text_doc_edges(doc, ngram, max_length)
        slot = {w: i for i, w in enumerate(nodes)}
        if not edges:
            outs.append(torch.zeros(node_hidden.shape[1], dtype=node_hidden.dtype))
            continue
        src = torch.tensor([e[0] for e in edges])
        dst = torch.tensor([slot[e[1]] for e in edges])
        eid = torch.tensor([edge_id(e[0], e[1]) for e in edges])
        msg = node_hidden[src] * w_flat[eid].unsqueeze(1)
        agg = torch.zeros(len(nodes), node_hidden.shape[1], dtype=node_hidden.dtype)
        agg = agg.scatter_reduce(0, dst.unsqueeze(1).expand_as(msg), msg, reduce='amax', include_self=False)
        outs.append(agg.sum(0))
    out = torch.stack(outs)
    return F.relu(out) if relu else out


# ------------------------------------------------------------------------------ whole head
def lstm_memory_bank(P, text, text_lens, hidden_size, num_layers, prefix='lstm.'):
    """models/Multi_GCN_Multihead_att.py:366-398 — embedding, packed bi-LSTM, re-pad to L."""
    emb = F.embedding(text, P['embedding.weight'], padding_idx=0)
    lstm = torch.nn.LSTM(emb.shape[-1], hidden_size, num_layers=num_layers, bidirectional=True, batch_first=True)
    lstm = lstm.to(emb.dtype)
    params = {n: P[prefix + n] for n, _ in lstm.named_parameters()}
    packed = torch.nn.utils.rnn.pack_padded_sequence(emb, text_lens.cpu(), batch_first=True, enforce_sorted=False)
    out, _ = torch.func.functional_call(lstm, params, (packed,))
    bank, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=text.shape[1])
    return bank


def label_channel(P, fmap, inp, A, att_prefix, lin5, xlin, query, n_heads=5):
    """models/Multi_GCN_Multihead_att.py:452-479 (object) / :484-506 (place)."""
    B = fmap.shape[0]
    flat = fmap.reshape(B, fmap.shape[1], -1)
    bank_in = flat.permute(0, 2, 1)                                   # [B,196,2048]
    pooled = flat.max(dim=2)[0]                                       # MaxPool2d(14,14)
    adj = gen_adj(A).detach()
    x = graph_convolution(inp, adj, P['gc1.weight'])
    x = F.leaky_relu(x, 0.2)
    x = graph_convolution(x, adj, P['gc2.weight']).transpose(0, 1)    # [2048,N]
    scores = torch.matmul(pooled, x)                                  # [B,N]
    att = label_attention(P, att_prefix, query, scores, scores, n_heads)
    att = F.linear(att, P[lin5 + '.weight'], P[lin5 + '.bias']).reshape(B, -1)
    return F.linear(att, P[xlin + '.weight'], P[xlin + '.bias']), bank_in, pooled


def model_forward(P, text, text_lens, text_mask, object_fmap, place_fmap, object_inp, place_inp, query, edge_id,
                  cfg, return_intermediates=False):
    """models/Multi_GCN_Multihead_att.py:431-567, eval mode, on pre-extracted trunk feature maps.

    P: state_dict-named tensors.  cfg: dict(ngram, n_head, d_kv, stack_num, hidden_size, num_layers).
    """
    text_feature = text_gcn_forward(text, P['text_features.node_hidden.weight'],
                                    P['text_features.seq_edge_w.weight'], edge_id, cfg['ngram'])
    text_bank = lstm_memory_bank(P, text, text_lens, cfg['hidden_size'], cfg['num_layers'])

    obj_att, obj_in, obj_pooled = label_channel(P, object_fmap, object_inp, P['object_A'], 'object_attention.',
                                                'object_linear_5', 'object_x_linear', query)
    obj_bank = F.linear(obj_in, P['liner_img_object.weight'], P['liner_img_object.bias'])
    plc_att, plc_in, plc_pooled = label_channel(P, place_fmap, place_inp, P['place_A'], 'place_attention.',
                                                'place_linear_5', 'place_x_linear', query)
    plc_bank = F.linear(plc_in, P['liner_img_place.weight'], P['liner_img_place.bias'])

    H, dkv = cfg['n_head'], cfg['d_kv']

    def stack(name, q, bank, mask):
        for i in range(cfg['stack_num']):
            q, _ = my_multi_head_attention(P, '%s.%d.' % (name, i), q, bank, bank, mask, H, dkv)
        return q

    iot = stack('img_object_text_multi_head_att', obj_att, text_bank, text_mask)
    ipt = stack('img_place_text_multi_head_att', plc_att, text_bank, text_mask)
    tio = stack('text_img_object_multi_head_att', text_feature, obj_bank, None)
    tip = stack('text_img_place_multi_head_att', text_feature, plc_bank, None)
    feat = torch.cat([tio, tip, iot, ipt], dim=1)
    feat = F.linear(feat, P['multi_linear_1.weight'], P['multi_linear_1.bias'])
    logits = F.linear(feat, P['multi_linear_2.weight'], P['multi_linear_2.bias'])
    if return_intermediates:
        return logits, dict(text_feature=text_feature, text_bank=text_bank, obj_att=obj_att, plc_att=plc_att,
                            obj_bank=obj_bank, plc_bank=plc_bank, obj_pooled=obj_pooled, plc_pooled=plc_pooled,
                            iot=iot, ipt=ipt, tio=tio, tip=tip)
    return logits
