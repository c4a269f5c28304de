"""Shared builders for the parity tests (same seeds as oracle/make_golden.py)."""
import numpy as np
import torch

from mgnns_b200 import synth

MODEL_CFG = dict(B=6, V=300, L=100, ngram=4, n_head=4, d_kv=128, stack_num=2, hidden_size=150, num_layers=2,
                 object_t=0.4, place_t=0.3, num_labels=7, seed=11)


def model_inputs(cfg=MODEL_CFG):
    B, V, L, seed = cfg['B'], cfg['V'], cfg['L'], cfg['seed']
    text, lens, mask = synth.make_texts(B, V, L, seed=seed)
    fo, fp = synth.make_fmaps(B, seed=seed), synth.make_fmaps(B, seed=seed + 1)
    oinp, pinp = synth.label_inputs(B)
    labels = synth.make_labels(B, cfg['num_labels'], seed=seed)
    return text, lens, mask, fo, fp, oinp, pinp, labels


def edge_lookup_from_golden(z, V):
    """dict-based edges_matrix lookup from the (rows, cols, ids) stored in a pmi golden file."""
    table = {(int(r), int(c)): int(i) for r, c, i in zip(z['rows'], z['cols'], z['ids'])}
    return lambda u, v: table.get((int(u), int(v)), 0)


def edge_map_from_golden(z, V):
    from mgnns_b200.api.pmi import SparseEdgeMap
    rows, cols, ids = z['rows'].astype(np.int64), z['cols'].astype(np.int64), z['ids'].astype(np.int64)
    rowptr = np.zeros(V + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return SparseEdgeMap(np.cumsum(rowptr), cols, V, eid=ids)


def state_shapes(cfg=MODEL_CFG, edge_count=1):
    """Parameter names/shapes of the model head (state_dict contract) without building any module."""
    d, H, dkv, hs = 300, cfg['n_head'], cfg['d_kv'], cfg['hidden_size']
    S = {}
    S['text_features.node_hidden.weight'] = (cfg['V'], d)
    S['text_features.seq_edge_w.weight'] = (edge_count, 1)
    S['embedding.weight'] = (cfg['V'], d)
    for layer in range(cfg['num_layers']):
        inp = d if layer == 0 else 2 * hs
        for suf in ('', '_reverse'):
            S['lstm.weight_ih_l%d%s' % (layer, suf)] = (4 * hs, inp)
            S['lstm.weight_hh_l%d%s' % (layer, suf)] = (4 * hs, hs)
            S['lstm.bias_ih_l%d%s' % (layer, suf)] = (4 * hs,)
            S['lstm.bias_hh_l%d%s' % (layer, suf)] = (4 * hs,)
    for stack in ('img_object_text', 'img_place_text', 'text_img_object', 'text_img_place'):
        for i in range(cfg['stack_num']):
            p = '%s_multi_head_att.%d.' % (stack, i)
            for w in ('w_qs', 'w_ks', 'w_vs'):
                S[p + 'slf_attn.%s.weight' % w] = (H * dkv, d)
                S[p + 'slf_attn.%s.bias' % w] = (H * dkv,)
            S[p + 'slf_attn.fc.weight'] = (d, H * dkv)
            S[p + 'slf_attn.fc.bias'] = (d,)
            S[p + 'slf_attn.layer_norm.gamma'] = (d,)
            S[p + 'slf_attn.layer_norm.beta'] = (d,)
            S[p + 'pos_ffn.w_1.weight'] = (d, d, 1)
            S[p + 'pos_ffn.w_1.bias'] = (d,)
            S[p + 'pos_ffn.w_2.weight'] = (d, d, 1)
            S[p + 'pos_ffn.w_2.bias'] = (d,)
            S[p + 'pos_ffn.layer_norm.gamma'] = (d,)
            S[p + 'pos_ffn.layer_norm.beta'] = (d,)
    S['liner_img_object.weight'] = (d, 2048); S['liner_img_object.bias'] = (d,)
    S['liner_img_place.weight'] = (d, 2048); S['liner_img_place.bias'] = (d,)
    S['gc1.weight'] = (300, 1024); S['gc2.weight'] = (1024, 2048)
    for kind, n in (('object', cfg.get('n_obj', 80)), ('place', cfg.get('n_plc', 365))):
        a = kind + '_attention.'
        S[a + 'w_q.weight'] = (300, 300); S[a + 'w_q.bias'] = (300,)
        S[a + 'w_k.weight'] = (300, n); S[a + 'w_k.bias'] = (300,)
        S[a + 'w_v.weight'] = (300, n); S[a + 'w_v.bias'] = (300,)
        S[a + 'fc.weight'] = (300, 300); S[a + 'fc.bias'] = (300,)
        S[kind + '_linear_5.weight'] = (100, 300); S[kind + '_linear_5.bias'] = (100,)
        S[kind + '_x_linear.weight'] = (300, 100 * cfg['num_labels']); S[kind + '_x_linear.bias'] = (300,)
    S['multi_linear_1.weight'] = (300, 1200); S['multi_linear_1.bias'] = (300,)
    S['multi_linear_2.weight'] = (cfg['num_labels'], 300); S['multi_linear_2.bias'] = (cfg['num_labels'],)
    return S


def oracle_params(cfg=MODEL_CFG, edge_count=1, dtype=torch.float32):
    """Deterministic parameters for the oracle, identical to synth.fill_parameters on the real module."""
    P = {n: torch.zeros(s) for n, s in state_shapes(cfg, edge_count).items()}
    synth.fill_parameters(P, seed=cfg['seed'])
    P['embedding.weight'][0].zero_()
    from oracle import mgnns_oracle as O
    for kind, n, t, default in (('object', cfg.get('n_obj', 80), cfg['object_t'], 80),
                                ('place', cfg.get('n_plc', 365), cfg['place_t'], 365)):
        # shipped label graphs at the default sizes, synth.synthetic_label_graph otherwise (cfg 5)
        a = synth.adj_dict(kind) if n == default else synth.synthetic_label_graph(n, seed=default)
        P[kind + '_A'] = torch.from_numpy(O.gen_A(n, t, a['adj'], a['nums'])).float()
    return {k: v.to(dtype) if v.dtype.is_floating_point else v for k, v in P.items()}
