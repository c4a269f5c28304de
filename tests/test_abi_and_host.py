"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic
(adjacency builders, PMI float stage, edge map, module construction/state_dict contract)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mgnns_test_helpers as H
from mgnns_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, 'include', 'mgnns_b200.h')) as f:
        src = re.sub(r'/\*.*?\*/', '', f.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(mgnns_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from mgnns_b200 import _abi
    lib = ctypes.CDLL(_abi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
        assert n in _abi.SIGNATURES or n in _abi.OPTIONAL_SIGNATURES, 'no ctypes binding for ' + n
    assert lib.mgnns_abi_version() == 1


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'mgnns_b200')):
        for fn in files:
            if fn.endswith('.py'):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn


def test_ops_reject_cpu_tensors():
    from mgnns_b200 import ops  # noqa: F401
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.mgnns.mm(torch.randn(2, 2), torch.randn(2, 2), None, False, False, 0, 0.0)


def test_gen_A_gen_adj_api_bit_exact_vs_reference(golden):
    from mgnns_b200.api.graph_util import gen_A, gen_adj
    z = golden('adj.npz')
    for kind, n, ts in (('object', 80, (0.3, 0.4, 0.6)), ('place', 365, (0.3, 0.5))):
        for t in ts:
            tag = '%s_t%02d' % (kind, int(t * 10))
            A, nums = gen_A(n, t, synth.adj_dict(kind))
            ref = np.zeros((n, n))
            ref[z[tag + '_A_rows'], z[tag + '_A_cols']] = z[tag + '_A_vals']
            assert np.array_equal(A, ref) and nums.shape == (n, 1)
            adj = gen_adj(torch.from_numpy(A).float()).numpy()
            refadj = np.zeros((n, n), dtype=np.float32)
            refadj[z[tag + '_adj_rows'], z[tag + '_adj_cols']] = z[tag + '_adj_vals']
            assert np.array_equal(adj, refadj)


def test_sparse_edge_map_lookup_and_roundtrip():
    from mgnns_b200.api.pmi import SparseEdgeMap
    rs = np.random.RandomState(0)
    dense = np.zeros((40, 40), dtype=np.int64)
    idx = rs.choice(1600, 200, replace=False)
    dense.flat[np.sort(idx)] = np.arange(1, 201)
    m = SparseEdgeMap.from_dense(dense)
    assert np.array_equal(m.toarray(), dense)
    for i, j in rs.randint(0, 40, (300, 2)):
        assert m[i, j] == dense[i, j]
    implicit = SparseEdgeMap(m.rowptr, m.col, 40)          # ids = 1 + CSR position
    assert np.array_equal(implicit.toarray(), dense)


def test_pmi_host_stage_matches_reference(golden):
    """Host float64 stage (pmi_from_counts) fed with oracle integer counts reproduces the reference's
    edge set, ids and weights on real text."""
    from mgnns_b200.api import pmi
    from oracle import pmi_oracle as PO
    z = golden('pmi_val400.npz')
    texts, vocab = list(z['texts']), list(z['vocab'])
    ids, pad_id = pmi.encode_corpus(texts, vocab)
    ids_o, pad_o = PO.encode(texts, vocab)
    assert np.array_equal(ids, ids_o) and pad_id == pad_o == 0
    pair, wc = PO.counts_numpy(ids, pad_id, len(vocab), int(z['window']))
    pair[pair < int(z['min_cooc'])] = 0
    rows, cols = np.nonzero(pair)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=len(vocab)))])
    keep, vals = pmi.pmi_from_counts(rowptr, cols, pair[rows, cols], wc)
    assert np.array_equal(rows[keep], z['rows']) and np.array_equal(cols[keep], z['cols'])
    np.testing.assert_allclose(vals[keep].astype(np.float32), z['weights'][1:, 0], rtol=1e-6)


def test_text_padding_drops_long_texts():
    from mgnns_b200.api import pmi
    texts = ['a b', ' '.join(['x'] * 99), ' '.join(['x'] * 100), 'a  b']
    padded = pmi.text_padding(texts)
    assert len(padded) == 3 and all(len(p) == 100 for p in padded)
    assert padded[2][:3] == ['a', '', 'b']


def _build_cpu_model(cfg):
    from mgnns_b200.api.multi_gcn import IdentityTrunk, Multi_GCN_Multihead_Att
    from mgnns_b200.api.text_gcn import Model as TextModel
    emap, count = synth.synthetic_edge_map(cfg['V'], seed=1, docs=200)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, cfg['V'])]
    tm = TextModel(7, 300, vocab, cfg['ngram'], 0.5, count, emap, pmi=torch.zeros(count, 1))
    opt = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5,
               emb_type='random', vocab_size=cfg['V'], stack_num=2, n_head=4, d_kv=128, is_regu=False)
    torch.manual_seed(0)
    return Multi_GCN_Multihead_Att(opt, 7, tm, IdentityTrunk(), IdentityTrunk(), 80, 365, object_t=0.4, place_t=0.3,
                                   in_channel=300, object_adj_file=synth.adj_dict('object'),
                                   place_adj_file=synth.adj_dict('place')), count


def test_state_dict_contract_and_reference_initialisation(golden):
    """Same parameter names/shapes as the reference and — because modules are constructed in the
    reference's order with its initialisers — identical values under torch.manual_seed(0)."""
    z = golden('model_init_seed0.npz')
    model, count = _build_cpu_model(H.MODEL_CFG)
    sd = model.state_dict()
    ref_names = set(z.files)
    mine = {n for n in sd if not n.startswith('text_features.')}
    assert mine == ref_names, (sorted(mine - ref_names)[:5], sorted(ref_names - mine)[:5])
    for n in sorted(ref_names):
        stat = z[n]
        assert tuple(int(v) for v in stat[2:]) == tuple(sd[n].shape), n
        np.testing.assert_allclose([sd[n].double().sum().item(), sd[n].double().abs().sum().item()], stat[:2],
                                   rtol=1e-9, atol=1e-9, err_msg=n)
    for n, shape in H.state_shapes(H.MODEL_CFG, count).items():
        assert tuple(sd[n].shape) == tuple(shape), n


def test_optimizer_groups_omit_the_never_stepped_parameters():
    model, _ = _build_cpu_model(H.MODEL_CFG)
    groups = model.get_config_optim(1e-4, 0.1)
    assert len(groups) == 12
    stepped = {id(p) for g in groups for p in g['params']}
    named = dict(model.named_parameters())
    for n in ('multi_linear_1.weight', 'multi_linear_2.weight', 'liner_img_object.weight', 'object_linear_5.weight',
              'place_x_linear.weight', 'embedding.weight', 'object_A'):
        assert id(named[n]) not in stepped, n
    for n in ('gc1.weight', 'lstm.weight_ih_l0', 'text_features.node_hidden.weight',
              'img_object_text_multi_head_att.0.slf_attn.w_qs.weight', 'object_attention.w_q.weight'):
        assert id(named[n]) in stepped, n


def test_lstm_plan_host_schedule_sorted_tiles_and_compact_indices():
    """Host-side schedule of the packed bi-LSTM (ref: pack_padded_sequence at model:376): compact rows, length-sorted
    tiles of 8, first/last token rows, fixed-capacity refresh for CUDA-graph replays."""
    from mgnns_b200 import ops
    lens = torch.tensor([3, 0, 7, 1, 100, 5, 2, 2, 9, 4], dtype=torch.int64)
    L = 100
    plan = ops.LstmPlan(lens, L, torch.device('cpu'), capacity=256)
    n = int(lens.sum())
    assert plan.N == n and plan.capacity == 256 and plan.n_tiles == 2
    off = plan.offsets.numpy()
    assert off[0] == 0 and off[-1] == n and np.array_equal(np.diff(off), lens.numpy())
    tiles = plan.tiles.numpy()
    order = tiles[tiles >= 0]
    assert sorted(order.tolist()) == [i for i in range(10) if lens[i] > 0]           # empty sequences are skipped
    assert np.all(np.diff(lens.numpy()[order]) <= 0)                                  # longest first
    tok = plan.tok_idx.numpy()[:n]
    rows = np.repeat(np.arange(10), lens.numpy())
    assert np.array_equal(tok // L, rows) and np.array_equal(tok % L, np.arange(n) - off[rows])
    assert np.all(plan.flat_idx.numpy()[n:] == 10 * L)                                # padding rows -> dummy bank row
    assert np.array_equal(plan.last_idx.numpy()[lens.numpy() > 0], (off[1:] - 1)[lens.numpy() > 0])
    plan.update_(torch.tensor([1] * 10, dtype=torch.int64))                           # refresh in place, same capacity
    assert plan.N == 10 and plan.capacity == 256 and int(plan.offsets[-1]) == 10
    with pytest.raises(RuntimeError):
        ops.LstmPlan(torch.tensor([100, 100, 100]), L, torch.device('cpu'), capacity=256)


def test_cfg2_word_graph_and_csr_transpose_on_host():
    """SURVEY §8d cfg-2 adjacency generator: self loops, sorted unique columns, rows normalised to 1; and the host CSR
    transpose used by the backward SpMM."""
    from mgnns_b200.api.graph_util import CSRAdjacency
    N = 300
    rowptr, cols, val = synth.cfg2_word_graph(N, mean_degree=10, seed=2)
    assert rowptr[0] == 0 and rowptr[-1] == cols.shape[0] == val.shape[0]
    for i in (0, 17, N - 1):
        c = cols[rowptr[i]:rowptr[i + 1]]
        assert i in c and np.all(np.diff(c) > 0)
        np.testing.assert_allclose(val[rowptr[i]:rowptr[i + 1]].sum(), 1.0, rtol=1e-5)
    csr = CSRAdjacency.from_scipy_like(rowptr, cols, val, N, torch.device('cpu'))
    dense = np.zeros((N, N), dtype=np.float32)
    dense[np.repeat(np.arange(N), np.diff(rowptr)), cols] = val
    t = np.zeros((N, N), dtype=np.float32)
    tr, tc, tv = csr.t_rowptr.numpy(), csr.t_col.numpy(), csr.t_val.numpy()
    t[np.repeat(np.arange(N), np.diff(tr)), tc] = tv
    assert np.array_equal(t, dense.T) and csr.nnz == cols.shape[0]


def test_launch_list_summariser(tmp_path, capsys):
    """scripts/summarize_launches.py: one step = the span between the last two anchor launches."""
    import importlib.util
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC",'
            '"Section Name","Metric Name","Metric Unit","Metric Value"']
    names = ['mgnns::text_maxagg_fwd_kernel(a)', 'void mgnns::gemm_ffma_kernel<1, 1>(p)', 'void at::native::foo<int>(x)'] * 3
    for i, n in enumerate(names):
        rows.append('"%d","1","python","h","%s","1","7","(128, 1, 1)","(10, 1, 1)","0","10.0","s","gpu__time_duration.sum","ns","%d"'
                    % (i, n, 1000 * (i + 1)))
    p = tmp_path / 'l.csv'
    p.write_text("==PROF== noise\n" + "\n".join(rows) + "\n")
    spec = importlib.util.spec_from_file_location('summ', os.path.join(ROOT, 'scripts', 'summarize_launches.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import sys
    argv, sys.argv = sys.argv, ['summarize_launches.py', str(p)]
    try:
        mod.main()
    finally:
        sys.argv = argv
    out = capsys.readouterr().out
    assert 'launches [3, 6)' in out and '3 launches' in out and 'mgnns::gemm_ffma_kernel' in out


def test_heavy_plan_parser_and_flat_gradient_alloc_hook():
    """Host logic of the captured step: MGNNS_HEAVY_PLAN entries (CTA cap / gate per deferred heavy job) and the
    FlatGradients allocation hook that lets the flat gradient buffer live in peer-mapped memory."""
    import torch
    from mgnns_b200 import ops
    from mgnns_b200.optim import FlatGradients
    assert ops._parse_heavy_plan("g0,g1") == [(0, 0), (0, 1)]
    assert ops._parse_heavy_plan("c100,g0") == [(100, -1), (0, 0)]
    assert ops._parse_heavy_plan("c74g1") == [(74, 1)]
    assert ops._parse_heavy_plan("off") is None
    with pytest.raises(ValueError):
        ops._parse_heavy_plan("x3")
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    for p in ps[:2]:
        p.grad = torch.randn_like(p)
    asked = []

    def alloc(n):
        asked.append(n)
        return torch.zeros(n)
    fg = FlatGradients(ps, alloc=alloc)
    assert asked == [16 + 8] and len(fg.params) == 2 and fg.offsets == [0, 16]      # 15 -> 16, 7 -> 8 floats; no-grad param left out
    want = [p.grad.clone() for p in ps[:2]]
    fg.pack()
    for p, w, v in zip(ps[:2], want, fg.views):
        assert torch.equal(p.grad, w) and p.grad.data_ptr() == v.data_ptr()
    with pytest.raises(RuntimeError):
        FlatGradients(ps, alloc=lambda n: torch.zeros(n + 4))
