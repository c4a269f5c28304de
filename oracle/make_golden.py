"""Generate golden vectors by running the REFERENCE's own code (read-only at /root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the authoring container only:  python -m oracle.make_golden
Outputs (committed):
  mgnns_b200/data/label_graphs.npz  label-graph fixtures re-packed from data/adj/*.pkl, data/glove/*.pkl
  tests/golden/pmi_*.npz            utils/pmi.py cal_PMI outputs
  tests/golden/adj.npz              utils/util.py gen_A / gen_adj outputs
  tests/golden/modules.npz          models/submodules.py, models/moudles.py, GraphConvolution, Attention
  tests/golden/model.npz            Multi_GCN_Multihead_Att forward + gradients (B=6, V=300)

Nothing is copied from the reference: its modules are imported (or exec'd from where they lie with
the literal 'cuda:0' / '.cuda()' redirected to the CPU) behind these shims:
  * models.multi_head_att.submodules -> models.submodules      (moved file, moudles.py:4-5)
  * np.int = int; gen_A's missing 4th argument defaults to 0.2 (util.py:382 vs model:338)
  * word2vec: stub whose load() returns zero vectors (GloVe file absent; weights are overwritten)
  * dgl: ~60-line stand-in implementing exactly the calls models/Text_GCN.py makes.  It encodes
    DGL's documented semantics (builtin max reducer, zero fill for nodes without in-edges), which
    is why the text channel stays "parity unpinned" at that boundary.
"""
import importlib
import importlib.util
import io
import json
import os
import pickle
import sys
import tempfile
import types
from contextlib import redirect_stdout

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

from mgnns_b200 import synth  # noqa: E402  (deterministic inputs shared with the tests)


# ----------------------------------------------------------------------------- shims
def _mini_dgl():
    dgl = types.ModuleType('dgl')

    class Graph:
        def __init__(self):
            self.n = 0
            self.src, self.dst = [], []
            self.ndata, self.edata = {}, {}
            self.sizes = None

        def to(self, _device):
            return self

        def add_nodes(self, n):
            self.n += int(n)

        def add_edges(self, srcs, dsts):
            self.src += list(srcs)
            self.dst += list(dsts)

        def update_all(self, message_func, reduce_func):
            src = torch.as_tensor(self.src)
            dst = torch.as_tensor(self.dst)
            msg = message_func(self.ndata, self.edata, src)
            name, out = reduce_func
            red = torch.zeros(self.n, msg.shape[1], dtype=msg.dtype)       # zero fill: no in-edge -> 0
            red = red.scatter_reduce(0, dst.unsqueeze(1).expand_as(msg), msg, reduce='amax', include_self=False)
            self.ndata[out] = red

    def batch(graphs):
        g = Graph()
        off = 0
        for s in graphs:
            g.src += [a + off for a in s.src]
            g.dst += [a + off for a in s.dst]
            off += s.n
        g.n = off
        g.sizes = [s.n for s in graphs]
        g.ndata = {k: torch.cat([s.ndata[k] for s in graphs]) for k in graphs[0].ndata}
        g.edata = {k: torch.cat([s.edata[k] for s in graphs]) for k in graphs[0].edata}
        return g

    def sum_nodes(g, feat):
        return torch.stack([c.sum(0) for c in torch.split(g.ndata[feat], g.sizes)])

    fn = types.ModuleType('dgl.function')
    fn.src_mul_edge = lambda h, w, out: (lambda nd, ed, src: nd[h][src] * ed[w])
    fn.max = lambda msg, out: (msg, out)
    dgl.DGLGraph = Graph
    dgl.batch = batch
    dgl.sum_nodes = sum_nodes
    dgl.function = fn
    return dgl, fn


def install_shims():
    np.int = int
    if REF not in sys.path:
        sys.path.insert(0, REF)
    dgl, fn = _mini_dgl()
    sys.modules['dgl'] = dgl
    sys.modules['dgl.function'] = fn
    w2v = types.ModuleType('word2vec')

    class _Vec(dict):
        def __getitem__(self, k):
            return np.zeros(300, dtype=np.float32)
    w2v.load = lambda path: _Vec()
    sys.modules['word2vec'] = w2v
    import models.submodules as sub
    pkg = types.ModuleType('models.multi_head_att')
    pkg.submodules = sub
    sys.modules['models.multi_head_att'] = pkg
    sys.modules['models.multi_head_att.submodules'] = sub
    import utils.util as ref_util
    if not getattr(ref_util, '_patched', False):
        orig = ref_util.gen_A
        ref_util.gen_A = lambda n, t, f, gama=0.2: orig(n, t, f, gama)
        ref_util._patched = True


def exec_reference_module(name, relpath, replacements):
    """Execute a reference source file from where it lies, with device literals redirected."""
    with open(os.path.join(REF, relpath)) as f:
        src = f.read()
    for a, b in replacements:
        src = src.replace(a, b)
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, relpath)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    return mod


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ----------------------------------------------------------------------------- fixtures re-pack
def pack_label_graphs():
    def load(p):
        with open(os.path.join(REF, p), 'rb') as f:
            return pickle.load(f)
    obj, plc = load('data/adj/tumblr_objects_adj.pkl'), load('data/adj/tumblr_resnet50_places_adj.pkl')
    out = dict(
        object_adj=np.asarray(obj['adj']).astype(np.int32), object_nums=np.asarray(obj['nums']).astype(np.float64),
        place_adj=np.asarray(plc['adj']).astype(np.int32), place_nums=np.asarray(plc['nums']).astype(np.float64),
        object_glove=np.asarray(load('data/glove/object_glove_word2vec.pkl'), dtype=np.float32),
        place_glove=np.asarray(load('data/glove/place_glove_word2vec.pkl'), dtype=np.float32),
        label_glove=np.asarray(load('data/tumblr_label_glove.pkl'), dtype=np.float64),
    )
    assert np.array_equal(out['object_adj'], np.asarray(obj['adj'])) and np.array_equal(out['place_adj'], plc['adj'])
    os.makedirs(os.path.join(ROOT, 'mgnns_b200', 'data'), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, 'mgnns_b200', 'data', 'label_graphs.npz'), **out)


# ----------------------------------------------------------------------------- PMI
KAT_VOCAB = ['PAD', 'UNK', 'a', 'b', 'c', 'd', 'e']
KAT_DOCS = ["a b c a d", "b c d e", "a a b zzz c", "e d c b a b c"]


def write_corpus(root, texts, vocab, min_count=5):
    os.makedirs(os.path.join(root, 'all_anno_json'), exist_ok=True)
    os.makedirs(os.path.join(root, 'vocab'), exist_ok=True)
    with open(os.path.join(root, 'all_anno_json', 'train_all_anno.json'), 'w') as f:
        for t in texts:
            f.write(json.dumps({'text': t}) + '\n')
    with open(os.path.join(root, 'vocab', 'vocab-%d.txt' % min_count), 'w') as f:
        f.write('\n'.join(vocab))


def ref_cal_pmi(texts, vocab, window, min_cooc):
    import utils.pmi as ref_pmi
    with tempfile.TemporaryDirectory() as d:
        write_corpus(d, texts, vocab)
        w, m, c = quiet(ref_pmi.cal_PMI, d, d, 5, 'train', window, min_cooc)
    return w.numpy(), np.asarray(m), int(c)


def synth_corpus(n_docs, V, seed):
    ids, lens, _ = synth.make_texts(n_docs, V, 100, seed=seed)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, V)]
    texts = [' '.join(vocab[t] for t in row[:n]) for row, n in zip(ids.tolist(), lens.tolist())]
    return texts, vocab


def golden_pmi():
    out = {}
    for tag, window, mc in (('w2m1', 2, 1), ('w2m2', 2, 2)):
        w, m, c = ref_cal_pmi(KAT_DOCS, KAT_VOCAB, window, mc)
        out['kat_%s_weights' % tag], out['kat_%s_map' % tag], out['kat_%s_count' % tag] = w, m, c
    np.savez_compressed(os.path.join(GOLD, 'pmi_kat.npz'), **out)
    # real text: first 400 val posts, vocabulary = first 300 words of vocab-5 (keeps the V^2 loops short)
    with open(os.path.join(REF, 'data/all_anno_json/val_all_anno.json')) as f:
        texts = [json.loads(l)['text'] for l in list(f)[:400]]
    with open(os.path.join(REF, 'data/vocab/vocab-5.txt')) as f:
        vocab = f.read().split('\n')[:300]
    # make it interesting: add the 200 most frequent words of these texts
    from collections import Counter
    cnt = Counter(w for t in texts for w in t.split(' '))
    for wd, _ in cnt.most_common(260):
        if wd not in vocab:
            vocab.append(wd)
    w, m, c = ref_cal_pmi(texts, vocab, 5, 2)
    rows, cols = np.nonzero(m)
    np.savez_compressed(os.path.join(GOLD, 'pmi_val400.npz'), texts=np.array(texts, dtype=object),
                        vocab=np.array(vocab, dtype=object), weights=w, rows=rows, cols=cols, ids=m[rows, cols],
                        count=c, window=5, min_cooc=2)
    # synthetic corpus used by the model golden (V=300)
    texts, vocab = synth_corpus(600, 300, seed=5)
    w, m, c = ref_cal_pmi(texts, vocab, 6, 2)
    rows, cols = np.nonzero(m)
    np.savez_compressed(os.path.join(GOLD, 'pmi_synth300.npz'), weights=w, rows=rows, cols=cols, ids=m[rows, cols],
                        count=c, window=6, min_cooc=2, n_docs=600, V=300, seed=5)
    return m, c


# ----------------------------------------------------------------------------- adjacency
def golden_adj():
    import utils.util as ref_util
    out = {}
    for kind, n, path, ts in (('object', 80, 'data/adj/tumblr_objects_adj.pkl', (0.3, 0.4, 0.6)),
                              ('place', 365, 'data/adj/tumblr_resnet50_places_adj.pkl', (0.3, 0.5))):
        for t in ts:
            A, _ = quiet(ref_util.gen_A, n, t, os.path.join(REF, path))
            adj = ref_util.gen_adj(torch.from_numpy(A).float())
            r, c = np.nonzero(adj.numpy())
            tag = '%s_t%02d' % (kind, int(t * 10))
            out[tag + '_A_rows'], out[tag + '_A_cols'] = np.nonzero(A)
            out[tag + '_A_vals'] = A[np.nonzero(A)]
            out[tag + '_adj_rows'], out[tag + '_adj_cols'], out[tag + '_adj_vals'] = r, c, adj.numpy()[r, c]
    np.savez_compressed(os.path.join(GOLD, 'adj.npz'), **out)


# ----------------------------------------------------------------------------- modules
def golden_modules(model_mod):
    import models.moudles as ref_moudles
    import models.submodules as ref_sub
    out = {}
    g = torch.Generator().manual_seed(7)
    B, L, d = 5, 9, 300
    # LayerNorm
    ln = ref_sub.LayerNorm(d)
    synth.fill_parameters(ln, seed=1)
    x = torch.randn(B, 1, d, generator=g)
    out['ln_x'], out['ln_y'] = x.numpy(), ln(x).detach().numpy()
    # MyMultiHeadAttention (slf_attn + pos_ffn), masked and unmasked, eval mode
    layer = quiet(ref_moudles.MyMultiHeadAttention, 4, d, 128, dropout=0.5, need_mask=False)
    synth.fill_parameters(layer, seed=2)
    layer.eval()
    q = torch.randn(B, d, generator=g)
    bank = torch.randn(B, L, d, generator=g)
    mask = (torch.arange(L).unsqueeze(0) < torch.tensor([9, 1, 4, 7, 2]).unsqueeze(1)).float()
    y_m, a_m = layer(q, bank, bank, mask)
    y_u, a_u = layer(q, bank, bank, None)
    out.update(mha_q=q.numpy(), mha_bank=bank.numpy(), mha_mask=mask.numpy(), mha_out_masked=y_m.detach().numpy(),
               mha_attn_masked=a_m.detach().numpy(), mha_out_unmasked=y_u.detach().numpy(),
               mha_attn_unmasked=a_u.detach().numpy())
    # gradients of sum(y_m * r) w.r.t. q, bank and every parameter (norms + full small ones)
    r = torch.randn(B, d, generator=g)
    q2, bank2 = q.clone().requires_grad_(), bank.clone().requires_grad_()
    y, _ = layer(q2, bank2, bank2, mask)
    (y * r).sum().backward()
    out.update(mha_r=r.numpy(), mha_gq=q2.grad.numpy(), mha_gbank=bank2.grad.numpy())
    for n, p in layer.named_parameters():
        out['mha_g_' + n] = p.grad.numpy() if p.numel() <= 512 else np.array(p.grad.norm().item())
    # GraphConvolution on the real object label graph
    z = synth.label_graphs()
    import utils.util as ref_util
    A, _ = quiet(ref_util.gen_A, 80, 0.4, os.path.join(REF, 'data/adj/tumblr_objects_adj.pkl'))
    adj = ref_util.gen_adj(torch.from_numpy(A).float())
    gc = model_mod.GraphConvolution(300, 64)
    synth.fill_parameters(gc, seed=3)
    inp = torch.from_numpy(z['object_glove']).float()
    out['gc_out'] = gc(inp, adj).detach().numpy()
    gcb = model_mod.GraphConvolution(300, 32, bias=True)
    synth.fill_parameters(gcb, seed=4)
    xb = torch.randn(3, 80, 300, generator=g)
    out['gcb_x'], out['gcb_out'] = xb.numpy(), gcb(xb, adj).detach().numpy()
    # label Attention (7 labels, 5 heads), eval mode
    att = model_mod.Attention(hid_dim=300, image_dim=80, n_heads=5, dropout=0.5)
    synth.fill_parameters(att, seed=5)
    att.eval()
    key = torch.randn(B, 80, generator=g)
    query = torch.from_numpy(z['label_glove'])
    out['latt_key'], out['latt_out'] = key.numpy(), att(query, key, key).detach().numpy()
    np.savez_compressed(os.path.join(GOLD, 'modules.npz'), **out)


# ----------------------------------------------------------------------------- whole model
MODEL_CFG = dict(B=6, V=300, L=100, ngram=4, n_head=4, d_kv=128, stack_num=2, hidden_size=150, num_layers=2,
                 object_t=0.4, place_t=0.3, num_labels=7, seed=11)


def model_inputs(cfg=MODEL_CFG):
    B, V, L, seed = cfg['B'], cfg['V'], cfg['L'], cfg['seed']
    text, lens, mask = synth.make_texts(B, V, L, seed=seed)
    fo, fp = synth.make_fmaps(B, seed=seed), synth.make_fmaps(B, seed=seed + 1)
    oinp, pinp = synth.label_inputs(B)
    labels = synth.make_labels(B, cfg['num_labels'], seed=seed)
    return text, lens, mask, fo, fp, oinp, pinp, labels


def golden_model(model_mod, text_mod, edge_matrix, edge_count):
    cfg = MODEL_CFG
    opt = dict(emb_path='', bidirectional=True, hidden_size=cfg['hidden_size'], emb_size=300,
               num_layers=cfg['num_layers'], dropout=0.5, emb_type='random', vocab_size=cfg['V'],
               stack_num=cfg['stack_num'], n_head=cfg['n_head'], d_kv=cfg['d_kv'], is_regu=False)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, cfg['V'])]
    text_model = quiet(text_mod.Model, 7, 300, vocab, cfg['ngram'], 0.5, edge_count, edge_matrix,
                       pmi=torch.zeros(edge_count, 1), cuda=False)

    class Trunk(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for n in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'):
                setattr(self, n, torch.nn.Identity())
    model = quiet(model_mod.Multi_GCN_Multihead_Att, opt, 7, text_model, Trunk(), Trunk(), 80, 365,
                  object_t=cfg['object_t'], place_t=cfg['place_t'], in_channel=300,
                  object_adj_file=os.path.join(REF, 'data/adj/tumblr_objects_adj.pkl'),
                  place_adj_file=os.path.join(REF, 'data/adj/tumblr_resnet50_places_adj.pkl'))
    synth.fill_parameters(model, seed=cfg['seed'])
    model.eval()
    text, lens, mask, fo, fp, oinp, pinp, labels = model_inputs(cfg)
    logits = model(text, lens, mask, fo, fp, oinp, pinp)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    out = dict(logits=logits.detach().numpy(), loss=np.array(loss.item()), labels=labels.numpy())
    names, norms = [], []
    for n, p in model.named_parameters():
        if p.grad is not None:
            names.append(n)
            norms.append(p.grad.norm().item())
    out['grad_names'] = np.array(names, dtype=object)
    out['grad_norms'] = np.array(norms)
    for n in ('multi_linear_2.weight', 'multi_linear_2.bias', 'object_attention.w_q.bias',
              'img_object_text_multi_head_att.1.slf_attn.layer_norm.gamma', 'place_x_linear.bias'):
        out['grad::' + n] = dict(model.named_parameters())[n].grad.numpy()
    out['grad::text_features.seq_edge_w.weight'] = text_model.seq_edge_w.weight.grad.numpy()
    gh = text_model.node_hidden.weight.grad
    out['grad::node_hidden_rowsum'] = gh.sum(1).numpy()
    out['text_feature'] = model.text_features(text).detach().numpy()
    np.savez_compressed(os.path.join(GOLD, 'model.npz'), **out)
    # reference default initialisation under torch.manual_seed(0): parameter statistics, used to check
    # that the mirror modules construct their parameters in the same order / with the same inits
    torch.manual_seed(0)
    m2 = quiet(model_mod.Multi_GCN_Multihead_Att, dict(opt), 7, text_model, Trunk(), Trunk(), 80, 365,
               object_t=cfg['object_t'], place_t=cfg['place_t'], in_channel=300,
               object_adj_file=os.path.join(REF, 'data/adj/tumblr_objects_adj.pkl'),
               place_adj_file=os.path.join(REF, 'data/adj/tumblr_resnet50_places_adj.pkl'))
    init = {}
    for n, p in m2.state_dict().items():
        if n.startswith('text_features.'):
            continue
        init[n] = np.array([p.double().sum().item(), p.double().abs().sum().item()] + list(p.shape), dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, 'model_init_seed0.npz'), **init)


def main():
    os.makedirs(GOLD, exist_ok=True)
    install_shims()
    pack_label_graphs()
    # the model file reads data/glove/tumblr_label_glove.pkl relative to the cwd at import time (model:20-27)
    work = tempfile.mkdtemp()
    os.makedirs(os.path.join(work, 'data', 'glove'))
    os.symlink(os.path.join(REF, 'data', 'tumblr_label_glove.pkl'),
               os.path.join(work, 'data', 'glove', 'tumblr_label_glove.pkl'))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        text_mod = exec_reference_module('models.Text_GCN', 'models/Text_GCN.py',
                                         [(".to('cuda:0')", ".to('cpu')"), ('.cuda()', '.cpu()')])
        model_mod = exec_reference_module('models.Multi_GCN_Multihead_att', 'models/Multi_GCN_Multihead_att.py',
                                          [("'cuda:0'", "'cpu'")])
    finally:
        os.chdir(cwd)
    edge_matrix, edge_count = golden_pmi()
    golden_adj()
    golden_modules(model_mod)
    golden_model(model_mod, text_mod, edge_matrix, edge_count)
    print('golden vectors written to', GOLD)


if __name__ == '__main__':
    main()
