"""GPU tests of the loop-side pieces (b, f4): the per-batch engine protocol driving the FULL model
(real torchvision ResNet-101 / ResNet-50 trunks, images in) through the reference's hook sequence and `state` keys;
device-side confusion counts and label co-occurrence counts, bit-exact against plain Python restatements of
engine/Multi_GCN_Multihead_Att_engine.py:829-838 and utils/util.py:336-357."""
import numpy as np
import pytest
import torch

import mgnns_test_helpers as H
from mgnns_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device('cuda', 0)


def test_confusion_count_bit_exact_and_metrics_match_sklearn(dev):
    from sklearn.metrics import accuracy_score, f1_score
    from mgnns_b200 import ops
    from mgnns_b200.api.engine import metrics_from_confusion
    g = torch.Generator().manual_seed(3)
    for B, C in ((64, 7), (1024, 3), (5, 7), (300, 40)):
        scores = torch.randn(B, C, generator=g)
        scores[0, :] = 0.25                                   # a full tie: first index wins
        if B > 4:
            scores[3, 1] = float('nan')                       # NaN never wins
        target = torch.randint(0, C, (B,), generator=g)
        conf = torch.zeros(C, C, dtype=torch.int32, device=dev)
        pred = torch.empty(B, dtype=torch.int64, device=dev)
        ops.confusion_count(scores.to(dev), target.to(dev), conf, pred)
        ops.confusion_count(scores.to(dev), target.to(dev), conf, None)     # accumulates
        ref_pred = torch.where(torch.isnan(scores), torch.full_like(scores, -float('inf')), scores).argmax(1)
        assert ref_pred[0] == 0
        assert torch.equal(pred.cpu(), ref_pred)
        ref = np.zeros((C, C), dtype=np.int64)
        np.add.at(ref, (target.numpy(), ref_pred.numpy()), 2)
        assert np.array_equal(conf.cpu().numpy(), ref)
        got = metrics_from_confusion(ref // 2)
        t, p = target.numpy(), ref_pred.numpy()
        want = (accuracy_score(t, p), f1_score(t, p, average='micro'), f1_score(t, p, average='macro', zero_division=0),
                f1_score(t, p, average='weighted', zero_division=0))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_label_cooccurrence_matches_reference_loops(dev):
    from mgnns_b200.api import engine as E
    rs = np.random.RandomState(0)
    for C, n_img in ((80, 500), (365, 300), (4096, 2000)):
        objects = []
        for _ in range(n_img):
            k = rs.randint(0, 9)
            objects.append(list(set(rs.randint(0, C, size=k).tolist())))
        objects[3] = []
        nums = np.zeros(C)
        adj = np.zeros((C, C))
        for obj in objects:                                   # utils/util.py:336-357, restated
            for j in obj:
                nums[j] += 1
            for a in obj:
                for b in obj:
                    if a != b:
                        adj[a][b] += 1
        g_nums, g_adj = E.label_counts(objects, C, device=dev)
        assert np.array_equal(g_nums, nums) and np.array_equal(g_adj, adj)
        assert np.array_equal(E.generate_Adj(objects, C), adj) and np.array_equal(E.generate_nums(objects, C), nums)
    all_nums, all_adj = E.get_Adj_from_lists([objects[:1000], objects[1000:]], C)
    assert np.array_equal(all_adj, adj) and (all_nums >= 1).all() and np.array_equal(all_nums[nums > 0], nums[nums > 0])
    # the counted adjacency feeds gen_A / gen_adj exactly like the shipped pickles
    from mgnns_b200.api.graph_util import gen_A
    A, _ = gen_A(C, 0.05, {'adj': all_adj, 'nums': all_nums})
    assert A.shape == (C, C) and np.isfinite(A).all()


class _TumblrShapedLoader:
    """Batches shaped like Tumblr_Dataset.__getitem__ collated by a DataLoader (ref: utils/Multi_GCN_Co_att_dataset.py
    :230-266; engine:853-865 documents the tuple): (id, text, content ids, len, mask, image, image path, object_inp,
    place_inp), target."""

    def __init__(self, n_batches, B, V, num_labels, seed):
        self.batches = []
        for i in range(n_batches):
            text, lens, mask = synth.make_texts(B, V, 100, seed=seed + i)
            g = torch.Generator().manual_seed(seed + 50 + i)
            img = torch.randn(B, 3, 448, 448, generator=g)
            oinp, pinp = synth.label_inputs(B)
            ids = ['post_%d_%d' % (i, j) for j in range(B)]
            inp = (ids, ['text'] * B, text, lens, mask, img, ['img.jpg'] * B, oinp.contiguous(), pinp.contiguous())
            self.batches.append((inp, synth.make_labels(B, num_labels, seed=seed + i)))

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def test_engine_protocol_full_model_with_real_resnet_trunks(dev):
    """16 synthetic posts (2 batches of 8) with 448x448 images through train() and validate(): the model is the
    reference-shaped factory output with REAL ResNet-101 (object) and ResNet-50/365 (scene) trunks, so the whole
    path — images -> trunks (cuDNN) -> hand-written head kernels -> loss -> backward -> clip -> Adam — runs the way
    engine:792-851 drives it.  Checks the state keys against sklearn on the gathered predictions, that only the
    optimizer's parameter groups move, and that full model == head fed with the trunks' own feature maps."""
    import torchvision.models as models
    from sklearn.metrics import accuracy_score, f1_score
    from mgnns_b200.api.engine import GCNMultiClassEngine
    from mgnns_b200.api.multi_gcn import Multi_GCN_Multihead_Att
    from mgnns_b200.api.text_gcn import Model as TextModel
    V, B = 400, 8
    torch.manual_seed(0)
    emap, count = synth.synthetic_edge_map(V, seed=5, docs=500)
    vocab = ['PAD', 'UNK'] + ['w%d' % i for i in range(2, V)]
    tm = TextModel(7, 300, vocab, 4, 0.5, count, emap, pmi=torch.zeros(count, 1))
    opt = dict(emb_path='', bidirectional=True, hidden_size=150, emb_size=300, num_layers=2, dropout=0.5,
               emb_type='random', vocab_size=V, stack_num=2, n_head=4, d_kv=128, is_regu=False)
    model = Multi_GCN_Multihead_Att(opt, 7, tm, models.resnet101(weights=None), models.resnet50(weights=None, num_classes=365),
                                    80, 365, object_t=0.4, place_t=0.3, in_channel=300,
                                    object_adj_file=synth.adj_dict('object'), place_adj_file=synth.adj_dict('place'))
    model = model.to(dev)
    optimizer = torch.optim.Adam(model.get_config_optim(5e-5, 0.1), lr=5e-5, weight_decay=1e-5)
    criterion = torch.nn.CrossEntropyLoss()
    loader = _TumblrShapedLoader(2, B, V, 7, seed=9)
    eng = GCNMultiClassEngine({'use_gpu': True, 'print_freq': 0})
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    out = eng.train(loader, model, criterion, optimizer, epoch=0)
    loss, acc, micro, macro, weighted, ids, targets, preds = out
    assert np.isfinite(loss) and len(ids) == len(targets) == len(preds) == 2 * B
    assert ids[0] == 'post_0_0' and ids[-1] == 'post_1_%d' % (B - 1)
    st = eng.state
    for key in ('batch_acc_list', 'batch_micro_f1_list', 'batch_macro_f1_list', 'batch_weighted_f1_list'):
        assert len(st[key]) == 2
    for i in range(2):
        t, p = targets[i * B:(i + 1) * B], preds[i * B:(i + 1) * B]
        assert st['batch_acc_list'][i] == accuracy_score(t, p)
        np.testing.assert_allclose(st['batch_macro_f1_list'][i], f1_score(t, p, average='macro', zero_division=0), atol=1e-12)
        np.testing.assert_allclose(st['batch_weighted_f1_list'][i], f1_score(t, p, average='weighted', zero_division=0), atol=1e-12)
    assert st['epoch_acc'] == sum(st['batch_acc_list']) / len(loader) == acc
    assert [int(x) for x in targets] == torch.cat([b[1] for b in loader.batches]).tolist()
    moved = {n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])}
    assert any(n.startswith('object_features.') for n in moved) and any(n.startswith('gc1.') for n in moved)
    assert not any(n.startswith(('multi_linear_', 'liner_img_', 'embedding')) for n in moved)    # never stepped
    vloss, vacc, vmicro, vmacro, vweighted = eng.validate(loader, model, criterion)
    assert np.isfinite(vloss) and 0.0 <= vacc <= 1.0 and vmicro == vacc
    # full model == head on the trunks' own outputs (same parameters, eval mode)
    model.eval()
    inp, _ = loader.batches[0]
    with torch.no_grad():
        img = inp[5].to(dev)
        full = model(inp[2].to(dev), inp[3], inp[4].to(dev), img, img, inp[7].to(dev), inp[8].to(dev))
        fo, fp = model.object_features(img), model.place_features(img)
        assert fo.shape == (B, 2048, 14, 14) and fp.shape == (B, 2048, 14, 14)
        trunks = (model.object_features, model.place_features)
        model.object_features, model.place_features = torch.nn.Identity(), torch.nn.Identity()
        try:
            head = model(inp[2].to(dev), inp[3], inp[4].to(dev), fo, fp, inp[7].to(dev), inp[8].to(dev))
        finally:
            model.object_features, model.place_features = trunks
    assert torch.allclose(full, head, rtol=0, atol=1e-6)
