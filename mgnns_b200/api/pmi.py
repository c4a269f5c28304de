"""PMI word-graph builder (ref: utils/pmi.py:28-105) with the integer co-occurrence count on the GPU.

Host work that stays on the host: reading the JSONL corpus, whitespace tokenisation, the
vocabulary lookup (strings), and the float64 PMI arithmetic over the *non-zero* cells only (done
with numpy in the reference's operation order so the pmi>0 edge set is identical).  The O(tokens x
window) counting loop and the three O(V^2) Python loops of the reference are replaced by the table-free
device count (ops.pmi_count: row buckets + shared-memory column counters, csrc/pmi_sparse.cu), which
emits the kept cells directly in the reference's row-major order.  `save_pmi` / `load_pmi` are the
on-disk CSR form (.npz) of what the reference keeps as a dense int [V,V] array.
"""
import json
import os

import numpy as np
import torch

from .. import ops
from ..edge_map import SparseEdgeMap  # noqa: F401  (numpy-only class; this module is its reference-facing home)

MAX_LEN = 100   # ref: utils/pmi.py:13-14 (texts are padded to 100 tokens, longer ones dropped)


def text_padding(content):
    """Whitespace split, drop texts with >= 100 tokens, pad with 'PAD' (ref: utils/pmi.py:8-16)."""
    out = []
    for text in content:
        sentence = text.split(' ')
        if len(sentence) < MAX_LEN:
            out.append(sentence + ['PAD'] * (MAX_LEN - len(sentence)))
    return out


def get_content(data_root_path):
    """(ref: utils/pmi.py:18-26)"""
    texts = []
    with open(os.path.join(data_root_path, 'all_anno_json', 'train_all_anno.json'), 'r') as f:
        for line in f:
            texts.append(json.loads(line)['text'])
    return texts


def encode_corpus(texts, vocab):
    """Texts -> (int32 [D,100] vocabulary indices with -1 for out-of-vocabulary, pad_id)."""
    index = dict(zip(vocab, range(len(vocab))))     # later duplicates win, like the reference's dict(zip(...))
    padded = text_padding(texts)
    ids = np.full((len(padded), MAX_LEN), -1, dtype=np.int32)
    for r, sentence in enumerate(padded):
        ids[r] = [index.get(w, -1) for w in sentence]
    return ids, index.get('PAD', -1)


def pmi_from_counts(rowptr, col, cnt, word_count):
    """Float64 PMI over the kept cells, reference operation order (ref: utils/pmi.py:69-87).

    Returns the mask of cells with pmi > 0 and their float64 values.
    """
    word_count = np.asarray(word_count, dtype=np.int64)
    total = np.sum(word_count)
    p_word = word_count / total
    p_pair = np.asarray(cnt, dtype=np.int64) / total
    rows = np.repeat(np.arange(rowptr.shape[0] - 1), np.diff(rowptr))
    denom = p_word[rows] * p_word[col]
    pmi = np.zeros(col.shape[0], dtype=np.float64)
    ok = (denom != 0) & (p_pair != 0)
    with np.errstate(divide='ignore', invalid='ignore'):
        pmi[ok] = np.log(p_pair[ok] / denom[ok])
    pmi = np.maximum(np.nan_to_num(pmi), 0.0)
    return pmi != 0, pmi


def cal_PMI_from_ids(ids, pad_id, vocab_size, window_size=6, min_cooccurence=2, device=None):
    """Core: encoded corpus -> (edges_weights, edges_mappings, count) like cal_PMI.

    ids: int32 [D, L] numpy array or CUDA tensor (vocabulary indices, -1 = OOV).
    """
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    tok = ids if isinstance(ids, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32))
    tok = tok.to(device=device, dtype=torch.int32)
    rowptr, col, cnt, wc = ops.pmi_count(tok, vocab_size, window_size, pad_id, min_cooccurence)
    rowptr = rowptr.cpu().numpy().astype(np.int64)
    col = col.cpu().numpy().astype(np.int64)
    cnt = cnt.cpu().numpy().astype(np.int64)
    wc = wc.cpu().numpy()
    keep, pmi = pmi_from_counts(rowptr, col, cnt, wc)
    rows = np.repeat(np.arange(vocab_size), np.diff(rowptr))[keep]
    e_rowptr = np.zeros(vocab_size + 1, dtype=np.int64)
    np.add.at(e_rowptr, rows + 1, 1)
    edge_map = SparseEdgeMap(np.cumsum(e_rowptr), col[keep], vocab_size)
    weights = np.concatenate([[0.0], pmi[keep]]).reshape(-1, 1)
    count = int(keep.sum()) + 1
    edge_map.pair_counts = (rowptr, col, cnt)      # kept for inspection / tests (counts >= min_cooccurence)
    edge_map.word_count = wc
    return torch.Tensor(weights), edge_map, count


def cal_PMI_from_texts(texts, vocab, window_size=6, min_cooccurence=2, device=None):
    ids, pad_id = encode_corpus(texts, vocab)
    return cal_PMI_from_ids(ids, pad_id, len(vocab), window_size, min_cooccurence, device)


def cal_PMI(data_root_path, vocab_root_path, min_count, phase='train', window_size=6, min_cooccurence=2):
    """Drop-in for utils.pmi.cal_PMI (ref: utils/pmi.py:28).

    Returns (edges_weights FloatTensor[count,1], edges_mappings, count).  `edges_mappings` is a
    SparseEdgeMap (indexable like the reference's dense [V,V] array) instead of a dense matrix.
    """
    from .vocab import get_vocab_list
    vocab = get_vocab_list(data_root_path, vocab_root_path, min_count)
    return cal_PMI_from_texts(get_content(data_root_path), vocab, window_size, min_cooccurence)


def save_pmi(path, edges_weights, edges_mappings, count):
    """Persist cal_PMI's result as one compressed .npz of CSR arrays: a few MB instead of the
    reference's dense int64 [V,V] `edges_mappings` (3.25 GB at V=20k, 20 GB at V=50k; ref: utils/pmi.py:89-105)."""
    w = edges_weights.detach().cpu().numpy() if torch.is_tensor(edges_weights) else np.asarray(edges_weights)
    if w.shape[0] != count:
        raise ValueError("save_pmi: %d weights for count=%d" % (w.shape[0], count))
    edges_mappings.save(path, weights=w)


def load_pmi(path):
    """-> (edges_weights FloatTensor[count,1], edges_mappings SparseEdgeMap, count), as cal_PMI returns them."""
    m, w = SparseEdgeMap.load(path)
    if w is None:
        raise ValueError("load_pmi: %s holds no edge weights" % path)
    return torch.Tensor(w), m, int(w.shape[0])
