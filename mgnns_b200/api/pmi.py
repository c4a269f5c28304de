"""PMI word-graph builder (ref: utils/pmi.py:28-105) with the integer co-occurrence count on the GPU.

Host work that stays on the host: reading the JSONL corpus, whitespace tokenisation, the
vocabulary lookup (strings), and the float64 PMI arithmetic over the *non-zero* cells only (done
with numpy in the reference's operation order so the pmi>0 edge set is identical).  The O(tokens x
window) counting loop and the three O(V^2) Python loops of the reference are replaced by
mgnns_pmi_count + an ordered CSR compaction on the device.
"""
import json
import os

import numpy as np
import torch

from .. import ops

MAX_LEN = 100   # ref: utils/pmi.py:13-14 (texts are padded to 100 tokens, longer ones dropped)


class SparseEdgeMap:
    """CSR stand-in for the reference's dense int `edges_mappings[V,V]` (3.25 GB at V=20k).

    Supports the two things the reference does with the matrix: `m[i, j]` lookups
    (models/Text_GCN.py:160,:164) and `.shape`; `toarray()` densifies for small V.
    Edge ids are 1 + CSR position (row-major enumeration, ref: utils/pmi.py:92-97); 0 = no edge.
    """

    def __init__(self, rowptr, col, n, eid=None):
        self.rowptr = np.asarray(rowptr, dtype=np.int64)
        self.col = np.asarray(col, dtype=np.int64)
        self.eid = None if eid is None else np.asarray(eid, dtype=np.int64)
        self.shape = (n, n)

    @property
    def nnz(self):
        return int(self.col.shape[0])

    def __getitem__(self, ij):
        i, j = int(ij[0]), int(ij[1])
        lo, hi = self.rowptr[i], self.rowptr[i + 1]
        k = lo + np.searchsorted(self.col[lo:hi], j)
        if k < hi and self.col[k] == j:
            return int(k + 1) if self.eid is None else int(self.eid[k])
        return 0

    def toarray(self):
        out = np.zeros(self.shape, dtype=np.int64)
        rows = np.repeat(np.arange(self.shape[0]), np.diff(self.rowptr))
        out[rows, self.col] = (np.arange(self.nnz) + 1) if self.eid is None else self.eid
        return out

    @classmethod
    def from_dense(cls, m):
        m = np.asarray(m)
        rows, cols = np.nonzero(m)            # row-major order
        rowptr = np.zeros(m.shape[0] + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        return cls(np.cumsum(rowptr), cols, m.shape[0], eid=m[rows, cols])


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
