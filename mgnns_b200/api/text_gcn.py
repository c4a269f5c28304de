"""TextLevelGCN text channel with the reference's constructor/forward (ref: models/Text_GCN.py:36-275).

The reference builds one DGL graph per document in Python every forward (seq_to_graph,
Text_GCN.py:168-211, after a device->host copy of the ids, :232).  Here the PMI edge-id map lives on
the device in CSR form and mgnns::text_maxagg does windowing, edge-id lookup, the max aggregation
and the sum readout in one kernel per batch, with no host round trip and no graph object.
"""
import os
import warnings

import numpy as np
import torch

from .pmi import SparseEdgeMap


class Model(torch.nn.Module):
    def __init__(self, class_num, hidden_size_node, vocab, n_gram, drop_out, edges_num, edges_matrix,
                 max_length=100, trainable_edges=True, pmi=None, cuda=True, is_padding=True):
        super().__init__()
        self.is_cuda = cuda
        self.is_padding = is_padding
        self.vocab = vocab
        self.node_hidden = torch.nn.Embedding(len(vocab), hidden_size_node)
        self.edges_num = edges_num
        if trainable_edges:
            # ref: Text_GCN.py:67-69 — trainable scalar per PMI edge, initialised to ONE (PMI values unused)
            self.seq_edge_w = torch.nn.Embedding.from_pretrained(torch.ones(edges_num, 1), freeze=False)
        else:
            self.seq_edge_w = torch.nn.Embedding.from_pretrained(pmi, freeze=False)
        self.hidden_size_node = hidden_size_node
        glove = self.load_word2vec('glove/glove.6B.300d.txt')
        if glove is not None:
            self.node_hidden.weight.data.copy_(torch.tensor(glove))
        else:
            # the GloVe text file is not shipped (ref: Text_GCN.py:76); BASELINE configs use random GloVe-300
            torch.nn.init.normal_(self.node_hidden.weight, mean=0.0, std=0.4)
        self.node_hidden.weight.requires_grad = True
        self.len_vocab = len(vocab)
        self.ngram = n_gram
        self.d = dict(zip(self.vocab, range(len(self.vocab))))
        self.max_length = max_length
        self.edges_matrix = edges_matrix
        self.dropout = torch.nn.Dropout(p=drop_out)
        self.activation = torch.nn.ReLU()
        self.Linear = torch.nn.Linear(hidden_size_node, class_num, bias=True)   # ref: :95, never used in forward

        emap = edges_matrix if isinstance(edges_matrix, SparseEdgeMap) else SparseEdgeMap.from_dense(edges_matrix)
        if emap.shape[0] != len(vocab):
            raise ValueError("edges_matrix is %s but the vocabulary has %d words" % (emap.shape, len(vocab)))
        self.register_buffer('pmi_rowptr', torch.from_numpy(emap.rowptr.astype(np.int32)), persistent=False)
        self.register_buffer('pmi_col', torch.from_numpy(emap.col.astype(np.int32)), persistent=False)
        if emap.eid is None:
            self.pmi_eid = None
        else:
            self.register_buffer('pmi_eid', torch.from_numpy(emap.eid.astype(np.int32)), persistent=False)

    def word2id(self, word):
        return self.d.get(word, self.d.get('UNK'))

    def load_word2vec(self, word2vec_file):
        """GloVe initialisation (ref: Text_GCN.py:105-121); returns None when the file or the `word2vec`
        package is unavailable (random init is used instead)."""
        if not os.path.exists(word2vec_file):
            return None
        try:
            import word2vec
            model = word2vec.load(word2vec_file)
        except Exception as exc:      # pragma: no cover - optional dependency
            warnings.warn("GloVe file present but word2vec could not load it: %s" % exc)
            return None
        rows = []
        for word in self.vocab:
            try:
                rows.append(model[word])
            except KeyError:
                rows.append(model['the'])
        return np.array(rows)

    def add_seq_edges(self, doc_ids: list, old_to_new: dict):
        """Window edge list of one document, host-side (ref: Text_GCN.py:142-166).  Debug/inspection
        only — forward() never calls it."""
        seq = [t for t in doc_ids if t != 0]
        edges, ids = [], []
        for p, src_word in enumerate(seq):
            for q in range(max(0, p - self.ngram), min(p + self.ngram + 1, len(seq))):
                edges.append([old_to_new[src_word], old_to_new[seq[q]]])
                ids.append(self.edges_matrix[src_word, seq[q]])
            edges.append([old_to_new[src_word], old_to_new[src_word]])
            ids.append(self.edges_matrix[src_word, src_word])
        return edges, ids

    def seq_to_graph(self, doc_ids):
        raise NotImplementedError("mgnns_b200 never materialises per-document DGL graphs; "
                                  "see mgnns::text_maxagg (ref: models/Text_GCN.py:168-211)")

    def forward(self, doc_ids, is_20ng=None):
        if not torch.is_tensor(doc_ids):
            doc_ids = torch.as_tensor(doc_ids)
        dev = self.node_hidden.weight.device
        doc_ids = doc_ids.to(device=dev, dtype=torch.int64)
        h = torch.ops.mgnns.text_maxagg(doc_ids, self.node_hidden.weight, self.seq_edge_w.weight,
                                        self.pmi_rowptr, self.pmi_col, self.pmi_eid,
                                        int(self.ngram), int(self.max_length), True)
        # ref order is dropout -> ReLU (:270-271); dropout scales by a non-negative factor, so
        # ReLU (fused in the kernel) commutes with it exactly.
        return self.dropout(h)
