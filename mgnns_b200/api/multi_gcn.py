"""MGNNS model assembly with the reference's constructors, forward signature and state_dict names
(ref: models/Multi_GCN_Multihead_att.py).  All arithmetic named by the north star runs in the mgnns
ops (the bi-LSTM included: ops.packed_bilstm); only the ResNet trunks (torchvision/cuDNN) stay torch modules.
"""
import math
import os
import pickle

import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter

from .. import ops
from .graph_util import CSRAdjacency, as_csr, gen_A, gen_adj
from .layers import CoAttention, MyAnotherMultiHeadAttention, MyMultiHeadAttention  # noqa: F401
from .text_gcn import Model as Text_GCN_Model

# forward schedule knob (see forward()): 1 = the scene channel's image-bank kernel waits for the last LSTM layer's launch
_PLACE_AFTER_LAST_LAYER = os.environ.get('MGNNS_PLACE_AFTER_LAST_LAYER', '1') == '1'
# side streams of forward() that run at normal priority: the two image channels (persistent image-bank kernels +
# image-query stacks); the LSTM / text-bank stacks / label channels are the critical path and get priority -1
_NORMAL_PRIORITY = tuple(int(x) for x in os.environ.get('MGNNS_NORMAL_PRIORITY_STREAMS', '1,2').split(',') if x != '')
_PKG_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data', 'label_graphs.npz')

# ref: models/Multi_GCN_Multihead_att.py:20-27 loads the label GloVe matrix at import time from a path
# that does not exist in the shipped tree; here it is resolved lazily.
save_label_pkl_path = 'data/glove/tumblr_label_glove.pkl'
glove_label_embedding = None


def get_glove_embedding(glove_file):
    with open(glove_file, 'rb') as f:
        return pickle.load(f)


def _default_label_embedding():
    global glove_label_embedding
    if glove_label_embedding is None:
        for path in (save_label_pkl_path, 'data/tumblr_label_glove.pkl'):
            if os.path.exists(path):
                glove_label_embedding = torch.from_numpy(np.array(get_glove_embedding(path)))
                break
        else:
            glove_label_embedding = torch.from_numpy(np.load(_PKG_DATA)['label_glove'])
    return glove_label_embedding


class GraphConvolution(nn.Module):
    """Kipf-style layer out = Â·(X·W) (+bias) (ref: models/Multi_GCN_Multihead_att.py:30-63).

    `adj` may be the dense Â tensor the reference passes or a CSRAdjacency.  Â is applied with the CSR
    SpMM kernel on whichever side of W is narrower ((Â·X)·W when in<=out, else Â·(X·W)); the two are
    equal up to fp32 rounding.
    """

    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.Tensor(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.Tensor(1, 1, out_features))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, input, adj, act=ops.ACT_NONE, slope=0.0):
        csr = as_csr(adj)
        bias = None if self.bias is None else self.bias.view(-1)
        needs_grad = torch.is_grad_enabled() and (input.requires_grad or self.weight.requires_grad)
        if (not needs_grad and input.dim() == 3 and input.shape[0] * csr.n_rows >= 2048
                and csr.n_rows == csr.n_cols == input.shape[1] and ops.gcn_fused_ok(input, self.weight)):
            # batched node features, forward only: one kernel, the aggregated rows never reach HBM
            return ops.gcn_fused(csr.fused_plan(input.shape[2]), input, self.weight, bias, act, slope)
        if self.in_features <= self.out_features:
            return ops.matmul_nn(csr.spmm(input), self.weight, bias, act, slope)
        out = csr.spmm(ops.matmul_nn(input, self.weight))
        if bias is not None:
            out = out + self.bias
        if act == ops.ACT_RELU:
            out = torch.relu(out)
        elif act == ops.ACT_LEAKY:
            out = torch.nn.functional.leaky_relu(out, slope)
        return out

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class Attention(nn.Module):
    """Label-query element-wise attention (ref: models/Multi_GCN_Multihead_att.py:65-133).

    out[b,c] = fc( dropout(softmax_d(Q[c,h,d]·K[b,h,d]/sqrt(d_h))) ⊙ V[b,h,d] ); the reference's
    O(B²) torch.cat loop (:114-115) and its hard-coded 7 labels are replaced by the closed form for
    any number of query rows.
    """

    def __init__(self, hid_dim, image_dim, n_heads, dropout):
        super().__init__()
        self.hid_dim = hid_dim
        self.n_heads = n_heads
        assert hid_dim % n_heads == 0
        self.w_q = nn.Linear(hid_dim, hid_dim)
        self.w_k = nn.Linear(image_dim, hid_dim)
        self.w_v = nn.Linear(image_dim, hid_dim)
        self.fc = nn.Linear(hid_dim, hid_dim)
        self.do = nn.Dropout(dropout)
        self.scale = torch.sqrt(torch.FloatTensor([hid_dim // n_heads]))

    def forward(self, query, key, value, mask=None):
        if mask is not None:
            raise NotImplementedError("mgnns_b200 Attention: the mask argument is never used by the model "
                                      "(ref: models/Multi_GCN_Multihead_att.py:476,:503)")
        dev = self.w_q.weight.device
        Q = ops.linear(query.to(device=dev, dtype=torch.float32), self.w_q.weight, self.w_q.bias)
        if key is value:
            kv = ops.linear(key, torch.cat([self.w_k.weight, self.w_v.weight], 0),
                            torch.cat([self.w_k.bias, self.w_v.bias], 0))
        else:
            kv = torch.cat([ops.linear(key, self.w_k.weight, self.w_k.bias),
                            ops.linear(value, self.w_v.weight, self.w_v.bias)], 1)
        p_drop = self.do.p if (self.training and self.do.p > 0) else 0.0
        seed = ops.new_seed() if p_drop > 0 else 0
        x = torch.ops.mgnns.label_attn(Q, kv, self.n_heads, 1.0 / float(self.scale), p_drop, seed)
        return ops.linear(x, self.fc.weight, self.fc.bias)


class IdentityTrunk(nn.Module):
    """Stand-in for a ResNet trunk when pre-extracted [B,2048,14,14] feature maps are fed (head-only
    runs).  Exposes the attribute names the model constructor reads (ref: model:274-294)."""

    def __init__(self):
        super().__init__()
        for name in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'):
            setattr(self, name, nn.Identity())


class Multi_GCN_Multihead_Att(nn.Module):
    def __init__(self, opt, num_labels, text_model, object_model, place_model,
                 object_num_classes, place_num_classes, object_t=0, place_t=0, in_channel=300,
                 object_adj_file=None, place_adj_file=None):
        super().__init__()
        self.emb_path = opt['emb_path']
        self.bidirectional = opt['bidirectional']
        self.num_directions = 2 if self.bidirectional else 1
        self.hidden_size = opt['hidden_size']
        self.bi_hidden_size = self.num_directions * opt['hidden_size']
        opt['bi_hidden_size'] = self.bi_hidden_size
        self.d_model = self.bi_hidden_size
        self.pad_idx = 0
        self.stack_num = opt['stack_num']
        self.n_head = opt['n_head']
        self.d_kv = opt['d_kv']
        self.is_regu = opt['is_regu']

        self.embedding = nn.Embedding(opt['vocab_size'], opt['emb_size'], padding_idx=self.pad_idx)
        self.init_weights(opt['emb_type'], self.pad_idx)

        rnn_kw = dict(input_size=opt['emb_size'], hidden_size=opt['hidden_size'], num_layers=opt['num_layers'],
                      bidirectional=opt['bidirectional'], batch_first=True, dropout=opt['dropout'])
        self.rnn = nn.GRU(**rnn_kw)      # ref: model:172, constructed but unused
        self.lstm = nn.LSTM(**rnn_kw)

        self.object_gate = nn.Linear(self.bi_hidden_size * 2, self.bi_hidden_size)
        self.place_gate = nn.Linear(self.bi_hidden_size * 2, self.bi_hidden_size)

        def stack(need_mask, kind):
            return nn.ModuleList([MyMultiHeadAttention(self.n_head, self.d_model, self.d_kv, dropout=opt['dropout'],
                                                       need_mask=need_mask, is_regu=self.is_regu,
                                                       interaction_type=kind) for _ in range(self.stack_num)])

        self.img_object_text_multi_head_att = stack(True, 'img_object_text')
        self.text_object_text_multi_head_att = MyAnotherMultiHeadAttention(
            self.n_head, self.d_model, self.d_kv, dropout=opt['dropout'], need_mask=False,
            interaction_type='text_object_text')
        self.img_place_text_multi_head_att = stack(True, 'img_place_text')
        self.text_place_text_multi_head_att = MyAnotherMultiHeadAttention(
            self.n_head, self.d_model, self.d_kv, dropout=opt['dropout'], need_mask=False,
            interaction_type='text_place_text')
        self.text_img_object_multi_head_att = stack(False, 'text_img_object')
        self.text_img_place_multi_head_att = stack(False, 'text_img_place')

        self.liner_img_object = nn.Linear(2048, self.bi_hidden_size)
        self.liner_img_place = nn.Linear(2048, self.bi_hidden_size)

        self.text_features = text_model
        self.object_features = nn.Sequential(
            object_model.conv1, object_model.bn1, object_model.relu, object_model.maxpool,
            object_model.layer1, object_model.layer2, object_model.layer3, object_model.layer4)
        self.place_features = nn.Sequential(
            place_model.conv1, place_model.bn1, place_model.relu, place_model.maxpool,
            place_model.layer1, place_model.layer2, place_model.layer3, place_model.layer4)

        self.num_labels = num_labels
        self.object_num_classes = object_num_classes
        self.place_num_classes = place_num_classes
        self.object_t = object_t
        self.place_t = place_t

        self.pooling = nn.MaxPool2d(14, 14)
        self.gc1 = GraphConvolution(in_channel, 1024)
        self.gc2 = GraphConvolution(1024, 2048)
        self.leakyrelu = nn.LeakyReLU(0.2)
        self.tanh = nn.Tanh()
        self.relu = nn.ReLU()
        self.object_attention = Attention(hid_dim=300, image_dim=self.object_num_classes, n_heads=5, dropout=0.5)
        self.place_attention = Attention(hid_dim=300, image_dim=self.place_num_classes, n_heads=5, dropout=0.5)

        self.object_linear_1 = nn.Linear(2048, 1024)
        self.object_linear_2 = nn.Linear(1024, 512)
        self.object_linear_3 = nn.Linear(512, 256)
        self.object_linear_5 = nn.Linear(300, 100)
        # ref hard-codes 700 = 7 labels x 100 (model:321,:329); generalised to num_labels x 100
        self.object_x_linear = nn.Linear(100 * num_labels, 300)
        self.place_linear_1 = nn.Linear(2048, 1024)
        self.place_linear_2 = nn.Linear(1024, 512)
        self.place_linear_3 = nn.Linear(512, 256)
        self.place_linear_5 = nn.Linear(300, 100)
        self.place_x_linear = nn.Linear(100 * num_labels, 300)

        self.dropout = nn.Dropout(0.5)
        self.multi_linear_1 = nn.Linear(1200, self.bi_hidden_size)
        self.multi_linear_2 = nn.Linear(self.bi_hidden_size, num_labels)

        object_adj, _ = gen_A(object_num_classes, self.object_t, object_adj_file)
        self.object_A = Parameter(torch.from_numpy(object_adj).float())
        place_adj, _ = gen_A(place_num_classes, self.place_t, place_adj_file)
        self.place_A = Parameter(torch.from_numpy(place_adj).float())

        self.image_normalization_mean = [0.485, 0.456, 0.406]
        self.image_normalization_std = [0.229, 0.224, 0.225]
        self.label_query = None          # optional override of the label GloVe query matrix
        self._adj_cache = {}

    # ------------------------------------------------------------------ helpers with reference names
    def init_weights(self, emb_type, pad_idx):
        """(ref: model:353-364)"""
        if emb_type == 'random':
            self.embedding.weight.data.uniform_(-0.1, 0.1)
        else:
            with open(self.emb_path, 'rb') as f:
                weights = pickle.load(f)
            self.embedding.weight.data = torch.Tensor(weights)
        self.embedding.weight.data[pad_idx] = 0

    def get_text_memory_bank(self, text, text_lens, return_last_state=True, after_first_projection=None):
        """Embedding -> 2-layer bi-LSTM over the valid tokens only -> zero-padded bank [B,L,300]
        (ref: model:366-398: pack_padded_sequence + cuDNN LSTM + pad_packed_sequence).  The recurrence
        runs in mgnns::lstm_rec over compacted tokens; nn.LSTM only holds the parameters."""
        batch_size, max_text_len = list(text.size())
        dev = self.embedding.weight.device
        plan = self.make_text_plan(text_lens, max_text_len)
        tokens = text.to(dev).reshape(-1).index_select(0, plan.tok_idx)
        fused_glue = (self.embedding.weight.shape[1] % 4 == 0 and self.bi_hidden_size % 4 == 0
                      and self.embedding.max_norm is None and not self.embedding.sparse)
        if fused_glue:
            text_embed = ops.embedding_rows(self.embedding.weight, tokens, self.embedding.padding_idx)   # [N, emb]
        else:
            text_embed = self.embedding(tokens)
        y = ops.packed_bilstm(self.lstm, text_embed, plan, self.training, after_first_projection)   # [N, 2H]
        if fused_glue:
            memory_bank = ops.pad_text_bank(y, plan, batch_size, max_text_len)     # one pass: valid rows + zero padding
        else:
            # scatter back to the padded layout; plan padding rows land in one dummy row that is sliced off
            memory_bank = y.new_zeros(batch_size * max_text_len + 1, y.shape[1]).index_copy(0, plan.flat_idx, y)
            memory_bank = memory_bank[:batch_size * max_text_len].view(batch_size, max_text_len, y.shape[1])
        assert memory_bank.size() == torch.Size([batch_size, max_text_len, self.bi_hidden_size])
        if not return_last_state:
            return memory_bank
        H = self.hidden_size
        if self.bidirectional:
            # ref: cat(enc_final_state[-1] (reverse), enc_final_state[-2] (forward)), model:392
            last = torch.cat((y.index_select(0, plan.first_idx)[:, H:], y.index_select(0, plan.last_idx)[:, :H]), 1)
        else:
            last = y.index_select(0, plan.last_idx)
        return memory_bank, last

    def make_text_plan(self, text_lens, max_text_len, capacity=None):
        """Length-sorted tile schedule + compact-token indices for the LSTM (host work on the CPU
        lengths, one small H2D).  Memoised on the lengths tensor, so a data pipeline can call this
        while prefetching a batch (on its copy stream) and forward() will find it."""
        lens = text_lens if torch.is_tensor(text_lens) else torch.as_tensor(text_lens)
        dev = self.embedding.weight.device
        key = (lens.data_ptr(), lens._version, tuple(lens.shape), int(max_text_len), dev)
        cache = self.__dict__.setdefault('_text_plans', {})
        hit = cache.get(key)
        if hit is not None and hit[0]() is lens and (capacity is None or hit[1].capacity >= int(capacity)):
            return hit[1]
        import weakref
        plan = ops.LstmPlan(lens, max_text_len, dev, capacity)
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        cache[key] = (weakref.ref(lens), plan)
        return plan

    def _img_bank(self, feats, linear):
        bank, pooled, _ = torch.ops.mgnns.imgbank(feats, linear.weight, linear.bias)
        if not feats.requires_grad:
            # pooled depends on the feature map only: with frozen / absent trunks nothing upstream wants its gradient,
            # and leaving it attached would make the image-bank backward (the weight gradient) wait for the whole
            # label-channel backward to deliver a gradient nobody reads
            pooled = pooled.detach()
        return bank, pooled

    def get_img_object_memory_bank(self, img_object_feats):
        """[B,2048,14,14] -> [B,196,300] (ref: model:400-416)"""
        return self._img_bank(img_object_feats, self.liner_img_object)[0]

    def get_img_place_memory_bank(self, img_place_feats):
        """(ref: model:418-428)"""
        return self._img_bank(img_place_feats, self.liner_img_place)[0]

    def _adj_csr(self, name) -> CSRAdjacency:
        """Â = gen_adj(A).detach() in CSR.  The reference recomputes the dense Â every forward
        (model:461,:490); A never receives a gradient, so Â is rebuilt only when A's storage changes."""
        A = getattr(self, name)
        key = (A.data_ptr(), A._version, A.device)
        hit = self._adj_cache.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                csr = CSRAdjacency.from_dense(gen_adj(A.detach()))
            hit = (key, csr)
            self._adj_cache[name] = hit
        return hit[1]

    def _label_channel(self, pooled, inp, adj_name, attention, linear_5, x_linear, query):
        """Label-graph GCN -> per-sample class scores -> label attention (ref: model:459-479 / :488-506)."""
        dev = pooled.device
        inp = inp[0].to(device=dev, dtype=torch.float32)
        csr = self._adj_csr(adj_name)
        x = self.gc1(inp, csr, ops.ACT_LEAKY, self.leakyrelu.negative_slope)
        x = self.gc2(x, csr)                                                     # [N, 2048]
        scores = torch.ops.mgnns.mm(pooled, x, None, False, True, ops.ACT_NONE, 0.0)   # [B, N] = pooled @ x^T
        att = attention(query=query, key=scores, value=scores)                   # [B, C, 300]
        att = ops.linear(att, linear_5.weight, linear_5.bias)                    # [B, C, 100]
        att = att.reshape(pooled.size(0), -1)
        return ops.linear(att, x_linear.weight, x_linear.bias)                  # [B, 300]

    def _query(self):
        """Label GloVe query matrix on the model's device (cached: the reference re-uploads it every
        forward, model:462-468)."""
        q = self.label_query if self.label_query is not None else _default_label_embedding()
        if q.shape[0] < self.num_labels:
            raise ValueError("label query matrix has %d rows, model has %d labels" % (q.shape[0], self.num_labels))
        dev = self.gc1.weight.device
        hit = self.__dict__.get('_query_dev')
        if hit is None or hit[0] is not q or hit[1].device != dev:
            hit = (q, q[:self.num_labels].to(device=dev, dtype=torch.float32))
            self.__dict__['_query_dev'] = hit
        return hit[1]

    # ------------------------------------------------------------------ forward (ref: model:431-567)
    def forward(self, text, text_lens, text_mask, object_feature, place_feature, object_inp, place_inp,
                return_last_state=True):
        query = self._query()
        text_mask = text_mask.to(torch.float32)

        def text_gcn():
            return self.text_features(text)                                           # [B, 300]

        def text_bank(hook=None):
            return self.get_text_memory_bank(text, text_lens, True, hook)[0]          # [B, L, 300]

        def object_channel():
            self.object_feature = self.object_features(object_feature)                # [B, 2048, 14, 14]
            bank, pooled = self._img_bank(self.object_feature, self.liner_img_object)
            return bank, self._label_channel(pooled, object_inp, 'object_A', self.object_attention,
                                             self.object_linear_5, self.object_x_linear, query)

        def place_channel():
            self.place_feature = self.place_features(place_feature)
            bank, pooled = self._img_bank(self.place_feature, self.liner_img_place)
            return bank, self._label_channel(pooled, place_inp, 'place_A', self.place_attention,
                                             self.place_linear_5, self.place_x_linear, query)

        def stack(layers, q, bank, mask=None):
            x = q
            for layer in layers:
                x = layer(q=x, k=bank, v=bank, mask=mask)[0]
            return x

        streams = self._branch_streams(7)
        if streams is None:
            # one stream, the reference's order (ref: model:444-546)
            ops.set_concurrent_streams(False)
            text_feature = text_gcn()
            text_memory_bank = text_bank()
            img_object_memory_bank, object_x_attention = object_channel()
            img_place_memory_bank, place_x_attention = place_channel()
            img_object_text = stack(self.img_object_text_multi_head_att, object_x_attention, text_memory_bank, text_mask)
            img_place_text = stack(self.img_place_text_multi_head_att, place_x_attention, text_memory_bank, text_mask)
            text_img_object = stack(self.text_img_object_multi_head_att, text_feature, img_object_memory_bank)
            text_img_place = stack(self.text_img_place_multi_head_att, text_feature, img_place_memory_bank)
        else:
            # The forward is a small dependency graph: four independent channels, then four attention stacks that
            # each need two of them.  Every chain stays on its own stream and waits only for the tensor it needs
            # (events), so e.g. the image-query stacks overlap the LSTM.  Autograd replays each backward op on its
            # forward stream, so the backward pass forks the same way; a CUDA-graph capture records parallel paths.
            # The label-graph half of an image channel (pooled -> label GCN scores -> label attention) has its own
            # stream: in the backward pass it is fed late (through the text-bank stacks), and on a shared stream it
            # would sit in front of the image-bank weight gradient, which is ready much earlier.
            main, (s_txt, s_obj, s_plc, s_obj_lab, s_plc_lab, s_txt2, s_txt3) = streams
            ops.set_concurrent_streams(True)                # stays on: the backward pass of this forward forks the same way
            for side in (s_txt, s_obj, s_plc, s_obj_lab, s_plc_lab, s_txt2, s_txt3):
                side.wait_stream(main)                      # fork: after everything already enqueued on main
            # The LSTM recurrence is latency-bound and leaves most SMs idle, while the image-bank kernels are
            # persistent and take every SM they can get: the image channels therefore start once the LSTM's first
            # input projection is done, i.e. when the recurrence is being launched on its high-priority stream, so
            # the recurrence gets its SMs first and the (dynamically scheduled) tensor-core kernels fill the rest.
            lstm_ready, helpers = [], []

            def before_recurrence():
                ev, helper = ops.delayed_event(dev_main, 'fwd_gate%d' % len(lstm_ready))
                lstm_ready.append(ev)
                if helper is not None:
                    helpers.append(helper)
            dev_main = text_mask.device
            with torch.cuda.stream(s_txt):
                text_memory_bank = text_bank(before_recurrence)
                ev_bank = s_txt.record_event()

            def image_channel(s_img, s_lab, gates, trunk, feature, linear, inp, adj, attention, linear_5, x_linear, attr):
                with torch.cuda.stream(s_img):
                    for ev in gates:
                        s_img.wait_event(ev)
                    fmap = trunk(feature)
                    setattr(self, attr, fmap)                                         # side-effect attrs of the reference
                    bank, pooled = self._img_bank(fmap, linear)
                    ev_img = s_img.record_event()
                with torch.cuda.stream(s_lab):
                    s_lab.wait_event(ev_img)
                    x_att = self._label_channel(pooled, inp, adj, attention, linear_5, x_linear, query)
                    ev_lab = s_lab.record_event()
                pooled.record_stream(s_lab)
                return bank, x_att, ev_lab

            # the scene channel's image-bank kernel is released behind the LAST layer's recurrence launch: a
            # persistent kernel that already holds every SM would make that recurrence wait for all of it
            img_object_memory_bank, object_x_attention, ev_obj = image_channel(
                s_obj, s_obj_lab, lstm_ready[:1], self.object_features, object_feature, self.liner_img_object, object_inp,
                'object_A', self.object_attention, self.object_linear_5, self.object_x_linear, 'object_feature')
            img_place_memory_bank, place_x_attention, ev_plc = image_channel(
                s_plc, s_plc_lab, lstm_ready if _PLACE_AFTER_LAST_LAYER else lstm_ready[:1], self.place_features,
                place_feature, self.liner_img_place, place_inp, 'place_A', self.place_attention, self.place_linear_5,
                self.place_x_linear, 'place_feature')
            text_feature = text_gcn()
            ev_tf = main.record_event()
            with torch.cuda.stream(s_obj):
                s_obj.wait_event(ev_tf)
                text_img_object = stack(self.text_img_object_multi_head_att, text_feature, img_object_memory_bank)
            with torch.cuda.stream(s_plc):
                s_plc.wait_event(ev_tf)
                text_img_place = stack(self.text_img_place_multi_head_att, text_feature, img_place_memory_bank)
            # NOT on s_txt, the stream that produced the text bank: in the backward pass autograd sums the four gradients
            # of the bank on that stream, each sum waiting for its producer — a stack queued behind those sums would
            # wait for the OTHER stack's backward (measured: the two text-bank stacks ran back to back, 0.85 ms)
            with torch.cuda.stream(s_txt3):
                s_txt3.wait_event(ev_bank)
                s_txt3.wait_event(ev_obj)
                img_object_text = stack(self.img_object_text_multi_head_att, object_x_attention, text_memory_bank,
                                        text_mask)
            # both text-bank stacks sit between the LSTM forward and the LSTM backward — the critical path of the step —
            # so both run on high-priority streams (as do the label channels that produce their queries): their small
            # kernels get SM slots ahead of the image-query stacks and the deferred weight gradients
            with torch.cuda.stream(s_txt2):
                s_txt2.wait_event(ev_bank)
                s_txt2.wait_event(ev_plc)
                img_place_text = stack(self.img_place_text_multi_head_att, place_x_attention, text_memory_bank,
                                       text_mask)
            for side in (s_txt, s_obj, s_plc, s_obj_lab, s_plc_lab, s_txt2, s_txt3, *helpers):
                main.wait_stream(side)                      # join
            # tensors that crossed streams: tell the caching allocator about every stream that read them
            for t, readers in ((text_feature, (s_obj, s_plc)), (object_x_attention, (s_txt3,)),
                               (text_memory_bank, (s_txt2, s_txt3)), (place_x_attention, (s_txt2,)),
                               (text_img_object, (main,)), (text_img_place, (main,)), (img_object_text, (main,)),
                               (img_place_text, (main,)),
                               (text_mask, (s_txt2, s_txt3)), (query, (s_obj_lab, s_plc_lab))):
                for r in readers:
                    t.record_stream(r)

        multi_feature = torch.cat([text_img_object, text_img_place, img_object_text, img_place_text], dim=1)
        multi_feature = ops.linear(multi_feature, self.multi_linear_1.weight, self.multi_linear_1.bias)
        multi_feature = self.dropout(multi_feature)
        return ops.linear(multi_feature, self.multi_linear_2.weight, self.multi_linear_2.bias)

    # ------------------------------------------------------------------ side streams for the independent chains
    def _branch_streams(self, n):
        """(current stream, n side streams) when `branch_streams` is on (attribute, or env MGNNS_BRANCH_STREAMS=1)
        and the model lives on a CUDA device; None otherwise.  Results are identical either way: same kernels, same
        order within a chain."""
        enabled = self.__dict__.get('branch_streams')
        if enabled is None:
            enabled = os.environ.get('MGNNS_BRANCH_STREAMS', '0') == '1'
        dev = self.gc1.weight.device
        if not enabled or dev.type != 'cuda':
            return None
        pool = self.__dict__.get('_branch_pool')
        if pool is None or pool[0] != dev or len(pool[1]) < n:
            # the first side stream carries the LSTM chain (latency-bound, leaves most SMs idle): high priority, so
            # its CTAs are placed first and the dynamically scheduled tensor-core kernels fill the rest of the GPU
            pool = (dev, [torch.cuda.Stream(device=dev, priority=(0 if i in _NORMAL_PRIORITY else -1)) for i in range(n)])
            self.__dict__['_branch_pool'] = pool
            # gc1/gc2 are shared by the object and place channels: their AccumulateGrad nodes see gradients from
            # two streams by design
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        return torch.cuda.current_stream(dev), tuple(pool[1][:n])

    def get_config_optim(self, lr, lrp):
        """Optimiser groups exactly as the reference (model:569-585): the classifier tail, image-bank
        Linears, label-attention tail and the LSTM embedding are NOT in any group (never stepped)."""
        return [
            {'params': self.text_features.parameters(), 'lr': lr * 10},
            {'params': self.object_features.parameters(), 'lr': lr * lrp},
            {'params': self.place_features.parameters(), 'lr': lr * lrp},
            {'params': self.gc1.parameters(), 'lr': lr},
            {'params': self.gc2.parameters(), 'lr': lr},
            {'params': self.object_attention.parameters(), 'lr': lr},
            {'params': self.place_attention.parameters(), 'lr': lr},
            {'params': self.lstm.parameters(), 'lr': lr * 10},
            {'params': self.img_object_text_multi_head_att.parameters(), 'lr': lr},
            {'params': self.img_place_text_multi_head_att.parameters(), 'lr': lr},
            {'params': self.text_img_object_multi_head_att.parameters(), 'lr': lr},
            {'params': self.text_img_place_multi_head_att.parameters(), 'lr': lr},
        ]


def place_resnet(arch='resnet50'):
    """Places365 ResNet (ref: model:586-595).  The checkpoint is not shipped; random init when absent."""
    import torchvision.models as models
    model = models.__dict__[arch](num_classes=365)
    model_file = 'weights/%s_places365.pth.tar' % arch
    if os.path.exists(model_file):
        checkpoint = torch.load(model_file, map_location=lambda storage, loc: storage)
        model.load_state_dict({k.replace('module.', ''): v for k, v in checkpoint['state_dict'].items()})
    return model


def Text_model(data_root_path, vocab_root_path, text_min_count, window_size, num_labels, ngram, text_dropout,
               min_cooccurence):
    """(ref: model:598-615)"""
    from .pmi import cal_PMI
    from .vocab import get_vocab_list
    vocab = get_vocab_list(data_root_path, vocab_root_path, text_min_count)
    edges_weights, edges_mappings, count = cal_PMI(data_root_path, vocab_root_path, min_count=text_min_count,
                                                   phase='train', window_size=window_size,
                                                   min_cooccurence=min_cooccurence)
    return Text_GCN_Model(num_labels, hidden_size_node=300, vocab=vocab, n_gram=ngram, drop_out=text_dropout,
                          edges_matrix=edges_mappings, edges_num=count, pmi=edges_weights, cuda=True,
                          trainable_edges=True)


def multi_gcn_multihead_att_model(opt, num_labels, object_num_classes, place_num_classes, object_t, place_t,
                                  data_root_path, vocab_root_path, text_min_count, window_size, ngram,
                                  min_cooccurence, text_dropout=0.5, pretrained=True, object_adj_file=None,
                                  place_adj_file=None, in_channel=300):
    """Factory with the reference's keyword arguments (ref: model:619-642, called at entry:144-157)."""
    import torchvision.models as models
    try:
        object_model = models.resnet101(pretrained=pretrained)
    except Exception:           # no network / no cached ImageNet weights -> random init (BASELINE configs)
        object_model = models.resnet101(weights=None)
    place_model = place_resnet()
    text_model = Text_model(data_root_path, vocab_root_path, text_min_count, window_size, num_labels, ngram,
                            text_dropout, min_cooccurence)
    return Multi_GCN_Multihead_Att(opt, num_labels, text_model=text_model, object_model=object_model,
                                   place_model=place_model, object_num_classes=object_num_classes,
                                   place_num_classes=place_num_classes, in_channel=in_channel,
                                   object_t=object_t, place_t=place_t, object_adj_file=object_adj_file,
                                   place_adj_file=place_adj_file)
