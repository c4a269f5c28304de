"""Attention / FFN / LayerNorm layers with the reference's constructors, forward signatures and
state_dict names (ref: models/submodules.py, models/moudles.py), computing through the mgnns ops.

On-path classes (hand-written kernels): LayerNorm, MultiHeadAttention, PositionwiseFeedForward,
MyMultiHeadAttention.  Off-path classes that the reference only constructs (their call sites are
commented out, models/Multi_GCN_Multihead_att.py:517-519,:530-532) or never uses keep their
signatures in plain torch so that state_dicts round-trip: ScaledDotProductAttention (generic form),
PositionalEncoding, AnotherMultiHeadAttention, MyAnotherMultiHeadAttention, CoAttention,
masked_mean, masked_max, MaskedSoftmax.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class LayerNorm(nn.Module):
    """gamma * (x - mean) / (std_unbiased + eps) + beta  (ref: models/submodules.py:142-156).

    Not nn.LayerNorm: unbiased std and eps added to std, not to the variance.
    """

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x, residual=None):
        return torch.ops.mgnns.add_layernorm(x, residual, self.gamma, self.beta, self.eps)


class ScaledDotProductAttention(nn.Module):
    """Generic scaled dot-product attention (ref: models/submodules.py:97-119).

    Kept for interface parity; MultiHeadAttention does not route through it (the single-query
    case is fused in mgnns::attn_q1).  Plain torch, off the hot path.
    """

    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)
        self.softmax = nn.Softmax(dim=2)

    def forward(self, q, k, v, mask=None):
        attn = torch.bmm(q, k.transpose(1, 2)) / self.temperature
        if mask is not None:
            attn = attn.masked_fill(mask == 0.0, float("-inf"))
        attn = self.dropout(self.softmax(attn))
        return torch.bmm(attn, v), attn


class MultiHeadAttention(nn.Module):
    """Single-query multi-head attention block (ref: models/submodules.py:15-94).

    With len_q == 1 the projections re-associate:
        score_h(l) = <W_k,h^T (W_q,h q + b_q,h), k_l> / sqrt(d_k)   (+ a per-head constant from b_k that
                                                                     softmax cancels, so w_ks.bias is unused)
        out_h      = W_v,h (sum_l p_l k_l) + (sum_l p_l) b_v,h
    so the [B*L,300]x[300,512] key/value projections disappear and the memory bank is streamed
    once by mgnns::attn_q1.  Requires k is v (the only way the model calls it, model:512-513).
    """

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, is_regu=False):
        super().__init__()
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.is_regu = is_regu

        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))

        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5))
        self.layer_norm = LayerNorm(d_model)

        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.dropout = nn.Dropout(dropout)

    def diff_outputs(self, inputs):
        """Head-diversity regulariser (ref: models/submodules.py:38-53); off by default (is_regu=False)."""
        x = F.normalize(inputs, p=2, dim=-1)
        assert x.size(1) == 1, 'in our work, the sequence len for the query is only 1'
        x1 = x.squeeze(1)
        cos = torch.bmm(x1, x1.permute(0, 2, 1)) ** 2
        n_head = inputs.size(2)
        idx = torch.arange(0, n_head)
        cos[:, idx, idx] = 0
        return torch.sum(cos, dim=[1, 2]).div_(n_head * (n_head - 1))

    def forward(self, q, k, v, mask=None):
        d_k, d_v, n_head = self.d_k, self.d_v, self.n_head
        sz_b, len_q, d_model = q.size()
        if len_q != 1:
            raise NotImplementedError("mgnns_b200 MultiHeadAttention: only single-query attention (len_q == 1) "
                                      "is on the MGNNS path (ref: models/submodules.py:41)")
        if k is not v and not (k.data_ptr() == v.data_ptr() and k.shape == v.shape and k.stride() == v.stride()):
            raise NotImplementedError("mgnns_b200 MultiHeadAttention: keys and values must be the same memory bank "
                                      "(ref: models/Multi_GCN_Multihead_att.py:512-513)")
        bank = k
        residual = q.reshape(sz_b, d_model)
        mask2d = None
        if mask is not None:
            mask2d = mask.reshape(sz_b, -1).to(torch.float32)

        qp = ops.linear(residual, self.w_qs.weight, self.w_qs.bias)                    # [B, H*dk]
        u = torch.ops.mgnns.head_mm(qp, self.w_ks.weight, n_head, 0)                    # [B, H*D]
        p_drop = self.attention.dropout.p if (self.training and self.attention.dropout.p > 0) else 0.0
        seed = ops.new_seed() if p_drop > 0 else 0
        ctx, attn, psum, _ = torch.ops.mgnns.attn_q1(u.view(sz_b, n_head, d_model), bank, mask2d,
                                                     1.0 / float(self.attention.temperature), p_drop, seed)
        o = torch.ops.mgnns.head_mm(ctx.view(sz_b, n_head * d_model), self.w_vs.weight, n_head, 1)  # [B, H*dv]
        o = torch.addcmul(o.view(sz_b, n_head, d_v), psum.unsqueeze(-1), self.w_vs.bias.view(1, n_head, d_v))
        if self.is_regu:
            regu_term = self.diff_outputs(o.view(sz_b, 1, n_head, d_v))
        out = ops.linear(o.view(sz_b, n_head * d_v), self.fc.weight, self.fc.bias)
        out = self.dropout(out)
        out = self.layer_norm(out, residual).view(sz_b, 1, d_model)
        if self.is_regu:
            return out, attn, regu_term
        return out, attn


class PositionwiseFeedForward(nn.Module):
    """LN(x + dropout(W2 relu(W1 x))) with kernel-size-1 Conv1d weights (ref: models/submodules.py:122-139)."""

    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        h = ops.linear(x2, self.w_1.weight.squeeze(-1), self.w_1.bias, ops.ACT_RELU)
        o = ops.linear(h, self.w_2.weight.squeeze(-1), self.w_2.bias)
        o = self.dropout(o)
        return self.layer_norm(o, x2).view(shape)


class PositionalEncoding(nn.Module):
    """Sinusoid table (ref: models/submodules.py:159-182); unused by the model."""

    def __init__(self, d_hid, n_position=200):
        super().__init__()
        pos = np.arange(n_position)[:, None] / np.power(10000, 2 * (np.arange(d_hid)[None, :] // 2) / d_hid)
        pos[:, 0::2] = np.sin(pos[:, 0::2])
        pos[:, 1::2] = np.cos(pos[:, 1::2])
        self.register_buffer('pos_table', torch.FloatTensor(pos).unsqueeze(0))

    def forward(self, x):
        return x + self.pos_table[:, :x.size(1)].clone().detach()


class MyMultiHeadAttention(nn.Module):
    """One cross-modal attention layer: slf_attn then pos_ffn (ref: models/moudles.py:198-230)."""

    def __init__(self, n_head, d_model, d_kv, dropout=0.1, need_mask=False, is_regu=False, interaction_type=None):
        super().__init__()
        self.need_mask = need_mask
        self.is_regu = is_regu
        self.interaction_type = interaction_type
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_kv, d_kv, dropout=dropout, is_regu=is_regu)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_model, dropout=dropout)

    def forward(self, q, k, v, mask=None):
        if len(q.shape) == 2:
            q = q.unsqueeze(1)
        if mask is not None:
            mask = mask.unsqueeze(1)
        if self.need_mask:
            assert mask is not None, 'Please pass the attention mask to the multi-head'
        if self.is_regu:
            enc_output, enc_slf_attn, head_diff = self.slf_attn(q, k, v, mask)
        else:
            enc_output, enc_slf_attn = self.slf_attn(q, k, v, mask)
        enc_output = self.pos_ffn(enc_output).squeeze(1)
        if self.is_regu:
            return enc_output, enc_slf_attn, head_diff
        return enc_output, enc_slf_attn


# --------------------------------------------------------------------------- off-path, plain torch
class _TorchLayerNorm(nn.Module):
    """Same formula as LayerNorm, in torch ops, for the off-path modules (usable on any device)."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        return self.gamma * (x - x.mean(-1, keepdim=True)) / (x.std(-1, keepdim=True) + self.eps) + self.beta


class _TorchFFN(nn.Module):
    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = _TorchLayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        o = self.w_2(F.relu(self.w_1(x.transpose(1, 2)))).transpose(1, 2)
        return self.layer_norm(self.dropout(o) + x)


class AnotherMultiHeadAttention(nn.Module):
    """Batch-major variant (ref: models/moudles.py:232-288); constructed by the model, never called."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5))
        self.layer_norm = _TorchLayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None):
        H, dk, dv = self.n_head, self.d_k, self.d_v
        B, lq, _ = q.size()
        lk, lv = k.size(1), v.size(1)
        residual = q
        qh = self.w_qs(q).view(B, lq, H, dk).permute(0, 2, 1, 3).reshape(-1, lq, dk)
        kh = self.w_ks(k).view(B, lk, H, dk).permute(0, 2, 1, 3).reshape(-1, lk, dk)
        vh = self.w_vs(v).view(B, lv, H, dv).permute(0, 2, 1, 3).reshape(-1, lv, dv)
        if mask is not None:
            mask = mask.repeat(H, 1, 1)
        out, attn = self.attention(qh, kh, vh, mask=mask)
        out = out.view(B, H, lq, dv).permute(0, 2, 1, 3).reshape(B, lq, -1)
        out = self.layer_norm(self.dropout(self.fc(out)) + residual)
        return out, attn


class MyAnotherMultiHeadAttention(nn.Module):
    """(ref: models/moudles.py:292-324); constructed by the model (model:210,:236), never called."""

    def __init__(self, n_head, d_model, d_kv, dropout=0.1, need_mask=False, interaction_type=None):
        super().__init__()
        self.need_mask = need_mask
        self.slf_attn = AnotherMultiHeadAttention(n_head, d_model, d_kv, d_kv, dropout=dropout)
        self.pos_ffn = _TorchFFN(d_model, d_model, dropout=dropout)

    def forward(self, q, k, v, mask=None):
        q, k, v = [t.unsqueeze(1) if t.dim() == 2 else t for t in (q, k, v)]
        if mask is not None:
            mask = mask.unsqueeze(1)
        if self.need_mask:
            assert mask is not None, 'Please pass the attention mask to the multi-head'
        out, attn = self.slf_attn(q, k, v, mask)
        return self.pos_ffn(out).squeeze(1), attn


def masked_mean(input, mask=None, dim=1):
    """(ref: models/moudles.py:9-20); never called by the model."""
    if mask is None:
        return torch.mean(input, dim=dim)
    m = mask.unsqueeze(-1)
    return (input * m).sum(dim=dim) / m.sum(dim=1)


def masked_max(input, mask=None, dim=1):
    """(ref: models/moudles.py:23-34); never called by the model."""
    if mask is not None:
        input = input.masked_fill(mask.unsqueeze(-1).expand_as(input) == 0.0, float('-inf'))
    return torch.max(input, dim=dim)[0]


class MaskedSoftmax(nn.Module):
    """(ref: models/moudles.py:37-49); never called by the model."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, logit, mask=None):
        dist = F.softmax(logit - torch.max(logit, dim=self.dim, keepdim=True)[0], dim=self.dim)
        if mask is not None:
            dist = dist * mask
            dist = dist / dist.sum(self.dim, keepdim=True)
        return dist


class CoAttention(nn.Module):
    """Additive co-attention alternative (ref: models/moudles.py:51-196); imported by the model file
    (model:15) but never constructed.  Plain torch, same parameter names."""

    def __init__(self, text_feat_size, img_object_feat_size, img_place_feat_size, interaction_type='co_att'):
        super().__init__()
        t, o, p = text_feat_size, img_object_feat_size, img_place_feat_size
        self.text_feat_size, self.img_object_feat_size, self.img_place_feat_size = t, o, p
        self.interaction_type = interaction_type
        self.v_text_object = nn.Linear(t, 1, bias=False)
        self.v_text_place = nn.Linear(t, 1, bias=False)
        self.v_img_object = nn.Linear(o, 1, bias=False)
        self.v_img_place = nn.Linear(p, 1, bias=False)
        self.text2img_object_project = nn.Linear(t, o, bias=False)
        self.text2img_place_project = nn.Linear(t, p, bias=False)
        self.img_object2text_project = nn.Linear(o, t, bias=False)
        self.img_place2text_project = nn.Linear(p, t, bias=False)
        self.img_object_project = nn.Linear(o, o)
        self.img_place_project = nn.Linear(p, p)
        self.text_object_project = nn.Linear(t, t)
        self.text_place_project = nn.Linear(t, t)
        self.dropout = nn.Dropout(0.5)
        self.softmax = MaskedSoftmax(dim=1)
        self.linear = nn.Linear(t * 2 + o + p, t)

    def _scores(self, single, many, proj_many, proj_single, v):
        # additive attention: v^T tanh(P_many many_i + P_single single)
        return v(torch.tanh(proj_many(many) + proj_single(single).unsqueeze(1))).squeeze(-1)

    def forward(self, text_feat, text_feats, img_object_feat, img_object_feats, img_place_feat, img_place_feats,
                src_mask):
        a_o = self.softmax(self._scores(text_feat, img_object_feats, self.img_object2text_project,
                                        self.text_object_project, self.v_text_object))
        a_p = self.softmax(self._scores(text_feat, img_place_feats, self.img_place2text_project,
                                        self.text_place_project, self.v_text_place))
        a_to = self.softmax(self._scores(img_object_feat, text_feats, self.text2img_object_project,
                                         self.img_object_project, self.v_img_object), mask=src_mask)
        a_tp = self.softmax(self._scores(img_place_feat, text_feats, self.text2img_place_project,
                                         self.img_place_project, self.v_img_place), mask=src_mask)
        ctx = [torch.bmm(a_o.unsqueeze(1), img_object_feats).squeeze(1),
               torch.bmm(a_p.unsqueeze(1), img_place_feats).squeeze(1),
               torch.bmm(a_to.unsqueeze(1), text_feats).squeeze(1),
               torch.bmm(a_tp.unsqueeze(1), text_feats).squeeze(1)]
        return self.dropout(self.linear(torch.cat(ctx, dim=1)))
