"""Cross-attention scorer / aggregator with the reference's parameter names.

Mirrors architecture/transformer.py of the reference (pos_enc_1d :6-18,
MultiHeadCrossAttention :43-109, MLP :111-132, Transformer :134-152) so a
state_dict moves between the two unchanged.  The no-grad scoring path
(`get_scores`) runs on the library's CUDA kernels; it exploits that the
pre-softmax logit of a patch is linear in its embedding (z = emb . U), so keys
are never materialised.
"""
import math

import torch
from torch import nn

from . import ops
from .autograd import Linear, LayerNormFn, CrossAttentionFn


def pos_enc_1d(D, len_seq):
    """Sin/cos table (len_seq, D): even columns sin(n * w_j), odd columns cos(n * w_j), w_j = 10000^(-2j/D); the values
    of transformer.py:6-18."""
    if D % 2 != 0:
        raise ValueError('Cannot use sin/cos positional encoding with odd dim (got dim={:d})'.format(D))
    freq = torch.exp(torch.arange(0, D, 2, dtype=torch.float) * -(math.log(10000.0) / D))       # (D/2,)
    angle = torch.arange(0, len_seq).unsqueeze(1).float() * freq                                  # (len_seq, D/2)
    return torch.stack((torch.sin(angle), torch.cos(angle)), dim=-1).reshape(len_seq, D)


class _Temperature(nn.Module):
    """Parameter-free holder kept so `crs_attn.attention.temperature` exists as in the reference."""

    def __init__(self, temperature, attn_dropout):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)


class MultiHeadCrossAttention(nn.Module):
    def __init__(self, n_token, H, D, D_k, D_v, attn_dropout=0.1, dropout=0.1):
        super().__init__()
        self.n_token, self.H, self.D_k, self.D_v = n_token, H, D_k, D_v
        self.q = nn.Parameter(torch.empty((1, n_token, D)))
        bound = math.sqrt(1 / D_k)
        nn.init.uniform_(self.q, a=-bound, b=bound)
        self.q_w = Linear(D, H * D_k, bias=False)
        self.k_w = Linear(D, H * D_k, bias=False)
        self.v_w = Linear(D, H * D_v, bias=False)
        self.fc = Linear(H * D_v, D, bias=False)
        self.attention = _Temperature(D_k ** 0.5, attn_dropout)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(D, eps=1e-6)

    # -- no-grad scoring path (CUDA kernels) ------------------------------------------
    def score_basis(self):
        """U (D, H*T) with z[n, h*T+t] = emb_n . U[:, h*T+t]."""
        return ops.score_basis(self.q.detach().contiguous(), self.q_w.weight.detach().contiguous(),
                               self.k_w.weight.detach().contiguous(), self.H, self.D_k)

    @torch.no_grad()
    def get_logits(self, x):
        """(B,L,D) -> (B,L,H*T) pre-softmax attention logits."""
        B, L, D = x.shape
        z = ops.logits(x.reshape(B * L, D).contiguous().float(), self.score_basis())
        return z.view(B, L, -1)

    @torch.no_grad()
    def get_attn(self, x):
        """(B,L,D) -> (B,H,T,L) attention weights (transformer.py:71-83), eval semantics."""
        B, L = x.shape[:2]
        z = self.get_logits(x).view(B, L, self.H, self.n_token).permute(0, 2, 3, 1)
        return torch.softmax(z, dim=-1)

    # -- grad-mode aggregation -------------------------------------------------------------
    def forward(self, x):
        B, L = x.shape[:2]
        H, Dk, Dv, T = self.H, self.D_k, self.D_v, self.n_token
        if x.is_cuda:                       # library kernels forward and backward (autograd.py)
            q = self.q_w(self.q)[0] / self.attention.temperature                      # (T, H*Dk)
            k, v = self.k_w(x), self.v_w(x)                                             # (B, L, H*Dk / H*Dv)
            mask, keep = None, 1.0
            pd = self.attention.dropout.p
            if self.training and pd > 0:
                mask = (torch.rand((B, H, T, L), device=x.device) >= pd).float()
                keep = 1.0 / (1.0 - pd)
            out = CrossAttentionFn.apply(q, k, v, mask, keep, H, Dk, Dv)                # (B, T, H*Dv)
            out = self.dropout(self.fc(out)) + self.q
            return LayerNormFn.apply(out, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        # host path (CPU tensors: golden-vector tests of the module structure)
        qh = (self.q_w(self.q) / self.attention.temperature).view(T, H, Dk)
        kh = self.k_w(x).view(B, L, H, Dk)
        vh = self.v_w(x).view(B, L, H, Dv)
        attn = self.attention.dropout(torch.softmax(torch.einsum('thd,blhd->bhtl', qh, kh), dim=-1))
        out = torch.einsum('bhtl,blhd->bthd', attn, vh).reshape(B, T, H * Dv)
        out = self.dropout(self.fc(out)) + self.q
        return self.layer_norm(out)


class MLP(nn.Module):
    def __init__(self, D, D_inner, dropout=0.1):
        super().__init__()
        self.w_1 = Linear(D, D_inner)
        self.w_2 = Linear(D_inner, D)
        self.layer_norm = nn.LayerNorm(D, eps=1e-6)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        h = self.dropout(self.w_2(torch.relu(self.w_1(x)))) + x
        if x.is_cuda:
            return LayerNormFn.apply(h, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        return self.layer_norm(h)


class Transformer(nn.Module):
    def __init__(self, n_token, H, D, D_k, D_v, D_inner, attn_dropout=0.1, dropout=0.1):
        super().__init__()
        self.crs_attn = MultiHeadCrossAttention(n_token, H, D, D_k, D_v, attn_dropout=attn_dropout, dropout=dropout)
        self.mlp = MLP(D, D_inner, dropout=dropout)

    @torch.no_grad()
    def get_scores(self, x):
        """(B,L,D) -> (B,L): softmax over L per (head, token), mean over heads then tokens
        (transformer.py:143-148).  Runs entirely on the library's kernels."""
        z = self.crs_attn.get_logits(x)
        return ops.scores_from_logits(z.contiguous(), self.crs_attn.H, self.crs_attn.n_token)

    def forward(self, x):
        return self.mlp(self.crs_attn(x))
