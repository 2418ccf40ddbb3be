"""Set-transformer object encoder.  Stays in PyTorch (small dense matmuls; BASELINE.json north_star).

Module tree / state-dict names as in the reference's set_transformer.py:24-223 (q/k/v/o projectors, ln0/ln1, fc,
seeds, inducing points ``I``).  The presence handling reproduces the reference: ``(1 - presence) * 1e32`` is
subtracted from the attention logits before the 1/sqrt(d) scaling (set_transformer.py:41-43).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import skinny


def _layer_norm(x, ln):
    """nn.LayerNorm through ``ops.layer_norm``: for the tiny feature width used here (d = 16) ATen's kernels spend
    60 us forward and 0.24 ms backward per call on 2.6 MB of data (profiles/r01a, r01p); the thread-per-row kernel in
    csrc/support.cu does either direction, affine part and parameter gradients included, in one pass."""
    from . import ops
    if len(ln.normalized_shape) == 1 and ln.elementwise_affine:
        return ops.layer_norm(x, ln.weight, ln.bias, ln.eps)
    return ln(x)


def qkv_attention(queries, keys, values, presence=None):
    if queries.dim() == 3 and queries.shape[0] > 1 and queries.stride(0) == 0:
        # batch-shared queries: one (B*M, d) x (d, N) GEMM instead of B small ones
        logits = torch.matmul(keys, queries[0].t()).transpose(1, 2)
    else:
        logits = torch.matmul(queries, keys.transpose(1, 2))
    if presence is not None:
        logits = logits - (1. - presence.unsqueeze(-2)) * 1e32
    weights = F.softmax(logits / math.sqrt(queries.shape[-1]), -1)
    return torch.matmul(weights, values)


class MultiHeadQKVAttention(nn.Module):
    def __init__(self, d_k, d_v, n_heads):
        super().__init__()
        self.d_k, self.d_v, self.n_heads = d_k, d_v, n_heads
        d_k_p = int(math.ceil(d_k / n_heads)) * n_heads
        d_v_p = int(math.ceil(d_v / n_heads)) * n_heads
        self.q_projector = nn.Linear(d_k, d_k_p)
        self.k_projector = nn.Linear(d_k, d_k_p)
        self.v_projector = nn.Linear(d_v, d_v_p)
        self.o_projector = nn.Linear(d_v_p, d_v)

    def _split_heads(self, t):
        B, L, _ = t.shape
        H = self.n_heads
        return t if H == 1 else t.view(B, L, H, -1).permute(2, 0, 1, 3).reshape(H * B, L, -1)

    def forward(self, queries, keys, values, presence=None):
        assert queries.shape[2] == keys.shape[2]
        assert keys.shape[1] == values.shape[1]
        if presence is not None:
            assert values.shape[:2] == presence.shape
        B, N, _ = queries.shape
        H = self.n_heads
        if B > 1 and queries.stride(0) == 0:
            # batch-shared queries (the learnt seeds / inducing points, ``S.expand(B, -1, -1)``): project them once
            # instead of B times -- same values; the backward sums the batch before the (now tiny) weight-gradient GEMM
            q = self.q_projector(queries[:1]).expand(B, -1, -1)
        else:
            q = skinny.linear(queries, self.q_projector)
        q = self._split_heads(q)
        k = self._split_heads(skinny.linear(keys, self.k_projector))
        v = self._split_heads(skinny.linear(values, self.v_projector))
        if presence is not None and H > 1:
            presence = presence.repeat(H, 1)
        o = qkv_attention(q, k, v, presence)
        if H > 1:
            o = o.view(H, B, N, -1).permute(1, 2, 0, 3).reshape(B, N, -1)
        return skinny.linear(o, self.o_projector)


class MAB(nn.Module):
    def __init__(self, d, n_heads, layer_norm=False):
        super().__init__()
        self.layer_norm = layer_norm
        self.mqkv = MultiHeadQKVAttention(d_k=d, d_v=d, n_heads=n_heads)
        if layer_norm:
            self.ln0 = nn.LayerNorm(d)
            self.ln1 = nn.LayerNorm(d)
        self.fc = nn.Linear(d, d)

    def forward(self, queries, keys, presence=None):
        h = self.mqkv(queries, keys, keys, presence) + queries
        if presence is not None:
            assert presence.shape[1] == queries.shape[1] == keys.shape[1]
            h = h * presence.unsqueeze(-1)
        if self.layer_norm:
            h = _layer_norm(h, self.ln0)
        h = h + F.relu(skinny.linear(h, self.fc))
        if self.layer_norm:
            h = _layer_norm(h, self.ln1)
        return h


class SAB(nn.Module):
    def __init__(self, d, n_heads, layer_norm=False):
        super().__init__()
        self.mab = MAB(d=d, n_heads=n_heads, layer_norm=layer_norm)

    def forward(self, x, presence=None):
        from . import ops
        fused = ops.set_attention_block(x, presence, self.mab)    # one kernel per direction (csrc/sab.cu)
        return fused if fused is not None else self.mab(x, x, presence)


class ISAB(nn.Module):
    def __init__(self, d, n_heads, n_inducing_points, layer_norm=False):
        super().__init__()
        self.mab0 = MAB(d=d, n_heads=n_heads, layer_norm=layer_norm)
        self.mab1 = MAB(d=d, n_heads=n_heads, layer_norm=layer_norm)
        self.I = nn.Parameter(nn.init.xavier_uniform_(torch.zeros(1, n_inducing_points, d)))

    def forward(self, x, presence=None):
        h = self.mab0(self.I.expand(x.shape[0], -1, -1), x, presence)
        return self.mab1(x, h)


class PMA(nn.Module):
    def __init__(self, d, n_heads, n_seeds, layer_norm=False):
        super().__init__()
        self.mab = MAB(d=d, n_heads=n_heads, layer_norm=layer_norm)
        self.S = nn.Parameter(nn.init.xavier_uniform_(torch.zeros(1, n_seeds, d)))

    def forward(self, x, presence=None):
        return self.mab(self.S.expand(x.shape[0], -1, -1), x, presence)


class SetTransformer(nn.Module):
    """(B, M, dim_in) part descriptions -> (B, n_outputs, dim_out) object encodings."""

    def __init__(self, dim_in, dim_hidden, dim_out, n_outputs, n_layers, n_heads, layer_norm=False,
                 n_inducing_points: int = None):
        super().__init__()
        self.fc1 = nn.Linear(dim_in, dim_hidden)
        if n_inducing_points is None:
            blocks = [SAB(d=dim_hidden, n_heads=n_heads, layer_norm=layer_norm) for _ in range(n_layers)]
        else:
            blocks = [ISAB(d=dim_hidden, n_heads=n_heads, layer_norm=layer_norm,
                           n_inducing_points=n_inducing_points) for _ in range(n_layers)]
        self.sabs = nn.ModuleList(blocks)
        self.fc2 = nn.Linear(dim_hidden, dim_out)
        self.seeds = nn.Parameter(nn.init.xavier_uniform_(torch.zeros(1, n_outputs, dim_out)))
        self.multi_head_attention = MultiHeadQKVAttention(d_k=dim_out, d_v=dim_out, n_heads=n_heads)

    def forward(self, x, presence=None):
        h = skinny.linear(x, self.fc1)
        for block in self.sabs:
            h = block(h, presence)
        att = self.multi_head_attention
        if att.n_heads == 1 and h.shape[0] > 1:
            return self._pooled_output(h, presence)
        z = skinny.linear(h, self.fc2)
        return att(self.seeds.expand(x.shape[0], -1, -1), z, z, presence)

    def _pooled_output(self, h, presence):
        """``multi_head_attention(seeds, z, z)`` with ``z = fc2(h)`` (set_transformer.py:218-223) evaluated WITHOUT ever
        forming z, the keys or the values -- the same function of the same parameters, re-associated:

          keys   = z Wk^T + bk = h (Wk W2)^T + (Wk b2 + bk)            (fc2 and the projector are both affine)
          logits = q keys^T   = (q Wk W2) h^T + q (Wk b2 + bk)          -> a d_hidden-wide contraction
          out    = (A values) Wo^T + bo = (A h) (Wo Wv W2)^T + (Wo (Wv b2 + bv) + bo)   (rows of A = softmax sum to 1)

        The reference spends four (B*M, 256) x (256, 256) GEMMs here forward (and eight backward) on activations that
        have rank d_hidden = 16; composing the weights first -- three 256x256x16 products -- leaves (B*M, 16)-sized
        work and one (B*n_outputs, 16) x (16, 256) GEMM.  Parameters, state_dict and gradients are unchanged (autograd
        differentiates through the composed weights); values agree to fp32 rounding (tests/test_gpu_plumbing.py).
        """
        att = self.multi_head_attention
        w2, b2 = self.fc2.weight, self.fc2.bias                                   # (D, d), (D,)
        wk, bk = att.k_projector.weight, att.k_projector.bias
        wv, bv = att.v_projector.weight, att.v_projector.bias
        wo, bo = att.o_projector.weight, att.o_projector.bias
        q = att.q_projector(self.seeds[0])                                       # (N, D), batch-shared
        qk = q @ (wk @ w2)                                                       # (N, d)
        qb = q @ (wk @ b2 + bk)                                                  # (N,)
        logits = torch.matmul(h, qk.t()).transpose(1, 2) + qb.unsqueeze(-1)      # (B, N, M)
        if presence is not None:
            logits = logits - (1. - presence.unsqueeze(-2)) * 1e32
        weights = F.softmax(logits / math.sqrt(q.shape[-1]), -1)
        pooled = torch.matmul(weights, h)                                        # (B, N, d)
        wov = wo @ (wv @ w2)                                                     # (D, d)
        bov = wo @ (wv @ b2 + bv) + bo                                           # (D,)
        return torch.matmul(pooled, wov.t()) + bov
