"""Transformer blocks of the matching modules — PyTorch host code with the reference's parameter
names (core/unopose/model/transformer.py).  These layers sit inside the matching modules but are dense
cuBLAS work; `SparseToDenseTransformer._sample_feats` touches the hot path (a row gather,
`transformer.py:655-662`) and `GeometricStructureEmbedding` runs the fused tcgen05 kernels of "next" row f2
 on CUDA.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..model_utils import _gather_rows, pairwise_distance
from .linear import linear, linear_multi


def _heads(x, h):
    """(B, N, h*c) -> (B, h, N, c)"""
    b, n, _ = x.shape
    return x.reshape(b, n, h, -1).permute(0, 2, 1, 3)


def _merge(x):
    """(B, h, N, c) -> (B, N, h*c)"""
    b, h, n, c = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, n, h * c)


def _rpe_scores(embed, q2):
    """(B,N,M,C) embedding x (B,N,C,h) projected queries -> (B,h,N,M): one pass over the embedding (`upk_rpe_scores`)
    on CUDA without autograd, torch.matmul + permute otherwise."""
    if embed.is_cuda and embed.dtype == torch.float32 and not (torch.is_grad_enabled() and (embed.requires_grad or q2.requires_grad)):
        from .. import _lib as L

        B, N, M, C = embed.shape
        h = q2.shape[3]
        e, q = embed.contiguous(), q2.contiguous().float()
        out = torch.empty((B, h, N, M), dtype=torch.float32, device=embed.device)
        with torch.cuda.device(embed.device):
            rc = L.load().upk_rpe_scores(L.ptr(e), L.ptr(q), B, N, M, C, h, L.ptr(out), L.stream_ptr(e))
        if rc == 0:
            return out
        if rc != -2:
            L.check(rc, "rpe_scores")
    return torch.matmul(embed, q2).permute(0, 3, 1, 2)


class _FeedForward(nn.Module):
    """expand -> ReLU -> squeeze -> residual LayerNorm (transformer.py:186-201, `AttentionOutput`)."""

    def __init__(self, d_model):
        super().__init__()
        self.expand = nn.Linear(d_model, d_model * 2)
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x):
        return self.norm(x + linear(self.squeeze, linear(self.expand, x, relu=True)))


class _DotAttention(nn.Module):
    """Scaled dot-product attention, optionally with a relative positional term q.(W_p e_qk)
    (transformer.py:94-151 `MultiHeadAttention`, :353-407 `RPEMultiHeadAttention`)."""

    def __init__(self, d_model, num_heads, relative):
        super().__init__()
        if d_model % num_heads:
            raise ValueError("`d_model` ({}) must be a multiple of `num_heads` ({}).".format(d_model, num_heads))
        self.num_heads = num_heads
        self.head_dim = d_model // num_heads
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        if relative:
            self.proj_p = nn.Linear(d_model, d_model)

    def forward(self, x_q, x_kv, embed_qk=None, key_masks=None):
        h = self.num_heads
        if x_q is x_kv:     # self-attention: one operand split and one GEMM for the three projections
            q, k, v = linear_multi((self.proj_q, self.proj_k, self.proj_v), x_q)
        else:
            q = linear(self.proj_q, x_q)
            k, v = linear_multi((self.proj_k, self.proj_v), x_kv)
        q, k, v = _heads(q, h), _heads(k, h), _heads(v, h)
        scores = q @ k.transpose(-1, -2)
        if embed_qk is not None:
            # q.(W_p e + b_p) = (W_p^T q).e + q.b_p: the reference projects the (B,N,M,C) embedding with W_p
            # (transformer.py:392,394: 2 N M C^2 FLOPs and a second (B,N,M,C) tensor per layer call); projecting the
            # (B,h,N,c) queries instead costs 2 N C^2 and one pass over the embedding (32x fewer FLOPs at h = 8, c = 32).
            c = self.head_dim
            wp = self.proj_p.weight.view(h, c, -1)                                  # (h, c, C)
            q2 = torch.einsum("bhnc,hck->bnkh", q, wp)                              # (B, N, C, h)
            qb = torch.einsum("bhnc,hc->bhn", q, self.proj_p.bias.view(h, c))
            scores = scores + _rpe_scores(embed_qk, q2) + qb.unsqueeze(-1)
        scores = scores / self.head_dim ** 0.5
        if key_masks is not None:
            scores = scores.masked_fill(key_masks[:, None, None, :], float("-inf"))
        attn = F.softmax(scores, dim=-1)
        return _merge(attn @ v), attn


class _AttentionBlock(nn.Module):
    """attention -> linear -> residual LayerNorm (transformer.py:154-183 / :410-439)."""

    def __init__(self, d_model, num_heads, relative):
        super().__init__()
        self.attention = _DotAttention(d_model, num_heads, relative)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x, memory, embed=None, memory_masks=None):
        h, attn = self.attention(x, memory, embed, memory_masks)
        return self.norm(linear(self.linear, h) + x), attn


class TransformerLayer(nn.Module):
    """transformer.py:204-229 (relative=False) / :442-467 `RPETransformerLayer` (relative=True)."""

    def __init__(self, d_model, num_heads, relative=False):
        super().__init__()
        self.attention = _AttentionBlock(d_model, num_heads, relative)
        self.output = _FeedForward(d_model)

    def forward(self, x, memory, embed=None, memory_masks=None):
        h, attn = self.attention(x, memory, embed, memory_masks)
        return self.output(h), attn


class GeometricTransformer(nn.Module):
    """Alternating self (with geometric RPE) / cross blocks over two token sets (transformer.py:470-515)."""

    def __init__(self, blocks, d_model, num_heads, dropout=None, activation_fn="ReLU", return_attention_scores=False,
                 parallel=False):
        super().__init__()
        if dropout:
            raise NotImplementedError("dropout is None in every UNOPose config")
        for b in blocks:
            if b not in ("self", "cross"):
                raise ValueError('Unsupported block type "{}".'.format(b))
        self.blocks = list(blocks)
        self.layers = nn.ModuleList([TransformerLayer(d_model, num_heads, relative=(b == "self")) for b in self.blocks])
        self.return_attention_scores = return_attention_scores
        self.parallel = parallel

    def forward(self, feats0, embeddings0, feats1, embeddings1, masks0=None, masks1=None):
        all_scores = []
        for layer, kind in zip(self.layers, self.blocks):
            if kind == "self":
                feats0, s0 = layer(feats0, feats0, embeddings0, masks0)
                feats1, s1 = layer(feats1, feats1, embeddings1, masks1)
            elif self.parallel:
                new0, s0 = layer(feats0, feats1, None, masks1)
                feats1, s1 = layer(feats1, feats0, None, masks0)
                feats0 = new0
            else:
                feats0, s0 = layer(feats0, feats1, None, masks1)
                feats1, s1 = layer(feats1, feats0, None, masks0)
            if self.return_attention_scores:
                all_scores.append([s0, s1])
        if self.return_attention_scores:
            return feats0, feats1, all_scores
        return feats0, feats1


class _FocusedLinearAttention(nn.Module):
    """Kernelised O(N d^2) attention of the dense tokens over the sparse ones (transformer.py:517-568)."""

    def __init__(self, d_model, num_heads, focusing_factor=3):
        super().__init__()
        if d_model % num_heads:
            raise ValueError("`d_model` ({}) must be a multiple of `num_heads` ({}).".format(d_model, num_heads))
        self.num_heads = num_heads
        self.focusing_factor = focusing_factor
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        self.scale = nn.Parameter(torch.zeros(size=(1, 1, d_model)))

    def _focus(self, x, scale):
        x = (F.relu(x) + 1e-6) / scale
        n = x.norm(dim=-1, keepdim=True)
        x = x ** self.focusing_factor
        return (x / x.norm(dim=-1, keepdim=True)) * n

    def forward(self, x_q, x_kv):
        h = self.num_heads
        scale = F.softplus(self.scale)
        kp, vp = linear_multi((self.proj_k, self.proj_v), x_kv)
        q = _heads(self._focus(linear(self.proj_q, x_q), scale), h)   # (B,h,i,c)
        k = _heads(self._focus(kp, scale), h)                         # (B,h,j,c)
        v = _heads(vp, h)                                             # (B,h,j,d)
        z = 1.0 / (torch.einsum("bhic,bhc->bhi", q, k.sum(dim=2)) + 1e-6)
        i, j, c, d = q.shape[2], k.shape[2], k.shape[3], v.shape[3]
        if i * j * (c + d) > c * d * (i + j):
            kv = torch.einsum("bhjc,bhjd->bhcd", k, v)
            x = torch.einsum("bhic,bhcd,bhi->bhid", q, kv, z)
        else:
            x = torch.einsum("bhij,bhjd,bhi->bhid", q @ k.transpose(-1, -2), v, z)
        return _merge(x)


class _LinearAttentionBlock(nn.Module):
    def __init__(self, d_model, num_heads, focusing_factor):
        super().__init__()
        self.attention = _FocusedLinearAttention(d_model, num_heads, focusing_factor)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x, memory):
        return self.norm(linear(self.linear, self.attention(x, memory)) + x)


class LinearTransformerLayer(nn.Module):
    """transformer.py:599-612."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn="ReLU", focusing_factor=3):
        super().__init__()
        self.attention = _LinearAttentionBlock(d_model, num_heads, focusing_factor)
        self.output = _FeedForward(d_model)

    def forward(self, x, memory):
        return self.output(self.attention(x, memory))


class SparseToDenseTransformer(nn.Module):
    """Geometric transformer on the FPS-sampled tokens, then linear attention dense <- sparse
    (transformer.py:615-671)."""

    def __init__(self, d_model, sparse_blocks, num_heads=4, dropout=None, activation_fn="ReLU", parallel=False,
                 focusing_factor=3, with_bg_token=True, replace_bg_token=True):
        super().__init__()
        self.with_bg_token = with_bg_token
        self.replace_bg_token = replace_bg_token
        self.sparse_layer = GeometricTransformer(sparse_blocks, d_model, num_heads, dropout=dropout,
                                                 activation_fn=activation_fn, parallel=parallel)
        self.dense_layer = LinearTransformerLayer(d_model, num_heads, focusing_factor=focusing_factor)

    def _sample_feats(self, dense_feats, fps_idx):
        # NOTE: the FPS indices are NOT shifted for the background token at row 0 — kept as is.
        feats = _gather_rows(dense_feats, fps_idx)
        if self.with_bg_token:
            feats = torch.cat([dense_feats[:, 0:1, :], feats], dim=1)
        return feats

    def _to_dense(self, dense_feats, feats):
        if self.with_bg_token and self.replace_bg_token:
            out = self.dense_layer(dense_feats[:, 1:, :].contiguous(), feats[:, 1:, :].contiguous())
            return torch.cat([feats[:, 0:1, :], out], dim=1)
        return self.dense_layer(dense_feats, feats)

    def forward(self, dense_feats0, embeddings0, fps_idx0, dense_feats1, embeddings1, fps_idx1, masks0=None, masks1=None):
        feats0 = self._sample_feats(dense_feats0, fps_idx0)
        feats1 = self._sample_feats(dense_feats1, fps_idx1)
        feats0, feats1 = self.sparse_layer(feats0, embeddings0, feats1, embeddings1, masks0, masks1)
        return self._to_dense(dense_feats0, feats0), self._to_dense(dense_feats1, feats1)


class _Sinusoid(nn.Module):
    """transformer.py:261-284; holds the `div_term` buffer."""

    def __init__(self, d_model):
        super().__init__()
        if d_model % 2:
            raise ValueError(f"Sinusoidal positional encoding with odd d_model: {d_model}")
        self.d_model = d_model
        self.register_buffer("div_term", torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model)))

    def forward(self, idx):
        w = idx.unsqueeze(-1) * self.div_term                      # (*, d/2)
        return torch.stack([torch.sin(w), torch.cos(w)], dim=-1).reshape(*idx.shape, self.d_model).detach()


class GeometricStructureEmbedding(nn.Module):
    """Pair-wise distance + triplet-wise angle embedding (transformer.py:287-350)."""

    def __init__(self, cfg):
        super().__init__()
        g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
        self.sigma_d, self.sigma_a, self.angle_k = g("sigma_d"), g("sigma_a"), g("angle_k")
        self.factor_a = 180.0 / (self.sigma_a * math.pi)
        self.embedding = _Sinusoid(g("hidden_dim"))
        self.proj_d = nn.Linear(g("hidden_dim"), g("hidden_dim"))
        self.proj_a = nn.Linear(g("hidden_dim"), g("hidden_dim"))
        self.reduction_a = g("reduction_a")
        if self.reduction_a not in ("max", "mean"):
            raise ValueError(f"Unsupported reduction mode: {self.reduction_a}.")

    def _fused(self, points):
        """CUDA inference goes to the fused kernels (modules/geo.py); CPU tensors, autograd and hidden sizes the
        kernel does not cover take the torch-op sequence below (cuBLAS on CUDA)."""
        if not points.is_cuda:
            return False
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return False
        from . import geo
        return geo.supported(self.proj_d.weight.shape[0], self.angle_k, points.shape[1])

    @torch.no_grad()
    def get_embedding_indices(self, points):
        if self._fused(points):
            from . import geo
            return geo.geometric_embedding_indices(points, self.sigma_d, self.factor_a, self.angle_k)
        B, N, _ = points.shape
        dist = torch.sqrt(pairwise_distance(points, points))
        knn = dist.topk(k=self.angle_k + 1, dim=2, largest=False)[1][:, :, 1:]               # (B,N,k)
        knn_pts = torch.gather(points.unsqueeze(1).expand(B, N, N, 3), 2, knn.unsqueeze(3).expand(B, N, self.angle_k, 3))
        ref = (knn_pts - points.unsqueeze(2)).unsqueeze(2).expand(B, N, N, self.angle_k, 3)   # neighbour directions
        anc = (points.unsqueeze(1) - points.unsqueeze(2)).unsqueeze(3).expand(B, N, N, self.angle_k, 3)
        sin = torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1)
        cos = torch.sum(ref * anc, dim=-1)
        return dist / self.sigma_d, torch.atan2(sin, cos) * self.factor_a

    def forward(self, points):
        if self._fused(points):
            from . import geo
            return geo.geometric_embedding(points, self.embedding.div_term, self.proj_d.weight, self.proj_d.bias,
                                           self.proj_a.weight, self.proj_a.bias, self.sigma_d, self.factor_a,
                                           self.angle_k, self.reduction_a)
        d_idx, a_idx = self.get_embedding_indices(points)
        d = self.proj_d(self.embedding(d_idx))
        a = self.proj_a(self.embedding(a_idx))
        a = a.max(dim=3)[0] if self.reduction_a == "max" else a.mean(dim=3)
        return d + a
