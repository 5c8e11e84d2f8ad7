"""AttentionFusion with the reference's constructor, sub-module names and forward signature
(/root/reference/model/attention_fusion.py:99-154), executed by the CUDA library.

Supported configuration = what IMFNet instantiates (model/resunet.py:91-99): depth=0, cross_heads=1, mask=None.
Anything else raises instead of silently running a different path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib


class PreNorm(nn.Module):
    """Container with the reference's names: .fn, .norm, .norm_context (attention_fusion.py:32-37)."""

    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if context_dim is not None else None


class FeedForward(nn.Module):
    """.net.0 = Linear(dim, 8*dim), .net.1 = GEGLU (no parameters), .net.2 = Linear(4*dim, dim) (attention_fusion.py:53-60)."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), nn.Identity(), nn.Linear(dim * mult, dim))


class Attention(nn.Module):
    """.to_q, .to_kv (no bias), .to_out (attention_fusion.py:65-75)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, query_dim)


class AttentionFusion(nn.Module):
    def __init__(self, depth, dim, latent_dim=512, cross_heads=1, latent_heads=8, cross_dim_head=64, latent_dim_head=64,
                 weight_tie_layers=False):
        super().__init__()
        if depth != 0 or cross_heads != 1:
            raise NotImplementedError("imfnet_b200 implements the IMFNet configuration: depth=0, cross_heads=1")
        self.dim, self.latent_dim, self.inner = int(dim), int(latent_dim), int(cross_dim_head) * int(cross_heads)
        self.cross_attend_blocks = nn.ModuleList([
            PreNorm(latent_dim, Attention(latent_dim, dim, heads=cross_heads, dim_head=cross_dim_head), context_dim=dim),
            PreNorm(latent_dim, FeedForward(latent_dim)),
        ])
        self.layers = nn.ModuleList([])
        self._packed = None

    # -- weights as a C struct of device pointers ------------------------------------------------
    def packed(self):
        ca, ff = self.cross_attend_blocks
        tensors = [ca.norm.weight, ca.norm.bias, ca.norm_context.weight, ca.norm_context.bias, ca.fn.to_q.weight,
                   ca.fn.to_kv.weight, ca.fn.to_out.weight, ca.fn.to_out.bias, ff.norm.weight, ff.norm.bias,
                   ff.fn.net[0].weight, ff.fn.net[0].bias, ff.fn.net[2].weight, ff.fn.net[2].bias]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is None or self._packed[0] != key:
            keep = [t.detach().contiguous() for t in tensors]
            for t in keep:
                _lib.require_cuda(t, "attention-fusion weights")
                if t.dtype != torch.float32:
                    raise TypeError("attention-fusion weights must be float32")
            w = _lib.AttnWeights(*[t.data_ptr() for t in keep], self.latent_dim, self.dim, self.inner)
            self._packed = (key, w, keep)
        return self._packed[1]

    def packed_h2(self):
        """The module's matrices pre-packed for the fp16 hi/lo tensor-core GEMM (imf_h2_gemm): W^T as a one-offset "convolution kernel"
        [1, K, N] through imf_sparse_conv_h2_pack, scaled by a power of two into fp16 range; net.0 with its value / gate rows interleaved
        per 128-column tile so that the GEGLU product is an epilogue.  None when the shapes do not fit that kernel."""
        import math
        self.packed()
        key = self._packed[0]
        hit = getattr(self, "_packed_h2", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        ca, ff = self.cross_attend_blocks
        lat, inner, dim = self.latent_dim, self.inner, self.dim
        ok = inner == 128 and lat % 128 == 0 and dim % 64 == 0
        w = None
        keep = []
        if ok:
            L = _lib.lib()
            dev = ca.fn.to_q.weight.device

            def pk(Wt):          # Wt [N, K] (a Linear weight) -> packed slabs + scale
                N, K = Wt.shape
                W = Wt.detach().float().t().contiguous().reshape(1, K, N)
                wmax = float(W.abs().max())
                wmul = 2.0 ** math.floor(math.log2(2048.0 / wmax)) if wmax > 0 else 1.0
                buf = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, K, N, 64)), dtype=torch.uint8, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), 1, K, N, 64, wmul, buf.data_ptr(), _lib.cur_stream()))
                    torch.cuda.current_stream().synchronize()          # W is a temporary
                keep.append(buf)
                return buf.data_ptr(), wmul

            w1 = ff.fn.net[0].weight.detach()
            half = w1.shape[0] // 2
            w1i = torch.stack([w1[:half].reshape(half // 64, 64, -1), w1[half:].reshape(half // 64, 64, -1)], dim=1).reshape(2 * half, -1)
            (q, mq), (kv, mkv), (o, mo), (a1, m1), (a2, m2) = (pk(ca.fn.to_q.weight), pk(ca.fn.to_kv.weight), pk(ca.fn.to_out.weight), pk(w1i),
                                                                pk(ff.fn.net[2].weight))
            w = _lib.AttnPacked(q, kv, o, a1, a2, mq, mkv, mo, m1, m2)
        self._packed_h2 = (key, w, keep)
        return w

    def _apply(self, fn, *a, **k):
        self._packed = None
        self._packed_h2 = None
        return super()._apply(fn, *a, **k)

    # -- kernels ---------------------------------------------------------------------------------
    def project_context(self, tokens: torch.Tensor, channel_major: bool) -> torch.Tensor:
        """Projected context of one image (opaque kv buffer: K then V^T): tokens [dim, L] (channel_major) or [L, dim]."""
        L = _lib.lib()
        w = self.packed()
        tokens = tokens.contiguous()
        n_tok = tokens.shape[1] if channel_major else tokens.shape[0]
        kv = torch.empty(int(L.imf_attention_kv_bytes(n_tok, self.inner)), dtype=torch.uint8, device=tokens.device)
        kv.n_tokens = n_tok
        ws_bytes = int(L.imf_attention_kv_workspace_bytes(n_tok, self.dim))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=tokens.device)
        with torch.cuda.device(tokens.device):
            _lib.check(L.imf_attention_kv(w, _lib.ptr(tokens), n_tok, 1 if channel_major else 0, _lib.ptr(kv), _lib.ptr(ws),
                                          ws_bytes, _lib.cur_stream()))
        return kv

    def fuse(self, queries: torch.Tensor, kv: torch.Tensor, out: torch.Tensor | None = None, arena=None) -> torch.Tensor:
        """queries [M, latent] (row stride = queries.stride(0)) x kv (from project_context) -> [M, latent]."""
        L = _lib.lib()
        w = self.packed()
        M, n_tok = queries.shape[0], kv.n_tokens
        if queries.stride(1) != 1:
            queries = queries.contiguous()
        if out is None:
            out = torch.empty((M, self.latent_dim), dtype=torch.float32, device=queries.device)
        ws_bytes = int(L.imf_attention_workspace_bytes(M, n_tok, self.latent_dim, self.inner))
        ws = arena.take(ws_bytes) if arena is not None else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=queries.device)
        with torch.cuda.device(queries.device):
            _lib.check(L.imf_attention_fusion_fwd(w, _lib.ptr(queries), queries.stride(0) if M > 0 else self.latent_dim, M,
                                                  _lib.ptr(kv), n_tok, _lib.ptr(out), out.stride(0) if M > 0 else self.latent_dim,
                                                  _lib.ptr(ws), ws_bytes, _lib.cur_stream()))
        return out

    def forward(self, data, mask=None, queries_encoder=None):
        """data [B, L, dim] image tokens, queries_encoder [B, M, latent_dim] point tokens -> [B, M, latent_dim]."""
        if mask is not None:
            raise NotImplementedError("mask is not used by the IMFNet path")
        _lib.require_cuda(data, "data")
        if torch.is_grad_enabled() and (data.requires_grad or queries_encoder.requires_grad):
            raise NotImplementedError("imfnet_b200 implements the inference forward only (no autograd)")
        data, q = data.float(), queries_encoder.float()
        outs = [self.fuse(q[b], self.project_context(data[b], False)) for b in range(data.shape[0])]
        return torch.stack(outs, dim=0)
