"""B200-native mirrors of the two AF2-style modules the denoiser uses.

Same class names, constructor arguments, parameter names/shapes and call signatures as the
reference's ``ProteinReDiff/models/AF2_modules.py`` (SPAttention :369-473, OuterProductUpdate
:476-545, Attention :189-367, LayerNorm :161-182, Linear :94-159), so reference state-dicts load
with ``strict=True``.  The forward passes enqueue hand-written sm_100a kernels through the C ABI
(``prd_spattention_fwd``, ``prd_opm_project_fwd`` + ``prd_pair_embed_fwd``); there is no
PyTorch/CPU compute path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import ops
from .._packing import PackCache, half, split_k


def _trunc_normal_(w: torch.Tensor, scale: float) -> None:
    fan_in = w.shape[1]
    std = math.sqrt(scale / max(1, fan_in)) / 0.87962566103423978  # std of N(0,1) truncated to [-2, 2]
    nn.init.trunc_normal_(w, 0.0, std, -2.0 * std, 2.0 * std)


class Linear(nn.Linear):
    """nn.Linear with AlphaFold's named initialisers (reference AF2_modules.py:94-159)."""

    def __init__(self, in_dim: int, out_dim: int, bias: bool = True, init: str = "default", init_fn=None):
        super().__init__(in_dim, out_dim, bias=bias)
        with torch.no_grad():
            if bias:
                self.bias.zero_()
            if init_fn is not None:
                init_fn(self.weight, self.bias)
            elif init == "default":
                _trunc_normal_(self.weight, 1.0)
            elif init == "relu":
                _trunc_normal_(self.weight, 2.0)
            elif init == "glorot":
                nn.init.xavier_uniform_(self.weight, gain=1)
            elif init == "gating":
                self.weight.zero_()
                if bias:
                    self.bias.fill_(1.0)
            elif init == "normal":
                nn.init.kaiming_normal_(self.weight, nonlinearity="linear")
            elif init == "final":
                self.weight.zero_()
            else:
                raise ValueError("Invalid init string.")


class LayerNorm(nn.Module):
    """Affine LayerNorm parameter holder (reference AF2_modules.py:161-182); applied inside the fused ops."""

    def __init__(self, c_in: int, eps: float = 1e-5):
        super().__init__()
        self.c_in = (c_in,)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(c_in))
        self.bias = nn.Parameter(torch.zeros(c_in))


class Attention(nn.Module):
    """Parameter container for SPAttention.mha (reference AF2_modules.py:189-249)."""

    def __init__(self, c_q: int, c_k: int, c_v: int, c_hidden: int, no_heads: int, gating: bool = True):
        super().__init__()
        self.c_q, self.c_k, self.c_v = c_q, c_k, c_v
        self.c_hidden, self.no_heads, self.gating = c_hidden, no_heads, gating
        self.linear_q = Linear(c_q, c_hidden * no_heads, bias=False, init="glorot")
        self.linear_k = Linear(c_k, c_hidden * no_heads, bias=False, init="glorot")
        self.linear_v = Linear(c_v, c_hidden * no_heads, bias=False, init="glorot")
        self.linear_o = Linear(c_hidden * no_heads, c_q, init="final")
        self.linear_g = Linear(c_q, c_hidden * no_heads, init="gating") if gating else None


class SPAttention(nn.Module):
    """Single-representation attention with pair bias (reference AF2_modules.py:369-473).

    ``forward(m, z, mask)``: m [B, N, c_in], z [B, N, N, c_z] -> LN(m) + mha(LN(m), bias(z)).
    As in the reference the mask is accepted but not used (its mask_bias is dead code).
    """

    def __init__(self, c_in, c_hidden, no_heads, pair_bias=False, c_z=None, inf=1e9):
        super().__init__()
        if not pair_bias or c_z is None:
            raise ValueError("the B200 SPAttention is built for pair_bias=True with c_z given")
        if c_hidden != c_in:
            raise ValueError("the B200 SPAttention is built for c_hidden == c_in (modules.py:366-371)")
        self.c_in, self.c_hidden, self.no_heads = c_in, c_hidden, no_heads
        self.pair_bias, self.c_z, self.inf = pair_bias, c_z, inf
        self.layer_norm_m = LayerNorm(c_in)
        self.linear_z = nn.Sequential(LayerNorm(c_z), Linear(c_z, no_heads, bias=False, init="normal"))
        self.mha = Attention(c_in, c_in, c_in, c_hidden, no_heads)
        self._pack = PackCache()

    def packed_weights(self):
        srcs = [self.layer_norm_m.weight, self.layer_norm_m.bias, self.linear_z[0].weight, self.linear_z[0].bias,
                self.linear_z[1].weight, self.mha.linear_q.weight, self.mha.linear_k.weight, self.mha.linear_v.weight,
                self.mha.linear_g.weight, self.mha.linear_g.bias, self.mha.linear_o.weight, self.mha.linear_o.bias]

        def build():
            f = lambda t: t.detach().float().contiguous()
            return [f(srcs[0]), f(srcs[1]), f(srcs[2]), f(srcs[3]), f(srcs[4]), split_k(srcs[5]), split_k(srcs[6]),
                    split_k(srcs[7]), split_k(srcs[8]), f(srcs[9]), split_k(srcs[10]), f(srcs[11])]

        return self._pack.get(srcs, build)

    def bias_projection(self):
        """(ln_weight, ln_bias, w_z, None) of linear_z for ops.pair_bias."""
        w = self.packed_weights()
        return (w[2], w[3], w[4], None)

    def forward(self, m: torch.Tensor, z: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                cfg=None, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None):
        if z is None:
            raise ValueError("pair embedding z is required")
        cfg = cfg if cfg is not None else _MiniCfg(self.c_in, self.c_z, self.no_heads)
        return ops.spattention(cfg, m.contiguous(), z.contiguous(), self.packed_weights(), out=out, bias=bias)


class OuterProductUpdate(nn.Module):
    """Outer-product update (reference AF2_modules.py:476-545); no contracted index.

    ``forward(m, mask)`` returns ``linear_out(a_i * b_j) / (mask_i mask_j + eps)`` as [B, N, N, c_z].
    """

    def __init__(self, c_m, c_z, c_hidden, eps=1e-3):
        super().__init__()
        if abs(eps - 1e-3) > 0:
            raise ValueError("the B200 OuterProductUpdate is built for eps=1e-3")
        if c_hidden * 4 != c_m:
            raise ValueError("the B200 OuterProductUpdate is built for c_hidden == c_m // 4 (modules.py:372-374)")
        self.c_m, self.c_z, self.c_hidden, self.eps = c_m, c_z, c_hidden, eps
        self.layer_norm = nn.LayerNorm(c_m)
        self.linear_1 = Linear(c_m, c_hidden)
        self.linear_2 = Linear(c_m, c_hidden)
        self.linear_out = Linear(c_hidden, c_z, init="final")
        self._pack = PackCache()

    def packed_weights(self):
        srcs = [self.layer_norm.weight, self.layer_norm.bias, self.linear_1.weight, self.linear_1.bias,
                self.linear_2.weight, self.linear_2.bias, self.linear_out.weight, self.linear_out.bias]

        def build():
            f = lambda t: t.detach().float().contiguous()
            proj = [f(srcs[0]), f(srcs[1]), split_k(srcs[2]), f(srcs[3]), split_k(srcs[4]), f(srcs[5])]
            out = [half(srcs[6]), f(srcs[7])]
            return proj, out

        return self._pack.get(srcs, build)

    def project(self, cfg, m, mask, out_a=None, out_b=None):
        proj, _ = self.packed_weights()
        return ops.opm_project(cfg, m, mask, proj, out_a, out_b)

    def forward(self, m: torch.Tensor, mask: Optional[torch.Tensor] = None, cfg=None):
        cfg = cfg if cfg is not None else _MiniCfg(self.c_m, self.c_z, 4)
        m = m.contiguous()
        if mask is None:
            mask = m.new_ones(m.shape[:-1])
        mask = mask.contiguous()
        a, b = self.project(cfg, m, mask)
        _, (w_o, b_o) = self.packed_weights()
        B, N, _ = m.shape
        out = torch.empty(B, N, N, self.c_z, dtype=torch.float32, device=m.device)
        z = torch.zeros(B, N, 3, dtype=torch.float32, device=m.device)
        ops.pair_embed(cfg, None, z, mask, None, a, b, [None, None, None, None, w_o, b_o], out, flags=3)
        return out


class _MiniCfg:
    """Enough of the hyper-parameter set to size a stand-alone module call."""

    def __init__(self, single_dim, pair_dim, num_heads, head_dim=16, transition_factor=4):
        self.single_dim, self.pair_dim, self.num_heads, self.head_dim = single_dim, pair_dim, num_heads, head_dim
        self.transition_factor = transition_factor
        self.esm_dim, self.time_dim, self.dist_dim = 1280, 256, 256
        self.max_bond_distance, self.max_relpos, self.num_steps = 7, 32, 64
