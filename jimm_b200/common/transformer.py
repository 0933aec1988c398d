"""Mirror of jimm.common.transformer (reference: src/jimm/common/transformer.py): same class names and constructor
kwargs.  These are parameter-tree nodes; the arithmetic of a block runs inside the CUDA library
(LayerNorm -> fused QKV tcgen05 GEMM -> flash attention -> out-proj GEMM + residual -> LayerNorm -> FC1 GEMM + GELU ->
FC2 GEMM + residual), driven per tower by VisionTransformerBase / CLIP / SigLIP."""

from __future__ import annotations

import torch

from .. import nn


def quickgelu(x: torch.Tensor) -> torch.Tensor:
    """x * sigmoid(1.702 x) (common/transformer.py:12-19) -- provided for API parity; the fused kernel version is the
    FC1 epilogue."""
    return x * torch.sigmoid(1.702 * x)


class TransformerEncoder(nn.Module):
    """common/transformer.py:22-132."""

    def __init__(self, hidden_size: int, mlp_dim: int, num_heads: int, layernorm_epsilon=1e-5, dropout_rate: float = 0.0,
                 attn_mask=None, use_quick_gelu: bool = False, dtype=None, param_dtype=None, rngs=None, mesh=None) -> None:
        super().__init__()
        object.__setattr__(self, "attn_mask", attn_mask)
        object.__setattr__(self, "layernorm_epsilon", layernorm_epsilon)
        object.__setattr__(self, "use_quick_gelu", use_quick_gelu)
        g = nn._gen(rngs)
        self.add_child("norm1", nn.LayerNorm(hidden_size, layernorm_epsilon))
        self.add_child("attn", nn.MultiHeadAttention(num_heads, hidden_size, rngs=g_wrap(g)))
        self.add_child("norm2", nn.LayerNorm(hidden_size, layernorm_epsilon))
        # nnx.Sequential [Linear, act, Dropout, Linear, Dropout] -> param indices 0 and 3 (:92-114)
        self.add_child("mlp", nn.Sequential(nn.Linear(hidden_size, mlp_dim, rngs=g_wrap(g)), None, None,
                                            nn.Linear(mlp_dim, hidden_size, rngs=g_wrap(g)), None))

    def __call__(self, x):
        raise NotImplementedError("TransformerEncoder runs inside the CUDA library as part of a tower forward; call the owning "
                                  "VisionTransformerBase / VisionTransformer / CLIP / SigLIP instead")


class _G(nn.Rngs):
    def __init__(self, g):
        self.seed = 0
        self._gen = g


def g_wrap(g):
    return _G(g)


class Transformer(nn.Module):
    """common/transformer.py:135-196.  NOTE the default layernorm_epsilon=1e-6 (:142) is what every tower gets, because
    VisionTransformerBase and CLIP never forward their own epsilon."""

    def __init__(self, width: int, mlp_dim: int, layers: int, num_heads: int, layernorm_epsilon=1e-6, dropout_rate: float = 0.0,
                 attn_mask=None, use_quick_gelu: bool = False, dtype=None, param_dtype=None, rngs=None, mesh=None):
        super().__init__()
        for k, v in dict(width=width, layers=layers, num_heads=num_heads, dropout_rate=dropout_rate, mlp_dim=mlp_dim,
                         layernorm_epsilon=layernorm_epsilon, use_quick_gelu=use_quick_gelu, attn_mask=attn_mask).items():
            object.__setattr__(self, k, v)
        g = nn._gen(rngs)
        blocks = nn.Module()
        blocks.add_child("layers", nn.ModuleList([
            TransformerEncoder(width, mlp_dim, num_heads, layernorm_epsilon, dropout_rate, attn_mask, use_quick_gelu, dtype,
                               param_dtype, g_wrap(g), mesh) for _ in range(layers)]))
        self.add_child("blocks", blocks)

    def __call__(self, x):
        raise NotImplementedError("Transformer runs inside the CUDA library as part of a tower forward")
