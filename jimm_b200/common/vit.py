"""Mirror of jimm.common.vit (reference: src/jimm/common/vit.py): VisionTransformerBase and
MultiHeadAttentionPoolingHead with the reference's constructor kwargs and parameter tree.  `__call__` runs the whole
tower in the CUDA library (patchify + tcgen05 GEMMs + flash attention + LayerNorm + CLS|MAP pooling)."""

from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _lib, nn
from .._runtime import NativeModel, default_max_batch
from .transformer import Transformer, _SubModuleRunner, g_wrap


class MultiHeadAttentionPoolingHead(_SubModuleRunner, nn.Module):
    """common/vit.py:12-101.  Inside a tower it is evaluated as part of the tower forward (probe query precomputed at finalize, k/v
    projection GEMM over all tokens, single-query attention, LN + MLP + residual); called on its own it runs the same kernels through
    jimm_map_head_forward."""

    def __init__(self, hidden_size: int, intermediate_size: int, num_heads: int, layernorm_epsilon: float = 1e-6, rngs=None,
                 dtype=None, param_dtype=None, mesh=None):
        nn.Module.__init__(self)
        self._sub_init(dtype)
        if intermediate_size != 4 * hidden_size:
            raise ValueError("the MAP head kernels take intermediate_size == 4 * hidden_size (the only value the reference uses, common/vit.py:175)")
        g = nn._gen(rngs)
        object.__setattr__(self, "layernorm_epsilon", layernorm_epsilon)
        object.__setattr__(self, "_dims", (hidden_size, num_heads))
        self.add_param("probe", nn.zeros((1, 1, hidden_size)))
        self.add_child("attn", nn.MultiHeadAttention(num_heads, hidden_size, rngs=g_wrap(g)))
        self.add_child("layernorm", nn.LayerNorm(hidden_size, layernorm_epsilon))
        # nnx.Sequential [Linear, gelu, Linear] -> param indices 0 and 2 (:65-85)
        self.add_child("mlp", nn.Sequential(nn.Linear(hidden_size, intermediate_size, rngs=g_wrap(g)), None,
                                            nn.Linear(intermediate_size, hidden_size, rngs=g_wrap(g))))

    def _sub_config(self, max_seq):
        D, H = self._dims
        cfg = _lib.Config()
        cfg.kind = _lib.KIND_MAPHEAD
        cfg.v_width, cfg.v_heads, cfg.v_mlp, cfg.v_layers = D, H, 4 * D, 0
        cfg.v_eps_outer = cfg.v_eps_block = float(self.layernorm_epsilon)
        cfg.ctx_len = int(max_seq)
        cfg.compute_dtype = self._sub_dtype
        return cfg

    def __call__(self, hidden_state):
        """[batch, seq, hidden] -> [batch, hidden] (common/vit.py:87-101)."""
        return self._run(hidden_state)


class _NativeOwner:
    """Mixin for top-level runnable models: lazily builds (and rebuilds when parameters or the batch bound change) the
    native handle from the flat parameter tree."""

    def _native_init(self, dtype):
        object.__setattr__(self, "_native", None)
        object.__setattr__(self, "_preproc", None)
        object.__setattr__(self, "_compute_dtype", nn.compute_dtype_code(dtype))
        object.__setattr__(self, "_max_batch", default_max_batch())

    @staticmethod
    def _release_native(n):
        """Free a native handle.  If it owns an NVLink gather buffer, the peers have it mapped through CUDA IPC and may still be storing
        into it: every rank rebuilds at the same point of the program (same batch on every rank), so all of them drain their streams
        and meet at a barrier before anybody frees."""
        if getattr(n, "_comm", None) is not None:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                torch.cuda.synchronize(n.device)
                dist.barrier()
        n.close()

    def _invalidate(self):
        n = getattr(self, "_native", None)
        if n is not None:
            self._release_native(n)
        object.__setattr__(self, "_native", None)

    def _native_config(self) -> _lib.Config:
        raise NotImplementedError

    def native(self, batch: int = 1, require: bool = False) -> NativeModel:
        """Native handle whose workspace holds `max_batch` samples per call; larger vision / text batches are chunked by
        the library, the contrastive head needs the whole batch resident (`require=True`)."""
        n = self._native
        if n is not None and (not require or batch <= n.max_batch):
            return n
        if n is not None:
            self._release_native(n)
        mb = max(self._max_batch, int(batch) if require else 1)
        n = NativeModel(self._native_config(), self.flat_params(raw=True), mb)
        n.preproc = self._preproc
        object.__setattr__(self, "_native", n)
        return n

    def set_preprocessor(self, preprocessor):
        """Attach a `jimm_b200.preprocess.ImagePreprocessor`: the model then also accepts raw uint8 RGB frames [B,H,W,3] (host or
        CUDA) -- examples/vit_inference.py:27-37's `processor(images=...)` + transpose runs on the GPU, and host batches cross PCIe
        as bytes instead of fp32 pixel values."""
        object.__setattr__(self, "_preproc", preprocessor)
        if self._native is not None:
            self._native.preproc = preprocessor
        return self

    def set_max_batch(self, max_batch: int):
        """Bound of samples per native call (workspace is sized for it at finalize)."""
        object.__setattr__(self, "_max_batch", int(max_batch))
        self._invalidate()
        return self


def tower_config_fields(cfg: _lib.Config, *, img_size, patch_size, in_channels, hidden_size, num_layers, num_heads, mlp_dim,
                        pooling_type, use_quick_gelu, use_pre_norm, use_patch_bias, layernorm_epsilon):
    cfg.img_size, cfg.patch, cfg.in_ch = img_size, patch_size, in_channels
    cfg.v_width, cfg.v_layers, cfg.v_heads, cfg.v_mlp = hidden_size, num_layers, num_heads, mlp_dim
    cfg.pooling = _lib.POOL_CLS if pooling_type == "CLS" else _lib.POOL_MAP
    cfg.pre_norm, cfg.patch_bias = int(use_pre_norm), int(use_patch_bias)
    cfg.v_act = _lib.ACT_QUICK_GELU if use_quick_gelu else _lib.ACT_GELU_TANH
    cfg.v_eps_outer = layernorm_epsilon
    cfg.v_eps_block = 1e-6  # Transformer default; VisionTransformerBase never forwards its epsilon (common/vit.py:193-204)
    return cfg


class VisionTransformerBase(_NativeOwner, nn.Module):
    """common/vit.py:104-248."""

    def __init__(self, img_size: int, patch_size: int, in_channels: int, hidden_size: int, num_layers: int, num_heads: int,
                 mlp_dim: int, pooling_type: str = "CLS", dropout_rate: float = 0.0, use_quick_gelu: bool = False,
                 use_pre_norm: bool = False, use_patch_bias: bool = True, layernorm_epsilon: float = 1e-5, rngs=None, dtype=None,
                 param_dtype=None, mesh=None):
        nn.Module.__init__(self)
        self._native_init(dtype)
        g = nn._gen(rngs)
        n_patches = (img_size // patch_size) ** 2
        hp = dict(img_size=img_size, patch_size=patch_size, in_channels=in_channels, hidden_size=hidden_size, num_layers=num_layers,
                  num_heads=num_heads, mlp_dim=mlp_dim, pooling_type=pooling_type, use_quick_gelu=use_quick_gelu,
                  use_pre_norm=use_pre_norm, use_patch_bias=use_patch_bias, layernorm_epsilon=layernorm_epsilon)
        object.__setattr__(self, "_hp", hp)
        object.__setattr__(self, "use_pre_norm", use_pre_norm)
        object.__setattr__(self, "pooling_type", pooling_type)
        object.__setattr__(self, "dropout_rate", dropout_rate)
        self.add_child("patch_embeddings", nn.Conv(in_channels, hidden_size, (patch_size, patch_size), use_patch_bias, g_wrap(g)))
        if pooling_type == "CLS":
            self.add_param("cls_token", nn.zeros((1, 1, hidden_size)))
            pos = nn.truncated_normal(g, (1, n_patches + 1, hidden_size), 0.02)
        elif pooling_type == "MAP":
            pos = nn.truncated_normal(g, (1, n_patches, hidden_size), 0.02)
            self.add_child("MAPHead", MultiHeadAttentionPoolingHead(hidden_size, 4 * hidden_size, num_heads, layernorm_epsilon,
                                                                      rngs=g_wrap(g)))
        else:
            raise ValueError("pooling_type must be either MAP or CLS.")  # common/vit.py:178
        self.add_param("position_embeddings", pos)
        if use_pre_norm:
            self.add_child("ln_pre", nn.LayerNorm(hidden_size, layernorm_epsilon))
        self.add_child("transformer", Transformer(width=hidden_size, mlp_dim=mlp_dim, layers=num_layers, num_heads=num_heads,
                                                  dropout_rate=dropout_rate, use_quick_gelu=use_quick_gelu, rngs=g_wrap(g)))
        self.add_child("ln_post", nn.LayerNorm(hidden_size, layernorm_epsilon))

    def _native_config(self) -> _lib.Config:
        cfg = _lib.Config()
        cfg.kind = _lib.KIND_TOWER
        tower_config_fields(cfg, **self._hp)
        cfg.num_classes = 0
        cfg.compute_dtype = self._compute_dtype
        return cfg

    def __call__(self, img) -> torch.Tensor:
        """[batch, height, width, channels] -> [batch, hidden_size] (CLS token or MAP head output)."""
        B = img.shape[0]
        return self.native(B).vision(img)

    def forward_async(self, img):
        """Asynchronous dispatch for host inputs (the reference's calls return before the device finishes, examples/vit_inference.py:54
        only blocks when it reads the logits): returns a `PendingResult`; back-to-back calls overlap their copies with compute."""
        return self.native(img.shape[0]).vision_async(img)
