"""Mirror of jimm.models.siglip (reference: src/jimm/models/siglip.py)."""

from __future__ import annotations

from typing import Set

import torch

from .. import _lib, nn
from ..common.transformer import g_wrap
from ..common.utils import load_params_and_config
from ..common.vit import VisionTransformerBase, tower_config_fields
from ._dual import DualTower, build_text_tower, hf_block_mapping, transform_attn


class SigLIP(DualTower):
    """models/siglip.py:15-174."""

    def __init__(self, image_resolution: int, vision_layers: int, vision_width: int, vision_patch_size: int, context_length: int,
                 vocab_size: int, transformer_width: int, transformer_heads: int, transformer_layers: int, rngs=None,
                 dtype=torch.float32, param_dtype=torch.float32, mesh=None):
        self._init_common(image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
                          transformer_width, transformer_heads, transformer_layers, dtype)
        g = nn._gen(rngs)
        object.__setattr__(self, "vision_heads", vision_width // 64)  # models/siglip.py:59
        # models/siglip.py:60-78: MAP pooling, patch bias, tanh-GELU, eps 1e-6
        self.add_child("vision_model", VisionTransformerBase(
            img_size=image_resolution, patch_size=vision_patch_size, in_channels=3, hidden_size=vision_width,
            num_layers=vision_layers, num_heads=self.vision_heads, mlp_dim=vision_width * 4, use_pre_norm=False,
            use_patch_bias=True, use_quick_gelu=False, pooling_type="MAP", layernorm_epsilon=1e-6, dtype=dtype, rngs=g_wrap(g)))
        build_text_tower(self, g, head_bias=True, layernorm_epsilon=1e-6, use_quick_gelu=False, attn_mask=None)
        self.add_param("logit_scale", nn.ones(()))
        self.add_param("logit_bias", nn.ones(()))

    def _native_config(self) -> _lib.Config:
        cfg = _lib.Config()
        cfg.kind = _lib.KIND_SIGLIP
        tower_config_fields(cfg, **self.vision_model._hp)
        cfg.num_classes = 0
        # text: no mask, tanh-GELU, ln_final eps 1e-6 (:104), last-token pooling (:151), Linear head with bias (:111-119,152)
        return self._text_config(cfg, act=_lib.ACT_GELU_TANH, causal=0, pool=_lib.TPOOL_LAST, head_bias=1, eps_outer=1e-6)

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, use_pytorch: bool = False, mesh=None, dtype=torch.float32) -> "SigLIP":
        """Load a HF `SiglipModel` checkpoint (models/siglip.py:176-385): shapes always inferred from the tensors, except
        `image_size`, which must come from config["vision_config"] (:210)."""
        params_fstate, config = load_params_and_config(model_name_or_path, use_pytorch)
        vision_patch_size = params_fstate["vision_model.embeddings.patch_embedding.weight"].shape[3]
        vision_width = params_fstate["vision_model.embeddings.patch_embedding.bias"].shape[0]
        vision_num_layers = 0
        for k in params_fstate:
            if k.startswith("vision_model.encoder.layers.") and k.endswith(".mlp.fc2.bias"):
                vision_num_layers = max(vision_num_layers, int(k.split(".")[3]) + 1)
        context_length = params_fstate["text_model.embeddings.position_embedding.weight"].shape[0]
        vocab_size, text_hidden = params_fstate["text_model.embeddings.token_embedding.weight"].shape
        text_num_layers = 0
        for k in params_fstate:
            if k.startswith("text_model.encoder.layers.") and k.endswith(".self_attn.q_proj.weight"):
                text_num_layers = max(text_num_layers, int(k.split(".")[3]) + 1)
        with nn.deferred_init():  # every parameter is overwritten below (and asserted to be)
            model = cls(image_resolution=config["vision_config"]["image_size"], vision_layers=vision_num_layers, vision_width=vision_width,
                        vision_patch_size=vision_patch_size, context_length=context_length, vocab_size=vocab_size,
                        transformer_width=text_hidden, transformer_heads=text_hidden // 64, transformer_layers=text_num_layers, mesh=mesh,
                        dtype=dtype, param_dtype=dtype)
        flax_params = model.flat_params()
        v_, m_ = "vision_model.", "vision_model.MAPHead."
        mapping = {
            "logit_scale": "logit_scale",
            "logit_bias": "logit_bias",
            "positional_embedding": "text_model.embeddings.position_embedding.weight",
            "token_embedding.embedding": "text_model.embeddings.token_embedding.weight",
            "ln_final.scale": "text_model.final_layer_norm.weight",
            "ln_final.bias": "text_model.final_layer_norm.bias",
            "text_projection.kernel": "text_model.head.weight",
            "text_projection.bias": "text_model.head.bias",
            v_ + "patch_embeddings.kernel": v_ + "embeddings.patch_embedding.weight",
            v_ + "patch_embeddings.bias": v_ + "embeddings.patch_embedding.bias",
            v_ + "position_embeddings": v_ + "embeddings.position_embedding.weight",
            v_ + "ln_post.scale": v_ + "post_layernorm.weight",
            v_ + "ln_post.bias": v_ + "post_layernorm.bias",
            m_ + "probe": v_ + "head.probe",
            m_ + "layernorm.scale": v_ + "head.layernorm.weight",
            m_ + "layernorm.bias": v_ + "head.layernorm.bias",
            m_ + "mlp.layers.0.kernel": v_ + "head.mlp.fc1.weight",
            m_ + "mlp.layers.0.bias": v_ + "head.mlp.fc1.bias",
            m_ + "mlp.layers.2.kernel": v_ + "head.mlp.fc2.weight",
            m_ + "mlp.layers.2.bias": v_ + "head.mlp.fc2.bias",
            m_ + "attn.out.kernel": v_ + "head.attention.out_proj.weight",
            m_ + "attn.out.bias": v_ + "head.attention.out_proj.bias",
        }
        for y in ("query", "key", "value"):
            mapping[m_ + f"attn.{y}.kernel"] = v_ + "head.attention.in_proj_weight"
            mapping[m_ + f"attn.{y}.bias"] = v_ + "head.attention.in_proj_bias"
        for i in range(text_num_layers):
            hf_block_mapping(mapping, f"text_model.blocks.layers.{i}.", f"text_model.encoder.layers.{i}.")
        for i in range(vision_num_layers):
            hf_block_mapping(mapping, f"vision_model.transformer.blocks.layers.{i}.", f"vision_model.encoder.layers.{i}.")

        nonvisited = set(flax_params.keys())
        used_hf_keys: Set[str] = set()
        Hv, dv = model.vision_heads, vision_width // model.vision_heads
        for dst, src in mapping.items():
            nonvisited.discard(dst)
            used_hf_keys.add(src)
            v = params_fstate[src].to(torch.float32)
            is_text = "text_model" in src
            hidden = model.transformer_width if is_text else vision_width
            heads = model.transformer_heads if is_text else Hv
            if dst == v_ + "patch_embeddings.kernel":
                v = v.permute(2, 3, 1, 0)
            elif dst == v_ + "position_embeddings":
                v = v.reshape(1, v.shape[0], v.shape[1])
            elif dst in ("logit_scale", "logit_bias"):
                v = v.squeeze()
            elif ".self_attn." in src or src.endswith("out_proj.weight"):
                v = transform_attn(v, src, hidden, heads)
            elif src.endswith("in_proj_weight"):
                q_w, k_w, v_w = torch.chunk(v, 3, dim=0)  # :352-357
                v = {"query": q_w, "key": k_w, "value": v_w}[dst.split(".")[-2]].T.reshape(vision_width, Hv, dv)
            elif src.endswith("in_proj_bias"):
                q_b, k_b, v_b = torch.chunk(v, 3, dim=0)  # :358-363
                v = {"query": q_b, "key": k_b, "value": v_b}[dst.split(".")[-2]].reshape(Hv, dv)
            elif src.endswith("weight") and v.ndim == 2:
                if "position_embedding" not in src and "token_embedding" not in src:
                    v = v.T
            if tuple(v.shape) != tuple(flax_params[dst].shape):
                raise ValueError(f"Shape mismatch for {dst} (Flax) vs {src} (HF): {tuple(flax_params[dst].shape)} (expected) != {tuple(v.shape)} (actual)")
            model.set_flat_param(dst, v)
        leftover = set(params_fstate.keys()) - used_hf_keys
        known_unused = {"text_model.embeddings.position_ids", "vision_model.embeddings.position_ids"}
        unexpected = leftover - known_unused
        assert len(unexpected) == 0, f"Some unexpected HuggingFace checkpoint parameters were not used: {sorted(list(unexpected))}"
        return model
