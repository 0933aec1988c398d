"""Mirror of jimm.models.clip (reference: src/jimm/models/clip.py)."""

from __future__ import annotations

from typing import Any, Set

import torch

from .. import _lib, nn
from ..common.transformer import g_wrap
from ..common.utils import load_params_and_config
from ..common.vit import VisionTransformerBase, tower_config_fields
from ._dual import DualTower, build_text_tower, hf_block_mapping, transform_attn


class CLIP(DualTower):
    """models/clip.py:15-188."""

    def __init__(self, image_resolution: int, vision_layers: int, vision_width: int, vision_patch_size: int, context_length: int,
                 vocab_size: int, transformer_width: int, transformer_heads: int, transformer_layers: int, rngs=None,
                 dtype=torch.float32, param_dtype=torch.float32, mesh=None):
        self._init_common(image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
                          transformer_width, transformer_heads, transformer_layers, dtype)
        g = nn._gen(rngs)
        vision_heads = vision_width // 64  # models/clip.py:60
        object.__setattr__(self, "attn_mask", torch.tril(torch.ones(context_length, context_length)))  # :62
        # models/clip.py:64-81: pre-norm, no patch bias, QuickGELU, CLS, eps 1e-5
        self.add_child("vision_model", VisionTransformerBase(
            img_size=image_resolution, patch_size=vision_patch_size, in_channels=3, hidden_size=vision_width,
            num_layers=vision_layers, num_heads=vision_heads, mlp_dim=vision_width * 4, use_pre_norm=True, use_patch_bias=False,
            use_quick_gelu=True, pooling_type="CLS", layernorm_epsilon=1e-5, dtype=dtype, rngs=g_wrap(g)))
        self.add_child("visual_projection", nn.Linear(vision_width, transformer_width, use_bias=False, rngs=g_wrap(g)))
        build_text_tower(self, g, head_bias=False, layernorm_epsilon=1e-6, use_quick_gelu=True, attn_mask=self.attn_mask)
        self.add_param("logit_scale", nn.ones(()))

    def _native_config(self) -> _lib.Config:
        cfg = _lib.Config()
        cfg.kind = _lib.KIND_CLIP
        tower_config_fields(cfg, **self.vision_model._hp)
        cfg.num_classes = 0
        # text: causal tril mask (:62,:98), QuickGELU (:99), ln_final eps 1e-5 (:117), EOT = argmax(ids) pooling (:164), bias-free projection (:166)
        return self._text_config(cfg, act=_lib.ACT_QUICK_GELU, causal=1, pool=_lib.TPOOL_EOT_ARGMAX, head_bias=0, eps_outer=1e-5)

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, use_pytorch: bool = False, mesh=None, dtype=torch.float32) -> "CLIP":
        """Load a HF `CLIPModel` checkpoint (models/clip.py:190-416)."""
        params_fstate, config = load_params_and_config(model_name_or_path, use_pytorch)
        if config == {}:
            if not use_pytorch:
                tw = params_fstate["text_model.embeddings.token_embedding.weight"]
                text_hidden, text_vocab = tw.shape[1], tw.shape[0]
                text_ctx = params_fstate["text_model.embeddings.position_embedding.weight"].shape[0]
                text_layers = 0
                for k in params_fstate:
                    if k.startswith("text_model.encoder.layers.") and k.endswith(".self_attn.q_proj.weight"):
                        text_layers = max(text_layers, int(k.split(".")[3]) + 1)
                vis_hidden = params_fstate["vision_model.embeddings.class_embedding"].shape[0]
                vis_patch = params_fstate["vision_model.embeddings.patch_embedding.weight"].shape[2]
                vis_img = int((params_fstate["vision_model.embeddings.position_embedding.weight"].shape[0] - 1) ** 0.5) * vis_patch
                vis_layers = 0
                for k in params_fstate:
                    if k.startswith("vision_model.encoder.layers.") and k.endswith(".self_attn.q_proj.weight"):
                        vis_layers = max(vis_layers, int(k.split(".")[3]) + 1)
                config = {
                    "text_config": {"hidden_size": text_hidden, "num_attention_heads": text_hidden // 64, "num_hidden_layers": text_layers,
                                    "max_position_embeddings": text_ctx, "vocab_size": text_vocab},
                    "vision_config": {"hidden_size": vis_hidden, "num_attention_heads": vis_hidden // 64, "num_hidden_layers": vis_layers,
                                      "image_size": vis_img, "patch_size": vis_patch},
                }
            else:
                raise ValueError(f"Configuration could not be loaded for PyTorch model {model_name_or_path}")
        tc, vc = config["text_config"], config["vision_config"]
        with nn.deferred_init():  # every parameter is overwritten below (and asserted to be)
            model = cls(image_resolution=vc["image_size"], vision_layers=vc["num_hidden_layers"], vision_width=vc["hidden_size"],
                        vision_patch_size=vc["patch_size"], context_length=tc["max_position_embeddings"], vocab_size=tc["vocab_size"],
                        transformer_width=tc["hidden_size"], transformer_heads=tc["num_attention_heads"],
                        transformer_layers=tc["num_hidden_layers"], mesh=mesh, dtype=dtype, param_dtype=dtype)
        flax_params = model.flat_params()
        mapping = {
            "logit_scale": "logit_scale",
            "positional_embedding": "text_model.embeddings.position_embedding.weight",
            "token_embedding.embedding": "text_model.embeddings.token_embedding.weight",
            "ln_final.scale": "text_model.final_layer_norm.weight",
            "ln_final.bias": "text_model.final_layer_norm.bias",
            "text_projection.kernel": "text_projection.weight",
            "vision_model.cls_token": "vision_model.embeddings.class_embedding",
            "vision_model.position_embeddings": "vision_model.embeddings.position_embedding.weight",
            "vision_model.patch_embeddings.kernel": "vision_model.embeddings.patch_embedding.weight",
            "vision_model.ln_pre.scale": "vision_model.pre_layrnorm.weight",
            "vision_model.ln_pre.bias": "vision_model.pre_layrnorm.bias",
            "vision_model.ln_post.scale": "vision_model.post_layernorm.weight",
            "vision_model.ln_post.bias": "vision_model.post_layernorm.bias",
            "visual_projection.kernel": "visual_projection.weight",
        }
        for i in range(tc["num_hidden_layers"]):
            hf_block_mapping(mapping, f"text_model.blocks.layers.{i}.", f"text_model.encoder.layers.{i}.")
        for i in range(vc["num_hidden_layers"]):
            hf_block_mapping(mapping, f"vision_model.transformer.blocks.layers.{i}.", f"vision_model.encoder.layers.{i}.")

        nonvisited = set(flax_params.keys())
        used_hf_keys: Set[str] = set()
        for dst, src in mapping.items():
            if dst not in flax_params or src not in params_fstate:
                continue
            used_hf_keys.add(src)
            nonvisited.discard(dst)
            v = params_fstate[src].to(torch.float32)
            is_text = dst.startswith("text_model")
            hidden = tc["hidden_size"] if is_text else vc["hidden_size"]
            heads = tc["num_attention_heads"] if is_text else vc["hidden_size"] // 64
            if dst == "vision_model.patch_embeddings.kernel":
                v = v.permute(2, 3, 1, 0)
            elif dst == "vision_model.cls_token":
                v = v.reshape(1, 1, -1)
            elif dst == "vision_model.position_embeddings":
                v = v.reshape(1, v.shape[0], v.shape[1])
            elif ".self_attn." in src:
                v = transform_attn(v, src, hidden, heads)
            elif dst in ("token_embedding.embedding", "positional_embedding"):
                pass
            elif src.endswith("weight") and v.ndim == 2:
                v = v.T
            if tuple(v.shape) != tuple(flax_params[dst].shape):
                raise ValueError(f"Shape mismatch for {dst} (Flax) vs {src} (HF): {tuple(flax_params[dst].shape)} (expected) != {tuple(v.shape)} (actual)")
            model.set_flat_param(dst, v)
        assert len(nonvisited) == 0, f"Some Flax CLIP model parameters were not visited: {sorted(list(nonvisited))}"
        leftover = set(params_fstate.keys()) - used_hf_keys
        known_unused = {"text_model.embeddings.position_ids", "vision_model.embeddings.position_ids"}
        unexpected = leftover - known_unused
        assert len(unexpected) == 0, f"Some unexpected HuggingFace checkpoint parameters were not used: {sorted(list(unexpected))}"
        return model
