"""Mirror of jimm.models.siglip (reference: src/jimm/models/siglip.py)."""

from __future__ import annotations

import torch

from .. import _lib, nn
from ..common.transformer import g_wrap
from ..common import hf_loader as L
from ..common.utils import load_params_and_config
from ..common.vit import VisionTransformerBase, tower_config_fields
from ._dual import DualTower, build_text_tower


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
        hf, config = load_params_and_config(model_name_or_path, use_pytorch)

        def depth(tower, suffix):
            return max((int(k.split(".")[3]) + 1 for k in hf if k.startswith(f"{tower}.encoder.layers.") and k.endswith(suffix)), default=0)

        pw = hf["vision_model.embeddings.patch_embedding.weight"]
        vision_width, vision_patch = pw.shape[0], pw.shape[3]
        vocab_size, text_width = hf["text_model.embeddings.token_embedding.weight"].shape
        v_layers, t_layers = depth("vision_model", ".mlp.fc2.bias"), depth("text_model", ".self_attn.q_proj.weight")
        with nn.deferred_init():  # every parameter is replaced below
            model = cls(image_resolution=config["vision_config"]["image_size"], vision_layers=v_layers, vision_width=vision_width,
                        vision_patch_size=vision_patch, context_length=hf["text_model.embeddings.position_embedding.weight"].shape[0],
                        vocab_size=vocab_size, transformer_width=text_width, transformer_heads=text_width // 64, transformer_layers=t_layers,
                        mesh=mesh, dtype=dtype, param_dtype=dtype)
        v, mh = "vision_model.", "vision_model.MAPHead."
        rules = [
            ("logit_scale", "logit_scale", L.ASIS),                                      # (1,) -> ()          models/siglip.py:322-323
            ("logit_bias", "logit_bias", L.ASIS),
            ("positional_embedding", "text_model.embeddings.position_embedding.weight", L.ASIS),
            ("token_embedding.embedding", "text_model.embeddings.token_embedding.weight", L.ASIS),
            ("ln_final.scale", "text_model.final_layer_norm.weight", L.ASIS),
            ("ln_final.bias", "text_model.final_layer_norm.bias", L.ASIS),
            ("text_projection.kernel", "text_model.head.weight", L.LINEAR),
            ("text_projection.bias", "text_model.head.bias", L.ASIS),
            (v + "patch_embeddings.kernel", v + "embeddings.patch_embedding.weight", L.CONV),
            (v + "patch_embeddings.bias", v + "embeddings.patch_embedding.bias", L.ASIS),
            (v + "position_embeddings", v + "embeddings.position_embedding.weight", L.ASIS),  # (S,D) -> (1,S,D)    :320-321
            (v + "ln_post.scale", v + "post_layernorm.weight", L.ASIS),
            (v + "ln_post.bias", v + "post_layernorm.bias", L.ASIS),
            (mh + "probe", v + "head.probe", L.ASIS),
            (mh + "layernorm.scale", v + "head.layernorm.weight", L.ASIS),
            (mh + "layernorm.bias", v + "head.layernorm.bias", L.ASIS),
            (mh + "mlp.layers.0.kernel", v + "head.mlp.fc1.weight", L.LINEAR),
            (mh + "mlp.layers.0.bias", v + "head.mlp.fc1.bias", L.ASIS),
            (mh + "mlp.layers.2.kernel", v + "head.mlp.fc2.weight", L.LINEAR),
            (mh + "mlp.layers.2.bias", v + "head.mlp.fc2.bias", L.ASIS),
            (mh + "attn.out.kernel", v + "head.attention.out_proj.weight", L.OUT_W),
            (mh + "attn.out.bias", v + "head.attention.out_proj.bias", L.ASIS),
        ]
        # the MAP head's q/k/v live packed in one (3D, D) / (3D) pair: row blocks 0, 1, 2 (models/siglip.py:249-256,352-363)
        for i, role in enumerate(("query", "key", "value")):
            rules.append((mh + f"attn.{role}.kernel", v + "head.attention.in_proj_weight", L.QKV_W, (i, 3)))
            rules.append((mh + f"attn.{role}.bias", v + "head.attention.in_proj_bias", L.QKV_B, (i, 3)))
        for i in range(t_layers):
            rules += L.block_rules(f"text_model.blocks.layers.{i}.", f"text_model.encoder.layers.{i}.", L.CLIP_BLOCK)
        for i in range(v_layers):
            rules += L.block_rules(f"{v}transformer.blocks.layers.{i}.", f"{v}encoder.layers.{i}.", L.CLIP_BLOCK)
        L.apply_mapping(model, hf, rules, missing="strict", shape_error=ValueError, what="SigLIP ")
        return model
