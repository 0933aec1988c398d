"""Mirror of jimm.models.vit (reference: src/jimm/models/vit.py): VisionTransformer with the reference's constructor,
`__call__` and `from_pretrained`, running on the B200 CUDA library."""

from __future__ import annotations

import os
from typing import Any, Set

import torch

from .. import _lib, nn
from ..common.transformer import g_wrap
from ..common.utils import load_params_and_config
from ..common.vit import VisionTransformerBase, _NativeOwner, tower_config_fields


class VisionTransformer(_NativeOwner, nn.Module):
    """Vision Transformer for image classification (models/vit.py:16-103)."""

    def __init__(self, num_classes: int = 1000, in_channels: int = 3, img_size: int = 224, patch_size: int = 16, num_layers: int = 12,
                 num_heads: int = 12, mlp_dim: int = 3072, hidden_size: int = 768, dropout_rate: float = 0.1,
                 use_quick_gelu: bool = False, do_classification: bool = True, dtype=torch.float32, param_dtype=torch.float32,
                 rngs=None, mesh=None) -> None:
        nn.Module.__init__(self)
        self._native_init(dtype)
        g = nn._gen(rngs)
        object.__setattr__(self, "do_classification", do_classification)
        object.__setattr__(self, "num_classes", num_classes)
        object.__setattr__(self, "dtype", dtype)
        # models/vit.py:61-78: CLS pooling, no pre-norm, patch bias, layernorm_epsilon=1e-12
        self.add_child("encoder", VisionTransformerBase(
            img_size=img_size, patch_size=patch_size, in_channels=in_channels, hidden_size=hidden_size, num_layers=num_layers,
            num_heads=num_heads, mlp_dim=mlp_dim, dropout_rate=dropout_rate, use_quick_gelu=use_quick_gelu, use_pre_norm=False,
            use_patch_bias=True, layernorm_epsilon=1e-12, rngs=g_wrap(g), dtype=dtype, param_dtype=param_dtype, mesh=mesh))
        if do_classification:
            self.add_child("classifier", nn.Linear(hidden_size, num_classes, rngs=g_wrap(g)))

    def _native_config(self) -> _lib.Config:
        cfg = _lib.Config()
        cfg.kind = _lib.KIND_VIT
        tower_config_fields(cfg, **self.encoder._hp)
        cfg.num_classes = self.num_classes if self.do_classification else 0
        cfg.compute_dtype = self._compute_dtype
        return cfg

    def __call__(self, x) -> torch.Tensor:
        """[batch, height, width, channels] -> logits [batch, num_classes] (models/vit.py:91-103).

        Inference semantics (dropout is the identity), like the reference after `.eval()`."""
        return self.native(x.shape[0]).vision(x)

    def forward_async(self, x):
        """Asynchronous dispatch for host inputs (JAX dispatches asynchronously; examples/vit_inference.py:54-58 only blocks when
        it reads the logits): returns a `PendingResult`; back-to-back calls overlap their H2D copies with the previous forward."""
        return self.native(x.shape[0]).vision_async(x)

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, use_pytorch: bool = False, mesh=None, dtype=torch.float32) -> "VisionTransformer":
        """Load a HF `ViTForImageClassification` checkpoint (models/vit.py:105-273): same config parsing, shape
        inference without a config, HF->flax name map, layout transforms and strict visit checks."""
        params_fstate, config = load_params_and_config(model_name_or_path, use_pytorch)
        use_quick_gelu_val = False
        if config:
            hidden_size_val = config["hidden_size"]
            num_classes_val = len(config["id2label"]) if "id2label" in config else config.get("num_labels", 1000)
            num_layers_val = config["num_hidden_layers"]
            num_heads_val = config["num_attention_heads"]
            mlp_dim_val = config["intermediate_size"]
            patch_size_val = config["patch_size"]
            img_size_val = config["image_size"]
            if "hidden_act" in config and config["hidden_act"] == "quick_gelu":
                use_quick_gelu_val = True
            elif "hidden_act" in config and config["hidden_act"] != "gelu":
                print(f"Warning: Unexpected hidden_act '{config['hidden_act']}' in config, defaulting to standard GELU.")
        elif not use_pytorch and (os.path.exists(model_name_or_path) and os.path.isfile(model_name_or_path)):
            hidden_size_val = params_fstate["vit.embeddings.cls_token"].shape[-1]
            num_classes_val = params_fstate["classifier.bias"].shape[0]
            max_layer_idx = -1
            for k in params_fstate:
                if k.startswith("vit.encoder.layer."):
                    max_layer_idx = max(max_layer_idx, int(k.split(".")[3]))
            num_layers_val = max_layer_idx + 1
            mlp_dim_val = params_fstate["vit.encoder.layer.0.intermediate.dense.weight"].shape[0]
            num_heads_val = hidden_size_val // 64
            patch_size_val = params_fstate["vit.embeddings.patch_embeddings.projection.weight"].shape[2]
            num_patches = params_fstate["vit.embeddings.position_embeddings"].shape[1] - 1
            img_size_val = int(num_patches ** 0.5) * patch_size_val
        else:
            raise ValueError(f"Could not load or infer configuration for {model_name_or_path}")

        with nn.deferred_init():  # every parameter is overwritten below (and asserted to be)
            model = cls(num_classes=num_classes_val, img_size=img_size_val, patch_size=patch_size_val, num_layers=num_layers_val,
                        num_heads=num_heads_val, mlp_dim=mlp_dim_val, hidden_size=hidden_size_val, use_quick_gelu=use_quick_gelu_val,
                        mesh=mesh, dtype=dtype, param_dtype=dtype)
        flax_params = model.flat_params()

        def hf_param_name(name: str) -> str:
            return "weight" if name in ["kernel", "scale"] else name

        head_dim = hidden_size_val // num_heads_val
        mapping = {
            "encoder.cls_token": "vit.embeddings.cls_token",
            "encoder.position_embeddings": "vit.embeddings.position_embeddings",
            "encoder.patch_embeddings.kernel": "vit.embeddings.patch_embeddings.projection.weight",
            "encoder.patch_embeddings.bias": "vit.embeddings.patch_embeddings.projection.bias",
            "classifier.kernel": "classifier.weight",
            "classifier.bias": "classifier.bias",
            "encoder.ln_post.scale": "vit.layernorm.weight",
            "encoder.ln_post.bias": "vit.layernorm.bias",
        }
        for i in range(num_layers_val):
            fb, hb = f"encoder.transformer.blocks.layers.{i}.", f"vit.encoder.layer.{i}."
            for p in ("kernel", "bias"):
                for y in ("key", "value", "query"):
                    mapping[fb + f"attn.{y}.{p}"] = hb + f"attention.attention.{y}.{hf_param_name(p)}"
                mapping[fb + f"attn.out.{p}"] = hb + f"attention.output.dense.{hf_param_name(p)}"
                mapping[fb + f"mlp.layers.0.{p}"] = hb + f"intermediate.dense.{hf_param_name(p)}"
                mapping[fb + f"mlp.layers.3.{p}"] = hb + f"output.dense.{hf_param_name(p)}"
            for p in ("scale", "bias"):
                mapping[fb + f"norm1.{p}"] = hb + f"layernorm_before.{hf_param_name(p)}"
                mapping[fb + f"norm2.{p}"] = hb + f"layernorm_after.{hf_param_name(p)}"

        nonvisited = set(flax_params.keys())
        used_hf_keys: Set[str] = set()
        for dst, src in mapping.items():
            assert dst in flax_params, dst
            used_hf_keys.add(src)
            assert src in params_fstate, f"HF key '{src}' (from Flax key {dst}) not found in loaded weights."
            nonvisited.remove(dst)
            v = params_fstate[src].to(torch.float32)
            tail = src.split(".")
            if dst == "encoder.patch_embeddings.kernel":
                v = v.permute(2, 3, 1, 0)
            elif tail[-1] == "weight" and tail[-2] in ("key", "value", "query"):
                v = v.T.reshape(hidden_size_val, num_heads_val, head_dim)
            elif tail[-1] == "bias" and tail[-2] in ("key", "value", "query"):
                v = v.reshape(num_heads_val, head_dim)
            elif tail[-4:] == ["attention", "output", "dense", "weight"]:
                v = v.T.reshape(num_heads_val, head_dim, hidden_size_val)
            elif tail[-1] == "weight" and v.ndim == 2:
                v = v.T
            assert tuple(v.shape) == tuple(flax_params[dst].shape), \
                f"Shape mismatch for {dst} (Flax) vs {src} (HF): {tuple(flax_params[dst].shape)} != {tuple(v.shape)}"
            model.set_flat_param(dst, v)
        assert len(nonvisited) == 0, f"Some Flax model parameters were not visited: {nonvisited}"
        leftover = set(params_fstate.keys()) - used_hf_keys
        known_unused = {"text_model.embeddings.position_ids", "vision_model.embeddings.position_ids"}
        unexpected = leftover - known_unused
        assert len(unexpected) == 0, f"Some unexpected HuggingFace checkpoint parameters were not used: {sorted(list(unexpected))}"
        return model
