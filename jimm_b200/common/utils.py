"""Mirror of jimm.common.utils (reference: src/jimm/common/utils.py)."""

from __future__ import annotations

import json
import os
from typing import Any, Dict, Tuple

import torch


def sharded_init(init, spec=None, mesh=None):
    """Reference: common/utils.py:14-25.  Parameters are replicated on every GPU in the B200 build (data parallel
    only, SURVEY.md 2 'TP: metadata only'), so the partition spec is accepted and ignored."""
    return init


def load_params_and_config(
    model_name_or_path: str,
    use_pytorch: bool = False,
    default_config_filename: str = "config.json",
    default_pytorch_filename: str = "pytorch_model.bin",
    default_safetensors_filename: str = "model.safetensors",
) -> Tuple[Dict[str, torch.Tensor], Dict[str, Any]]:
    """Load HF-named parameters and the config dict from (a) a local dir / hub repo with `pytorch_model.bin`,
    (b) a local `.safetensors` file with a sibling (or parent-of-`model/`) `config.json`, (c) the HF hub.
    Same search order, return contract and ValueError as common/utils.py:28-107 (tensors are torch, not jax)."""
    params: Dict[str, torch.Tensor] | None = None
    config: Dict[str, Any] = {}
    config_file_path = None
    weights_file_path = None

    def _hub(filename):
        from huggingface_hub import hf_hub_download

        return hf_hub_download(repo_id=model_name_or_path, filename=filename)

    if use_pytorch:
        if os.path.isdir(model_name_or_path):
            config_file_path = os.path.join(model_name_or_path, default_config_filename)
            weights_file_path = os.path.join(model_name_or_path, default_pytorch_filename)
        else:
            config_file_path = _hub(default_config_filename)
            weights_file_path = _hub(default_pytorch_filename)
        if config_file_path and os.path.exists(config_file_path):
            with open(config_file_path, "r") as f:
                config = json.load(f)
        if weights_file_path and os.path.exists(weights_file_path):
            state_dict = torch.load(weights_file_path, map_location="cpu", weights_only=True)
            params = {k: v for k, v in state_dict.items()}
    else:
        if os.path.exists(model_name_or_path) and os.path.isfile(model_name_or_path):
            weights_file_path = model_name_or_path
            attempt1 = os.path.join(os.path.dirname(model_name_or_path), default_config_filename)
            if os.path.exists(attempt1):
                config_file_path = attempt1
            else:
                if os.path.basename(os.path.dirname(model_name_or_path)) == "model":
                    attempt2 = os.path.join(os.path.dirname(os.path.dirname(model_name_or_path)), default_config_filename)
                    if os.path.exists(attempt2):
                        config_file_path = attempt2
            if config_file_path and os.path.exists(config_file_path):
                with open(config_file_path, "r") as f:
                    config = json.load(f)
        else:
            try:
                config_file_path = _hub(default_config_filename)
                with open(config_file_path, "r") as f:
                    config = json.load(f)
            except Exception:
                config = {}
            try:
                weights_file_path = _hub(default_safetensors_filename)
            except Exception:
                weights_file_path = None
        if weights_file_path and os.path.exists(weights_file_path):
            from safetensors.torch import load_file

            params = load_file(weights_file_path)

    if params is None:
        raise ValueError(f"Could not load parameters from {model_name_or_path} (use_pytorch={use_pytorch})")
    return params, config
