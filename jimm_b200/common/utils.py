"""Mirror of jimm.common.utils (reference: src/jimm/common/utils.py): `load_params_and_config` with the reference's search order and
return contract, rebuilt for zero-copy ingestion -- tensors come back as views of the memory-mapped checkpoint file in their stored
dtype (fp32 / fp16 / bf16); nothing is converted or transposed on the CPU (the CUDA library casts and packs at finalize)."""

from __future__ import annotations

import json
import mmap
import os
import struct
from typing import Any, Dict, Optional, Tuple

import torch

_ST_DTYPES = {"F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16, "F64": torch.float64, "I64": torch.int64, "I32": torch.int32,
              "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}


def sharded_init(init, spec=None, mesh=None):
    """Reference: common/utils.py:14-25.  Parameters are replicated on every GPU in the B200 build (data parallel only, SURVEY.md 2
    'TP: metadata only'), so the partition spec is accepted and ignored."""
    return init


def read_safetensors_mmap(path: str) -> Dict[str, torch.Tensor]:
    """The safetensors container read directly: 8-byte little-endian header length, JSON header {name: {dtype, shape, data_offsets}},
    then the raw tensor bytes.  Every tensor is a `torch.frombuffer` view of one shared read-only mmap (no copy; pages are faulted in
    when finalize streams them to the GPU)."""
    f = open(path, "rb")
    try:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    finally:
        f.close()
    (hlen,) = struct.unpack("<Q", mm[:8])
    header = json.loads(mm[8:8 + hlen].decode("utf-8"))
    base = 8 + hlen
    buf = memoryview(mm)
    out: Dict[str, torch.Tensor] = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        dt = _ST_DTYPES.get(meta["dtype"])
        if dt is None:
            raise ValueError(f"unsupported safetensors dtype {meta['dtype']} for {name}")
        lo, hi = meta["data_offsets"]
        shape = tuple(meta["shape"])
        n = 1
        for d in shape:
            n *= d
        if n == 0:
            out[name] = torch.empty(shape, dtype=dt)
            continue
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # the buffer is read-only by design; nothing on this path writes to it
            t = torch.frombuffer(buf, dtype=dt, count=n, offset=base + lo)
        out[name] = t.reshape(shape)
    return out


def _first_existing(*paths: Optional[str]) -> Optional[str]:
    for p in paths:
        if p and os.path.exists(p):
            return p
    return None


def load_params_and_config(
    model_name_or_path: str,
    use_pytorch: bool = False,
    default_config_filename: str = "config.json",
    default_pytorch_filename: str = "pytorch_model.bin",
    default_safetensors_filename: str = "model.safetensors",
) -> Tuple[Dict[str, torch.Tensor], Dict[str, Any]]:
    """HF-named parameters + config dict (common/utils.py:28-107).  Sources, in the reference's order:
      use_pytorch=True : `<dir>/pytorch_model.bin` + `<dir>/config.json`, or the same two files from the HF hub;
      otherwise        : a local `.safetensors` FILE whose config.json sits beside it (or one level up when the file lives in `model/`),
                         or `model.safetensors` + `config.json` from the hub (a missing config gives {} and the caller infers the shapes).
    Raises ValueError when no weights could be read.  Tensors keep their stored dtype -- bf16 `pytorch_model.bin` files load
    (the reference's `.numpy()` at common/utils.py:71 cannot take them) -- and safetensors are memory-mapped views."""

    def hub(filename: str) -> str:
        from huggingface_hub import hf_hub_download

        return hf_hub_download(repo_id=model_name_or_path, filename=filename)

    config: Dict[str, Any] = {}
    params: Optional[Dict[str, torch.Tensor]] = None
    if use_pytorch:
        local = os.path.isdir(model_name_or_path)
        cfg_path = os.path.join(model_name_or_path, default_config_filename) if local else hub(default_config_filename)
        bin_path = os.path.join(model_name_or_path, default_pytorch_filename) if local else hub(default_pytorch_filename)
        if os.path.exists(cfg_path):
            with open(cfg_path) as fh:
                config = json.load(fh)
        if os.path.exists(bin_path):
            try:
                params = dict(torch.load(bin_path, map_location="cpu", weights_only=True, mmap=True))
            except (RuntimeError, ValueError):  # legacy (non-zip) pickles cannot be mapped
                params = dict(torch.load(bin_path, map_location="cpu", weights_only=True))
    elif os.path.isfile(model_name_or_path):
        here = os.path.dirname(model_name_or_path)
        up = os.path.dirname(here) if os.path.basename(here) == "model" else None
        cfg_path = _first_existing(os.path.join(here, default_config_filename), up and os.path.join(up, default_config_filename))
        if cfg_path:
            with open(cfg_path) as fh:
                config = json.load(fh)
        params = read_safetensors_mmap(model_name_or_path)
    else:
        try:
            with open(hub(default_config_filename)) as fh:
                config = json.load(fh)
        except Exception:
            config = {}
        try:
            st_path = hub(default_safetensors_filename)
        except Exception:
            st_path = None
        if st_path and os.path.exists(st_path):
            params = read_safetensors_mmap(st_path)
    if params is None:
        raise ValueError(f"Could not load parameters from {model_name_or_path} (use_pytorch={use_pytorch})")
    return params, config
