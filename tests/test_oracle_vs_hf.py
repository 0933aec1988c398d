"""CPU: pin the oracle (a) against HuggingFace transformers on random-init models -- the offline mirror of the
reference's own tests (tests/test_vit.py, test_clip.py, test_siglip.py) -- and (b) against the committed golden vectors."""

import os

import numpy as np
import pytest
import torch

import check_vs_hf as H
import jimm_oracle as O


def test_vit_tiny_matches_hf():
    r = H.check_vit()
    assert r["hf_rel"] < 1e-9, r  # HF semantics: layout / head split / patch order exact
    assert r["jimm_abs"] < 0.05, r  # the reference's own tolerance (tests/test_vit.py:49-52)
    assert r["argmax_equal"], r


def test_clip_tiny_matches_hf():
    r = H.check_clip()
    assert r["hf_rel"] < 1e-9, r
    assert r["jimm_abs"] < 1e-1, r  # tests/test_clip.py:48


def test_siglip_tiny_matches_hf():
    r = H.check_siglip()
    assert r["img_rel"] < 1e-9 and r["txt_rel"] < 1e-9 and r["logits_rel"] < 1e-9, r  # jimm semantics == HF semantics here
    assert r["logits_abs"] < 1e-2, r  # tests/test_siglip.py:69


def test_vit_b16_fp32_matches_hf():
    """Full-size ViT-B/16 (config 1 shape) in fp32 at the reference's tolerance."""
    from transformers import ViTConfig

    r = H.check_vit(ViTConfig(num_labels=1000), B=1, dtype=torch.float32)
    assert r["hf_rel"] < 1e-4, r
    assert r["jimm_abs"] < 0.05 and r["argmax_equal"], r


def _load(golden_dir, name):
    from safetensors.torch import load_file

    d = os.path.join(golden_dir, name)
    return load_file(os.path.join(d, "model.safetensors")), dict(np.load(os.path.join(d, "io.npz")))


def test_golden_vit(golden_dir):
    sd, io = _load(golden_dir, "tiny_vit")
    oc = O.ViTCfg(num_classes=10, img_size=32, patch_size=8, num_layers=2, num_heads=2, mlp_dim=256, hidden_size=128)
    p = O.cast_params(O.hf_to_flax_vit(sd, 2, 2), torch.float64)
    img = torch.from_numpy(io["images"]).double()
    out = O.vit_forward(p, oc, img).float().numpy()
    np.testing.assert_allclose(out, io["oracle_logits"], rtol=0, atol=1e-6)
    assert np.abs(out - io["hf_logits"]).max() < 0.05
    out_hf = O.vit_forward(p, oc, img, O.Semantics(gelu="erf", block_eps=1e-12)).float().numpy()
    np.testing.assert_allclose(out_hf, io["hf_logits"], rtol=0, atol=2e-5)


def test_golden_clip(golden_dir):
    sd, io = _load(golden_dir, "tiny_clip")
    oc = O.DualCfg(32, 2, 128, 8, 16, 100, 128, 2, 2)
    p = O.cast_params(O.hf_to_flax_clip(sd, oc), torch.float64)
    img, txt = torch.from_numpy(io["images"]).double(), torch.from_numpy(io["tokens"]).long()
    out = O.clip_forward(p, oc, img, txt).float().numpy()
    np.testing.assert_allclose(out, io["oracle_logits"], rtol=0, atol=1e-5)
    assert np.abs(out - io["hf_logits"]).max() < 1e-1
    out_hf = O.clip_forward(p, oc, img, txt, O.Semantics(block_eps=1e-5)).float().numpy()
    np.testing.assert_allclose(out_hf, io["hf_logits"], rtol=0, atol=5e-5)


def test_golden_siglip(golden_dir):
    sd, io = _load(golden_dir, "tiny_siglip")
    oc = O.DualCfg(32, 2, 128, 8, 16, 100, 128, 2, 2)
    p = O.cast_params(O.hf_to_flax_siglip(sd, oc), torch.float64)
    img, txt = torch.from_numpy(io["images"]).double(), torch.from_numpy(io["tokens"]).long()
    np.testing.assert_allclose(O.siglip_encode_image(p, oc, img).float().numpy(), io["hf_image_embeds"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(O.siglip_encode_text(p, oc, txt).float().numpy(), io["hf_text_embeds"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(O.siglip_forward(p, oc, img, txt).float().numpy(), io["hf_logits"], rtol=0, atol=1e-4)


def test_operand_rounding_emulator():
    x = torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -10, 3.14159265], dtype=torch.float64)
    t = O.round_operand(x, "tf32")
    assert t[0] == 1.0 and t[1] == 1.0 + 2 ** -10  # 10 explicit mantissa bits, round-to-nearest-even
    assert abs(float(O.round_operand(x, "fp16")[2]) - float(torch.tensor(3.14159265).half())) == 0
    assert abs(float(O.round_operand(x, "bf16")[2]) - float(torch.tensor(3.14159265).bfloat16())) == 0


def test_edge_semantics():
    """Quirks the CUDA path must reproduce."""
    # fast-variance LN clamps at zero: constant row -> (x-mean)*rsqrt(eps)
    x = torch.full((1, 8), 3.0, dtype=torch.float64)
    y = O.layer_norm(x, torch.ones(8, dtype=torch.float64), torch.zeros(8, dtype=torch.float64), 1e-6)
    assert torch.all(y == 0)
    # CLIP pooling is argmax over token ids (first max), not "last token"
    oc = O.DualCfg(32, 1, 64, 8, 8, 50, 64, 1, 1)
    p = O.random_dual_params(oc, "clip", seed=3, dtype=torch.float64)
    t1 = torch.tensor([[5, 49, 7, 49, 1, 1, 1, 1]])
    t2 = torch.tensor([[5, 49, 7, 3, 1, 1, 1, 1]])
    # causal mask: pooled position 1 only sees tokens 0..1, so later tokens must not matter
    assert torch.allclose(O.clip_encode_text(p, oc, t1), O.clip_encode_text(p, oc, t2), atol=1e-12)
