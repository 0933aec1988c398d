"""Generate the committed golden fixtures (run from the repo root: `python tests/golden/make_golden.py`).

For each of tiny ViT / CLIP / SigLIP:
  <name>/model.safetensors + config.json   a random-init HuggingFace checkpoint (perturbed biases / LN / cls / probe),
                                           i.e. exactly what the reference's from_pretrained consumes
                                           (src/jimm/common/utils.py:74-90 local-safetensors branch)
  <name>/io.npz                            seeded inputs, the HuggingFace model's outputs (what the reference's own tests
                                           compare against: tests/test_vit.py:49, test_clip.py:48, test_siglip.py:36,52,69),
                                           and the oracle's outputs in jimm semantics (fp64 -> fp32)
The reference itself (jax/flax) cannot run in this image, so its outputs are not in the fixtures; see oracle header.
"""

import json
import os
import sys

os.environ.setdefault("HF_HUB_OFFLINE", "1")
import numpy as np
import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import check_vs_hf as H  # noqa: E402
import jimm_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, model, config_dict, io):
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)
    sd = {k: v.detach().to(torch.float32).contiguous() for k, v in model.state_dict().items()}
    save_file(sd, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(config_dict, f, indent=1)
    np.savez_compressed(os.path.join(d, "io.npz"), **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in io.items()})
    print(name, {k: tuple(v.shape) for k, v in io.items()})


def main():
    from transformers import CLIPModel, SiglipModel, ViTForImageClassification

    # ---------------- ViT ----------------
    torch.manual_seed(0)
    cfg = H.tiny_vit_config()
    m = H.perturb_(ViTForImageClassification(cfg)).eval()
    oc = O.ViTCfg(num_classes=cfg.num_labels, img_size=cfg.image_size, patch_size=cfg.patch_size, num_layers=cfg.num_hidden_layers,
                  num_heads=cfg.num_attention_heads, mlp_dim=cfg.intermediate_size, hidden_size=cfg.hidden_size)
    img = O.synthetic_images(5, cfg.image_size)
    with torch.no_grad():
        hf = m(pixel_values=img.permute(0, 3, 1, 2)).logits
        p = O.cast_params(O.hf_to_flax_vit({k: v.detach() for k, v in m.state_dict().items()}, oc.num_layers, oc.num_heads), torch.float64)
        orc = O.vit_forward(p, oc, img.double()).float()
    save("tiny_vit", m, cfg.to_dict(), dict(images=img, hf_logits=hf, oracle_logits=orc))

    # ---------------- CLIP ----------------
    torch.manual_seed(1)
    cfg = H.tiny_clip_config()
    m = H.perturb_(CLIPModel(cfg)).eval()
    oc = H._dual_cfg(cfg)
    img = O.synthetic_images(4, oc.image_resolution)
    txt = O.synthetic_tokens(6, oc.context_length, oc.vocab_size, "clip")
    with torch.no_grad():
        hf = m(pixel_values=img.permute(0, 3, 1, 2), input_ids=txt).logits_per_image
        p = O.cast_params(O.hf_to_flax_clip({k: v.detach() for k, v in m.state_dict().items()}, oc), torch.float64)
        ie, te = O.clip_encode_image(p, oc, img.double()), O.clip_encode_text(p, oc, txt)
        lg = O.contrastive_logits(ie, te, p["logit_scale"])
    save("tiny_clip", m, cfg.to_dict(), dict(images=img, tokens=txt.to(torch.int32), hf_logits=hf, oracle_image_embeds=ie.float(),
                                             oracle_text_embeds=te.float(), oracle_logits=lg.float()))

    # ---------------- SigLIP ----------------
    torch.manual_seed(2)
    cfg = H.tiny_siglip_config()
    m = H.perturb_(SiglipModel(cfg)).eval()
    with torch.no_grad():
        m.logit_scale.fill_(2.3)
        m.logit_bias.fill_(-1.7)
    oc = H._dual_cfg(cfg)
    img = O.synthetic_images(4, oc.image_resolution)
    txt = O.synthetic_tokens(6, oc.context_length, oc.vocab_size, "siglip")
    with torch.no_grad():
        out = m(pixel_values=img.permute(0, 3, 1, 2), input_ids=txt)
        hf_i = m.vision_model(pixel_values=img.permute(0, 3, 1, 2)).pooler_output
        hf_t = m.text_model(input_ids=txt).pooler_output
        p = O.cast_params(O.hf_to_flax_siglip({k: v.detach() for k, v in m.state_dict().items()}, oc), torch.float64)
        ie, te = O.siglip_encode_image(p, oc, img.double()), O.siglip_encode_text(p, oc, txt)
        lg = O.contrastive_logits(ie, te, p["logit_scale"], p["logit_bias"])
    save("tiny_siglip", m, cfg.to_dict(), dict(images=img, tokens=txt.to(torch.int32), hf_logits=out.logits_per_image, hf_image_embeds=hf_i,
                                               hf_text_embeds=hf_t, oracle_image_embeds=ie.float(), oracle_text_embeds=te.float(),
                                               oracle_logits=lg.float()))


if __name__ == "__main__":
    main()
