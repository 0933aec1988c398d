"""Checkpoint ingestion time: HF checkpoint file -> from_pretrained -> native handle (create + parameter hand-off + finalize)
-> first forward, for ViT-B/16, ViT-L/16@384 and a SigLIP2-L/16@512-sized dual tower (882 M parameters, vocabulary 256000), from
fp32 safetensors and from a bf16 pytorch_model.bin (which the reference's `.numpy()` loader, common/utils.py:71, cannot read).

    python scripts/gpu_load_time.py [vit_b] [vit_l] [siglip2_l]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from safetensors.torch import save_file

from jimm_b200.models import SigLIP, VisionTransformer

which = sys.argv[1:] or ["vit_b", "vit_l", "siglip2_l"]
torch.cuda.init()
torch.zeros(1, device="cuda")
torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))


def make(name, d):
    from transformers import SiglipConfig, SiglipModel, SiglipTextConfig, SiglipVisionConfig, ViTConfig, ViTForImageClassification

    torch.manual_seed(0)
    if name == "vit_b":
        hf = ViTForImageClassification(ViTConfig(num_labels=1000))
    elif name == "vit_l":
        hf = ViTForImageClassification(ViTConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=384,
                                                 num_labels=1000))
    else:
        cfg = SiglipConfig(vision_config=SiglipVisionConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                                                             image_size=512, patch_size=16).to_dict(),
                           text_config=SiglipTextConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                                                        vocab_size=256000, max_position_embeddings=64, projection_size=1024).to_dict())
        hf = SiglipModel(cfg)
    sd = {k: v.detach().contiguous() for k, v in hf.state_dict().items()}
    save_file(sd, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(hf.config.to_dict(), f)
    torch.save({k: v.to(torch.bfloat16) for k, v in sd.items()}, os.path.join(d, "pytorch_model.bin"))
    n = sum(v.numel() for v in sd.values())
    del hf, sd
    return n


for name in which:
    d = tempfile.mkdtemp(prefix=f"jimm_{name}_")
    t0 = time.perf_counter()
    nparams = make(name, d)
    print(f"== {name}: {nparams/1e6:.1f} M parameters, checkpoint written in {time.perf_counter()-t0:.1f} s", flush=True)
    cls = SigLIP if name == "siglip2_l" else VisionTransformer
    dt = torch.bfloat16 if name != "vit_b" else torch.float16
    for src, kw in (("fp32 safetensors", dict(path=os.path.join(d, "model.safetensors"), use_pytorch=False)),
                    ("bf16 pytorch_model.bin", dict(path=d, use_pytorch=True))):
        for rep in range(2):
            t0 = time.perf_counter()
            try:
                m = cls.from_pretrained(kw["path"], use_pytorch=kw["use_pytorch"], dtype=dt)
            except Exception as e:  # noqa: BLE001
                print(f"{name} | {src}: from_pretrained failed: {type(e).__name__}: {str(e)[:120]}", flush=True)
                break
            t1 = time.perf_counter()
            n = m.native(4)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            res = m.cfg_img if hasattr(m, "cfg_img") else None
            img = getattr(m, "image_resolution", None) or m.encoder._hp["img_size"]
            x = torch.randn(2, img, img, 3, device="cuda")
            out = m.encode_image(x) if cls is SigLIP else m(x)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"{name} | {src} | rep {rep}: from_pretrained {t1-t0:.2f} s | create + hand-off + finalize {t2-t1:.2f} s | first forward {t3-t2:.2f} s | "
                  f"total {t3-t0:.2f} s", flush=True)
            del m, n, out
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
