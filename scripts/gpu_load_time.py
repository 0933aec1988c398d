"""Checkpoint ingestion time: HF safetensors -> from_pretrained -> native handle -> first forward."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from transformers import ViTConfig, ViTForImageClassification
from jimm_b200.models import VisionTransformer

d = tempfile.mkdtemp()
torch.manual_seed(0)
hf = ViTForImageClassification(ViTConfig(num_labels=1000))
hf.save_pretrained(d, safe_serialization=True)
del hf
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(2):
    t0 = time.perf_counter()
    m = VisionTransformer.from_pretrained(os.path.join(d, "model.safetensors"), dtype=torch.float16).eval()
    t1 = time.perf_counter()
    n = m.native(8)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = m(torch.randn(8, 224, 224, 3, device="cuda"))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"rep {rep}: from_pretrained (read + transforms) {t1-t0:.3f} s | create + set_param + finalize {t2-t1:.3f} s | first forward {t3-t2:.3f} s", flush=True)
    del m, n
