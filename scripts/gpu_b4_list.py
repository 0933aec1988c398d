import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
os.environ.setdefault("JIMM_GRAPH_MAX_BATCH", "0")
model, img_size, _ = bench.build_model("vit_b16", "float16")
x = torch.randn(int(os.environ.get("B", "4")), img_size, img_size, 3, device="cuda")
for _ in range(3): model(x)
torch.cuda.synchronize()
