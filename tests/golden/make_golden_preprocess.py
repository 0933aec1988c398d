"""Generate the preprocessing golden fixtures with the real third-party pipeline the reference's examples call.

    python tests/golden/make_golden_preprocess.py

Runs transformers' PIL-backend processors (`ViTImageProcessorPil`, `CLIPImageProcessorPil`, `SiglipImageProcessorPil`; same
arithmetic as the slow processors of the transformers 4.53.0 the reference pins, on the Pillow of this image) on small seeded
uint8 images and stores input + `pixel_values` transposed to NHWC (examples/vit_inference.py:36-37).  Needs transformers +
Pillow (present in the build container, not needed on the GPU box: the .npz files are committed).
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import preprocess_oracle as P  # only for the seeded synthetic images


def main():
    from transformers import CLIPImageProcessorPil, SiglipImageProcessorPil, ViTImageProcessorPil

    cases = {
        "vit": (ViTImageProcessorPil(size={"height": 48, "width": 48}), [(61, 83), (48, 48), (30, 100)]),
        "clip": (CLIPImageProcessorPil(size={"shortest_edge": 40}, crop_size={"height": 40, "width": 40}), [(61, 83), (90, 57), (40, 40)]),
        "siglip": (SiglipImageProcessorPil(size={"height": 64, "width": 64}), [(61, 83), (200, 150), (20, 24)]),
    }
    for name, (proc, sizes) in cases.items():
        data = {}
        for i, (h, w) in enumerate(sizes):
            img = P.synthetic_u8_images(1, h, w, seed=100 + i)[0]
            pv = proc(images=Image.fromarray(img), return_tensors="np")["pixel_values"][0]
            data[f"img{i}"] = img
            data[f"out{i}"] = np.ascontiguousarray(pv.transpose(1, 2, 0)).astype(np.float32)
        path = os.path.join(HERE, f"preprocess_{name}.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
