"""CPU oracle for the image front-end the reference's examples run on the host before the forward path.

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden_preprocess.py and bench.py's cpu_baseline leg, never by
the product (jimm_b200/).

What it restates.  The reference itself has no preprocessing code: `examples/vit_inference.py:27-37`,
`examples/clip_inference.py:35-38` and `examples/siglip_inference.py` call HuggingFace image processors and transpose the
result NCHW -> NHWC.  The arithmetic therefore lives in two un-vendored dependencies pinned by the reference's `uv.lock`:

* `transformers` 4.53.0 (`uv.lock:2679-2680`), slow (PIL/NumPy) processors — the reference's environment has no torchvision:
  `resize` (PIL `Image.resize`, no reducing gap) -> `center_crop` (CLIP only) -> `rescale` (`float64(u8) * factor`, then
  cast to float32) -> `normalize` (`(x - float32(mean)) / float32(std)` in float32).  Output sizes: exact (height, width)
  for ViT / SigLIP; CLIP resizes the SHORTEST edge to `size` keeping the aspect ratio (`int(size * long / short)` for the
  other edge) and centre-crops `crop_size` with `top = (h - ch) // 2`, `left = (w - cw) // 2`.
* `pillow` 11.3.0 (`uv.lock:1573-1574`), `src/libImaging/Resample.c`, the 8-bit path: per output coordinate a window
  `[xmin, xmin + xmax)` of the input and double-precision filter weights, normalised to sum 1, converted to fixed point
  with 22 fractional bits (round half away from zero), accumulated in int32 starting from 1 << 21, shifted right by 22 and
  clamped to [0, 255]; a horizontal pass into an 8-bit temporary image followed by a vertical pass over it.

Pinned (tests/test_preprocess_oracle.py, CPU): bit-exact against the Pillow installed in this image (`Image.resize`, bilinear and
bicubic, up- and down-scaling, odd sizes) and against transformers' PIL-backend processors (`ViTImageProcessorPil`,
`CLIPImageProcessorPil`, `SiglipImageProcessorPil`), plus committed golden fixtures (tests/golden/preprocess_*.npz).
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np

BILINEAR = 2  # PIL.Image.Resampling.BILINEAR
BICUBIC = 3  # PIL.Image.Resampling.BICUBIC
PRECISION_BITS = 32 - 8 - 2  # Resample.c: room for the 8-bit sample and two guard bits in an int32 accumulator


def _bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {BILINEAR: (_bilinear, 1.0), BICUBIC: (_bicubic, 2.0)}


def resample_coeffs(in_size: int, out_size: int, resample: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the full-image box: (first[out], count[out], kk[out, ksize] int32)."""
    f, fsupport = _FILTERS[resample]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    first = np.zeros(out_size, np.int32)
    count = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [f((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        first[xx], count[xx] = xmin, xmax
    return first, count, kk


def _pass(img: np.ndarray, axis_first: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an [H, W, C] image."""
    out_size, ksize = kk.shape
    n_in = img.shape[axis]
    idx = np.minimum(axis_first[:, None] + np.arange(ksize)[None, :], n_in - 1)  # taps past the window have weight 0
    src = img.astype(np.int64)
    if axis == 1:
        g = src[:, idx, :]  # [H, out, ksize, C]
        acc = (g * kk[None, :, :, None].astype(np.int64)).sum(axis=2)
    else:
        g = src[idx, :, :]  # [out, ksize, W, C]
        acc = (g * kk[:, :, None, None].astype(np.int64)).sum(axis=1)
    acc += 1 << (PRECISION_BITS - 1)
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_u8(img: np.ndarray, out_h: int, out_w: int, resample: int) -> np.ndarray:
    """`PIL.Image.resize((out_w, out_h), resample)` of an 8-bit [H, W, C] image (Resample.c `ImagingResample`): horizontal pass
    over the rows the vertical pass will read, then the vertical pass; a pass whose size does not change is skipped."""
    h, w, _ = img.shape
    fv, cv, kv = resample_coeffs(h, out_h, resample)
    out = img
    if out_w != w:
        fh, _, kh = resample_coeffs(w, out_w, resample)
        y0, y1 = int(fv[0]), int(fv[-1] + cv[-1])
        out = _pass(img[y0:y1], fh, kh, axis=1)
        fv = fv - y0
    if out_h != h:
        out = _pass(out, fv, kv, axis=0)
    return np.ascontiguousarray(out)


@dataclass
class PreprocessConfig:
    """The HF image-processor settings the three model families ship with (`preprocessor_config.json`)."""

    height: int = 224  # exact output size ...
    width: int = 224
    shortest_edge: int = 0  # ... or (CLIP) resize the shortest edge to this, keep the aspect ratio
    crop_h: int = 0  # centre crop (CLIP)
    crop_w: int = 0
    resample: int = BILINEAR
    rescale_factor: float = 1 / 255
    mean: Sequence[float] = (0.5, 0.5, 0.5)
    std: Sequence[float] = (0.5, 0.5, 0.5)

    @staticmethod
    def vit(size=224):
        return PreprocessConfig(height=size, width=size, resample=BILINEAR)

    @staticmethod
    def siglip(size=224):
        return PreprocessConfig(height=size, width=size, resample=BICUBIC)

    @staticmethod
    def clip(size=224):
        return PreprocessConfig(height=0, width=0, shortest_edge=size, crop_h=size, crop_w=size, resample=BICUBIC,
                                mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711))


def resized_size(cfg: PreprocessConfig, h: int, w: int) -> Tuple[int, int]:
    """transformers `get_resize_output_image_size(..., default_to_square=False)` for the shortest-edge mode."""
    if not cfg.shortest_edge:
        return cfg.height, cfg.width
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = cfg.shortest_edge, int(cfg.shortest_edge * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def rescale_normalize(u8: np.ndarray, cfg: PreprocessConfig) -> np.ndarray:
    """transformers `rescale` then `normalize` on an [..., C] uint8 array -> float32."""
    x = (u8.astype(np.float64) * cfg.rescale_factor).astype(np.float32)
    mean = np.array(cfg.mean, dtype=np.float32)
    std = np.array(cfg.std, dtype=np.float32)
    return (x - mean) / std


def preprocess(img: np.ndarray, cfg: PreprocessConfig) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [h, w, 3] NHWC sample (the examples' `transpose(0, 2, 3, 1)` of `pixel_values`)."""
    h, w, _ = img.shape
    rh, rw = resized_size(cfg, h, w)
    out = pil_resize_u8(img, rh, rw, cfg.resample) if (rh, rw) != (h, w) else img
    if cfg.crop_h:
        if cfg.crop_h > rh or cfg.crop_w > rw:
            raise ValueError("centre crop larger than the resized image (HF pads with zeros; not part of the path)")
        top, left = (rh - cfg.crop_h) // 2, (rw - cfg.crop_w) // 2
        out = out[top:top + cfg.crop_h, left:left + cfg.crop_w]
    return rescale_normalize(out, cfg)


def synthetic_u8_images(B: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """Smooth-ish random 8-bit images (low-frequency field plus noise) so that both filter lobes and the clamps are exercised."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(B, (h + 7) // 8 + 1, (w + 7) // 8 + 1, 3)).astype(np.float32)
    up = np.repeat(np.repeat(base, 8, axis=1), 8, axis=2)[:, :h, :w]
    noise = rng.integers(-96, 97, size=(B, h, w, 3)).astype(np.float32)
    return np.clip(up + noise, 0, 255).astype(np.uint8)


# ---- zero-shot / classification epilogue ----
def zero_shot_oracle(logits: np.ndarray):
    """examples/clip_inference.py:46-51 per row: un-shifted softmax in float32 and `argsort(scores)[::-1]` (stable ascending
    sort reversed: equal scores come out with the larger index first)."""
    x = np.asarray(logits, dtype=np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(x)  # float32 like jnp.exp on the float32 logits: overflows to inf above ~88.7, the row then holds NaN / 0
        probs = (e.astype(np.float64) / e.astype(np.float64).sum(axis=-1, keepdims=True)).astype(np.float32)
    order = np.argsort(x, axis=-1, kind="stable")[..., ::-1].astype(np.int32)
    return probs, order


def classify_oracle(logits: np.ndarray) -> np.ndarray:
    """examples/vit_inference.py:58: first maximum per row."""
    return np.argmax(np.asarray(logits, dtype=np.float32), axis=-1).astype(np.int32)
