"""CPU tests of the image front-end oracle (oracle/preprocess_oracle.py) and of the library's host-side resampling tables.

Pins, without a GPU: the oracle against Pillow itself, against transformers' PIL-backend processors (the arithmetic of the
slow processors of the transformers 4.53.0 the reference pins), against the committed golden fixtures; and the C++ recipe for
Pillow's windows / fixed-point weights (csrc/preprocess.cu make_table) against the oracle's tables."""
import os

import numpy as np
import pytest

import preprocess_oracle as P

SIZES = [(480, 640, 224, 224), (37, 53, 224, 224), (500, 333, 224, 336), (224, 224, 224, 224), (300, 224, 224, 224),
         (224, 300, 224, 224), (65, 1000, 96, 96), (17, 19, 5, 7), (3, 3, 32, 32), (1, 1, 4, 4), (256, 256, 255, 257)]


@pytest.mark.parametrize("resample", [P.BILINEAR, P.BICUBIC])
def test_oracle_resize_matches_pillow(resample):
    from PIL import Image

    for i, (h, w, oh, ow) in enumerate(SIZES):
        img = P.synthetic_u8_images(1, h, w, seed=7 * i + resample)[0]
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=resample))
        got = P.pil_resize_u8(img, oh, ow, resample)
        assert np.array_equal(ref, got), (h, w, oh, ow, resample)
    # saturating stripes: negative bicubic lobes hit both clamps
    img = np.zeros((64, 64, 3), np.uint8)
    img[::2] = 255
    ref = np.asarray(Image.fromarray(img).resize((23, 29), resample=resample))
    assert np.array_equal(ref, P.pil_resize_u8(img, 29, 23, resample))


def test_oracle_matches_hf_pil_processors():
    from PIL import Image
    from transformers import CLIPImageProcessorPil, SiglipImageProcessorPil, ViTImageProcessorPil

    cases = (("vit", ViTImageProcessorPil(), P.PreprocessConfig.vit()), ("clip", CLIPImageProcessorPil(), P.PreprocessConfig.clip()),
             ("siglip", SiglipImageProcessorPil(), P.PreprocessConfig.siglip()),
             ("siglip384", SiglipImageProcessorPil(size={"height": 384, "width": 384}), P.PreprocessConfig.siglip(384)))
    for name, proc, cfg in cases:
        for h, w in ((480, 640), (640, 480), (333, 500), (224, 224), (100, 80)):
            img = P.synthetic_u8_images(1, h, w, seed=h + w)[0]
            ref = proc(images=Image.fromarray(img), return_tensors="np")["pixel_values"][0].transpose(1, 2, 0)
            got = P.preprocess(img, cfg)
            assert ref.dtype == np.float32 and np.array_equal(ref, got), (name, h, w)


def _golden_cfg(name):
    if name == "vit":
        return P.PreprocessConfig.vit(48)
    if name == "siglip":
        return P.PreprocessConfig.siglip(64)
    return P.PreprocessConfig.clip(40)


@pytest.mark.parametrize("name", ["vit", "clip", "siglip"])
def test_oracle_matches_golden(name, golden_dir):
    z = np.load(os.path.join(golden_dir, f"preprocess_{name}.npz"))
    cfg = _golden_cfg(name)
    for i in range(3):
        assert np.array_equal(P.preprocess(z[f"img{i}"], cfg), z[f"out{i}"]), (name, i)


def test_shortest_edge_size_rule():
    cfg = P.PreprocessConfig.clip(224)
    assert P.resized_size(cfg, 480, 640) == (224, 298)
    assert P.resized_size(cfg, 640, 480) == (298, 224)
    assert P.resized_size(cfg, 333, 500) == (224, 336)
    assert P.resized_size(cfg, 224, 224) == (224, 224)


def test_library_resample_tables_match_oracle(lib):
    from jimm_b200 import preprocess as pp

    rng = np.random.default_rng(0)
    pairs = [(640, 224), (480, 224), (224, 224), (53, 224), (1000, 96), (1920, 384), (333, 224), (298, 224), (3, 32), (1, 4), (4000, 224)]
    pairs += [(int(a), int(b)) for a, b in zip(rng.integers(1, 2000, 40), rng.integers(1, 600, 40))]
    for n_in, n_out in pairs:
        for rs in (P.BILINEAR, P.BICUBIC):
            f0, c0, k0 = P.resample_coeffs(n_in, n_out, rs)
            f1, c1, k1 = pp.resample_coeffs(n_in, n_out, rs)
            assert np.array_equal(f0, f1) and np.array_equal(c0, c1) and np.array_equal(k0, k1), (n_in, n_out, rs)
            assert (np.diff(f0) >= 0).all()  # window starts are monotonic: a tile of output rows needs one contiguous input range


def test_library_exports_preprocess_symbols(lib):
    for sym in ("jimm_preproc_create", "jimm_preproc_output_size", "jimm_preproc_run", "jimm_preproc_destroy", "jimm_k_resample_coeffs"):
        assert hasattr(lib, sym)
