"""GPU parity tests of the fused image front-end kernel (csrc/preprocess.cu) against oracle/preprocess_oracle.py: bit-exact
(integer resampling, IEEE fp32 rescale / normalise), through the C ABI (jimm_preproc_*)."""
import os

import numpy as np
import pytest
import torch

import preprocess_oracle as P

pytestmark = pytest.mark.gpu


def _proc(kind, size, **kw):
    from jimm_b200.preprocess import ImagePreprocessor

    return getattr(ImagePreprocessor, kind)(size, **kw)


def _cfg(kind, size):
    return getattr(P.PreprocessConfig, kind)(size)


CASES = [("vit", 224, 480, 640), ("vit", 224, 224, 224), ("vit", 224, 37, 53), ("vit", 384, 500, 333), ("clip", 224, 480, 640),
         ("clip", 224, 640, 480), ("clip", 224, 333, 500), ("clip", 224, 224, 224), ("siglip", 256, 480, 640), ("siglip", 224, 100, 80),
         ("siglip", 512, 1080, 1920), ("vit", 32, 3, 3), ("clip", 75, 301, 227), ("siglip", 50, 2000, 35)]


@pytest.mark.parametrize("kind,size,h,w", CASES)
def test_kernel_bit_exact_fp32(lib, kind, size, h, w):
    B = 3
    imgs = P.synthetic_u8_images(B, h, w, seed=h * 7 + w)
    ref = np.stack([P.preprocess(im, _cfg(kind, size)) for im in imgs])
    proc = _proc(kind, size)
    out = proc(torch.from_numpy(imgs), dtype=torch.float32)
    assert out.shape == ref.shape and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), ref), np.abs(out.cpu().numpy() - ref).max()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_kernel_half_outputs(lib, dtype):
    imgs = P.synthetic_u8_images(2, 300, 450, seed=3)
    ref = torch.from_numpy(np.stack([P.preprocess(im, _cfg("clip", 224)) for im in imgs])).to(dtype)  # round-to-nearest-even
    out = _proc("clip", 224)(imgs, dtype=dtype)
    assert out.dtype == dtype and torch.equal(out.cpu(), ref)


def test_unaligned_input_and_odd_output_width(lib):
    """Input pointer off the 16-byte grid (byte-load staging) and an output row length that breaks the vector stores."""
    imgs = P.synthetic_u8_images(2, 97, 131, seed=9)
    flat = torch.zeros(imgs.size + 5, dtype=torch.uint8, device="cuda")
    flat[5:] = torch.from_numpy(imgs).reshape(-1).cuda()
    view = flat[5:].view(2, 97, 131, 3)
    from jimm_b200.preprocess import ImagePreprocessor

    proc = ImagePreprocessor(size={"height": 45, "width": 37}, resample=P.BICUBIC, image_mean=(0.1, 0.2, 0.3), image_std=(0.9, 0.8, 0.7))
    cfg = P.PreprocessConfig(height=45, width=37, resample=P.BICUBIC, mean=(0.1, 0.2, 0.3), std=(0.9, 0.8, 0.7))
    ref = np.stack([P.preprocess(im, cfg) for im in imgs])
    for dt in (torch.float32, torch.float16):
        out = proc(view, dtype=dt)
        assert torch.equal(out.cpu(), torch.from_numpy(ref).to(dt))


@pytest.mark.parametrize("name", ["vit", "clip", "siglip"])
def test_golden_fixtures(lib, name, golden_dir):
    """Outputs of transformers' PIL processors committed under tests/golden (made by make_golden_preprocess.py)."""
    z = np.load(os.path.join(golden_dir, f"preprocess_{name}.npz"))
    proc = {"vit": lambda: _proc("vit", 48), "clip": lambda: _proc("clip", 40), "siglip": lambda: _proc("siglip", 64)}[name]()
    for i in range(3):
        out = proc(z[f"img{i}"], dtype=torch.float32)[0]
        assert np.array_equal(out.cpu().numpy(), z[f"out{i}"]), (name, i)


def test_list_of_mixed_sizes_and_errors(lib):
    from jimm_b200.preprocess import ImagePreprocessor

    proc = ImagePreprocessor.clip(64)
    ims = [P.synthetic_u8_images(1, h, w, seed=h)[0] for h, w in ((80, 120), (130, 70), (64, 64))]
    out = proc(ims, dtype=torch.float32)
    ref = np.stack([P.preprocess(im, P.PreprocessConfig.clip(64)) for im in ims])
    assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(ValueError):
        proc(np.zeros((2, 8, 8, 3), np.float32))
    with pytest.raises(ValueError):
        proc(np.zeros((2, 8, 8, 4), np.uint8))
    with pytest.raises(ValueError):
        ImagePreprocessor(size={"height": 8, "width": 8}, resample=1)
    with pytest.raises(ValueError):
        ImagePreprocessor(size={"height": 8, "width": 8}, image_std=(0.5, 0.0, 0.5))
    # centre crop larger than the resized image is outside the path (HF zero-pads): rejected loudly
    big = ImagePreprocessor(size={"height": 16, "width": 16}, crop_size={"height": 32, "width": 32})
    with pytest.raises(ValueError):
        big(np.zeros((1, 20, 20, 3), np.uint8))


def test_front_end_feeds_the_tower(lib):
    """uint8 images -> GPU front-end (fp16 NHWC) -> ViT forward, against oracle front-end -> oracle forward."""
    import jimm_oracle as O
    from jimm_b200.models import VisionTransformer
    from jimm_b200.preprocess import ImagePreprocessor

    imgs = P.synthetic_u8_images(4, 90, 120, seed=11)
    pix = np.stack([P.preprocess(im, P.PreprocessConfig.vit(64)) for im in imgs])
    cfg = O.ViTCfg(num_classes=10, img_size=64, patch_size=16, num_layers=2, num_heads=2, mlp_dim=256, hidden_size=128)
    p = O.random_vit_params(cfg, seed=0)
    with torch.no_grad():
        ref = O.vit_forward(p, cfg, torch.from_numpy(pix))
    m = VisionTransformer(num_classes=10, img_size=64, patch_size=16, num_layers=2, num_heads=2, mlp_dim=256, hidden_size=128,
                          dtype=torch.float16)
    for k, v in p.items():
        m.set_flat_param(k, v.to(torch.float32))
    proc = ImagePreprocessor.vit(64)
    out = m.eval()(proc(imgs, dtype=torch.float16))
    r = float((out.double().cpu() - ref.double()).abs().max() / ref.double().abs().max())
    assert r < 1e-3, r
    assert torch.equal(out.argmax(-1).cpu(), ref.argmax(-1))
    # fp32 front-end output into the same tower gives the same bits: patchify rounds fp32 -> fp16 the same way
    assert torch.equal(m(proc(imgs, dtype=torch.float32)), out)


def test_uint8_frames_through_the_model():
    """Raw uint8 frames into the model (examples/vit_inference.py:27-58 end to end): host frames cross PCIe as bytes and go through the
    front-end + tower inside jimm_vit_forward_host_u8's sliced pipeline; results equal front-end-then-model bit for bit, for host and
    device frames, small (graph-replayed) and sliced (>= 128) batches, and the asynchronous dispatch."""
    from jimm_b200.models import VisionTransformer
    from jimm_b200.preprocess import ImagePreprocessor

    torch.manual_seed(0)
    m = VisionTransformer(num_classes=12, img_size=32, patch_size=8, num_layers=2, num_heads=2, mlp_dim=256, hidden_size=128,
                          dtype=torch.float16).eval().set_max_batch(160)
    proc = ImagePreprocessor.vit(32)
    frames = torch.randint(0, 256, (150, 48, 64, 3), dtype=torch.uint8)
    with pytest.raises(ValueError):
        m(frames[:2])  # no front-end attached
    m.set_preprocessor(proc)
    ref = m(proc(frames.cuda(), dtype=torch.float16))
    out_dev = m(frames.cuda())
    assert torch.equal(out_dev, ref)
    out_host = m(frames.pin_memory())
    assert not out_host.is_cuda and torch.equal(out_host, ref.cpu())
    for _ in range(3):  # small batch: staged + graph replay
        assert torch.equal(m(frames[:5].contiguous()), ref[:5].cpu())
    pend = [m.forward_async(frames.pin_memory()) for _ in range(3)]
    for q in pend:
        assert torch.equal(q.result(), ref.cpu())
    # float inputs still work on the same model, and a frame size the front-end maps elsewhere is rejected
    assert torch.equal(m(proc(frames[:4].cuda(), dtype=torch.float16).cpu()), ref[:4].cpu())
    bad = ImagePreprocessor.vit(40)
    m.set_preprocessor(bad)
    with pytest.raises(ValueError):
        m(frames[:2])


def test_uint8_frames_through_the_dual_towers():
    """Raw uint8 frames + token ids into CLIP / SigLIP (examples/clip_inference.py:35-47 end to end) and into encode_image: equal to
    front-end-then-model bit for bit, for host and device inputs."""
    from jimm_b200.models import CLIP, SigLIP
    from jimm_b200.preprocess import ImagePreprocessor

    torch.manual_seed(0)
    frames = torch.randint(0, 256, (6, 40, 56, 3), dtype=torch.uint8)
    for cls, proc in ((CLIP, ImagePreprocessor.clip(32)), (SigLIP, ImagePreprocessor.siglip(32))):
        m = cls(image_resolution=32, vision_layers=2, vision_width=128, vision_patch_size=8, context_length=16, vocab_size=100,
                transformer_width=64, transformer_heads=1, transformer_layers=2, dtype=torch.float16).eval()
        ids = torch.randint(1, 100, (6, 16), dtype=torch.int32)
        with pytest.raises(ValueError):
            m(frames, ids)  # no front-end attached
        m.set_preprocessor(proc)
        px = proc(frames.cuda(), dtype=torch.float16)
        ref = m(px, ids.cuda())
        assert torch.equal(m(frames.cuda(), ids.cuda()), ref)
        out_host = m(frames.pin_memory(), ids)
        assert not out_host.is_cuda and torch.equal(out_host, ref.cpu())
        emb = m.encode_image(px)
        assert torch.equal(m.encode_image(frames.cuda()), emb)
        assert torch.equal(m.encode_image(frames.pin_memory()), emb.cpu())


def test_integer_size_is_square_for_vit_and_shortest_edge_for_clip(tmp_path):
    """preprocessor_config.json with a legacy integer `size`: ViTImageProcessor / SiglipImageProcessor read it as a square
    (default_to_square=True), CLIPImageProcessor as the shortest edge."""
    import json

    from jimm_b200.preprocess import ImagePreprocessor

    (tmp_path / "vit").mkdir()
    (tmp_path / "clip").mkdir()
    json.dump({"image_processor_type": "ViTImageProcessor", "size": 32, "resample": 2, "image_mean": [0.5] * 3, "image_std": [0.5] * 3},
              open(tmp_path / "vit" / "preprocessor_config.json", "w"))
    json.dump({"image_processor_type": "CLIPImageProcessor", "size": 32, "crop_size": 32, "resample": 3, "image_mean": [0.5] * 3,
               "image_std": [0.5] * 3}, open(tmp_path / "clip" / "preprocessor_config.json", "w"))
    v = ImagePreprocessor.from_pretrained(str(tmp_path / "vit"))
    c = ImagePreprocessor.from_pretrained(str(tmp_path / "clip"))
    assert v.output_size(48, 64) == (32, 32)
    assert c.output_size(48, 64) == (32, 32)  # shortest edge 32 -> 32x43, centre crop 32x32
    x = torch.randint(0, 256, (2, 48, 64, 3), dtype=torch.uint8)
    sq = ImagePreprocessor(size={"height": 32, "width": 32}, resample=2)
    assert torch.equal(v(x), sq(x))
