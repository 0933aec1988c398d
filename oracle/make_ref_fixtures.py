"""Pin the oracle to the reference ITSELF: run pythoncrazy/jimm (flax/JAX) on the committed golden checkpoints and store its outputs.

TEST INFRASTRUCTURE (see oracle/jimm_oracle.py header).  JAX/flax are not installable in the build image, so this script cannot run
there; it is the ready-to-run kit for any host that has them (`pip install "jax==0.6.2" "flax==0.10.6" jaxtyping safetensors`):

    python oracle/make_ref_fixtures.py [--reference /path/to/jimm/src] [--dtypes float32,bfloat16]

For every tests/golden/<name>/ it loads `model.safetensors` + `config.json` through the reference's own local-file branch
(src/jimm/common/utils.py:74-90 via <Model>.from_pretrained), runs the reference forward on the inputs stored in `io.npz`
(NHWC images, int32 tokens -- exactly what tests/test_parity_gpu.py feeds the CUDA path) and writes `ref_io.npz` next to it with
    jimm_<dtype>_logits / jimm_<dtype>_image_embeds / jimm_<dtype>_text_embeds        (fp32 arrays)
plus the jax / flax / jimm versions.  tests/test_ref_fixtures.py consumes the file when present: the oracle must then match the
reference to 1e-5 (fp32) and the flax-bf16 restatement to the bf16 rounding noise, and the `-m gpu` parity tests gain an
`against="reference"` row in PARITY.md.  Until the file exists those tests skip with a loud reason and DESIGN.md says
"parity unpinned at the flax boundary"."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("JIMM_REFERENCE_SRC", "/root/reference/src"))
    ap.add_argument("--dtypes", default="float32,bfloat16")
    args = ap.parse_args()
    try:
        import flax
        import jax
        import jax.numpy as jnp
    except ImportError as e:  # the build image: say so and stop, never fake the fixtures
        print(f"make_ref_fixtures: {e}; run this on a host with jax==0.6.2 and flax==0.10.6 (reference uv.lock:332,586)", file=sys.stderr)
        return 2
    jax.config.update("jax_platform_name", "cpu")
    sys.path.insert(0, args.reference)
    from jimm.models.clip import CLIP
    from jimm.models.siglip import SigLIP
    from jimm.models.vit import VisionTransformer

    dts = {"float32": jnp.float32, "bfloat16": jnp.bfloat16}
    meta = {"jax": jax.__version__, "flax": flax.__version__, "reference": args.reference}
    for name, cls in (("tiny_vit", VisionTransformer), ("tiny_clip", CLIP), ("tiny_siglip", SigLIP)):
        d = os.path.join(GOLDEN, name)
        io = np.load(os.path.join(d, "io.npz"))
        out = {}
        for dn in args.dtypes.split(","):
            model = cls.from_pretrained(os.path.join(d, "model.safetensors"), dtype=dts[dn])
            model.eval()
            img = jnp.asarray(io["images"], dtype=dts[dn])
            if cls is VisionTransformer:
                out[f"jimm_{dn}_logits"] = np.asarray(model(img), dtype=np.float32)
            else:
                txt = jnp.asarray(io["tokens"], dtype=jnp.int32)
                out[f"jimm_{dn}_image_embeds"] = np.asarray(model.encode_image(img), dtype=np.float32)
                out[f"jimm_{dn}_text_embeds"] = np.asarray(model.encode_text(txt), dtype=np.float32)
                out[f"jimm_{dn}_logits"] = np.asarray(model(img, txt), dtype=np.float32)
        np.savez_compressed(os.path.join(d, "ref_io.npz"), meta=json.dumps(meta), **out)
        print(name, {k: v.shape for k, v in out.items()})
    return 0


if __name__ == "__main__":
    sys.exit(main())
