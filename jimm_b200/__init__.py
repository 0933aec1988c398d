"""jimm_b200 -- B200-native (sm_100a) drop-in for the inference forward path of pythoncrazy/jimm.

    from jimm_b200.models import VisionTransformer, CLIP, SigLIP      # == jimm.models
    from jimm_b200.common.vit import VisionTransformerBase            # == jimm.common.vit
    from jimm_b200 import Rngs                                        # stand-in for flax.nnx.Rngs
"""

from .nn import Rngs  # noqa: F401

__version__ = "0.1.0"
