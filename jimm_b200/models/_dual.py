"""Shared machinery of the CLIP / SigLIP mirrors: text-tower parameter tree, native config, multi-GPU contrastive head."""

from __future__ import annotations

import torch

from .. import _lib, nn
from ..common.transformer import Transformer, g_wrap
from ..common.vit import _NativeOwner, tower_config_fields


class DualTower(_NativeOwner, nn.Module):
    """Base of CLIP and SigLIP: holds the shared attributes and the encode / call plumbing."""

    _kind = None

    def _init_common(self, image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
                     transformer_width, transformer_heads, transformer_layers, dtype):
        nn.Module.__init__(self)
        self._native_init(dtype)
        for k, v in dict(image_resolution=image_resolution, vision_layers=vision_layers, vision_width=vision_width,
                         vision_patch_size=vision_patch_size, context_length=context_length, vocab_size=vocab_size,
                         transformer_width=transformer_width, transformer_heads=transformer_heads,
                         transformer_layers=transformer_layers, dtype=dtype).items():
            object.__setattr__(self, k, v)
        object.__setattr__(self, "_comm_mode", None)

    def _text_config(self, cfg: _lib.Config, *, act, causal, pool, head_bias, eps_outer):
        cfg.ctx_len, cfg.vocab, cfg.t_width = self.context_length, self.vocab_size, self.transformer_width
        cfg.t_heads, cfg.t_layers, cfg.t_mlp = self.transformer_heads, self.transformer_layers, self.transformer_width * 4
        cfg.t_act, cfg.t_causal, cfg.t_pool, cfg.t_head_bias = act, causal, pool, head_bias
        cfg.t_eps_outer = eps_outer
        cfg.t_eps_block = 1e-6  # Transformer default (common/transformer.py:142); CLIP does not forward 1e-5 (models/clip.py:92-104)
        cfg.compute_dtype = self._compute_dtype
        return cfg

    # ---- reference API ----
    def encode_image(self, image) -> torch.Tensor:
        return self.native(image.shape[0]).vision(image, encode=True)

    def encode_text(self, text) -> torch.Tensor:
        return self.native(text.shape[0]).text(text)

    def __call__(self, image, text) -> torch.Tensor:
        """Similarity logits.  Single process: [B_img, B_txt].  Under torch.distributed (one process per GPU, batch sharded
        over ranks like the reference's P("batch") inputs, examples/clip_inference.py:41-42): this rank's row block
        [B_local, world*B_local], embeddings exchanged over NVLink peer memory inside the fused logits kernel."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and self._comm_mode != "off":
            return self._call_distributed(image, text)
        return self.native(max(image.shape[0], text.shape[0]), require=True).dual(image, text)

    def set_comm(self, mode: str):
        """'peer' (default, fused NVLink peer-store kernel) | 'nccl' (torch.distributed all_gather baseline) | 'off'."""
        if mode not in ("peer", "nccl", "off"):
            raise ValueError(mode)
        object.__setattr__(self, "_comm_mode", mode)
        return self

    def _call_distributed(self, image, text) -> torch.Tensor:
        import torch.distributed as dist

        B = image.shape[0]
        n = self.native(B, require=True)
        x, ids = n._prep_images(image), n._prep_ids(text)
        host_in = not x.is_cuda and not ids.is_cuda
        with torch.cuda.device(n.device):
            cur = torch.cuda.current_stream(n.device)
            # host inputs: the ids go first (H2D copies share one engine), the images follow on a side stream and land while
            # the text tower runs -- the same overlap as the single-GPU host path (csrc/model.cu jimm_dual_forward_host)
            ids_d = ids if ids.is_cuda else ids.to(n.device, non_blocking=True)
            if x.is_cuda:
                x_d = x
            else:
                side = n.side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    x_d = x.to(n.device, non_blocking=True)
            if x.is_cuda:
                ie, te = n.dual_encode(x_d, ids_d)  # both towers concurrently (text forked onto the library's side stream)
            elif x.dtype == torch.uint8:
                # raw frames are a quarter of the bytes: take the short copy up front and run the two towers concurrently
                cur.wait_stream(side)
                x_d.record_stream(cur)
                ie, te = n.dual_encode(x_d, ids_d)
            else:
                te = n.text(ids_d)
                cur.wait_stream(side)
                x_d.record_stream(cur)
                ie = n.vision(x_d, encode=True)
            out = self._distributed_logits(n, ie, te, B)
            if not host_in:
                return out
            out_h = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
            out_h.copy_(out, non_blocking=True)
            cur.synchronize()
            return out_h

    def _distributed_logits(self, n, ie, te, B) -> torch.Tensor:
        import torch.distributed as dist

        if (self._comm_mode or "peer") == "peer":
            if n._comm is None or n._comm[2] < B:
                n.comm_setup(max(B, n.max_batch))
            return n.comm_logits(ie, te)
        # NCCL baseline: normalise locally, all-gather the packed [B, 2E] buffer, logits for the local rows
        world = dist.get_world_size()
        i_n = ie / torch.linalg.norm(ie, dim=-1, keepdim=True)
        t_n = te / torch.linalg.norm(te, dim=-1, keepdim=True)
        gathered = torch.empty((world * B, t_n.shape[1]), dtype=torch.float32, device=n.device)
        dist.all_gather_into_tensor(gathered, t_n.contiguous())
        scale = self.logit_scale.to(n.device).reshape(1)
        bias = self.logit_bias.to(n.device).reshape(1) if "logit_bias" in self._params else None
        import ctypes as C

        out = torch.empty((B, world * B), dtype=torch.float32, device=n.device)
        _lib.check(n.lib.jimm_k_logits(C.c_void_p(i_n.data_ptr()), C.c_void_p(gathered.data_ptr()), C.c_void_p(scale.data_ptr()),
                                       C.c_void_p(bias.data_ptr()) if bias is not None else None, C.c_void_p(out.data_ptr()), B,
                                       world * B, t_n.shape[1], world * B, C.c_void_p(torch.cuda.current_stream(n.device).cuda_stream)))
        return out


def build_text_tower(model: DualTower, g, *, head_bias: bool, layernorm_epsilon, use_quick_gelu: bool, attn_mask):
    Dt, T, V = model.transformer_width, model.context_length, model.vocab_size
    model.add_child("text_model", Transformer(width=Dt, mlp_dim=Dt * 4, layers=model.transformer_layers, num_heads=model.transformer_heads,
                                              dropout_rate=0.0, attn_mask=attn_mask, use_quick_gelu=use_quick_gelu,
                                              layernorm_epsilon=layernorm_epsilon, rngs=g_wrap(g)))
    model.add_child("token_embedding", nn.Embed(V, Dt, rngs=g_wrap(g)))
    model.add_param("positional_embedding", nn.truncated_normal(g, (T, Dt), 0.02))
    model.add_child("ln_final", nn.LayerNorm(Dt))
    model.add_child("text_projection", nn.Linear(Dt, Dt, use_bias=head_bias, rngs=g_wrap(g)))
