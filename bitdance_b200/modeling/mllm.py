"""API mirror of ``modeling/mllm.py`` (inference surface of the image path): ``MLLModel``.

The reference class is the training / evaluation model (``PreTrainedModel`` over an omegaconf config); the image path
touches five of its methods, mirrored here with the same names, signatures and attribute names:

  * ``gen_image`` / ``gen_image_block_causal``  (mllm.py:257-273, 387-501) — the same algorithm as
    ``BitDanceT2IPipeline.gen_image`` (prefill, 64x AR loop of vision_diffusion_head.sample -> sign -> embed_vision_mlp
    -> + pos-embed -> Qwen3 block), with the tokenizer's special ids taken from ``tokenizer.start_of_image_id`` /
    ``res_{n}_id`` / ``query_{i}_id`` (data/data_utils.py:95-124) when present;
  * ``encode_image(image_list)``  (mllm.py:899-930) — tokenizer encode (``vt_forward`` / ``vt_forward_maxpad``) -> +-1 bits ->
    ``embed_vision_mlp`` -> + 2-D sincos pos-embed: the image-conditioned context (editing / interleaved use);
  * ``decode_image`` (:932-941), ``get_2d_embed`` (:51-60).

Training (``forward_train``), text generation (``forward_inference*``: lm_head + top-k/p sampling) and the
``parallel_num == 1`` full-causal generator are outside the hot path and raise. omegaconf is not
needed: build with ``MLLModel.from_pipeline(pipe)`` or ``MLLModel.from_components(...)``; ``config`` is a light namespace
with the fields the mirrored methods read (``vit_patch_size``, ``encoder.vt_forward_func`` / ``max_bs``)."""
from __future__ import annotations

import types

import torch

from ..pipeline import pos_embed_2d


class _Cfg(dict):
    """dict with attribute access and ``.get`` — the subset of omegaconf's DictConfig the mirrored methods use"""
    __getattr__ = dict.__getitem__


class MLLModel:
    def __init__(self, config=None):
        raise NotImplementedError("construct with MLLModel.from_pipeline(pipe) or MLLModel.from_components(...): the "
                                  "reference constructor downloads checkpoints through omegaconf configs (mllm.py:23-29)")

    @classmethod
    def from_components(cls, *, tokenizer, llm_model, engine, vision_encoder, vision_diffusion_head, embed_vision_mlp,
                        vit_patch_size: int, device="cuda", vt_forward_func: str = "group", max_bs: int = 32):
        self = object.__new__(cls)
        self.device, self.tokenizer, self.llm_model, self.engine = device, tokenizer, llm_model, engine
        self.vision_encoder, self.vision_diffusion_head = vision_encoder, vision_diffusion_head
        self.embed_vision_mlp = embed_vision_mlp
        self.parallel_num, self.ps, self.hidden_size = engine.pn, engine.ps, engine.D
        self.pos_embed_1d = engine.pos_1d
        self.training = False
        self.config = _Cfg(vit_patch_size=vit_patch_size,
                           encoder=_Cfg(vt_forward_func=vt_forward_func, max_bs=max_bs),
                           head=_Cfg(vision_pred=_Cfg(parallel_num=engine.pn, type="diffusion_parallel_x")))
        return self

    @classmethod
    def from_pipeline(cls, pipe, **kw):
        """Share the engine and modules of a ``BitDanceT2IPipeline``."""
        return cls.from_components(tokenizer=pipe.tokenizer, llm_model=pipe.llm_model, engine=pipe.engine,
                                   vision_encoder=pipe.ae, vision_diffusion_head=pipe.vision_head,
                                   embed_vision_mlp=pipe.embed_vision_mlp, vit_patch_size=pipe.vae_patch_size,
                                   device=pipe.device, **kw)

    # ---- helpers ------------------------------------------------------------------------------------------------------
    def get_2d_embed(self, h, w, ps=1):
        return pos_embed_2d(self.pos_embed_1d, h, w, ps)

    def _special_id(self, alias: str, token: str) -> int:
        tok = self.tokenizer
        if hasattr(tok, alias):  # the reference's aliased tokenizer (set_special_token_aliases etc.)
            return int(getattr(tok, alias))
        return int(tok.convert_tokens_to_ids(token))

    # ---- generation ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def gen_image(self, cond_prompt, uncond_prompt=None, guidance_scale: float = 1.0, num_sampling_steps: int = 50,
                  max_length: int = 64, num_images: int = 1, image_size=[256, 256], show_progress: bool = False):
        if self.parallel_num > 1:
            return self.gen_image_block_causal(cond_prompt, uncond_prompt, guidance_scale, num_sampling_steps, max_length,
                                               num_images, image_size, show_progress)
        return self.gen_image_full_causal(cond_prompt, uncond_prompt, guidance_scale, num_sampling_steps, max_length,
                                          num_images, image_size, show_progress)

    def gen_image_full_causal(self, *a, **k):
        raise NotImplementedError("parallel_num == 1 (token-by-token) generation is not on the BitDance-14B image path")

    @torch.no_grad()
    def gen_image_block_causal(self, cond_prompt, uncond_prompt=None, guidance_scale: float = 1.0,
                               num_sampling_steps: int = 50, max_length: int = 64, num_images: int = 1,
                               image_size=[256, 256], show_progress: bool = False):
        tok, dev = self.tokenizer, self.device
        embed = self.llm_model.model.embed_tokens
        vps = self.config.vit_patch_size
        h, w = image_size[0] // vps, image_size[1] // vps
        if max_length != h * w:
            raise ValueError(f"max_length ({max_length}) must equal the token count of image_size ({h * w})")
        ids = lambda s: torch.tensor(tok.encode(s), device=dev, dtype=torch.long)
        cond_emb = embed(ids(cond_prompt))
        uncond_emb = embed(ids(uncond_prompt)) if guidance_scale > 1.0 else None
        start = [self._special_id("start_of_image_id", "<|vision_start|>"), self._special_id(f"res_{h}_id", f"<|res_{h}|>"),
                 self._special_id(f"res_{w}_id", f"<|res_{w}|>")]
        start += [self._special_id(f"query_{i}_id", f"<|query_{i}|>") for i in range(1, self.parallel_num)]
        start_emb = embed(torch.tensor(start, device=dev, dtype=torch.long))
        tokens, self.last_packed_tokens = self.engine.gen_tokens(
            cond_emb, uncond_emb, start_emb, h=h, w=w, num_images=num_images, guidance_scale=guidance_scale,
            num_sampling_steps=num_sampling_steps)
        return self.decode_image(tokens, [h, w], ps=self.ps)

    # ---- image-conditioned context ------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_image(self, image_list, packed_label_indexes_vision=None):
        """image_list: tensors [1, 3, H, W] (mixed sizes allowed). Returns (packed_vision_embedding [sum tokens, D] fp32 =
        embed_vision_mlp(bits) + pos-embed, packed_vision_latents [sum tokens, C] of +-1)."""
        if self.training and packed_label_indexes_vision is not None:
            raise NotImplementedError("training-time bit-flip perturbation (mllm.py:910-916) is out of scope")
        enc = self.config.encoder
        if enc.get("vt_forward_func", "group") == "maxpad":
            latents = self.vision_encoder.vt_forward_maxpad(image_list=image_list, max_bs=enc.get("max_bs", 32))
        else:
            latents = self.vision_encoder.vt_forward(image_list=image_list, max_bs=enc.get("max_bs", 32), ps=self.ps)
        emb = self.embed_vision_mlp(latents)                       # bf16 [N, D]
        vps = self.config.vit_patch_size
        pos = torch.cat([self.get_2d_embed(img.shape[-2] // vps, img.shape[-1] // vps, ps=self.ps) for img in image_list], dim=0)
        return emb.float() + pos, latents.clone().detach()         # bf16 + fp32 pos-embed -> fp32, as under autocast

    def decode_image(self, image_latents, image_size=None, ps=1):
        if image_size is None:
            h = w = int(image_latents.size(1) ** 0.5)
        else:
            h, w = image_size
        if ps != self.engine.ps:
            raise ValueError("ps must match the head's parallel block size")
        return self.engine.ae.decode_tokens(image_latents.to(torch.float32).contiguous(), h, w, ps)

    # ---- out of scope -------------------------------------------------------------------------------------------------
    def forward(self, *a, **k):
        raise NotImplementedError("MLLModel.forward (training / interleaved text inference) is outside the image hot path")

    forward_train = forward_inference = forward_inference_full_causal = forward_inference_block_causal = forward
