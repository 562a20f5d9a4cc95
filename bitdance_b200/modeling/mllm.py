"""API mirror of ``modeling/mllm.py`` (inference surface of the image path): ``MLLModel``.

The reference class is the training / evaluation model (``PreTrainedModel`` over an omegaconf config); the image path
touches these methods, mirrored here with the same names, signatures and attribute names:

  * ``gen_image`` / ``gen_image_block_causal``  (mllm.py:257-273, 387-501) — the same algorithm as
    ``BitDanceT2IPipeline.gen_image`` (prefill, 64x AR loop of vision_diffusion_head.sample -> sign -> embed_vision_mlp
    -> + pos-embed -> Qwen3 block), with the tokenizer's special ids taken from ``tokenizer.start_of_image_id`` /
    ``res_{n}_id`` / ``query_{i}_id`` (data/data_utils.py:95-124) when present;
  * ``encode_image(image_list)``  (mllm.py:899-930) — tokenizer encode (``vt_forward`` / ``vt_forward_maxpad``) -> +-1 bits ->
    ``embed_vision_mlp`` -> + 2-D sincos pos-embed: the image-conditioned context (editing / interleaved use);
  * ``decode_image`` (:932-941), ``get_2d_embed`` (:51-60).

* ``forward`` / ``forward_inference`` / ``forward_inference_block_causal`` (mllm.py:157-161, 503-524, 696-897) —
    interleaved text+image inference over a ``sequence_plan``: user text / user images build the
    context (``encode_image``, ``remove_first_user_block`` for the unconditional twin), a model image item runs the block
    generator on that context, a model text item decodes token by token: native Qwen3 step (causal, bf16 stream, paged KV
    that persists across items) -> ``lm_head`` on the tcgen05 GEMM -> ``sample_codebook`` (top-k 1200 / top-p 0.95).

Training (``forward_train``) and the ``parallel_num == 1`` full-causal generators are outside the hot path and raise.
omegaconf is not
needed: build with ``MLLModel.from_pipeline(pipe)`` or ``MLLModel.from_components(...)``; ``config`` is a light namespace
with the fields the mirrored methods read (``vit_patch_size``, ``encoder.vt_forward_func`` / ``max_bs``)."""
from __future__ import annotations

import torch

from ..pipeline import pos_embed_2d
from .utils import remove_first_user_block, sample_codebook

# the sampler settings the reference hard-codes at its call site (mllm.py:852-858)
TEXT_TOP_K, TEXT_TOP_P = 1200, 0.95


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

    # ---- interleaved text + image inference --------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        if self.training:
            return self.forward_train(*args, **kwargs)
        return self.forward_inference(*args, **kwargs)

    @torch.no_grad()
    def forward_inference(self, sequence_plan, text_list, image_list, do_sample: bool = True, max_length_text: int = 128,
                          max_length_vision: int = 64, temperature: float = 1.0, sample_steps: int = 50,
                          image_size=[256, 256], cfg_scale=7.5, *args, **kwargs):
        if self.parallel_num > 1:
            return self.forward_inference_block_causal(sequence_plan, text_list, image_list, do_sample, max_length_text,
                                                       max_length_vision, temperature, sample_steps, image_size, cfg_scale,
                                                       *args, **kwargs)
        return self.forward_inference_full_causal(sequence_plan, text_list, image_list, do_sample, max_length_text,
                                                  max_length_vision, temperature, sample_steps, image_size, cfg_scale,
                                                  *args, **kwargs)

    @torch.no_grad()
    def forward_inference_block_causal(self, sequence_plan, text_list, image_list, do_sample: bool = True,
                                       max_length_text: int = 128, max_length_vision: int = 64, temperature: float = 1.0,
                                       sample_steps: int = 50, image_size=[256, 256], cfg_scale=7.5, *args, **kwargs):
        """mllm.py:696-897. ``sequence_plan``: list of ``{"type": "text"|"image", "from": "user"|"model"}``; user items
        consume ``text_list`` / ``image_list`` front to back (both lists are popped, as in the reference). Returns
        ``{"generated_text": [str, ...], "generated_image": [Tensor[1, 3, H, W], ...]}``.

        What is the reference's and kept: the context bookkeeping (image items contribute ``<|vision_start|> <|res_h|>
        <|res_w|>`` + content + ``<|vision_end|>``; the unconditional context drops the first user turn of every text and
        keeps the images), the KV cache of the conditional stream persisting across items, ``context_embed`` being cleared
        only when a generated text ends with ``<|im_end|>`` (so a text cut off by ``max_length_text`` is followed by the
        whole context again on top of the cache — reproduced literally), the sampler (top-k 1200 / top-p 0.95 on the
        global torch generator) and the image generator (= ``gen_image_block_causal`` on the accumulated context).

        Where the reference cannot be followed: its text branch raises — without a cache at mllm.py:798 (it subscripts
        ``past_key_values`` = None before the first pass), with one (after a generated image) on the second token, because
        the sampled token's 2-D ``(1, hidden)`` embedding is fed back as ``inputs_embeds`` (:857), and it would concatenate
        1-D token tensors along dim 1 at the end (:867); tests/test_interleaved_vs_reference.py asserts both failures on
        the unmodified reference. Here the evident intent runs (one ``[1, 1, hidden]`` step per token). The reference
        ignores its ``do_sample`` / ``temperature`` arguments (hard-coded ``True`` / ``1.0`` at the call site); here they
        are honoured, and their defaults are the reference's constants. Its wasted unconditional LLM pass during TEXT
        decoding (computed, never read) is skipped. Not supported (raises): a generated image when the persistent cache is
        not empty, or anything generated after an image (the reference would re-feed the context on top of the image
        generation's cache)."""
        tok, dev, eng = self.tokenizer, self.device, self.engine
        embed = self.llm_model.model.embed_tokens
        emb_ids = lambda ids: embed(torch.tensor(list(ids), device=dev, dtype=torch.long))
        use_cfg = cfg_scale > 1.0
        vps = self.config.vit_patch_size
        out = {"generated_text": [], "generated_image": []}
        ctx, ctx_un = [], []
        cache, image_done = None, False
        capacity = self._interleaved_capacity(sequence_plan, text_list, image_list, max_length_text)
        for item in sequence_plan:
            kind, src = item["type"], item["from"]
            if kind not in ("text", "image") or src not in ("user", "model"):
                raise ValueError(f"bad sequence_plan item {item!r}")
            if image_done and src == "model":
                raise NotImplementedError("a generated image must be the last generated item of the plan")
            if kind == "image":
                gh, gw = image_size[0] // vps, image_size[1] // vps
                start = emb_ids([self._special_id("start_of_image_id", "<|vision_start|>"),
                                 self._special_id(f"res_{gh}_id", f"<|res_{gh}|>"),
                                 self._special_id(f"res_{gw}_id", f"<|res_{gw}|>")])
                ctx.append(start)
                ctx_un.append(start)
            if src == "user" and kind == "text":
                text = text_list.pop(0)
                ctx.append(emb_ids(tok.encode(text)))
                if use_cfg:
                    ctx_un.append(emb_ids(tok.encode(remove_first_user_block(text))))
            elif src == "user":
                img = image_list.pop(0).to(dev)
                pre = self.encode_image([img])[0]
                end = emb_ids([self._special_id("end_of_image_id", "<|vision_end|>")])
                ctx += [pre, end]
                if use_cfg:
                    ctx_un += [pre, end]
            elif kind == "image":
                if cache is not None:
                    raise NotImplementedError("image generation on top of a non-empty interleaved KV cache")
                if max_length_vision != gh * gw:
                    raise ValueError(f"max_length_vision ({max_length_vision}) must equal the token count of image_size "
                                     f"({gh * gw})")
                queries = emb_ids([self._special_id(f"query_{i}_id", f"<|query_{i}|>") for i in range(1, self.parallel_num)])
                # context = everything so far INCLUDING the three start tokens appended above; the engine wants them split
                cat = lambda parts: torch.cat([p.to(torch.float32) for p in parts], dim=0)
                start_emb = torch.cat([ctx[-1].to(torch.float32), queries.to(torch.float32)], dim=0)
                cond = cat(ctx[:-1]) if len(ctx) > 1 else start_emb[:0]
                uncond = (cat(ctx_un[:-1]) if len(ctx_un) > 1 else start_emb[:0]) if use_cfg else None
                ctx.append(queries)
                ctx_un.append(queries)
                tokens, self.last_packed_tokens = eng.gen_tokens(
                    cond.to(torch.bfloat16), None if uncond is None else uncond.to(torch.bfloat16),
                    start_emb.to(torch.bfloat16), h=gh, w=gw, num_images=1, guidance_scale=cfg_scale,
                    num_sampling_steps=sample_steps)
                out["generated_image"].append(self.decode_image(tokens, [gh, gw], ps=self.ps))
                image_done = True
            else:
                end_id = self._special_id("im_end_id", "<|im_end|>")
                if cache is None:
                    cache = eng.llm.new_cache(1, capacity)
                ids, ended = self._decode_text(torch.cat([p.to(torch.float32) for p in ctx], dim=0), cache, end_id,
                                               max_length_text, do_sample, temperature)
                if ended:
                    ctx = []    # the reference clears the conditional context only (mllm.py:865)
                full = torch.stack(ids)
                words = tok.convert_ids_to_tokens(full.tolist(), skip_special_tokens=True)
                out["generated_text"].append(tok.convert_tokens_to_string([t for t in words if t is not None]))
                self.last_text_ids = full
        return out

    def _interleaved_capacity(self, plan, text_list, image_list, max_length_text) -> int:
        """Upper bound of the tokens the persistent conditional cache can receive, computed before anything is consumed:
        every generated text feeds at most the whole context so far and decodes at most ``max_length_text`` tokens."""
        vps = self.config.vit_patch_size
        texts, images = list(text_list), list(image_list)
        n, n_gen = 0, 0
        for item in plan:
            if item["type"] == "image":
                n += 3
            if item["from"] == "model":
                n_gen += 1
                n += max_length_text if item["type"] == "text" else self.parallel_num
            elif item["type"] == "text" and texts:
                n += len(self.tokenizer.encode(texts.pop(0)))
            elif item["type"] == "image" and images:
                img = images.pop(0)
                n += (img.shape[-2] // vps) * (img.shape[-1] // vps) + 1
        cap = (n * max(1, n_gen) + 64 + 63) // 64 * 64
        limit = self.engine.llm.rope_cos.shape[0] // 64 * 64
        if cap > limit:
            raise ValueError(f"interleaved plan may need {cap} positions, the RoPE table holds {limit}")
        return cap

    @torch.no_grad()
    def _decode_text(self, context: torch.Tensor, cache, end_id: int, max_length: int, do_sample: bool, temperature: float):
        """context [L, D] -> (list of 0-d token tensors incl. the end token when sampled, ended?). One causal pass over the
        context on top of ``cache``, then one [1, 1, D] step per token: Qwen3 (bd_llm_forward, bf16 stream) -> last hidden ->
        lm_head (bd_gemm_bf16) -> sample_codebook."""
        llm, embed, lm_head = self.engine.llm, self.llm_model.model.embed_tokens, self.llm_model.lm_head
        D = context.shape[-1]
        x = context.to(torch.bfloat16).view(1, -1, D).contiguous()
        ids = []
        for step in range(max_length):
            hidden = llm.forward(x, cache, 0, 1, causal=True)[:, -1]            # [1, D] bf16 (final RMSNorm applied)
            # bf16 logits (a bf16 Linear's output), sampler math in fp32 — what softmax / cumsum do under autocast
            token, nxt = sample_codebook(lm_head(hidden).float(), "text", embed, do_sample=do_sample,
                                         temperature=temperature, top_k=TEXT_TOP_K, top_p=TEXT_TOP_P)
            ids.append(token[0])
            if int(token[0]) == end_id:
                return ids, True
            x = nxt.to(torch.bfloat16).view(1, 1, D).contiguous()
        return ids, False

    # ---- out of scope -------------------------------------------------------------------------------------------------
    def forward_train(self, *a, **k):
        raise NotImplementedError("MLLModel.forward_train (training) is outside the image hot path")

    def forward_inference_full_causal(self, *a, **k):
        raise NotImplementedError("parallel_num == 1 (token-by-token image) inference is not on the BitDance-14B path")
