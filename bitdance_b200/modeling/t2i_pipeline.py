"""API mirror of ``modeling/t2i_pipeline.py``: ``BitDanceT2IPipeline``.

Same constructor (``model_path, device``), the same files read (HF ``config.json`` + tokenizer + safetensors shards,
``ae_config.json``/``ae.safetensors``, ``vision_head_config.json``/``vision_head.safetensors``, ``projector.safetensors``;
reference :45-77), the same ``generate`` / ``gen_image`` / ``decode_image`` signatures and the same attributes callers
use (``vae_patch_size``, ``parallel_num``, ``ps``, ``tokenizer``, ``ae``, ``vision_head``, ``embed_vision_mlp``) — so
``example_t2i.py`` and ``eval/*.py`` run unchanged against it. The arithmetic is the B200-native engine
(bitdance_b200/pipeline.py); there is no eager / CPU fallback."""
from __future__ import annotations

import glob
import json
import os
import types

import numpy as np
import torch
from torch import nn

from ..llm import LlmRunner
from ..pipeline import T2IEngine
from .utils import MLPconnector
from .vision_encoder.autoencoder import VQModel
from .vision_head.flow_head_parallel_x import DiffHead

IMAGE_SIZE_LIST = [
    # 1024px area
    [2048, 512], [1920, 512], [1536, 640], [1280, 768], [1152, 896], [1024, 1024], [896, 1152], [768, 1280],
    [640, 1536], [512, 1920], [512, 2048],
    # 512px area
    [1024, 256], [896, 256], [640, 384], [512, 512], [384, 640], [256, 896], [256, 1024],
]


def _load_llm_state_dict(model_path: str) -> dict:
    from safetensors.torch import load_file
    files = sorted(glob.glob(os.path.join(model_path, "model*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no model*.safetensors under {model_path}")
    sd = {}
    for f in files:
        sd.update(load_file(f))
    return sd


class _EmbedOnlyLLM:
    """What callers touch on ``pipe.llm_model``: ``.model.embed_tokens`` and — interleaved text generation only,
    modeling/mllm.py:845 — ``.lm_head`` (the decoder layers live in the native runner). ``lm_head_weight`` ([vocab, hidden],
    may stay on the host) is prepacked for the GEMM on first use; with tied embeddings it is the embedding table."""

    def __init__(self, embed_weight: torch.Tensor, lm_head_weight: torch.Tensor | None = None):
        emb = nn.Embedding.from_pretrained(embed_weight, freeze=True)
        self.model = types.SimpleNamespace(embed_tokens=emb)
        self._lm_head_weight = lm_head_weight
        self._lm_head = None

    @property
    def lm_head(self):
        if self._lm_head is None:
            if self._lm_head_weight is None:
                raise AttributeError("this model directory has no lm_head.weight (and tie_word_embeddings is off): "
                                     "text generation is unavailable")
            from ..llm import LmHead
            self._lm_head = LmHead(self._lm_head_weight, device=self.model.embed_tokens.weight.device)
            self._lm_head_weight = None
        return self._lm_head


class BitDanceT2IPipeline:
    def __init__(self, model_path, device='cuda'):
        from safetensors.torch import load_file as load_sft
        from transformers import AutoTokenizer
        self.device = device
        self.tokenizer = AutoTokenizer.from_pretrained(model_path)
        with open(os.path.join(model_path, "config.json")) as f:
            self.llm_config = json.load(f)
        self.hidden_size = self.llm_config["hidden_size"]
        sd = _load_llm_state_dict(model_path)
        head_w = sd.get("lm_head.weight")
        if head_w is None and self.llm_config.get("tie_word_embeddings", False):
            head_w = sd["model.embed_tokens.weight"]
        self.llm_model = _EmbedOnlyLLM(sd["model.embed_tokens.weight"].to(device, torch.bfloat16), head_w)
        llm = LlmRunner(sd, self.llm_config, device=device,
                        max_positions=max(8192, int(self.llm_config.get("max_position_embeddings", 8192))))
        del sd

        with open(os.path.join(model_path, 'ae_config.json')) as f:
            self.ae_config = json.load(f)
        self.ae = VQModel(**self.ae_config).eval()
        self.ae.load_state_dict(load_sft(os.path.join(model_path, 'ae.safetensors')), strict=True, assign=True)
        self.ae.to(device)
        self.vae_patch_size = 2 ** (len(self.ae_config['ddconfig']['ch_mult']) - 1)

        with open(os.path.join(model_path, 'vision_head_config.json')) as f:
            self.vision_head_config = json.load(f)
        self.vision_head = DiffHead(**self.vision_head_config).eval()
        self.vision_head.load_state_dict(load_sft(os.path.join(model_path, 'vision_head.safetensors')), strict=True,
                                         assign=True)
        self.vision_head.to(device)
        self.parallel_num = self.vision_head_config['parallel_num']
        print(f'use {self.parallel_num}-token parallel prediction per step')
        self.ps = int(self.parallel_num ** 0.5)

        self.embed_vision_mlp = MLPconnector(self.ae_config['ddconfig']['z_channels'], self.hidden_size,
                                             "gelu_pytorch_tanh")
        self.embed_vision_mlp.load_state_dict(load_sft(os.path.join(model_path, 'projector.safetensors')), strict=True,
                                              assign=True)
        self.embed_vision_mlp.to(device)
        self._finish(llm)

    @classmethod
    def from_components(cls, *, tokenizer, embed_weight, llm: LlmRunner, ae: VQModel, vision_head: DiffHead,
                        embed_vision_mlp: MLPconnector, ae_config: dict, vision_head_config: dict, device='cuda',
                        lm_head_weight: torch.Tensor | None = None):
        """Assemble a pipeline from already-built parts (tests, synthetic weights)."""
        self = object.__new__(cls)
        self.device, self.tokenizer = device, tokenizer
        self.hidden_size = llm.cfg["hidden_size"]
        self.llm_config = dict(llm.cfg)
        self.llm_model = _EmbedOnlyLLM(embed_weight.to(device, torch.bfloat16), lm_head_weight)
        self.ae_config, self.vision_head_config = ae_config, vision_head_config
        self.ae, self.vision_head, self.embed_vision_mlp = ae, vision_head, embed_vision_mlp
        self.vae_patch_size = 2 ** (len(ae_config['ddconfig']['ch_mult']) - 1)
        self.parallel_num = vision_head_config['parallel_num']
        self.ps = int(self.parallel_num ** 0.5)
        self._finish(llm)
        return self

    @classmethod
    def from_engine(cls, engine: T2IEngine, *, tokenizer, embed_weight, device='cuda',
                    lm_head_weight: torch.Tensor | None = None):
        """The public surface over an already-built engine (bench.py: synthetic 14B weights generated on the device, so
        there are no nn.Module copies of them): ``generate`` / ``gen_image`` / ``decode_image`` work as usual; ``ae`` /
        ``vision_head`` / ``embed_vision_mlp`` expose only their native runners."""
        self = object.__new__(cls)
        self.device, self.tokenizer, self.engine = device, tokenizer, engine
        self.hidden_size = engine.D
        self.llm_config = dict(engine.llm.cfg)
        self.llm_model = _EmbedOnlyLLM(embed_weight.to(device, torch.bfloat16), lm_head_weight)
        self.ae = types.SimpleNamespace(runner=engine.ae)
        self.vision_head = types.SimpleNamespace(runner=engine.head)
        self.embed_vision_mlp = None
        self.vae_patch_size, self.parallel_num, self.ps = engine.vae_patch_size, engine.pn, engine.ps
        self.pos_embed_1d = engine.pos_1d
        return self

    def _finish(self, llm: LlmRunner):
        p = self.embed_vision_mlp
        self.engine = T2IEngine(llm, self.vision_head.runner, self.ae.runner, p.fc1.weight, p.fc1.bias, p.fc2.weight,
                                p.fc2.bias, parallel_num=self.parallel_num, vae_patch_size=self.vae_patch_size,
                                device=self.device)
        self.pos_embed_1d = self.engine.pos_1d

    def get_2d_embed(self, h, w, ps=1):
        from ..pipeline import pos_embed_2d
        return pos_embed_2d(self.pos_embed_1d, h, w, ps)

    @torch.no_grad()
    def generate(self, prompt: str, height: int = 1024, width: int = 1024, num_sampling_steps: int = 50,
                 guidance_scale: float = 7.5, num_images: int = 1, seed: int = 1234):
        from PIL import Image
        if seed is not None:
            from transformers import set_seed
            set_seed(seed)
        image_size = [height, width]
        if image_size not in IMAGE_SIZE_LIST:
            raise ValueError(f"image_size {image_size} is not supported. Please choose from {IMAGE_SIZE_LIST}")
        max_length = (height // self.vae_patch_size) * (width // self.vae_patch_size)
        imgs = self.gen_image(cond_prompt=f"<|im_start|>user\n{prompt}<|im_end|>\n<|im_start|>assistant\n",
                              uncond_prompt="<|im_start|>assistant\n", guidance_scale=guidance_scale,
                              num_sampling_steps=num_sampling_steps, num_images=num_images, image_size=image_size,
                              max_length=max_length, show_progress=True)
        u8 = torch.clamp(127.5 * imgs + 128.0, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8).numpy()
        return [Image.fromarray(np.ascontiguousarray(u8[i])) for i in range(u8.shape[0])]

    @torch.no_grad()
    def gen_image(self, cond_prompt, uncond_prompt=None, guidance_scale: float = 1.0, num_sampling_steps: int = 50,
                  max_length: int = 64, num_images: int = 1, image_size=[256, 256], show_progress: bool = False):
        tok, dev = self.tokenizer, self.device
        embed = self.llm_model.model.embed_tokens
        h, w = image_size[0] // self.vae_patch_size, image_size[1] // self.vae_patch_size
        if max_length != h * w:
            raise ValueError(f"max_length ({max_length}) must equal the token count of image_size ({h * w})")
        ids = lambda s: torch.tensor(tok.encode(s), device=dev, dtype=torch.long)
        cond_emb = embed(ids(cond_prompt))
        uncond_emb = embed(ids(uncond_prompt)) if guidance_scale > 1.0 else None
        start = [tok.convert_tokens_to_ids("<|vision_start|>"), tok.convert_tokens_to_ids(f"<|res_{h}|>"),
                 tok.convert_tokens_to_ids(f"<|res_{w}|>")]
        start += [tok.convert_tokens_to_ids(f"<|query_{i}|>") for i in range(1, self.parallel_num)]
        start_emb = embed(torch.tensor(start, device=dev, dtype=torch.long))
        tokens, self.last_packed_tokens = self.engine.gen_tokens(
            cond_emb, uncond_emb, start_emb, h=h, w=w, num_images=num_images, guidance_scale=guidance_scale,
            num_sampling_steps=num_sampling_steps)
        return self.decode_image(tokens, [h, w], ps=self.ps)

    def decode_image(self, image_latents, image_size=None, ps=1):
        """image_latents: [B, h*w, C] tokens in patch-raster order -> [B, 3, H, W]."""
        if image_size is None:
            h = w = int(image_latents.size(1) ** 0.5)
        else:
            h, w = image_size
        if ps != self.engine.ps:
            raise ValueError("ps must match the head's parallel block size")
        return self.engine.ae.decode_tokens(image_latents.to(torch.float32).contiguous(), h, w, ps)
