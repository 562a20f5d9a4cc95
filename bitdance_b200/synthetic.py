"""Random-init models of the published architectures (there is no network for checkpoints): used by bench.py,
``__graft_entry__.smoke()`` and the tests. Dimensions come from the reference configs:
  * BitDance-14B-64x / -16x : train/configs/bitdance_14b_64x.yaml:9-33 (ae_d16c32, Qwen3-14B, head 5120 x 6 blocks,
    2 adaLN, SwiGLU, parallel_num 64 / 16); Qwen3-14B dims from its HF config.json (hidden 5120, 40 layers,
    40 Q / 8 KV heads x 128, MLP 17408, rms eps 1e-6, rope theta 1e6).
"""
from __future__ import annotations

import torch

from .ae import AERunner, ae_spec
from .head import HeadRunner, head_spec
from .llm import LlmRunner
from .pipeline import T2IEngine
from .synth import synth_tensor

QWEN3_14B = dict(hidden_size=5120, intermediate_size=17408, num_hidden_layers=40, num_attention_heads=40,
                 num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=151936)
AE_D16C32 = dict(double_z=False, z_channels=32, in_channels=3, out_ch=3, ch=256, ch_mult=[1, 1, 2, 2, 4],
                 num_res_blocks=4)

MODELS = {
    "BitDance-14B-64x": dict(llm=QWEN3_14B, ae=AE_D16C32, parallel_num=64,
                             head=dict(ch_target=32, ch_cond=5120, ch_latent=5120, depth_latent=6, depth_adanln=2,
                                       use_swiglu=True)),
    "BitDance-14B-16x": dict(llm=QWEN3_14B, ae=AE_D16C32, parallel_num=16,
                             head=dict(ch_target=32, ch_cond=5120, ch_latent=5120, depth_latent=6, depth_adanln=2,
                                       use_swiglu=True)),
    # a few-second model with the same structure (smoke / CI)
    "tiny": dict(llm=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                          num_key_value_heads=2, head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6, vocab_size=512),
                 ae=dict(double_z=False, z_channels=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
                         num_res_blocks=1),
                 parallel_num=16,
                 head=dict(ch_target=32, ch_cond=256, ch_latent=256, depth_latent=2, depth_adanln=2, use_swiglu=True)),
}


def _gpu_state_dict(spec: dict, seed: int, device, std=0.02) -> dict:
    return {k: synth_tensor(k, v, seed, std, device=device, dtype=torch.bfloat16 if len(v) >= 2 else torch.float32)
            for k, v in spec.items()}


def build_synthetic_engine(model: str = "BitDance-14B-64x", device="cuda", seed: int = 0, with_ae: bool = True,
                           llm_stream: bool | None = None):
    """Returns (engine, embed_table bf16 [vocab, D]). llm_stream: also keep the stream-major copy of the decoder weights so
    that an AR block runs as ONE persistent launch (None: the BD_LLM_STREAM environment variable, default off)."""
    m = MODELS[model]
    dev = torch.device(device)
    llm = LlmRunner(None, m["llm"], device=dev, synthetic_seed=seed + 1, stream=llm_stream)
    hc = m["head"]
    sd_head = _gpu_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"],
                                        hc["depth_adanln"], hc["use_swiglu"]), seed + 2, dev)
    head = HeadRunner(sd_head, device=dev, **hc)
    del sd_head
    ae = None
    if with_ae:
        sd_ae = _gpu_state_dict(ae_spec(m["ae"]), seed + 3, dev)
        ae = AERunner(sd_ae, m["ae"], device=dev)
        del sd_ae
    D, zc = m["llm"]["hidden_size"], m["ae"]["z_channels"]
    proj = _gpu_state_dict({"fc1.weight": (D, zc), "fc1.bias": (D,), "fc2.weight": (D, D), "fc2.bias": (D,)}, seed + 4, dev)
    vps = 2 ** (len(m["ae"]["ch_mult"]) - 1)
    eng = T2IEngine(llm, head, ae, proj["fc1.weight"], proj["fc1.bias"], proj["fc2.weight"], proj["fc2.bias"],
                    parallel_num=m["parallel_num"], vae_patch_size=vps, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 5)
    embed = (torch.randn((m["llm"]["vocab_size"], D), generator=g, device=dev) * 0.02).to(torch.bfloat16)
    torch.cuda.empty_cache()
    return eng, embed


def tiny_state_dicts(seed: int = 0) -> dict:
    """CPU state dicts (reference key names) of the "tiny" model — reproducible on any machine, so the CPU oracle and the
    GPU engine can be loaded with identical weights (smoke(), tests)."""
    from .llm import llm_spec
    from .synth import synth_state_dict
    m = MODELS["tiny"]
    D, zc = m["llm"]["hidden_size"], m["ae"]["z_channels"]
    spec = llm_spec(m["llm"])
    spec["model.embed_tokens.weight"] = (m["llm"]["vocab_size"], D)
    hc = m["head"]
    sd_llm = {k: v.to(torch.bfloat16).float() for k, v in synth_state_dict(spec, seed=seed + 3, std=0.05).items()}
    return dict(
        llm=sd_llm,
        head=synth_state_dict(head_spec(hc["ch_target"], hc["ch_cond"], hc["ch_latent"], hc["depth_latent"],
                                        hc["depth_adanln"], hc["use_swiglu"]), seed=seed + 1, std=0.05),
        ae=synth_state_dict(ae_spec(m["ae"]), seed=seed + 2, std=0.05),
        proj=synth_state_dict({"fc1.weight": (D, zc), "fc1.bias": (D,), "fc2.weight": (D, D), "fc2.bias": (D,)},
                              seed=seed + 4, std=0.05),
    )


def engine_from_state_dicts(sds: dict, model: str = "tiny", device="cuda", parallel_num: int | None = None):
    m = MODELS[model]
    hc = m["head"]
    llm = LlmRunner(sds["llm"], m["llm"], device=device, max_positions=4096)
    head = HeadRunner(sds["head"], device=device, **hc)
    ae = AERunner(sds["ae"], m["ae"], device=device)
    vps = 2 ** (len(m["ae"]["ch_mult"]) - 1)
    p = sds["proj"]
    return T2IEngine(llm, head, ae, p["fc1.weight"], p["fc1.bias"], p["fc2.weight"], p["fc2.bias"],
                     parallel_num=parallel_num or m["parallel_num"], vae_patch_size=vps, device=device, pe_max_len=1024)


def special_tokens(pn: int, max_res: int = 160):
    """The added tokens the pipeline looks up (t2i_pipeline.py:182-192): chat markers, <|vision_start|>, <|res_N|>, <|query_i|>."""
    return (["<|im_start|>", "<|im_end|>", "<|vision_start|>", "<|vision_end|>"]
            + [f"<|res_{i}|>" for i in range(1, max_res + 1)] + [f"<|query_{i}|>" for i in range(1, pn)])


def synthetic_tokenizer(vocab_size: int, pn: int, n_words: int = 56, save_to: str | None = None):
    """An offline stand-in for the Qwen tokenizer files (no network): a word-level ``tokenizers`` model wrapped in
    ``PreTrainedTokenizerFast`` with the special tokens above. Returns (tokenizer, a prompt of ``n_words`` words)."""
    from tokenizers import Regex, Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    words = ["user", "assistant", "a", "photo", "of", "cat", "dog", "red", "blue", "the", "on", "table", "\n", "[UNK]"]
    words += [f"w{i}" for i in range(64)]
    sp = special_tokens(pn)
    assert len(words) + len(sp) <= vocab_size
    tok = Tokenizer(models.WordLevel(vocab={w: i for i, w in enumerate(words)}, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(r"\n"), behavior="isolated"),
                                                 pre_tokenizers.WhitespaceSplit()])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]")
    fast.add_special_tokens({"additional_special_tokens": sp})
    if save_to is not None:
        fast.save_pretrained(save_to)
    prompt = " ".join((["a", "photo", "of", "the", "red", "cat", "on", "table"] + [f"w{i}" for i in range(64)])[:n_words])
    return fast, prompt
