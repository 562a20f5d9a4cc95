"""Writes a complete BitDance T2I model directory with the reference's file layout (modeling/t2i_pipeline.py:45-77) for
the "tiny" synthetic weights: HF ``config.json`` + ``model.safetensors`` + tokenizer files, ``ae_config.json`` /
``ae.safetensors``, ``vision_head_config.json`` / ``vision_head.safetensors``, ``projector.safetensors``. No network: the
tokenizer is a word-level ``tokenizers`` model built here with the special tokens the pipeline looks up
(``<|vision_start|>``, ``<|res_N|>``, ``<|query_i|>``, chat markers)."""
from __future__ import annotations

import json
import os

import torch


def special_tokens(pn: int, max_res: int = 160):
    return (["<|im_start|>", "<|im_end|>", "<|vision_start|>", "<|vision_end|>"]
            + [f"<|res_{i}|>" for i in range(1, max_res + 1)] + [f"<|query_{i}|>" for i in range(1, pn)])


def write_tokenizer(path: str, vocab_size: int, pn: int):
    from tokenizers import Regex, Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    words = ["user", "assistant", "a", "photo", "of", "cat", "dog", "red", "blue", "the", "on", "table", "\n", "[UNK]"]
    sp = special_tokens(pn)
    assert len(words) + len(sp) <= vocab_size
    vocab = {w: i for i, w in enumerate(words)}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Split(Regex(r"\n|[^\s]+"), behavior="isolated")
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]")
    fast.add_special_tokens({"additional_special_tokens": sp})
    assert len(fast) <= vocab_size
    fast.save_pretrained(path)


def write_model_dir(path: str, seed: int = 0) -> dict:
    """Returns the state dicts written (CPU fp32, reference key names) + configs."""
    from safetensors.torch import save_file
    from bitdance_b200.synthetic import MODELS, tiny_state_dicts
    os.makedirs(path, exist_ok=True)
    m = MODELS["tiny"]
    sds = tiny_state_dicts(seed)
    llm_cfg = dict(m["llm"], model_type="qwen3", architectures=["Qwen3ForCausalLM"], max_position_embeddings=32768,
                   tie_word_embeddings=False, hidden_act="silu", torch_dtype="bfloat16")
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(llm_cfg, f)
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in sds["llm"].items()}, os.path.join(path, "model.safetensors"))
    write_tokenizer(path, m["llm"]["vocab_size"], m["parallel_num"])
    ae_config = {"ddconfig": m["ae"]}
    with open(os.path.join(path, "ae_config.json"), "w") as f:
        json.dump(ae_config, f)
    save_file({k: v.contiguous() for k, v in sds["ae"].items()}, os.path.join(path, "ae.safetensors"))
    head_config = dict(m["head"], parallel_num=m["parallel_num"], time_shift=1.0, P_mean=0.0, P_std=1.0)
    with open(os.path.join(path, "vision_head_config.json"), "w") as f:
        json.dump(head_config, f)
    save_file({k: v.contiguous() for k, v in sds["head"].items()}, os.path.join(path, "vision_head.safetensors"))
    save_file({k: v.contiguous() for k, v in sds["proj"].items()}, os.path.join(path, "projector.safetensors"))
    return dict(sds=sds, model=m, ae_config=ae_config, head_config=head_config)
