"""Writes a complete BitDance T2I model directory with the reference's file layout (modeling/t2i_pipeline.py:45-77) for
the "tiny" synthetic weights: HF ``config.json`` + ``model.safetensors`` + tokenizer files, ``ae_config.json`` /
``ae.safetensors``, ``vision_head_config.json`` / ``vision_head.safetensors``, ``projector.safetensors``. No network: the
tokenizer is a word-level ``tokenizers`` model built here with the special tokens the pipeline looks up
(``<|vision_start|>``, ``<|res_N|>``, ``<|query_i|>``, chat markers)."""
from __future__ import annotations

import json
import os

import torch


def write_tokenizer(path: str, vocab_size: int, pn: int):
    from bitdance_b200.synthetic import synthetic_tokenizer
    synthetic_tokenizer(vocab_size, pn, save_to=path)


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
    # lm_head (interleaved text generation): not part of tiny_state_dicts (the goldens predate it); same generator family
    from bitdance_b200.synth import synth_state_dict
    lm_head = synth_state_dict({"lm_head.weight": (m["llm"]["vocab_size"], m["llm"]["hidden_size"])}, seed=seed + 9, std=0.05)
    lm_head = {k: v.to(torch.bfloat16).float() for k, v in lm_head.items()}
    save_file({k: v.to(torch.bfloat16).contiguous() for k, v in {**sds["llm"], **lm_head}.items()},
              os.path.join(path, "model.safetensors"))
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
    return dict(sds=sds, model=m, ae_config=ae_config, head_config=head_config, lm_head=lm_head["lm_head.weight"])
