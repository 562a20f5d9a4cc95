"""Interleaved text+image inference on the GPU (modeling/mllm.py:696-897): the text decode step
(Qwen3 causal step over the persistent paged KV -> lm_head on the tcgen05 GEMM -> sampler) against the CPU oracle
(oracle/interleaved.py, autocast-bf16 policy), and ``MLLModel.forward_inference_block_causal`` through the drop-in
``modeling`` package on a model directory with the reference's file layout."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

U, M_ = {"from": "user"}, {"from": "model"}
TEXT = "<|im_start|>user\na photo of the red cat<|im_end|>\n<|im_start|>assistant\n"


def rel_err(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


def test_lm_head_gemm_padded_vocab():
    """LmHead: vocab not a multiple of 128 (zero-padded rows), M = 1 and 3 rows, vs torch fp32 with one bf16 rounding."""
    from bitdance_b200.llm import LmHead
    torch.manual_seed(0)
    V, D = 500, 256
    w = (torch.randn(V, D) * 0.05).to(torch.bfloat16)
    head = LmHead(w, device="cuda")
    assert head.vocab == V
    for rows in (1, 3):
        h = torch.randn(rows, D).to(torch.bfloat16)
        got = head(h.cuda()).float().cpu()
        ref = (h.float() @ w.float().t()).to(torch.bfloat16).float()
        assert got.shape == (rows, V)
        assert (got - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item() + 1e-3


def test_text_decode_steps_vs_oracle():
    """Teacher-forced decode on a 3-layer Qwen3: causal pass over a 37-token context, then six [1, 1, D] steps, then a second
    multi-token causal pass ON TOP of the cache (a later user turn) and two more steps: logits of every step and the K pages
    against the oracle. Covers causal attention with a non-empty past for S = 1 and S > 1."""
    from bitdance_b200.llm import LlmRunner, LmHead, llm_spec
    from bitdance_b200.synth import synth_state_dict
    from oracle import interleaved as oi
    from oracle import llm as ol
    cfg = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2,
               head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6)
    V, D = 500, 256
    spec = llm_spec(cfg)
    spec["model.embed_tokens.weight"] = (V, D)
    spec["lm_head.weight"] = (V, D)
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth_state_dict(spec, seed=11, std=0.05).items()}
    run = LlmRunner(sd, cfg, max_positions=1024)
    head = LmHead(sd["lm_head.weight"], device="cuda")
    emb = sd["model.embed_tokens.weight"]
    torch.manual_seed(2)
    ctx1 = torch.randn(37, D).to(torch.bfloat16).float()
    ctx2 = torch.randn(9, D).to(torch.bfloat16).float()
    forced1, forced2 = [7, 499, 3, 250, 128, 64], [11, 12]
    cache = run.new_cache(1, 128)
    ocache = [None] * cfg["num_hidden_layers"]

    def gpu_decode(ctx, forced):
        x = ctx.to(torch.bfloat16).cuda().view(1, -1, D).contiguous()
        lg = []
        for t in forced:
            hid = run.forward(x, cache, 0, 1, causal=True)[:, -1]
            lg.append(head(hid)[0].float().cpu())
            x = emb[t].to(torch.bfloat16).cuda().view(1, 1, D).contiguous()
        return torch.stack(lg)

    with torch.no_grad():
        for ctx, forced in ((ctx1, forced1), (ctx2, forced2)):
            got = gpu_decode(ctx, forced)
            # neither loop feeds the last forced token of a turn back (nobody consumes its logits)
            _, ref, _ = oi.decode_text(sd, cfg, sd["lm_head.weight"], emb, ctx, ocache, end_id=-1, max_length=len(forced),
                                       rnd=ol.bf16, forced=forced)
            for s in range(len(forced)):
                e = rel_err(got[s], ref[s])
                print(f"text decode, context {ctx.shape[0]} tokens, step {s}: logits rel err {e:.4f}")
                assert e < 3e-2
    n_gpu = 37 + 5 + 9 + 1
    assert cache.seq_lens.tolist() == [n_gpu] and ocache[0][0].shape[2] == n_gpu
    k_ref = ocache[0][0][0]                       # layer 0: [Hkv, L, hd]
    pages = cache.page_table[0].long()
    k_dev = cache.pool[0, 0][pages].permute(1, 0, 2, 3).reshape(cfg["num_key_value_heads"], -1, 128)[:, :n_gpu]
    assert rel_err(k_dev.float().cpu(), k_ref) < 2e-2


# ----------------------------------------------------------------------------------------------------------------------
# the public surface on a model directory
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    from _fake_model import write_model_dir
    d = str(tmp_path_factory.mktemp("bitdance_tiny_il"))
    return d, write_model_dir(d)


@pytest.fixture(scope="module")
def mll(model_dir):
    from modeling.mllm import MLLModel
    from modeling.t2i_pipeline import BitDanceT2IPipeline
    return MLLModel.from_pipeline(BitDanceT2IPipeline(model_dir[0], device="cuda"))


def test_interleaved_t2i_plan_equals_gen_image(mll):
    """plan [user text, model image] is text-to-image: bit-identical to gen_image_block_causal on the same prompt pair
    (the unconditional twin = remove_first_user_block) for the same seed."""
    from modeling.utils import remove_first_user_block
    torch.manual_seed(5)
    out = mll.forward_inference_block_causal([dict(type="text", **U), dict(type="image", **M_)], [TEXT], [],
                                             max_length_vision=64, sample_steps=3, image_size=[32, 32], cfg_scale=3.0)
    torch.manual_seed(5)
    ref = mll.gen_image_block_causal(TEXT, remove_first_user_block(TEXT), guidance_scale=3.0, num_sampling_steps=3,
                                     max_length=64, num_images=1, image_size=[32, 32])
    img = out["generated_image"][0]
    assert out["generated_text"] == [] and img.shape == (1, 3, 32, 32) and torch.equal(img, ref)
    # editing-style plan: a user image in the context (start + tokens + end in both streams) changes the result
    torch.manual_seed(5)
    src = (torch.rand(1, 3, 32, 32) * 2 - 1).cuda()
    torch.manual_seed(5)
    out2 = mll.forward_inference_block_causal([dict(type="text", **U), dict(type="image", **U), dict(type="image", **M_)],
                                              [TEXT], [src], max_length_vision=64, sample_steps=3, image_size=[32, 32],
                                              cfg_scale=3.0)
    img2 = out2["generated_image"][0]
    assert img2.shape == (1, 3, 32, 32) and torch.isfinite(img2).all() and not torch.equal(img2, img)
    with pytest.raises(NotImplementedError):
        mll.forward_inference_block_causal([dict(type="text", **U), dict(type="image", **M_), dict(type="text", **M_)],
                                           [TEXT], [], max_length_vision=64, sample_steps=2, image_size=[32, 32])


def test_interleaved_text_generation_vs_oracle(mll, model_dir):
    """plans with generated text (understanding / multi-turn chat). Greedy: every token the GPU picked is the oracle's argmax
    up to the logit tolerance (oracle teacher-forced with the GPU's tokens, cache persisting across turns; a turn that ends
    with <|im_end|> clears the context, so the next turn feeds only its own text on top of the cache). Sampling: reproducible
    for a seed, different across seeds."""
    from oracle import interleaved as oi
    from oracle import llm as ol
    info = model_dir[1]
    sd, cfg = info["sds"]["llm"], info["model"]["llm"]
    emb_w, head_w = sd["model.embed_tokens.weight"], info["lm_head"]
    tok = mll.tokenizer
    plan = [dict(type="text", **U), dict(type="text", **M_), dict(type="text", **U), dict(type="text", **M_)]
    t1, t2 = TEXT, "<|im_start|>user\nthe blue dog on table<|im_end|>\n<|im_start|>assistant\n"
    # pass 1 (greedy, 6 tokens per turn): find what the model says, then make the 3rd token of turn 1 the end token
    out = mll.forward_inference_block_causal(plan[:2], [t1], [], do_sample=False, max_length_text=6)
    ids1 = mll.last_text_ids.tolist()
    assert len(ids1) == 6 and isinstance(out["generated_text"][0], str)
    real_end = tok.convert_tokens_to_ids("<|im_end|>")
    assert real_end not in ids1, "unlucky seed: the random model emitted <|im_end|>"
    end_id = ids1[2]
    tok.im_end_id = end_id          # the reference's alias attribute (data/data_utils.py:95-109) takes precedence
    try:
        texts = [t1, t2]
        out = mll.forward_inference_block_causal(plan, texts, [], do_sample=False, max_length_text=6)
        assert texts == [] and len(out["generated_text"]) == 2
        ids2 = mll.last_text_ids.tolist()
        # oracle, teacher-forced with the GPU's choices
        first_end = ids1.index(end_id)
        turn1 = ids1[:first_end + 1]
        ocache = [None] * cfg["num_hidden_layers"]
        tol_hits = 0
        with torch.no_grad():
            for text, forced in ((t1, turn1), (t2, ids2)):
                ctx = emb_w[torch.tensor(tok.encode(text))]
                _, logits, _ = oi.decode_text(sd, cfg, head_w, emb_w, ctx, ocache, end_id=end_id, max_length=len(forced),
                                              rnd=ol.bf16, forced=forced)
                # decode_text stops feeding at the end token, exactly like the GPU loop
                for s, t in enumerate(forced):
                    lg = logits[s]
                    gap = (lg.max() - lg[t]).item()
                    tol = 6e-2 * lg.abs().max().item()
                    assert gap <= tol, f"token {t} at step {s} is not the oracle's argmax within tolerance: gap {gap}, tol {tol}"
                    tol_hits += int(gap > 0)
                if forced is turn1:
                    assert forced[-1] == end_id
        print(f"greedy text decode: {len(turn1) + len(ids2)} tokens, {tol_hits} within-tolerance ties vs the oracle")
    finally:
        del tok.im_end_id
    # sampling (the reference's default): reproducible per seed
    def sample(seed):
        torch.manual_seed(seed)
        o = mll.forward_inference_block_causal([dict(type="image", **U), dict(type="text", **U), dict(type="text", **M_)],
                                               [t2], [(torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(1))
                                                       * 2 - 1).cuda()], max_length_text=8, image_size=[32, 32])
        return o["generated_text"][0], mll.last_text_ids.tolist()
    a, b, c = sample(3), sample(3), sample(4)
    assert a == b and a[1] != c[1]
    assert len(a[1]) <= 8 and all(0 <= t < cfg["vocab_size"] for t in a[1])
