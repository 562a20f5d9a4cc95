"""bd_llm_forward (Qwen3 decoder over a paged KV cache) vs the CPU oracle (oracle/llm.py, autocast-bf16 policy)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2,
           head_dim=128, rms_norm_eps=1e-6, rope_theta=1e6)


def make(cfg, seed=3):
    from bitdance_b200.llm import LlmRunner, llm_spec
    from bitdance_b200.synth import synth_state_dict
    sd = synth_state_dict(llm_spec(cfg), seed=seed, std=0.05)
    sd = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}   # the model is stored in bf16
    return sd, LlmRunner(sd, cfg, max_positions=1024)


def rel_err(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


@pytest.mark.parametrize("head_dim", [128, 64])
def test_llm_prefill_and_ar_blocks(head_dim):
    from oracle import llm as ol
    cfg = dict(CFG, head_dim=head_dim)
    sd, run = make(cfg)
    torch.manual_seed(0)
    B, pn, D = 2, 16, cfg["hidden_size"]
    L = cfg["num_hidden_layers"]
    lens = [21, 5]                      # cond / uncond prompt lengths differ (t2i_pipeline.py:199 vs :224)
    cache = run.new_cache(2 * B, 256)
    ocache = [[None] * L for _ in range(2 * B)]
    prompts = [torch.randn(B, n, D).to(torch.bfloat16).float() for n in lens]
    block0 = torch.randn(B, pn, D).to(torch.bfloat16).float()
    outs = []
    # prefill: causal over the prompt, then the first block with an all-ones mask; cond group then uncond group
    for gi, n in enumerate(lens):
        r0 = gi * B
        run.forward(prompts[gi].to(torch.bfloat16).cuda(), cache, r0, B, causal=True)
        o = run.forward(block0.to(torch.bfloat16).cuda(), cache, r0, B, causal=False)
        outs.append(o.float().cpu())
        for b in range(B):
            ol.decoder_forward(sd, cfg, prompts[gi][b:b + 1], ocache[r0 + b], causal=True, rnd=ol.bf16, stream_f32=False)
            ref = ol.decoder_forward(sd, cfg, block0[b:b + 1], ocache[r0 + b], causal=False, rnd=ol.bf16,
                                     stream_f32=False)
            e = rel_err(outs[gi][b:b + 1], ref)
            assert e < 3e-2, f"prefill group {gi} seq {b}: rel err {e}"
    # AR blocks: fp32 stream, cond+uncond sequences batched in ONE pass with different past lengths
    pos = torch.randn(pn, D)
    for step in range(3):
        x = torch.randn(2 * B, pn, D)
        o = run.forward(x.clone().cuda(), cache, 0, 2 * B, causal=False, out_add=pos.cuda(), out_add_mod=pn)
        torch.cuda.synchronize()
        for r in range(2 * B):
            ref = ol.decoder_forward(sd, cfg, x[r:r + 1], ocache[r], causal=False, rnd=ol.bf16, stream_f32=True) + pos
            e = rel_err(o[r:r + 1].cpu(), ref)
            assert e < 3e-2, f"AR step {step} seq {r}: rel err {e}"
    assert cache.seq_lens.tolist() == [lens[0] + 4 * pn] * B + [lens[1] + 4 * pn] * B
    # K written to the pages equals the oracle's cache (bf16 RoPE'd keys), token by token
    k_ref = ocache[0][0][0][0]          # layer 0, seq 0: [Hkv, L, hd]
    Lk = k_ref.shape[1]
    pages = cache.page_table[0].long()
    k_dev = cache.pool[0, 0][pages].permute(1, 0, 2, 3).reshape(cfg["num_key_value_heads"], -1, head_dim)[:, :Lk]
    assert rel_err(k_dev.float().cpu(), k_ref) < 2e-2


def test_llm_split_kv_matches_unsplit():
    cfg = dict(CFG)
    sd, run = make(cfg)
    torch.manual_seed(1)
    D = cfg["hidden_size"]
    x0 = torch.randn(1, 700, D).to(torch.bfloat16).cuda()
    x1 = torch.randn(1, 64, D).cuda()
    outs = []
    for splits in (1, 3):
        cache = run.new_cache(1, 1024)
        run.forward(x0.clone(), cache, 0, 1, causal=True, attn_splits=1)
        outs.append(run.forward(x1.clone(), cache, 0, 1, causal=False, attn_splits=splits).cpu())
    assert rel_err(outs[1], outs[0]) < 5e-3


def test_llm_qwen3_14b_layer_vs_oracle():
    """One decoder layer at the Qwen3-14B dimensions (hidden 5120, 40 Q / 8 KV heads x 128, MLP 17408 — the shapes the
    bench runs 40 times per AR step): causal prefill of a 2085-token prompt (bf16 stream), the first image block
    (block-bidirectional, bf16 stream), then an fp32-stream AR block over BOTH sequences of a CFG pair in one pass with
    very different cache lengths (2213 vs 134 tokens: split-KV over 35 pages vs 3), against oracle/llm.py."""
    from bitdance_b200.llm import LlmRunner, llm_spec
    from bitdance_b200.synth import synth_state_dict
    from bitdance_b200.synthetic import QWEN3_14B
    from oracle import llm as ol
    cfg = {k: v for k, v in QWEN3_14B.items() if k != "vocab_size"}
    cfg["num_hidden_layers"] = 1
    sd = synth_state_dict(llm_spec(cfg), seed=5, std=0.02)
    sd = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}
    run = LlmRunner(sd, cfg, max_positions=4096)
    torch.manual_seed(0)
    D, pn = cfg["hidden_size"], 64
    lens = [2085, 6]
    cache = run.new_cache(2, lens[0] + 4 * pn)
    ocache = [[None], [None]]
    block0 = torch.randn(1, pn, D).to(torch.bfloat16).float()
    with torch.no_grad():
        for r, n in enumerate(lens):
            prompt = torch.randn(1, n, D).to(torch.bfloat16).float()
            run.forward(prompt.to(torch.bfloat16).cuda(), cache, r, 1, causal=True)
            o = run.forward(block0.to(torch.bfloat16).cuda(), cache, r, 1, causal=False).float().cpu()
            ol.decoder_forward(sd, cfg, prompt, ocache[r], causal=True, rnd=ol.bf16, stream_f32=False)
            ref = ol.decoder_forward(sd, cfg, block0, ocache[r], causal=False, rnd=ol.bf16, stream_f32=False)
            e = rel_err(o, ref)
            print(f"Qwen3-14B layer, prefill + first block, seq {r} ({n} tokens): rel err {e:.4f}")
            assert e < 3e-2
        for step in range(2):
            x = torch.randn(2, pn, D)
            o = run.forward(x.clone().cuda(), cache, 0, 2, causal=False).cpu()
            for r in range(2):
                ref = ol.decoder_forward(sd, cfg, x[r:r + 1], ocache[r], causal=False, rnd=ol.bf16, stream_f32=True)
                e = rel_err(o[r:r + 1], ref)
                print(f"Qwen3-14B layer, fp32-stream AR block {step}, seq {r}: rel err {e:.4f}")
                assert e < 3e-2
    assert cache.seq_lens.tolist() == [lens[0] + 3 * pn, lens[1] + 3 * pn]
    k_ref = ocache[0][0][0][0]          # [Hkv, L, hd]
    pages = cache.page_table[0].long()
    k_dev = cache.pool[0, 0][pages].permute(1, 0, 2, 3).reshape(cfg["num_key_value_heads"], -1, 128)[:, :k_ref.shape[1]]
    assert rel_err(k_dev.float().cpu(), k_ref) < 2e-2
