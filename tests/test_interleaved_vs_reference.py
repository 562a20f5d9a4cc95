"""Interleaved text+image inference against the UNMODIFIED reference ``MLLModel.forward_inference_block_causal``
(modeling/mllm.py:696-897), on CPU in fp32 with the tiny random model (dev container only: the reference's ``mllm.py``
imports its ``data`` package, which is not shipped to the GPU box).

What can be pinned and is:
  * plan [user text, model image] == the reference's own ``gen_image(cond, remove_first_user_block(cond))`` bit for bit — the
    equivalence the mirror's image item rests on (it hands the accumulated context to the same block generator);
  * plan [user text, user image, model image] (editing) == ``oracle/pipeline.py::gen_image`` fed with the context the MIRROR's
    bookkeeping builds (start tokens + encode_image + <|vision_end|> in BOTH streams, unconditional text =
    remove_first_user_block): pins that bookkeeping against the reference's.
What cannot: the reference's TEXT branch raises — without a cache at mllm.py:798 (``past_key_values[0][0]`` of None), after a
generated image on the second token (a 2-D ``(1, hidden)`` tensor fed back as ``inputs_embeds``, :857 -> rotary shape
error). Both are asserted here, as the evidence for DESIGN.md section 2c's "text loop parity-unpinned"."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.reference

TEXT = "<|im_start|>user\na photo of the red cat<|im_end|>\n<|im_start|>assistant\n"
U, M_ = {"from": "user"}, {"from": "model"}


class _Cfg(dict):
    __getattr__ = dict.__getitem__


@pytest.fixture(scope="module")
def world():
    import os
    if not os.path.isdir("/root/reference/data"):
        pytest.skip("needs the full reference checkout (modeling/mllm.py imports data.data_utils)")
    from bitdance_b200.synthetic import synthetic_tokenizer
    from oracle import ref_harness as rh
    from oracle import ref_runner as rr
    rh.import_reference()
    mllm = rh.import_reference_mllm()
    torch.manual_seed(0)
    pipe, info = rr.build_pipeline("tiny", "cpu")
    pipe.llm_model.float()                                   # exact-math pin: everything fp32
    tok, _ = synthetic_tokenizer(512, 16)
    for alias, t in (("im_start", "<|im_start|>"), ("im_end", "<|im_end|>"), ("start_of_image", "<|vision_start|>"),
                     ("end_of_image", "<|vision_end|>")):
        setattr(tok, alias + "_id", tok.convert_tokens_to_ids(t))      # data/data_utils.py:95-109
    for i in range(1, 161):
        setattr(tok, f"res_{i}_id", tok.convert_tokens_to_ids(f"<|res_{i}|>"))
    for i in range(1, 16):
        setattr(tok, f"query_{i}_id", tok.convert_tokens_to_ids(f"<|query_{i}|>"))
    pipe.tokenizer = tok
    M = mllm.MLLModel
    m = M.__new__(M)
    nn.Module.__init__(m)
    m.config = _Cfg(vit_patch_size=pipe.vae_patch_size, head=_Cfg(vision_pred=_Cfg(parallel_num=16)),
                    encoder=_Cfg(vt_forward_func="group", max_bs=32))
    m.tokenizer, m.llm_model, m.vision_head_type = tok, pipe.llm_model, "diffusion_parallel_x"
    m.vision_diffusion_head, m.embed_vision_mlp, m.vision_encoder = pipe.vision_head, pipe.embed_vision_mlp, pipe.ae
    m.parallel_num, m.ps, m.hidden_size = 16, 4, 256
    m.register_buffer("pos_embed_1d", m._get_1d_sincos_pos_embed(128, 64), persistent=False)
    m.eval()
    return m, pipe, tok


def _capture_noise(fn):
    rec = []
    o1, o2 = torch.randn, torch.randn_like

    def r1(*a, **k):
        t = o1(*a, **k)
        rec.append(t.clone())
        return t

    def r2(a, **k):
        t = o2(a, **k)
        rec.append(t.clone())
        return t

    torch.randn, torch.randn_like = r1, r2
    try:
        out = fn()
    finally:
        torch.randn, torch.randn_like = o1, o2
    return out, rec


def test_reference_t2i_plan_is_gen_image(world):
    from bitdance_b200.modeling.utils import remove_first_user_block
    m, pipe, tok = world
    kw = dict(max_length_vision=64, sample_steps=3, image_size=[32, 32], cfg_scale=3.0)
    with torch.no_grad():
        torch.manual_seed(5)
        out = m.forward_inference_block_causal([dict(type="text", **U), dict(type="image", **M_)], [TEXT], [], **kw)
        torch.manual_seed(5)
        ref = pipe.gen_image(TEXT, remove_first_user_block(TEXT), guidance_scale=3.0, num_sampling_steps=3, max_length=64,
                             num_images=1, image_size=[32, 32])
    img = out["generated_image"][0]
    assert out["generated_text"] == [] and img.shape == (1, 3, 32, 32) and torch.equal(img, ref)


def test_reference_editing_plan_vs_oracle_with_mirror_bookkeeping(world):
    from bitdance_b200.modeling.utils import remove_first_user_block
    from oracle import pipeline as op
    from oracle import ref_runner as rr
    m, pipe, tok = world
    S, guidance, pn = 3, 3.0, 16
    src = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1
    plan = [dict(type="text", **U), dict(type="image", **U), dict(type="image", **M_)]
    with torch.no_grad():
        torch.manual_seed(7)
        out, noise = _capture_noise(lambda: m.forward_inference_block_causal(
            plan, [TEXT], [src.clone()], max_length_vision=64, sample_steps=S, image_size=[32, 32], cfg_scale=guidance))
        img_ref = out["generated_image"][0]
        steps = 64 // pn
        assert len(noise) == steps * (S + 1)
        per_step = [noise[i * (S + 1):(i + 1) * (S + 1)] for i in range(steps)]
        # the context as bitdance_b200/modeling/mllm.py builds it
        embed = m.llm_model.model.embed_tokens.weight.detach().float()
        E = lambda ids: embed[torch.tensor(list(ids))]
        start3 = [tok.start_of_image_id, tok.res_8_id, tok.res_8_id]
        pre = m.encode_image([src.clone()])[0].float()
        assert pre.shape == (64, 256)
        end = E([tok.end_of_image_id])
        cond = torch.cat([E(tok.encode(TEXT)), E(start3), pre, end])
        uncond = torch.cat([E(tok.encode(remove_first_user_block(TEXT))), E(start3), pre, end])
        start = start3 + [getattr(tok, f"query_{i}_id") for i in range(1, pn)]
        c = rr.CONFIGS["tiny"]["llm"]
        cfg = {k: v for k, v in c.items() if k != "vocab_size"}
        f32 = lambda sd: {k: v.detach().float() for k, v in sd.items()}
        tokens, img = op.gen_image(sd_llm=f32(m.llm_model.state_dict()), cfg_llm=cfg, embed=embed,
                                   sd_head=f32(pipe.vision_head.state_dict()), sd_proj=f32(pipe.embed_vision_mlp.state_dict()),
                                   sd_ae=f32(pipe.ae.state_dict()), cond_ids=None, uncond_ids=None, cond_emb=cond,
                                   uncond_emb=uncond, start_ids=start, h=8, w=8, pn=pn, num_images=1, guidance=guidance, S=S,
                                   noise=per_step, head_dim=128)
    assert img.shape == img_ref.shape == (1, 3, 32, 32)
    err = (img - img_ref).abs().max().item()
    assert err < 1e-3 * max(1.0, img_ref.abs().max().item()), err


def test_reference_text_branch_raises(world):
    m, pipe, tok = world
    with torch.no_grad():
        # no cache yet: mllm.py:798 subscripts past_key_values = None before the first pass
        with pytest.raises(TypeError):
            m.forward_inference_block_causal([dict(type="text", **U), dict(type="text", **M_)], [TEXT], [], max_length_text=4)
        # after a generated image (a cache exists): the first token is sampled, then its (1, hidden) embedding is fed back
        # as inputs_embeds (mllm.py:857) and the decoder fails on the rotary shapes
        with pytest.raises(RuntimeError):
            m.forward_inference_block_causal([dict(type="text", **U), dict(type="image", **M_), dict(type="text", **U),
                                              dict(type="text", **M_)], [TEXT, "the blue dog"], [], max_length_text=4,
                                             max_length_vision=64, sample_steps=2, image_size=[32, 32], cfg_scale=3.0)
