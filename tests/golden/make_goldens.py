#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference, CPU, fp32) on the
"tiny" synthetic weights of bitdance_b200.synthetic.tiny_state_dicts() (regenerable anywhere from their seeds).

  python tests/golden/make_goldens.py        # dev container only; commits the small fixtures it writes

The reference ships no golden vectors: these are the pinned outputs of its own code."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def capture_noise(fn):
    rec = []
    o1, o2 = torch.randn, torch.randn_like
    torch.randn = lambda *a, **k: (rec.append(o1(*a, **k)) or rec[-1])
    torch.randn_like = lambda a, **k: (rec.append(o2(a, **k)) or rec[-1])
    try:
        out = fn()
    finally:
        torch.randn, torch.randn_like = o1, o2
    return out, [r.clone() for r in rec]


def main():
    from bitdance_b200.synthetic import MODELS, tiny_state_dicts
    from oracle import ref_harness as rh
    ref = rh.import_reference()
    from transformers import Qwen3Config, Qwen3ForCausalLM
    m = MODELS["tiny"]
    sds = tiny_state_dicts()
    pn = m["parallel_num"]
    with torch.no_grad():
        # ---- tokenizer ----
        ae = ref.ae.VQModel(m["ae"]).eval()
        ae.load_state_dict(sds["ae"])
        torch.manual_seed(2)
        img = torch.rand(2, 3, 32, 48) * 2 - 1
        lat = ae.encoder(img)
        quant = ae.encode(img)
        dec = ae.decode(quant)
        np.savez_compressed(os.path.join(OUT, "ae_tiny.npz"), image=img.numpy(), latent=lat.numpy(),
                            quant=quant.numpy().astype(np.int8), decoded=dec.numpy())
        # ---- head ----
        hc = m["head"]
        head = ref.fh.DiffHead(parallel_num=pn, **hc).eval()
        head.load_state_dict(sds["head"])
        torch.manual_seed(3)
        R = 4
        x, t, c = torch.randn(R, pn, 32), torch.rand(R), torch.randn(R, pn, hc["ch_cond"])
        xpred = head.net(x, t, c)
        torch.manual_seed(4)
        samp, noise = capture_noise(lambda: head.sample(c, cfg=3.0, num_sampling_steps=5))
        np.savez_compressed(os.path.join(OUT, "head_tiny.npz"), x=x.numpy(), t=t.numpy(), c=c.numpy(),
                            xpred=xpred.numpy(), sample=samp.numpy(), noise=torch.stack(noise).numpy(), cfg=3.0, S=5)
        # ---- LLM ----
        lc = {k: v for k, v in m["llm"].items()}
        hf = Qwen3ForCausalLM(Qwen3Config(max_position_embeddings=4096, tie_word_embeddings=False, **lc)).eval()
        missing = hf.load_state_dict(sds["llm"], strict=False)
        assert set(missing.missing_keys) <= {"lm_head.weight"}, missing
        torch.manual_seed(5)
        B = 2
        x0, x1, x2 = torch.randn(B, 9, 256), torch.randn(B, pn, 256), torch.randn(B, pn, 256)
        o = hf.model(inputs_embeds=x0, use_cache=True)
        pkv, h0 = o.past_key_values, o.last_hidden_state
        outs = []
        for xx in (x1, x2):
            mask = torch.ones(B, 1, pn, pn + pkv[0][0].shape[2], dtype=torch.bool)
            o = hf.model(inputs_embeds=xx, past_key_values=pkv, use_cache=True, attention_mask=mask)
            pkv = o.past_key_values
            outs.append(o.last_hidden_state)
        np.savez_compressed(os.path.join(OUT, "llm_tiny.npz"), x0=x0.numpy(), x1=x1.numpy(), x2=x2.numpy(), h0=h0.numpy(),
                            h1=outs[0].numpy(), h2=outs[1].numpy())
        # ---- whole pipeline ----
        proj = ref.mu.MLPconnector(32, 256, "gelu_pytorch_tanh").eval()
        proj.load_state_dict(sds["proj"])

        class Tok:
            def encode(self, s):
                return [5, 17, 33, 2, 90] if s == "cond" else [3, 4]

            def convert_tokens_to_ids(self, tk):
                if tk == "<|vision_start|>":
                    return 400
                if tk.startswith("<|res_"):
                    return 401
                return 410 + int(tk[8:-2])

        P = ref.t2i.BitDanceT2IPipeline
        pipe = object.__new__(P)
        pipe.device, pipe.tokenizer, pipe.llm_model = "cpu", Tok(), hf
        pipe.hidden_size, pipe.ae, pipe.vision_head, pipe.embed_vision_mlp = 256, ae, head, proj
        pipe.vae_patch_size, pipe.parallel_num, pipe.ps = 4, pn, 4
        pipe.build_pos_embed(max_len=1024)
        torch.manual_seed(11)
        S, guidance, Bimg = 3, 3.0, 1
        img_out, noise = capture_noise(lambda: pipe.gen_image("cond", "uncond", guidance_scale=guidance,
                                                              num_sampling_steps=S, max_length=64, num_images=Bimg,
                                                              image_size=[32, 32]))
        np.savez_compressed(os.path.join(OUT, "pipeline_tiny.npz"), image=img_out.numpy(), noise=torch.stack(noise).numpy(),
                            S=S, guidance=guidance, B=Bimg)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
