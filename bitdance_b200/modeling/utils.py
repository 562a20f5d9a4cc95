"""API mirror of the reference ``modeling/utils.py`` (hot-path part): ``MLPconnector``.

Reference: modeling/utils.py:9-20 — fc2(act(fc1(x))) with ``hidden_act='gelu_pytorch_tanh'`` on the T2I path. The two
Linears run on the tcgen05 GEMM with the activation fused in fc1's epilogue.

Interleaved text+image inference adds the token sampler around the decode step:
``top_k_top_p_filtering`` / ``sample_codebook`` (modeling/utils.py:64-124) and ``remove_first_user_block`` (:206-216).
They are thin sequences of torch calls on the logits the native lm_head GEMM produced, deliberately the SAME torch calls
in the same order as the reference (``topk`` / ``sort`` / ``softmax`` / ``cumsum`` / ``multinomial``): the sampler's
contract is the global torch CUDA generator (``set_seed``), so an identical call sequence is what makes a seed reproduce
the reference's token. The training-time helpers (mask builders, bit-flip augmentation) are out of scope."""
from __future__ import annotations

import torch

from .. import ops
from ._lazy import NativeModule

_ACTS = {"gelu_pytorch_tanh": "gelu_tanh", "gelu_tanh": "gelu_tanh", "silu": "silu"}


class MLPconnector(NativeModule):
    def __init__(self, in_dim: int, out_dim: int, hidden_act: str):
        if hidden_act not in _ACTS:
            raise NotImplementedError(f"hidden_act {hidden_act!r} is not on the BitDance image path")
        super().__init__({"fc1.weight": (out_dim, in_dim), "fc1.bias": (out_dim,),
                          "fc2.weight": (out_dim, out_dim), "fc2.bias": (out_dim,)})
        self.act = _ACTS[hidden_act]
        self.in_dim, self.out_dim = in_dim, out_dim

    def _build_runner(self, device):
        bf = lambda t: t.detach().to(device, torch.bfloat16).contiguous()
        return dict(fc1_w=bf(self.fc1.weight), fc1_b=bf(self.fc1.bias), fc2_w=bf(self.fc2.weight), fc2_b=bf(self.fc2.bias))

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        r = self.runner
        shp = hidden_states.shape
        x = hidden_states.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
        h = ops.gemm(x, r["fc1_w"], bias=r["fc1_b"], act=self.act)
        y = ops.gemm(h, r["fc2_w"], bias=r["fc2_b"])
        return y.view(*shp[:-1], self.out_dim)


def top_k_top_p_filtering(logits: torch.Tensor, top_k: int = 0, top_p: float = 1.0, filter_value: float = -float("Inf"),
                          min_tokens_to_keep: int = 1) -> torch.Tensor:
    """modeling/utils.py:64-92. Keeps, per row, the ``top_k`` largest logits (ties with the k-th value are kept, as in the
    reference: strict ``<`` against the k-th value) and then the nucleus: in descending order a token survives iff the
    cumulative probability of the tokens BEFORE it is <= ``top_p`` (so the token that crosses the threshold is kept, and
    the first ``min_tokens_to_keep`` always are). Everything else becomes ``filter_value``. Returns a new tensor."""
    if top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, filter_value)
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(torch.nn.functional.softmax(sorted_logits, dim=-1), dim=-1)
        over = cum > top_p
        if min_tokens_to_keep > 1:
            over[..., :min_tokens_to_keep] = False
        drop_sorted = torch.zeros_like(over)
        drop_sorted[..., 1:] = over[..., :-1]           # shifted by one: judged on the mass strictly before the token
        drop = torch.zeros_like(over).scatter(-1, sorted_idx, drop_sorted)
        logits = logits.masked_fill(drop, filter_value)
    return logits


def sample_codebook(pred_logits: torch.Tensor, cur_item_type, codebook, do_sample: bool = True, temperature: float = 1.0,
                    top_k: int = 0, top_p: float = 1.0):
    """modeling/utils.py:95-124. pred_logits [B, vocab] -> (tokens [B] int64, codebook(tokens) [B, D]).
    ``multinomial`` draws from the global generator of the logits' device — the reference's RNG contract."""
    logits = pred_logits / max(temperature, 1e-5)
    if top_k > 0 or top_p < 1.0:
        logits = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
    probs = torch.nn.functional.softmax(logits, dim=-1)
    if do_sample:
        tokens = torch.multinomial(probs, num_samples=1).squeeze(-1)
    else:
        tokens = torch.argmax(probs, dim=-1)
    return tokens, codebook(tokens)


def remove_first_user_block(x: str) -> str:
    """modeling/utils.py:206-216: the unconditional twin of a chat prompt = the prompt without its first user turn."""
    head, tail = "<|im_start|>user\n", "<|im_end|>\n"
    i = x.find(head)
    if i < 0:
        return x
    j = x.find(tail, i + len(head))
    if j < 0:
        return x
    return x[:i] + x[j + len(tail):]
