"""API mirror of the reference ``modeling/utils.py`` (hot-path part): ``MLPconnector``.

Reference: modeling/utils.py:9-20 — fc2(act(fc1(x))) with ``hidden_act='gelu_pytorch_tanh'`` on the T2I path. The two
Linears run on the tcgen05 GEMM with the activation fused in fc1's epilogue. The training-time helpers of the reference
file (mask builders, top-k/p sampling, bit-flip augmentation) are out of scope."""
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
