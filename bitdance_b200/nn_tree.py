"""Parameter containers that reproduce a reference module's state-dict key names without re-implementing its
Python classes: a tree of bare ``nn.Module`` nodes built from a flat ``{dotted.name: shape}`` spec."""
from __future__ import annotations

import torch
from torch import nn


class ParamNode(nn.Module):
    """A bare container; parameters/children are attached by ``build_param_tree``."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ParamNode holds weights only; compute goes through the bitdance_b200 C ABI")


def build_param_tree(spec: dict, dtype=torch.float32, device="meta") -> ParamNode:
    root = ParamNode()
    for name, shape in spec.items():
        parts = name.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamNode())
            node = node._modules[p]
        node.register_parameter(parts[-1], nn.Parameter(torch.empty(tuple(shape), dtype=dtype, device=device),
                                                        requires_grad=False))
    return root
