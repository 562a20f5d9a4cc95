"""Shared base for the API-mirror modules: an nn.Module that only HOLDS parameters under the reference's state-dict key
names and builds its native runner (prepacked bf16 weights + C-ABI calls) on first use / after (re)loading weights."""
from __future__ import annotations

import torch
from torch import nn

from ..nn_tree import build_param_tree


class NativeModule(nn.Module):
    def __init__(self, spec: dict):
        super().__init__()
        tree = build_param_tree(spec, dtype=torch.float32, device="cpu")
        for name, child in tree.named_children():
            self.add_module(name, child)
        for name, p in tree.named_parameters(recurse=False):
            self.register_parameter(name, p)
        self._runner = None
        self._runner_device = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._runner = None

    def _apply(self, fn, *a, **k):  # .to(device) / .half() etc. invalidate the prepack
        self._runner = None
        return super()._apply(fn, *a, **k)

    def _device(self):
        return next(self.parameters()).device

    def _build_runner(self, device):  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def runner(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("bitdance_b200 modules compute on a B200 only (move the module to 'cuda'); "
                               "there is no CPU fallback")
        if self._runner is None or self._runner_device != dev:
            self._runner = self._build_runner(dev)
            self._runner_device = dev
        return self._runner
