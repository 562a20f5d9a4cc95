"""API mirror of ``modeling/vision_head/flow_head_parallel_x.py``: ``DiffHead`` (inference surface).

Same constructor arguments and state-dict keys (``net.*``) as the reference (flow_head_parallel_x.py:32-68,254-300);
``sample(z, cfg, num_sampling_steps)`` (:107-120) runs the whole Euler–Maruyama sampler as one native call
(bitdance_b200/head.py -> bd_head_sample). ``forward`` (the training loss, :70-105) is out of scope."""
from __future__ import annotations

import torch

from ...head import HeadRunner, head_spec
from .._lazy import NativeModule


class DiffHead(NativeModule):
    def __init__(self, ch_target, ch_cond, ch_latent, depth_latent, depth_adanln, grad_checkpointing=False,
                 time_shift=1., time_schedule='logit_normal', P_mean=0., P_std=1., parallel_num=4, diff_batch_mul=1,
                 use_swiglu=False):
        super().__init__(head_spec(ch_target, ch_cond, ch_latent, depth_latent, depth_adanln, use_swiglu))
        self.ch_target, self.time_shift, self.time_schedule = ch_target, time_shift, time_schedule
        self.P_mean, self.P_std, self.diff_batch_mul = P_mean, P_std, diff_batch_mul
        self._cfg = dict(ch_target=ch_target, ch_cond=ch_cond, ch_latent=ch_latent, depth_latent=depth_latent,
                         depth_adanln=depth_adanln, use_swiglu=use_swiglu)
        self.parallel_num = parallel_num

    def _build_runner(self, device):
        return HeadRunner(self.state_dict(), device=device, time_shift=self.time_shift, **self._cfg)

    def forward(self, x, cond):
        raise NotImplementedError("DiffHead.forward is the training loss (reference :70-105): out of scope of the "
                                  "B200 inference path")

    @torch.no_grad()
    def sample(self, z, cfg, num_sampling_steps):
        """Returns cat([x] * cfg_mult) like sampling_x.euler_maruyama (:97)."""
        mult = 2 if cfg > 1.0 else 1
        x = self.runner.sample(z.to(torch.float32), float(cfg), int(num_sampling_steps))
        return torch.cat([x] * mult, dim=0)
