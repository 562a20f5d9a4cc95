"""API mirror of ``imagenet_gen/src/model_parallel.py`` (the surface ``sample_ddp_parallel.py`` uses, :83,94-96,159-164):
``get_model_args()``, ``create_model(args, device)``, ``BitDance_models`` and the ``BitDance`` module with
``load_state_dict(strict=True)``, ``load_vae_weight()``, ``sample(cond, sample_steps, cfg_scale, cfg_schedule, chunk_size)``.

The module only HOLDS the parameters under the reference's state-dict key names (``layers.N.attention.wqkv.weight``,
``head.net.*``, ``vae.encoder.* / vae.decoder.*`` ...); sampling runs on the B200-native engine
(bitdance_b200/imagenet.py). Training (``forward``: the diffusion loss) is out of scope and raises."""
from __future__ import annotations

import argparse

import torch
from torch import nn

from ...ae import AERunner, ae_spec
from ...imagenet import MODELS, ImageNetEngine, imagenet_spec
from ...modeling._lazy import NativeModule


def get_model_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--model", type=str, choices=list(BitDance_models.keys()), default="BitDance-L")
    parser.add_argument("--image-size", type=int, choices=[256, 512], default=256)
    parser.add_argument("--down-size", type=int, default=16, choices=[16])
    parser.add_argument("--patch-size", type=int, default=1, choices=[1, 2, 4])
    parser.add_argument("--num-classes", type=int, default=1000)
    parser.add_argument("--cls-token-num", type=int, default=64)
    parser.add_argument("--latent-dim", type=int, default=16)
    parser.add_argument("--diff-batch-mul", type=int, default=4)
    parser.add_argument("--grad-checkpointing", action="store_true")
    parser.add_argument("--trained-vae", type=str, default="")
    parser.add_argument("--drop-rate", type=float, default=0.0)
    parser.add_argument("--perturb-schedule", type=str, default="constant")
    parser.add_argument("--perturb-rate", type=float, default=0.0)
    parser.add_argument("--perturb-rate-max", type=float, default=0.3)
    parser.add_argument("--time-schedule", type=str, default='logit_normal')
    parser.add_argument("--time-shift", type=float, default=1.)
    parser.add_argument("--parallel-num", type=int, default=4)
    parser.add_argument("--P-std", type=float, default=0.8)
    parser.add_argument("--P-mean", type=float, default=-0.8)
    parser.add_argument("--parallel-mode", type=str, default='patch', choices=['standard', 'patch'])
    return parser


def create_model(args, device):
    return BitDance_models[args.model](
        resolution=args.image_size, down_size=args.down_size, patch_size=args.patch_size, latent_dim=args.latent_dim,
        diff_batch_mul=args.diff_batch_mul, cls_token_num=args.cls_token_num, num_classes=args.num_classes,
        grad_checkpointing=args.grad_checkpointing, trained_vae=args.trained_vae, drop_rate=args.drop_rate,
        perturb_schedule=args.perturb_schedule, perturb_rate=args.perturb_rate, perturb_rate_max=args.perturb_rate_max,
        time_schedule=args.time_schedule, time_shift=args.time_shift, parallel_num=args.parallel_num, P_std=args.P_std,
        P_mean=args.P_mean, parallel_mode=args.parallel_mode).to(device)


class _Vae(NativeModule):
    """``qae.VQModel(ddconfig, num_codebooks)``: encoder / decoder parameters (the GFQ quantiser has no persistent state)."""

    def __init__(self, ddconfig):
        super().__init__(ae_spec(ddconfig))
        self.ddconfig = dict(ddconfig)

    def _build_runner(self, device):
        return AERunner(self.state_dict(), self.ddconfig, device=device)

    @torch.no_grad()
    def decode(self, quant):
        return self.runner.decode(quant)

    @torch.no_grad()
    def encode(self, x):
        """-> (quant, emb_loss, info, loss_breakdown) like qae.VQModel.encode; info = GFQ indices [4, B*h*w]."""
        q, _, idx, _ = self.runner.encode(x, num_codebooks=4)
        return q, None, idx, None


class BitDance(NativeModule):
    def __init__(self, dim, n_layer, n_head, diff_layers, diff_dim, diff_adanln_layers, latent_dim, down_size, patch_size,
                 resolution, diff_batch_mul, grad_checkpointing=False, cls_token_num=16, num_classes: int = 1000,
                 class_dropout_prob: float = 0.1, trained_vae: str = "", drop_rate: float = 0.0,
                 perturb_schedule: str = "constant", perturb_rate: float = 0.0, perturb_rate_max: float = 0.3,
                 time_schedule: str = 'logit_normal', time_shift: float = 1., parallel_num: int = 4, P_std: float = 1.,
                 P_mean: float = 0., parallel_mode: str = 'standard'):
        self.cfg = dict(dim=dim, n_layer=n_layer, n_head=n_head, diff_layers=diff_layers, diff_dim=diff_dim,
                        diff_adanln_layers=diff_adanln_layers, latent_dim=latent_dim, down_size=down_size,
                        patch_size=patch_size, resolution=resolution, cls_token_num=cls_token_num, num_classes=num_classes,
                        parallel_num=parallel_num, parallel_mode=parallel_mode, time_shift=time_shift)
        super().__init__(imagenet_spec(self.cfg))
        self.n_layer, self.resolution, self.down_size, self.patch_size = n_layer, resolution, down_size, patch_size
        self.num_classes, self.cls_token_num, self.latent_dim = num_classes, cls_token_num, latent_dim
        self.trained_vae, self.parallel_num, self.parallel_mode = trained_vae, parallel_num, parallel_mode
        self.h = self.w = resolution // (down_size * patch_size)
        self.total_tokens = self.h * self.w + cls_token_num
        ddconfig = {"double_z": False, "z_channels": latent_dim, "in_channels": 3, "out_ch": 3, "ch": 256,
                    "ch_mult": [1, 1, 2, 2, 4], "num_res_blocks": 4}
        self.vae = _Vae(ddconfig)

    def load_vae_weight(self):
        state = torch.load(self.trained_vae, map_location="cpu")
        missing, unexpected = self.vae.load_state_dict(
            {k: v for k, v in state["state_dict"].items() if k.startswith(("encoder.", "decoder."))}, strict=False)
        print(f"loading vae, missing_keys: {missing}")

    def _build_runner(self, device):
        sd = {k: v for k, v in self.state_dict().items() if not k.startswith("vae.")}
        return ImageNetEngine(sd, self.cfg, ae=self.vae.runner, device=device)

    def forward(self, images, class_id, cached=False):
        raise NotImplementedError("BitDance.forward is the training loss (model_parallel.py:274-335): out of scope")

    @torch.no_grad()
    def sample(self, cond, sample_steps, cfg_scale=1.0, cfg_schedule="linear", chunk_size=0):
        return self.runner.sample(cond, sample_steps, cfg_scale, cfg_schedule, chunk_size)


def BitDance_H(**kwargs):
    return BitDance(**MODELS["BitDance-H"], **kwargs)


def BitDance_L(**kwargs):
    return BitDance(**MODELS["BitDance-L"], **kwargs)


def BitDance_B(**kwargs):
    return BitDance(**MODELS["BitDance-B"], **kwargs)


BitDance_models = {"BitDance-B": BitDance_B, "BitDance-L": BitDance_L, "BitDance-H": BitDance_H}
