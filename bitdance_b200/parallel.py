"""Multi-GPU: replicas only. Images/prompts are independent units, so the path shards by unit with NO data-path
collective. What exists:

  * the two sharding schemes the reference's drivers use, as pure functions:
      - ``contiguous_shard``  eval/eval_dpg.py:24-29  (ceil-split of the prompt list)
      - ``strided_batches``   imagenet_gen/sample_ddp_parallel.py:143-150 (iteration k, rank r takes
                              [world*n*k + r*n, ... + n))
      - ``rank_seed``         sample_ddp_parallel.py:72 / eval/base_evaluator.py:27  (seed * world + rank)
  * ``gather_token_grids``: ONE end-of-batch all-gather of the PACKED token grids (uint32 words, 16 KB per 1024px image)
    — the reference exchanges results through PNG files instead (sample_ddp_parallel.py:173-195);
  * ``broadcast_tensors``: optional start-up broadcast of prepacked weights from rank 0 (replaces N disk reads).
Works on NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def rank_seed(seed: int, world: int, rank: int) -> int:
    return seed * world + rank


def contiguous_shard(n_items: int, world: int, rank: int) -> range:
    per = math.ceil(n_items / world)
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def strided_batches(n_items: int, per_rank_batch: int, world: int, rank: int):
    """Yields index ranges: iteration k covers [world*n*k + rank*n, +n), clipped to n_items."""
    global_batch = per_rank_batch * world
    iters = math.ceil(n_items / global_batch)
    for k in range(iters):
        lo = global_batch * k + rank * per_rank_batch
        yield range(min(n_items, lo), min(n_items, lo + per_rank_batch))


def gather_token_grids(packed_local: torch.Tensor, group=None) -> torch.Tensor:
    """packed_local: int32 [B_local, hw, words] (same shape on every rank) -> [world * B_local, hw, words], rank-major."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed_local
    world = dist.get_world_size(group)
    out = torch.empty((world * packed_local.shape[0], *packed_local.shape[1:]), dtype=packed_local.dtype,
                      device=packed_local.device)
    dist.all_gather_into_tensor(out, packed_local.contiguous(), group=group)
    return out


def broadcast_tensors(tensors, src: int = 0, group=None, include_integer: bool = False) -> int:
    """Start-up broadcast of prepacked WEIGHTS from ``src``. Integer tensors are skipped unless asked for: the engines keep
    tables of DEVICE POINTERS (int64 addresses of the local process's allocations, e.g. the per-layer table of the
    one-launch Qwen3 block) next to their weights, and another rank's addresses are poison. Returns the bytes shipped."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    n = 0
    for t in tensors:
        if t is None or (not include_integer and not (t.is_floating_point() or t.is_complex())):
            continue
        dist.broadcast(t, src=src, group=group)
        n += t.numel() * t.element_size()
    return n
