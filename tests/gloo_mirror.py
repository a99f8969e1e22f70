"""Test-only gloo mirror of the library's multi-GPU data flow (the product is csrc/dist.{h,cpp}: NCCL through dlopen,
no torch).  Same shard rule, same single all-gather, same sharded zero-shot -- run on torch.distributed/gloo by
tests/test_distributed_cpu.py so the N > 1 logic is exercised on a box without GPUs, and compared with the library's own
clip_b200_debug_shard_bounds."""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous balanced shards [n*r//w, n*(r+1)//w) -- the rule the library uses in devices mode (csrc/dist.h)."""
    return n * rank // world, n * (rank + 1) // world


def all_gather_rows(dist, local, n_total: int, world: int):
    """local: [n_local, d] tensor on this rank -> [n_total, d] on every rank (ragged shards are padded to the largest)."""
    import torch
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    width = max(sizes)
    d = local.shape[1]
    pad = torch.zeros((width, d), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    rows = [out[r * width: r * width + sizes[r]] for r in range(world)]
    return torch.cat(rows, 0)


def zero_shot_sharded(dist, rank, world, img_local, txt_local, n_txt_total, top_k):
    """configs[4] pattern: each rank holds its image shard and its LABEL shard; text embeddings are all-gathered
    (the one collective), logits + softmax_with_sorting arithmetic (clip.cpp:1591-1622) run locally per image."""
    import torch
    txt = all_gather_rows(dist, txt_local, n_txt_total, world) if world > 1 else txt_local
    logits = img_local @ txt.T
    e = torch.exp(logits.double()).float() + 1e-9
    p = (e / e.double().sum(1, keepdim=True)).float()
    scores, idx = torch.sort(p, dim=1, descending=True, stable=True)
    return scores[:, :top_k], idx[:, :top_k]
