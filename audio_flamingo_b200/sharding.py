"""Batch sharding across the GPUs of one box and the final token gather.

Every 30 s window is independent through mel -> encoder -> projector and every sequence is independent through
prefill and decode, so the path shards by rows with weights replicated and NO collective inside the data path.
The only exchange is one all-gather of the generated token ids [B_local, L] (int64; 32 KB per rank at config 2),
issued over torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch


def shard_rows(n_rows: int, world_size: int, rank: int, weights: list[int] | None = None) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `n_rows` sequences for `rank`.

    Without weights: sizes differ by at most one (first ranks get the extra rows).  With weights (e.g. windows per
    sequence for ragged long-audio batches) the cut points balance the cumulative weight; all windows of a sequence
    stay on one rank.
    """
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    if weights is None:
        base, extra = divmod(n_rows, world_size)
        begin = rank * base + min(rank, extra)
        return begin, begin + base + (1 if rank < extra else 0)
    if len(weights) != n_rows:
        raise ValueError("weights must have one entry per row")
    total = sum(weights)
    cuts, acc, r = [0], 0, 1
    for i, w in enumerate(weights):
        acc += w
        while r < world_size and acc * world_size >= r * total and len(cuts) <= r:
            cuts.append(i + 1)
            r += 1
    while len(cuts) < world_size:
        cuts.append(n_rows)
    cuts.append(n_rows)
    return cuts[rank], cuts[rank + 1]


def gather_tokens(local_tokens: torch.Tensor, counts: list[int] | None = None, group=None, pad_token_id: int = 0,
                  assume_equal_length: bool = False) -> torch.Tensor:
    """All-gather generated ids [B_local, L_local] -> [sum B_local, max L] on every rank (rows in rank order).

    counts: rows per rank (needed only when they differ; shorter shards are padded for the collective and trimmed).
    Ranks may hold different L: generate() returns fewer columns on a rank whose rows all hit EOS early, and prompts are left
    padded per shard.  The column counts are exchanged first (one tiny all-gather), every shard is right-padded with
    `pad_token_id` to the longest one for the id all-gather, so the collective always sees equal shapes
    (assume_equal_length=True skips the exchange when the caller knows every rank returns the same L).
    Without an initialised process group (single GPU) this is the identity.
    """
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_tokens
    world = dist.get_world_size(group)
    if world == 1:
        return local_tokens
    if counts is None:
        counts = [local_tokens.shape[0]] * world
    mx = max(counts)
    dev = local_tokens.device
    if assume_equal_length:  # caller guarantees it (no EOS, same prompt length everywhere): skips the length exchange and its host sync
        L = local_tokens.shape[1]
    else:
        lens = torch.empty((world,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(lens, torch.tensor([local_tokens.shape[1]], dtype=torch.int64, device=dev), group=group)
        L = int(lens.max())
    buf = local_tokens
    if buf.shape[0] < mx or buf.shape[1] < L:
        buf = torch.full((mx, L), pad_token_id, dtype=local_tokens.dtype, device=dev)
        buf[: local_tokens.shape[0], : local_tokens.shape[1]] = local_tokens
    out = torch.empty((world * mx, L), dtype=local_tokens.dtype, device=dev)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)
