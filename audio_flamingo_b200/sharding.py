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
    prefix = [0]
    for w in weights:
        prefix.append(prefix[-1] + w)
    # boundary r goes where the cumulative weight is CLOSEST to r/world of the total (ties: the earlier row), never before the
    # previous boundary: every shard ends within half a sequence of its ideal load
    cuts = [0]
    for r in range(1, world_size):
        best = cuts[-1]
        for i in range(cuts[-1], n_rows + 1):
            if abs(prefix[i] * world_size - r * total) < abs(prefix[best] * world_size - r * total):
                best = i
        cuts.append(best)
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


def kv_bytes_per_token(n_layers: int, n_kv_heads: int, head_dim: int, dtype_bytes: int = 2) -> int:
    """K + V bytes one cached token costs per sequence (AF3-7B: 28 x 4 x 128 x 2 B x 2 = 57 344 B)."""
    return 2 * n_layers * n_kv_heads * head_dim * dtype_bytes


def plan_kv_capacity(prompt_lens: list[int], max_new_tokens: int, *, n_layers: int, n_kv_heads: int, head_dim: int,
                     hbm_free_bytes: int, activation_bytes_per_token: int = 0, prefill_chunk_size: int | None = None,
                     bucket: int = 256) -> dict:
    """Capacity planning for long-audio batches (up to 20 windows -> 15 000 audio tokens per sequence,
    0.86 GB of KV cache each): how many of the given prompts fit one GPU at once, and what the batch costs.

    The cache is one [B, Hkv, Tmax, D] block per layer with Tmax = the longest (prompt + max_new_tokens) rounded up to `bucket`
    rows (generate()'s capacity rule), so a batch costs B * Tmax * kv_bytes_per_token.  Prefill activations scale with the number
    of prompt rows in flight: B * min(S, prefill_chunk_size) * activation_bytes_per_token (pass the model's figure: for AF3-7B
    about (3584 * 4 + 2 * 18944 + 4608) * 2 B = 115 kB per token with the fused SwiGLU path).
    Returns {"fits", "max_batch", "tmax", "kv_bytes", "activation_bytes", "bytes_per_sequence"}: max_batch counts how many
    sequences as long as the LONGEST one fit, i.e. a safe per-GPU batch for this length class.
    """
    if not prompt_lens:
        return {"fits": True, "max_batch": 0, "tmax": 0, "kv_bytes": 0, "activation_bytes": 0, "bytes_per_sequence": 0}
    S = max(prompt_lens)
    tmax = -(-(S + max_new_tokens) // bucket) * bucket
    per_tok = kv_bytes_per_token(n_layers, n_kv_heads, head_dim)
    rows_in_flight = min(S, prefill_chunk_size) if prefill_chunk_size else S
    per_seq = tmax * per_tok + rows_in_flight * activation_bytes_per_token
    B = len(prompt_lens)
    return {"fits": B * per_seq <= hbm_free_bytes, "max_batch": int(hbm_free_bytes // per_seq) if per_seq else 0, "tmax": tmax,
            "kv_bytes": B * tmax * per_tok, "activation_bytes": B * rows_in_flight * activation_bytes_per_token,
            "bytes_per_sequence": per_seq}
