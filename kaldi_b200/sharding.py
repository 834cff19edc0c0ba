"""Multi-GPU plan: utterances are independent, so the path is
"replicas + sharding" — one process per GPU, each with a full replica of model,
extractor and HCLG; no collective on the data path.  torch.distributed is used
for the barrier, the max-over-ranks timing and the gather of (small) results."""
from __future__ import annotations

import numpy as np


def shard_utterances(lengths, rank: int, world: int) -> np.ndarray:
    """Greedy longest-first bin packing by audio length (utterance ids of `rank`);
    deterministic on every rank."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world, np.int64)
    owner = np.zeros(len(lengths), np.int64)
    for u in order:
        r = int(np.argmin(loads))
        owner[u] = r
        loads[r] += int(lengths[u])
    return np.nonzero(owner == rank)[0]


def gather_results(local: dict, world: int) -> dict:
    """All ranks' {utterance id: result} dicts merged on every rank."""
    import torch.distributed as dist
    if world == 1:
        return dict(local)
    parts = [None] * world
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged
