"""Multi-GPU ingest and egress (BASELINE north_star): utterances are independent units, so the path is
"replicas + scatter / gather" with NO exchange step inside it:

  ingest  rank 0 holds the packed int16 PCM of the whole step; every other rank receives its shard by grouped
          point-to-point sends (ncclSend / ncclRecv in one group = a scatter with uneven counts);
  compute every rank runs the single-GPU pipeline on its shard (model, extractor and HCLG are replicated);
  egress  every rank packs its finalized lattices on the device (b2k_dec_pack_lattices_async), the byte counts are
          all-gathered, and rank 0 receives the packed buffers with grouped sends / receives.

Sharding keeps all utterances of one speaker on one rank, in their original order, because the reference carries the
i-vector / CMVN adaptation state from one utterance of a speaker to the next (online2-wav-nnet3-latgen-faster.cc:199-221,
287-288): a speaker is the unit that is placed, longest speakers first.

Everything here is backend agnostic torch.distributed (NCCL with device tensors on the GPUs, gloo with host tensors in the
CPU tests, tests/test_multi_gloo.py); the per-rank compute is a callable."""
from __future__ import annotations

import numpy as np


def shard_speakers(speakers, lengths, world: int) -> list[np.ndarray]:
    """Utterance ids per rank.  A speaker's utterances stay together and keep their order (spk2utt order of the tool);
    speakers are placed longest total audio first on the least loaded rank (deterministic, ties by first utterance id)."""
    speakers = list(speakers)
    lengths = np.asarray(lengths, np.int64)
    assert len(speakers) == len(lengths)
    first, total, members = {}, {}, {}
    for u, s in enumerate(speakers):
        first.setdefault(s, u)
        total[s] = total.get(s, 0) + int(lengths[u])
        members.setdefault(s, []).append(u)
    order = sorted(members, key=lambda s: (-total[s], first[s]))
    loads = [0] * world
    shards = [[] for _ in range(world)]
    for s in order:
        r = min(range(world), key=lambda q: (loads[q], q))
        shards[r].append(s)
        loads[r] += total[s]
    out = []
    for r in range(world):
        ids = [u for s in sorted(shards[r], key=lambda s: first[s]) for u in members[s]]
        out.append(np.asarray(ids, np.int64))
    return out


def speaker_waves(speakers) -> list[np.ndarray]:
    """The order in which ONE rank decodes its utterances when adaptation state is carried from one utterance of a speaker to
    the next (online2-wav-nnet3-latgen-faster.cc:199-221,287): wave k = the k-th utterance of every speaker that has one, so
    the utterances of a wave are independent (one batch, or several) and a speaker's state is final before its next utterance
    starts.  Positions into `speakers`; speakers in order of first appearance."""
    seen, nth = {}, []
    for s in speakers:
        nth.append(seen.get(s, 0))
        seen[s] = nth[-1] + 1
    nth = np.asarray(nth, np.int64)
    return [np.nonzero(nth == k)[0] for k in range(int(nth.max()) + 1)] if len(nth) else []


def _wire(t):
    """The tensor as the transport sees it: torch's NCCL wrapper has no 16-bit integer type ("Input tensor data type is not
    supported for NCCL process group: Short"), so anything but bytes / floats travels as a byte view of the same storage."""
    import torch
    return t if t.dtype in (torch.uint8, torch.float32, torch.int64, torch.int32) else t.contiguous().view(torch.uint8)


def scatter_rows(rows0, counts, rank: int, world: int, device=None):
    """rank 0: rows0 = [sum(counts) x S] tensor already ordered by destination rank; every rank gets its [counts[rank] x S]
    block.  One group of point-to-point operations (ncclGroupStart / ncclSend... / ncclGroupEnd under NCCL)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return rows0, 0
    offs = np.concatenate([[0], np.cumsum(counts)])
    if rank == 0:
        ops = [dist.P2POp(dist.isend, _wire(rows0[offs[r]:offs[r + 1]]), r) for r in range(1, world) if counts[r] > 0]
        sent = int(sum(rows0[offs[r]:offs[r + 1]].numel() * rows0.element_size() for r in range(1, world)))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        for q in reqs:
            q.wait()
        return rows0[offs[0]:offs[1]], sent
    shape, dtype = scatter_rows.meta                        # set by broadcast_plan
    mine = torch.empty((int(counts[rank]),) + tuple(shape), dtype=dtype, device=device)
    if counts[rank] > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, _wire(mine), 0)]):       # (a view: the bytes land in `mine`)
            q.wait()
    return mine, 0


def broadcast_plan(counts, row_shape, dtype, rank: int, world: int, device=None):
    """Rank 0 decides the shard sizes; the others learn them (and the row shape) from one small broadcast."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(world + 2, dtype=torch.int64, device=device)
    if rank == 0:
        t[:world] = torch.as_tensor(np.asarray(counts, np.int64))
        t[world] = int(row_shape[0]) if len(row_shape) else 0
        t[world + 1] = {torch.int16: 0, torch.float32: 1, torch.uint8: 2}[dtype]
    if world > 1:
        dist.broadcast(t, 0)
    t = t.cpu().numpy()
    shape = (int(t[world]),) if t[world] > 0 else ()
    scatter_rows.meta = (shape, [torch.int16, torch.float32, torch.uint8][int(t[world + 1])])
    return t[:world].copy()


def gather_bytes(local, rank: int, world: int, device=None):
    """Variable-length gather to rank 0: all-gather of the byte counts, then one group of sends / receives.
    local: 1-D uint8 tensor.  Returns (list of per-rank uint8 tensors on rank 0 | None, bytes that crossed the link)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local], 0
    n = torch.tensor([local.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    if rank == 0:
        parts = [local] + [torch.empty(sizes[r], dtype=torch.uint8, device=device) for r in range(1, world)]
        ops = [dist.P2POp(dist.irecv, parts[r], r) for r in range(1, world) if sizes[r] > 0]
        for q in (dist.batch_isend_irecv(ops) if ops else []):
            q.wait()
        return parts, int(sum(sizes[1:]))
    if local.numel() > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, 0)]):
            q.wait()
    return None, 0


def run_step(pcm0, speakers, lengths, rank: int, world: int, compute, device=None):
    """One multi-GPU step.  rank 0: pcm0 = [U x S] int16 tensor (device tensor under NCCL) of the whole step, speakers /
    lengths describe it (None elsewhere).  compute(shard [n x S]) -> 1-D uint8 tensor (that rank's packed lattices).
    Returns on rank 0: (per-rank packed buffers, utterance ids per rank, collective bytes); elsewhere (None, None, 0)."""
    import torch
    shards = shard_speakers(speakers, lengths, world) if rank == 0 else None
    counts = [len(s) for s in shards] if rank == 0 else None
    row_shape = tuple(pcm0.shape[1:]) if rank == 0 else ()
    counts = broadcast_plan(counts, row_shape, pcm0.dtype if rank == 0 else torch.int16, rank, world, device)
    rows0 = None
    if rank == 0:
        order = np.concatenate(shards) if len(shards) else np.zeros(0, np.int64)
        rows0 = pcm0[torch.as_tensor(order, device=pcm0.device)] if world > 1 else pcm0[torch.as_tensor(shards[0], device=pcm0.device)]
    mine, sent = scatter_rows(rows0, counts, rank, world, device)
    packed = compute(mine)
    parts, recvd = gather_bytes(packed, rank, world, device)
    if rank == 0:
        return parts, shards, sent + recvd
    return None, None, 0
