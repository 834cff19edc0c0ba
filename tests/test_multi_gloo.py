"""World-size-2 gloo test of the host-side multi-GPU logic (CPU): utterance
sharding covers every utterance exactly once, per-rank results gather in order,
and the timing reduction takes the max over ranks (the contract bench.py uses)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kaldi_b200.sharding import shard_utterances, gather_results


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lengths, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_utterances(lengths, rank, world)
    # fake "decode": result = utterance id * 10 + rank-independent function of the length
    local = {int(u): int(lengths[u]) * 3 for u in mine}
    merged = gather_results(local, world)
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((sorted(merged.items()), float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_gather_world2():
    lengths = np.array([160000, 80000, 160000, 32000, 48000, 160000, 16000], np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, tmax = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert [k for k, _ in merged] == list(range(len(lengths)))
    assert [v for _, v in merged] == [int(x) * 3 for x in lengths]
    assert abs(tmax - 0.2) < 1e-9


def test_shards_are_balanced_and_disjoint():
    rng = np.random.default_rng(0)
    lengths = rng.integers(16000, 480000, size=101)
    for world in (1, 2, 4, 8):
        shards = [shard_utterances(lengths, r, world) for r in range(world)]
        allu = np.sort(np.concatenate(shards))
        assert np.array_equal(allu, np.arange(len(lengths)))
        loads = [lengths[s].sum() for s in shards]
        assert max(loads) - min(loads) <= lengths.max()          # greedy longest-first bin packing
