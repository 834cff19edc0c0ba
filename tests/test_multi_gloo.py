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


# ----------------------------------------------------------------------------- ingest / egress

def _ingest_worker(rank, world, port, q):
    from kaldi_b200 import ingest
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, S = 11, 64
    rng = np.random.default_rng(5)
    pcm = torch.from_numpy(rng.integers(-30000, 30000, size=(U, S)).astype(np.int16)) if rank == 0 else None
    speakers = ["a", "b", "a", "c", "b", "a", "d", "c", "e", "e", "a"] if rank == 0 else None
    lengths = [S] * U if rank == 0 else None

    def compute(shard):
        # stand-in for the per-rank pipeline: the "packed lattices" are a function of the rows a rank received, in order
        body = (shard.to(torch.int32) * 3 + 1).to(torch.int16).contiguous().view(torch.uint8).reshape(-1)
        head = torch.tensor([shard.shape[0]], dtype=torch.int64).view(torch.uint8)
        return torch.cat([head, body])
    parts, shards, nbytes = ingest.run_step(pcm, speakers, lengths, rank, world, compute)
    if rank == 0:
        out = {}
        for r, (buf, ids) in enumerate(zip(parts, shards)):
            n = int(buf[:8].view(torch.int64)[0])
            rows = buf[8:].view(torch.int16).reshape(n, S)
            for k, u in enumerate(ids):
                out[int(u)] = rows[k].numpy().copy()
        q.put((out, [s.tolist() for s in shards], nbytes, pcm.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_ingest_scatter_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ingest_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out, shards, nbytes, pcm = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    # every utterance came back once, computed from exactly its own samples
    assert sorted(out) == list(range(11))
    for u, row in out.items():
        np.testing.assert_array_equal(row, (pcm[u].astype(np.int32) * 3 + 1).astype(np.int16))
    # a speaker lives on one rank, in the tool's order (online2-wav-nnet3-latgen-faster.cc:199-221)
    spk = ["a", "b", "a", "c", "b", "a", "d", "c", "e", "e", "a"]
    owner = {}
    for r, ids in enumerate(shards):
        for u in ids:
            assert owner.setdefault(spk[u], r) == r
        for s in set(spk[u] for u in ids):
            mine = [u for u in ids if spk[u] == s]
            assert mine == sorted(mine)
    # bytes that crossed the link: rank 1's PCM shard out, its packed result back
    n1 = len(shards[1])
    assert nbytes == n1 * 64 * 2 + (8 + n1 * 64 * 2)


def test_speaker_shards_are_balanced():
    from kaldi_b200.ingest import shard_speakers
    rng = np.random.default_rng(1)
    spk = [f"s{int(x)}" for x in rng.integers(0, 40, size=300)]
    lengths = rng.integers(16000, 320000, size=300)
    for world in (1, 2, 4, 8):
        shards = shard_speakers(spk, lengths, world)
        assert np.array_equal(np.sort(np.concatenate(shards)), np.arange(300))
        per_spk = {}
        for u, s in enumerate(spk):
            per_spk[s] = per_spk.get(s, 0) + int(lengths[u])
        loads = [int(lengths[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= max(per_spk.values())


def test_speaker_waves_keep_a_speakers_utterances_in_successive_batches():
    """ingest.speaker_waves: wave k holds the k-th utterance of every speaker (so a batch never holds two utterances of one
    speaker and the adaptation state is final before the speaker's next utterance), every utterance exactly once, a speaker's
    utterances in their original order (online2-wav-nnet3-latgen-faster.cc:199-221)."""
    from kaldi_b200.ingest import speaker_waves
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(1, 40))
        spk = [f"s{int(rng.integers(0, 7))}" for _ in range(n)]
        waves = speaker_waves(spk)
        flat = np.concatenate(waves)
        assert sorted(flat.tolist()) == list(range(n))
        when = {}
        for k, w in enumerate(waves):
            assert len({spk[u] for u in w}) == len(w), "two utterances of a speaker in one wave"
            for u in w:
                when[u] = k
        for s in set(spk):
            ids = [u for u in range(n) if spk[u] == s]
            assert [when[u] for u in ids] == list(range(len(ids)))
    assert speaker_waves([]) == []
