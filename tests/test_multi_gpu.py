"""Two GPUs of one box, NCCL: the ingest / egress path of SURVEY.md §8e end to end.  Rank 0 holds the packed int16 PCM of the
whole list, scatters speaker-ordered shards (grouped ncclSend / ncclRecv), both ranks run the C-ABI pipeline on their shard
and pack their lattices on the device, rank 0 gathers the packed buffers (size all-gather + grouped send / recv) and unpacks
them.  The lattices that arrive on rank 0 must be bit-identical to a single-GPU run of the same list.
Needs 2 GPUs (`gpurun --gpus 2`): skipped on the 1-GPU lease the round-end test run uses; the log of a 2-GPU run is kept
under profiles/."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _build(B, S):
    from kaldi_b200 import ivector as IVM, nnet_model as NM, synth
    from kaldi_b200.decoder import CudaFst
    from kaldi_b200.feat import FeatureOptions
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.pipeline import NativeBatchedPipeline, PipelineConfig
    P = 200
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(150_000, num_pdfs=P, seed=4)
    cfg = PipelineConfig(feature_opts=FeatureOptions(max_lanes=max(B, 8)), max_batch=B, num_samples=S)
    T_feat = 1 + (S - 400) // 160
    ivx = IVM.IvectorExtractorGpu(IVM.make_synthetic_extractor(3, num_gauss=64, ivector_dim=100), B, T_feat)
    return NativeBatchedPipeline(cfg, KaldiModel.from_arch(arch, W), CudaFst(g), ivx), g


def _worker(rank, world, port, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from kaldi_b200 import _lib, ingest, synth
    from kaldi_b200.decoder import CudaDecoder, lattice_to_canonical
    from kaldi_b200.pipeline import NativeBatchedPipeline
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    U, S = 8, 32000
    pipe, g = _build(U, S)
    L = _lib.lib()
    L.b2k_dec_pack_header_bytes.restype = C.c_int64
    L.b2k_dec_pack_header_bytes.argtypes = [C.c_int32]
    speakers = ["a", "b", "a", "c", "b", "d", "c", "a"]
    pcm = None
    if rank == 0:
        pcm = torch.from_numpy(np.stack([synth.make_audio(S, seed=50 + i) for i in range(U)]).astype(np.int16)).to(dev)
    cap = 64 << 20
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def compute(shard):
        n = int(shard.shape[0])
        pipe.run_device_i16(shard.data_ptr(), n, stream)
        pipe.pack_device(n, buf.data_ptr(), cap, stream)
        hdr = buf[:int(L.b2k_dec_pack_header_bytes(n))].cpu().numpy().view(np.int64)
        status, need = NativeBatchedPipeline.packed_bytes_needed(hdr, n)
        assert status == 0, (rank, status, need)
        return buf[:need]
    parts, shards, nbytes = ingest.run_step(pcm, speakers if rank == 0 else None, [S] * U if rank == 0 else None, rank, world, compute, dev)
    if rank == 0:
        got = {}
        for p_, ids in zip(parts, shards):
            lats = CudaDecoder.SplitLattices(NativeBatchedPipeline.unpack_lattices(p_.cpu().numpy(), len(ids)))
            for lat, u in zip(lats, ids):
                got[int(u)] = lattice_to_canonical(lat)
        # the same list on one GPU, in list order
        pipe.run_device_i16(pcm.data_ptr(), U, stream)
        pipe.pack_device(U, buf.data_ptr(), cap, stream)
        torch.cuda.synchronize()
        hdr = buf[:int(L.b2k_dec_pack_header_bytes(U))].cpu().numpy().view(np.int64)
        status, need = NativeBatchedPipeline.packed_bytes_needed(hdr, U)
        assert status == 0
        single = CudaDecoder.SplitLattices(NativeBatchedPipeline.unpack_lattices(buf[:need].cpu().numpy(), U))
        same = all(all(np.array_equal(got[u][k], lattice_to_canonical(single[u])[k]) for k in got[u]) for u in range(U))
        q.put((sorted(got), [s.tolist() for s in shards], nbytes, same, sum(len(got[u]["states"]) for u in got)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_ingest_egress_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    import queue as _queue
    import time
    res, t0 = None, time.time()
    while res is None and time.time() - t0 < 240:            # a worker that died (exception before the gather) must not hang the test
        try:
            res = q.get(timeout=2)
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    if res is None:
        [p.kill() for p in procs if p.is_alive()]
        pytest.fail("a worker failed or timed out (exit codes %s)" % [p.exitcode for p in procs])
    ids, shards, nbytes, same, nstates = res
    [p.join(timeout=60) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert ids == list(range(8))
    assert len(shards[1]) > 0 and nbytes > len(shards[1]) * 32000 * 2        # PCM out + packed lattices back
    assert nstates > 0
    assert same, "lattices gathered on rank 0 differ from the single-GPU run of the same list"
    print(f"two-GPU ingest/egress: shards {shards}, {nbytes} bytes over NVLink, {nstates} lattice states, identical to the single-GPU run")
