"""Two-GPU check of the sharded path (skipped on a one-GPU box): two NCCL ranks sketch their
genomes with the CUDA sketcher, all-gather the signatures over NVLink, build the same index
replica, search their share of the queries and gather the answers; everything must equal the
single-GPU result in global order."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nfiles, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import gsearch_b200 as g
        from gsearch_b200 import sharding

        files = [g.synth.dna_genome(i, 150_000 + 1000 * i) for i in range(nfiles)]
        sk = g.Sketcher(g.SeqSketcherParams(21, 2048), device=rank)
        sig, nb = sharding.sketch_sharded(sk.sketch_files, files, rank, world, device=f"cuda:{rank}")
        # tohnsw: insertion does not shard -- every rank builds the same (deterministic) replica
        idx = g.Hnsw(g.HnswParams(max_nb_conn=8, ef=32), 2048, sig.dtype, device=rank)
        idx.parallel_insert(sig, np.arange(nfiles, dtype=np.uint64))
        # request: queries sharded over the ranks, answers gathered
        out, cnt = sharding.search_sharded(lambda qs: idx.search_raw(qs, 4, 32)[:2], sig, 4, rank, world,
                                           device=f"cuda:{rank}")
        q.put((rank, sig, nb, out["d_id"].copy(), out["distance"].copy(), cnt))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sketch_equals_one_gpu():
    import torch
    import torch.multiprocessing as mp
    import gsearch_b200 as g

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    nfiles = 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nfiles, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = [g.synth.dna_genome(i, 150_000 + 1000 * i) for i in range(nfiles)]
    want, wnb = g.Sketcher(g.SeqSketcherParams(21, 2048)).sketch_files(files)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=8, ef=32), 2048, want.dtype)
    idx.parallel_insert(want, np.arange(nfiles, dtype=np.uint64))
    wout, wcnt, _ = idx.search_raw(want, 4, 32)
    for rank, sig, nb, ids, dd, cnt in got:
        assert sig.tobytes() == want.tobytes(), f"rank {rank}"
        assert nb.tolist() == wnb.tolist()
        assert cnt.tolist() == wcnt.tolist() and ids.tolist() == wout["d_id"].tolist()
        assert dd.tobytes() == wout["distance"].tobytes()
