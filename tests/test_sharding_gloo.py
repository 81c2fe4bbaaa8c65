"""N > 1 host logic on CPU: two gloo processes shard genomes / queries as bench.py does on
GPUs, and the gathered result must equal the single-process result, in global order.  The
per-rank compute is the CPU oracle here (test infrastructure); on the GPU box the same
functions wrap the CUDA sketcher and index (tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nfiles, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _oracle as O
        import gsearch_b200 as g
        from gsearch_b200 import sharding
        from gsearch_b200.index import NEIGHBOUR_DTYPE

        files = [g.synth.dna_genome(i, 20_000 + 500 * i) for i in range(nfiles)]
        sig, nb = sharding.sketch_sharded(lambda fs: O.sketch_files(fs, 16, 128, nthreads=1), files, rank, world)
        # request: replicated graph, sharded queries
        h = O.Hnsw(8, 32, 128, np.uint32)
        h.insert(sig, np.arange(nfiles, dtype=np.uint64))

        def search(qs):
            out, cnt, _ = h.search(qs, 3, 20)
            o = np.zeros(out.shape, dtype=NEIGHBOUR_DTYPE)
            for k in ("d_id", "distance", "layer", "rank"):
                o[k] = out[k]
            return o, cnt

        res, cnt = sharding.search_sharded(search, sig, 3, rank, world)
        q.put((rank, sig, nb, res["d_id"].copy(), res["distance"].copy(), cnt))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nfiles", [(2, 7), (2, 8), (3, 10)])
def test_sharded_sketch_and_request_equal_single_process(oracle, world, nfiles):
    import gsearch_b200 as g

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nfiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = [g.synth.dna_genome(i, 20_000 + 500 * i) for i in range(nfiles)]
    want, wnb = oracle.sketch_files(files, 16, 128, nthreads=2)
    h = oracle.Hnsw(8, 32, 128, np.uint32)
    h.insert(want, np.arange(nfiles, dtype=np.uint64))
    wout, wcnt, _ = h.search(want, 3, 20)
    for rank, sig, nb, ids, dd, cnt in got:
        assert sig.dtype == want.dtype and sig.tobytes() == want.tobytes(), f"rank {rank}"
        assert nb.tolist() == wnb.tolist()
        assert cnt.tolist() == wcnt.tolist()
        assert ids.tolist() == wout["d_id"].tolist()
        assert dd.tobytes() == wout["distance"].tobytes()


def test_gather_rows_orders_and_pads():
    from gsearch_b200 import sharding

    assert sharding.shard_indices(7, 1, 3) == [1, 4]
    assert sharding.shard_rows(7, 3) == 3
    x = torch.arange(12).reshape(6, 2)
    assert torch.equal(sharding.gather_rows(x, 5, rank=0, world=1), x[:5])
