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


# ------------------------------------------------------------------------------------------------
# the torch-free product path: gsb_comm_* (NCCL behind the C ABI) + sharded HNSW insertion
def _sharded_worker(rank, world, uid, nfiles, npts, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import gsearch_b200 as g
    from gsearch_b200.comm import Comm, DeviceBuffer, shard_rows

    comm = Comm(uid, world, rank, rank)
    # (a) sketch shard -> all_gather_rows -> every rank holds all signatures in file order, on device
    files = [g.synth.dna_genome(i, 120_000 + 700 * i) for i in range(nfiles)]
    S = 1024
    sk = g.Sketcher(g.SeqSketcherParams(21, S), device=rank)
    per = shard_rows(nfiles, world)
    row = S * sk.elem_size
    d_local = DeviceBuffer(per * row, rank)
    d_local.upload(np.zeros(per * row, dtype=np.uint8))
    mine = files[rank::world]
    buf, offs = sk.concat(mine)
    sk.sketch_buffer_to_device(buf, offs, d_local.ptr)
    d_tmp, d_all = DeviceBuffer(world * per * row, rank), DeviceBuffer(nfiles * row, rank)
    comm.all_gather_rows(d_local.ptr, per, row, nfiles, d_tmp.ptr, d_all.ptr)
    sig = d_all.download(sk.dtype, nfiles * S).reshape(nfiles, S)
    # (b) sharded insertion of a larger synthetic signature set: same graph on every rank
    base = g.synth.signatures(npts, 256, np.uint64, seed=7)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=12, ef=48), 256, np.uint64, device=rank)
    idx.set_wave_max(64)
    idx.insert_sharded(comm, base, np.arange(npts, dtype=np.uint64))
    gr = idx.export_graph()
    out, cnt, _ = idx.search_raw(base[rank::world][:50], 5, 64)
    q.put((rank, sig, gr, out["d_id"].copy(), cnt))
    comm.close()


def test_sharded_insert_builds_the_single_gpu_graph(oracle):
    """gsb_index_insert_batch_sharded on 2 GPUs == gsb_index_insert_batch on 1 GPU == the oracle's wave
    insertion, bit for bit; signatures all-gathered through gsb_comm_all_gather_rows arrive in file order"""
    import multiprocessing as mp
    import gsearch_b200 as g

    if g.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, nfiles, npts = 2, 7, 1500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = g.comm.unique_id()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, uid, nfiles, npts, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = [g.synth.dna_genome(i, 120_000 + 700 * i) for i in range(nfiles)]
    want, _ = g.Sketcher(g.SeqSketcherParams(21, 1024)).sketch_files(files)
    base = g.synth.signatures(npts, 256, np.uint64, seed=7)
    one = g.Hnsw(g.HnswParams(max_nb_conn=12, ef=48), 256, np.uint64)
    one.set_wave_max(64)
    one.parallel_insert(base, np.arange(npts, dtype=np.uint64))
    g1 = one.export_graph()
    h = oracle.Hnsw(12, 48, 256, np.uint64)
    h.insert_waves(base, np.arange(npts, dtype=np.uint64), 64)
    go = h.export()
    for rank, sig, gr, ids, cnt in got:
        assert sig.tobytes() == want.tobytes(), f"rank {rank}: gathered signatures"
        for k in ("levels", "ranks", "ids", "nbr_offsets", "nbr_index", "nbr_dist"):
            assert np.array_equal(gr[k], g1[k]), f"rank {rank}: {k} differs from the single-GPU graph"
            assert np.array_equal(gr[k], go[k]), f"rank {rank}: {k} differs from the oracle graph"
        assert gr["entry_point"] == g1["entry_point"] == go["entry_point"]
        w, wc, _ = one.search_raw(base[rank::world][:50], 5, 64)
        assert ids.tolist() == w["d_id"].tolist() and cnt.tolist() == wc.tolist()


def test_cli_two_gpus_equals_one_gpu(tmp_path, monkeypatch):
    """gsearch --gpus 2 tohnsw / request write the same files as the single-GPU run"""
    import gsearch_b200 as g
    from gsearch_b200 import cli

    if g.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    db, qd, w1, w2 = tmp_path / "db", tmp_path / "q", tmp_path / "w1", tmp_path / "w2"
    for d in (db, qd, w1, w2):
        d.mkdir()
    for i in range(21):
        (db / f"g{i:03d}.fna").write_bytes(g.synth.dna_genome(i, 50_000 + 100 * i))
    for i in (3, 8, 30):
        (qd / f"q{i:03d}.fna").write_bytes(g.synth.dna_genome(i, 50_000 + 100 * i))
    monkeypatch.chdir(w1)
    cli.main("tohnsw -d {} -k 16 -s 512 -n 16 --ef 64 --algo prob".format(db).split())
    cli.main("request -b {} -r {} -n 5".format(w1, qd).split())
    monkeypatch.chdir(w2)
    cli.main("--gpus 2 tohnsw -d {} -k 16 -s 512 -n 16 --ef 64 --algo prob".format(db).split())
    cli.main("--gpus 2 request -b {} -r {} -n 5".format(w2, qd).split())
    for f in ("hnswdump.hnsw.graph", "hnswdump.hnsw.data", "seqdict.json", "parameters.json", "gsearch.neighbors.txt"):
        assert open(w1 / f, "rb").read() == open(w2 / f, "rb").read(), f
