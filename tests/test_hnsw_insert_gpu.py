"""GPU parity of K8 (on-device HNSW construction) through the C ABI: gsb_index_insert_batch with
waves of at most W points must build EXACTLY the graph of the oracle's wave insertion
(oracle/hnsw.c gso_hnsw_insert_waves; W = 1 is the sequential insertion of the reference
algorithm): same levels, ranks, neighbour lists, neighbour distances (bit patterns) and entry
point -- and therefore the same search results."""
import numpy as np
import pytest

import gsearch_b200 as g
from test_hnsw_gpu import family_sigs, tree_sigs

pytestmark = pytest.mark.gpu

KEYS = ("levels", "ranks", "ids", "nbr_offsets", "nbr_index")


def assert_same_graph(ga, gb):
    assert ga["entry_point"] == gb["entry_point"]
    for k in KEYS:
        assert np.array_equal(ga[k], gb[k]), k
    assert ga["nbr_dist"].tobytes() == gb["nbr_dist"].tobytes()


def both(oracle, sigs, M, ef_c, wave, scale=1.0, extend=True, chunks=1):
    n, S = sigs.shape
    ids = np.arange(n, dtype=np.uint64) * 3 + 7
    h = oracle.Hnsw(M, ef_c, S, sigs.dtype, scale=scale, extend=extend)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=M, ef=ef_c, scale_modification=scale, extend_candidates=extend),
                 S, sigs.dtype)
    idx.set_wave_max(wave)
    # the index may be fed in several calls (tohnsw inserts once, `add` later)
    cuts = np.linspace(0, n, chunks + 1).astype(int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        h.insert_waves(sigs[a:b], ids[a:b], wave)
        idx.parallel_insert(sigs[a:b], ids[a:b])
    return h, idx


@pytest.mark.parametrize("dt,S", [(np.uint64, 512), (np.uint32, 300), (np.float32, 256)])
@pytest.mark.parametrize("gen", ["family", "tree"])
@pytest.mark.parametrize("M,ef_c,wave,scale", [(8, 32, 1, 1.0), (12, 48, 16, 1.0), (16, 64, 148, 0.5),
                                               (24, 100, 300, 1.0)])
def test_insert_builds_the_oracle_graph(oracle, dt, S, gen, M, ef_c, wave, scale):
    rng = np.random.default_rng(S + M)
    n = 700
    sigs = family_sigs(rng, n, S, dt) if gen == "family" else tree_sigs(rng, n, S, dt)
    h, idx = both(oracle, sigs, M, ef_c, wave, scale)
    assert idx.get_nb_point() == n == h.nb_point()
    assert_same_graph(idx.export_graph(), h.export())
    q = sigs[::41]
    got, gc, ge = idx.search_raw(q, 6, 80)
    want, wc, we = h.search(q, 6, 80, nthreads=4)
    assert gc.tolist() == wc.tolist() and ge.tolist() == we.tolist()
    assert got["d_id"].tolist() == want["d_id"].tolist()
    assert got["distance"].tobytes() == want["distance"].tobytes()


def test_waves_longer_than_the_list_scratch(oracle):
    """a wave may hold more points than a CTA's list scratch (512): the wave mates are met 512 at a time, in
    order; wave = n / 4 while that is below wave_max, so 3 000 points reach waves of 700"""
    rng = np.random.default_rng(21)
    sigs = tree_sigs(rng, 3000, 128, np.uint32)
    h, idx = both(oracle, sigs, 12, 48, 700, 0.5)
    assert_same_graph(idx.export_graph(), h.export())
    with pytest.raises(g.GsbError):
        idx.set_wave_max(5000)


def test_insert_in_several_calls_and_without_extension(oracle):
    rng = np.random.default_rng(11)
    sigs = tree_sigs(rng, 500, 400, np.uint64)
    h, idx = both(oracle, sigs, 10, 40, 64, extend=False, chunks=3)
    assert_same_graph(idx.export_graph(), h.export())


def test_tiny_indexes(oracle):
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 9):
        sigs = family_sigs(rng, n, 64, np.uint32)
        h, idx = both(oracle, sigs, 4, 8, 148)
        assert_same_graph(idx.export_graph(), h.export())


def test_build_recall_at_reference_like_parameters(oracle):
    """a graph built on device answers like brute force (size-independent property)"""
    rng = np.random.default_rng(8)
    S = 2048
    sigs = tree_sigs(rng, 3000, S, np.uint64)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=32, ef=200), S, np.uint64)
    idx.parallel_insert(sigs, np.arange(3000, dtype=np.uint64))
    q = sigs[::60]
    res = idx.parallel_search(q, 8, 400)
    d = g.DistHamming().matrix(q, sigs)
    hits = 0
    for i, r in enumerate(res):
        assert r[0].distance == 0.0
        kth = np.sort(d[i])[7]
        hits += sum(1 for nb in r if nb.distance <= kth)
    assert hits / (8 * len(q)) >= 0.97


def test_dump_and_reload_round_trip(oracle, tmp_path):
    rng = np.random.default_rng(9)
    sigs = tree_sigs(rng, 400, 256, np.uint32)
    h, idx = both(oracle, sigs, 8, 32, 32)
    idx.file_dump(tmp_path, "hnswdump")
    assert (tmp_path / "hnswdump.hnsw.graph").exists() and (tmp_path / "hnswdump.hnsw.data").exists()
    idx2 = g.Hnsw(g.HnswParams(max_nb_conn=8, ef=32), 256, np.uint32)
    idx2.load(tmp_path, "hnswdump")
    assert idx2.get_nb_point() == 400
    assert_same_graph(idx2.export_graph(), idx.export_graph())
    a = idx.search_raw(sigs[:20], 5, 50)
    b = idx2.search_raw(sigs[:20], 5, 50)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tolist() == b[1].tolist()
    # a reloaded index can be extended (the `add` sub-command), and a dump of another shape is refused
    idx2.parallel_insert(sigs[:5], np.arange(5, dtype=np.uint64) + 10_000)
    assert idx2.get_nb_point() == 405
    with pytest.raises(g.GsbError):
        g.Hnsw(g.HnswParams(max_nb_conn=9, ef=32), 256, np.uint32).load(tmp_path, "hnswdump")


def test_rows_larger_than_shared_memory(oracle):
    """S = 65535 is legal (README.md:676); a u64 row is then 524 KB and cannot be staged in shared
    memory: distance, construction and search must still equal the oracle"""
    rng = np.random.default_rng(13)
    S, n = 65535, 80
    sigs = tree_sigs(rng, n, S, np.uint64, keep=0.9)
    d = g.DistHamming().matrix(sigs[:5], sigs)
    assert d.tobytes() == oracle.hamming_matrix(sigs[:5], sigs).tobytes()
    h, idx = both(oracle, sigs, 6, 24, 16)
    assert_same_graph(idx.export_graph(), h.export())
    got, gc, ge = idx.search_raw(sigs[::9], 4, 30)
    want, wc, we = h.search(sigs[::9], 4, 30)
    assert gc.tolist() == wc.tolist() and ge.tolist() == we.tolist()
    assert got["d_id"].tolist() == want["d_id"].tolist()
    assert got["distance"].tobytes() == want["distance"].tobytes()


def test_visit_stamps_fallback(oracle, monkeypatch):
    """indexes too large for the shared-memory visited bitmap use visit stamps in global memory;
    GSB_NO_BITMAP forces that path at test size"""
    monkeypatch.setenv("GSB_NO_BITMAP", "1")
    rng = np.random.default_rng(17)
    sigs = tree_sigs(rng, 500, 320, np.uint64)
    h, idx = both(oracle, sigs, 10, 40, 64)
    assert_same_graph(idx.export_graph(), h.export())
    got, gc, ge = idx.search_raw(sigs[::23], 5, 60)
    want, wc, we = h.search(sigs[::23], 5, 60)
    assert gc.tolist() == wc.tolist() and ge.tolist() == we.tolist()
    assert got["d_id"].tolist() == want["d_id"].tolist()
