"""hnswio-style dump (gsearch_b200/hnswio.py; behind GSB_DUMP_HNSWIO=1): the writer and the
reader are inverses of each other on a graph the oracle built, and a device-built index answers the same
after a trip through the files.  (Byte compatibility with stock hnsw_rs is NOT what is tested: there is
no upstream source to test against.)"""
import numpy as np
import pytest

import gsearch_b200 as g
from gsearch_b200 import hnswio


def graded(n, S, dtype, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 2**31, (n, S)).astype(dtype)
    for i in range(1, n):
        redraw = rng.random(S) >= 0.8
        base[i] = np.where(redraw, base[i], base[(i - 1) // 3])
    return base[rng.permutation(n)]


def lists_by_id(im):
    """{id: (level, rank, [[(neighbour id, dist bits) ...] per layer])} -- independent of point numbering"""
    first = np.concatenate([[0], np.cumsum(im["levels"].astype(np.int64) + 1)])
    out = {}
    for p, pid in enumerate(im["ids"]):
        ls = []
        for l in range(int(im["levels"][p]) + 1):
            a, b = int(im["nbr_offsets"][first[p] + l]), int(im["nbr_offsets"][first[p] + l + 1])
            ls.append(list(zip(im["ids"][im["nbr_index"][a:b]].tolist(),
                               im["nbr_dist"][a:b].view(np.uint32).tolist())))
        out[int(pid)] = (int(im["levels"][p]), int(im["ranks"][p]), ls)
    return out


@pytest.mark.parametrize("dtype", [np.uint32, np.uint64, np.uint16, np.float32])
def test_hnswio_round_trip_of_an_oracle_graph(oracle, tmp_path, dtype):
    base = graded(180, 96, np.uint32, 5).astype(dtype)
    h = oracle.Hnsw(12, 40, 96, dtype, scale=0.5)
    h.insert_waves(base, np.arange(180, dtype=np.uint64) * 3 + 7, 16)
    im = h.export()
    hnswio.dump(str(tmp_path), "hnswdump", im, base, 12, 40)
    assert hnswio.is_hnswio(str(tmp_path)) and not hnswio.is_hnswio(str(tmp_path), "other")
    back = hnswio.load(str(tmp_path))
    assert back["max_nb_connection"] == 12 and back["ef"] == 40 and back["dtype"] == np.dtype(dtype)
    assert lists_by_id(back) == lists_by_id(im)
    assert int(back["ids"][back["entry_point"]]) == int(im["ids"][im["entry_point"]])
    order = {int(i): p for p, i in enumerate(im["ids"])}
    for p, pid in enumerate(back["ids"]):
        assert back["sigs"][p].tobytes() == base[order[int(pid)]].tobytes()
    # t_name is what src/utils/reloadhnsw.rs:13-37 sniffs
    raw = open(tmp_path / "hnswdump.hnsw.graph", "rb").read(128)
    assert {np.uint16: b"u16", np.uint32: b"u32", np.uint64: b"u64", np.float32: b"f32"}[dtype] in raw


@pytest.mark.gpu
def test_device_index_through_hnswio_files(tmp_path, monkeypatch):
    base = graded(600, 256, np.uint64, 9)
    ids = np.arange(600, dtype=np.uint64) + 1000
    idx = g.Hnsw(g.HnswParams(max_nb_conn=16, ef=64), 256, np.uint64)
    idx.parallel_insert(base, ids)
    q = base[::7].copy()
    want = idx.search_raw(q, 10, 80)
    monkeypatch.setenv("GSB_DUMP_HNSWIO", "1")
    idx.file_dump(tmp_path, "hnswdump")
    assert hnswio.is_hnswio(str(tmp_path))
    idx2 = g.Hnsw(g.HnswParams(max_nb_conn=16, ef=64), 256, np.uint64)
    idx2.load(tmp_path, "hnswdump")
    got = idx2.search_raw(q, 10, 80)
    assert np.array_equal(got[1], want[1])
    for f in ("d_id", "distance", "layer", "rank"):
        assert np.array_equal(got[0][f], want[0][f])
    idx2.parallel_insert(graded(40, 256, np.uint64, 11), np.arange(40, dtype=np.uint64) + 5000)   # still extensible
    assert idx2.get_nb_point() == 640
    idx.close(); idx2.close()
