"""GPU parity of K7 (on-device HNSW search) through the C ABI: on the SAME graph (built by the
oracle, loaded as a graph image) ids, distances, point ids and evaluation counts are identical."""
import numpy as np
import pytest

import gsearch_b200 as g

pytestmark = pytest.mark.gpu


def family_sigs(rng, n, S, dt, fam=8):
    """signatures with planted family structure: family mates share each slot with prob p;
    different families are unrelated (distance ~1, massively tied -- the hard case for
    tie-breaking parity)"""
    out = np.zeros((n, S), dtype=dt)
    for f0 in range(0, n, fam):
        root = rng.integers(1, 2**40, S)
        for j in range(f0, min(n, f0 + fam)):
            p = [1.0, 0.95, 0.9, 0.8, 0.6, 0.4, 0.2, 0.05][j - f0]
            keep = rng.random(S) < p
            row = np.where(keep, root, rng.integers(1, 2**40, S))
            out[j] = (row % 1000 / 1000.0) if dt == np.float32 else row
    return out


def tree_sigs(rng, n, S, dt, keep=0.85):
    """signatures on a random binary phylogeny: each child re-draws a fraction 1-keep of its
    parent's slots, so distances are graded (1 - keep^path) and the HNSW graph is navigable"""
    out = np.zeros((n, S), dtype=np.uint64)
    out[0] = rng.integers(1, 2**40, S)
    for i in range(1, n):
        parent = (i - 1) // 2
        redraw = rng.random(S) >= keep
        out[i] = np.where(redraw, rng.integers(1, 2**40, S), out[parent])
    return (out % 1000 / 1000.0).astype(np.float32) if dt == np.float32 else out.astype(dt)


def build(oracle, sigs, M, ef_c, scale=1.0):
    h = oracle.Hnsw(M, ef_c, sigs.shape[1], sigs.dtype, scale=scale)
    ids = np.arange(len(sigs), dtype=np.uint64) + 1000
    h.insert(sigs, ids)
    return h, ids


def load(h, sigs, M, ef_c):
    gr = h.export()
    idx = g.Hnsw(g.HnswParams(max_nb_conn=M, ef=ef_c), sigs.shape[1], sigs.dtype)
    idx.load_graph(sigs, gr["ids"], gr["levels"], gr["ranks"], gr["nbr_offsets"], gr["nbr_index"],
                   gr["entry_point"], gr["nbr_dist"])
    return idx


@pytest.mark.parametrize("dt,S", [(np.uint64, 2000), (np.uint32, 512), (np.float32, 1200)])
@pytest.mark.parametrize("gen", ["family", "tree"])
@pytest.mark.parametrize("M,ef_c,scale", [(16, 64, 1.0), (48, 200, 0.25)])
def test_search_identical_to_oracle_on_same_graph(oracle, dt, S, M, ef_c, scale, gen):
    rng = np.random.default_rng(S + M)
    sigs = family_sigs(rng, 600, S, dt) if gen == "family" else tree_sigs(rng, 600, S, dt)
    h, ids = build(oracle, sigs, M, ef_c, scale)
    idx = load(h, sigs, M, ef_c)
    assert idx.get_nb_point() == 600
    queries = np.concatenate([sigs[::37], family_sigs(rng, 16, S, dt)])
    for knbn, ef in [(10, 50), (50, 400), (5, 1)]:
        got, gc, ge = idx.search_raw(queries, knbn, ef)
        want, wc, we = h.search(queries, knbn, ef, nthreads=4)
        assert gc.tolist() == wc.tolist()
        assert ge.tolist() == we.tolist()
        for i in range(len(queries)):
            n = gc[i]
            assert got["d_id"][i, :n].tolist() == want["d_id"][i, :n].tolist()
            assert got["distance"][i, :n].tobytes() == want["distance"][i, :n].tobytes()
            assert got["layer"][i, :n].tolist() == want["layer"][i, :n].tolist()
            assert got["rank"][i, :n].tolist() == want["rank"][i, :n].tolist()


@pytest.mark.parametrize("dt,S", [(np.uint64, 2000), (np.float32, 1200), (np.uint16, 4096), (np.uint32, 1028)])
def test_search_ring_path_identical_to_oracle(oracle, monkeypatch, dt, S):
    """K7's TMA-ring variant (candidate rows + query pieces by bulk copy, control warp / worker warps,
    speculative expansion of the predicted next candidate with roll-back): forced for every ef here
    (by default it serves ef_search > 2046), tie-heavy data, and an ef larger than the index"""
    monkeypatch.setenv("GSB_K7_RING", "1")
    rng = np.random.default_rng(S)
    sigs = tree_sigs(rng, 700, S, dt) if dt != np.uint16 else (tree_sigs(rng, 700, S, np.uint32) % 7).astype(np.uint16)
    h, ids = build(oracle, sigs, 24, 100, 0.5)
    idx = load(h, sigs, 24, 100)
    queries = np.concatenate([sigs[::41], sigs[:3][:, ::-1].copy()])
    for knbn, ef in [(10, 50), (50, 400), (5, 1), (20, 3000)]:
        got, gc, ge = idx.search_raw(queries, knbn, ef)
        want, wc, we = h.search(queries, knbn, ef, nthreads=4)
        assert gc.tolist() == wc.tolist() and ge.tolist() == we.tolist()
        for i in range(len(queries)):
            n = gc[i]
            assert got["d_id"][i, :n].tolist() == want["d_id"][i, :n].tolist()
            assert got["distance"][i, :n].tobytes() == want["distance"][i, :n].tobytes()
            assert got["rank"][i, :n].tolist() == want["rank"][i, :n].tolist()


def test_search_pointers_and_device_forms_agree(oracle):
    """the raw host-pointer form (pinned buffers in bench.py) and the device-pointer form return what
    search_raw returns"""
    from gsearch_b200.comm import DeviceBuffer, PinnedBuffer
    rng = np.random.default_rng(3)
    sigs = tree_sigs(rng, 400, 512, np.uint64)
    h, ids = build(oracle, sigs, 12, 48)
    idx = load(h, sigs, 12, 48)
    q = np.ascontiguousarray(sigs[::29])
    nq, knbn = len(q), 7
    want, wc, we = idx.search_raw(q, knbn, 100)
    pq, po, pc, pe = PinnedBuffer(q.nbytes), PinnedBuffer(nq * knbn * 24), PinnedBuffer(nq * 4), PinnedBuffer(nq * 8)
    pq.array[:q.nbytes] = q.view(np.uint8).reshape(-1)
    idx.search_pointers(pq.ptr, nq, knbn, 100, po.ptr, pc.ptr, pe.ptr)
    assert po.array[:nq * knbn * 24].tobytes() == want.tobytes()
    assert pc.array[:nq * 4].view(np.uint32).tolist() == wc.tolist()
    assert pe.array[:nq * 8].view(np.uint64).tolist() == we.tolist()
    dq, do, dc, de = DeviceBuffer(q.nbytes), DeviceBuffer(nq * knbn * 24), DeviceBuffer(nq * 4), DeviceBuffer(nq * 8)
    dq.upload(q)
    idx.search_device(dq.ptr, nq, knbn, 100, do.ptr, dc.ptr, de.ptr)
    assert do.download(np.uint8, nq * knbn * 24).tobytes() == want.tobytes()
    assert dc.download(np.uint32, nq).tolist() == wc.tolist()


def test_search_recall_against_brute_force(oracle):
    rng = np.random.default_rng(4)
    S = 4096
    sigs = tree_sigs(rng, 1500, S, np.uint64)
    h, ids = build(oracle, sigs, 32, 200)
    idx = load(h, sigs, 32, 200)
    q = sigs[::50]
    res = idx.parallel_search(q, 8, 500)
    d = g.DistHamming().matrix(q, sigs)
    hits = tot = 0
    for i, r in enumerate(res):
        assert r[0].distance == 0.0                       # the query itself is in the base
        kth = np.sort(d[i])[7]
        # recall up to ties at the k-th distance
        hits += sum(1 for n in r if n.distance <= kth)
        tot += 8
        ds = [n.distance for n in r]
        assert ds == sorted(ds)
    assert hits / tot >= 0.95


def test_empty_index_and_small_base(oracle):
    idx = g.Hnsw(g.HnswParams(max_nb_conn=8, ef=16), 64, np.uint32)
    out, counts, _ = idx.search_raw(np.zeros((3, 64), dtype=np.uint32), 5, 10)
    assert counts.tolist() == [0, 0, 0]
    rng = np.random.default_rng(0)
    sigs = family_sigs(rng, 3, 64, np.uint32)
    h, ids = build(oracle, sigs, 8, 16)
    idx = load(h, sigs, 8, 16)
    got, gc, _ = idx.search_raw(sigs, 5, 10)
    want, wc, _ = h.search(sigs, 5, 10)
    assert gc.tolist() == wc.tolist() == [3, 3, 3]
    assert got["d_id"][:, :3].tolist() == want["d_id"][:, :3].tolist()
    assert not got["d_id"][:, 3:].any()                  # unused entries are zeroed
