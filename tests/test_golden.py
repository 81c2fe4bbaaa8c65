"""Committed golden fixtures (tests/golden/, made by make_golden.py): the oracle on CPU, and the
CUDA path through the C ABI on the GPU, must both reproduce them bit for bit."""
import os

import numpy as np
import pytest

import gsearch_b200 as g
from golden.make_golden import CASES, hash_bytes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    files = CASES[name][0]()
    assert [hash_bytes(f) for f in files] == z["sha"].tolist(), "synthetic generator changed"
    return z, files


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    z, files = load(name)
    k, S, algo, data_t, block, _ = z["meta"].tolist()
    sig, nb = oracle.sketch_files(files, k, S, algo, data_t, bool(block))
    assert sig.tobytes() == z["sig"].tobytes() and nb.tolist() == z["nb"].tolist()


def test_oracle_hnsw_reproduces_golden(oracle):
    z = np.load(os.path.join(GOLD, "hnsw_u64_s256.npz"))
    h = oracle.Hnsw(16, 64, 256, np.uint64)
    h.insert(z["base"], np.arange(300, dtype=np.uint64) + 7)
    out, counts, neval = h.search(z["q"], 6, 80)
    assert out["d_id"].tolist() == z["d_id"].tolist()
    assert out["distance"].tobytes() == z["distance"].tobytes()
    assert counts.tolist() == z["counts"].tolist() and neval.tolist() == z["neval"].tolist()
    assert oracle.hamming_matrix(z["q"], z["base"]).tobytes() == z["dist_q_base"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden(name):
    z, files = load(name)
    k, S, algo, data_t, block, _ = z["meta"].tolist()
    sk = g.Sketcher(g.SeqSketcherParams(k, S, algo, data_t, bool(block)))
    sig, nb = sk.sketch_files(files)
    assert sig.dtype == z["sig"].dtype
    assert sig.tobytes() == z["sig"].tobytes() and nb.tolist() == z["nb"].tolist()


@pytest.mark.gpu
def test_gpu_hnsw_and_hamming_reproduce_golden():
    z = np.load(os.path.join(GOLD, "hnsw_u64_s256.npz"))
    assert g.DistHamming().matrix(z["q"], z["base"]).tobytes() == z["dist_q_base"].tobytes()
    idx = g.Hnsw(g.HnswParams(max_nb_conn=16, ef=64), 256, np.uint64)
    idx.load_graph(z["base"], z["ids"], z["levels"], z["ranks"], z["nbr_offsets"], z["nbr_index"],
                   int(z["entry"][0]))
    out, counts, neval = idx.search_raw(z["q"], 6, 80)
    assert out["d_id"].tolist() == z["d_id"].tolist()
    assert out["distance"].tobytes() == z["distance"].tobytes()
    assert out["layer"].tolist() == z["layer"].tolist() and out["rank"].tolist() == z["rank"].tolist()
    assert counts.tolist() == z["counts"].tolist() and neval.tolist() == z["neval"].tolist()


def _wave_fixture():
    return np.load(os.path.join(GOLD, "hnsw_wave_u32_s192.npz"))


def _same_graph(gr, z):
    assert gr["entry_point"] == int(z["entry"][0])
    for k in ("levels", "ranks", "ids", "nbr_offsets", "nbr_index"):
        assert np.array_equal(gr[k], z[k]), k
    assert gr["nbr_dist"].tobytes() == z["nbr_dist"].tobytes()


def test_oracle_wave_insertion_reproduces_golden(oracle):
    z = _wave_fixture()
    h = oracle.Hnsw(12, 40, 192, np.uint32, scale=0.5)
    h.insert_waves(z["base"], np.arange(260, dtype=np.uint64) * 2 + 1, 24)
    _same_graph(h.export(), z)


@pytest.mark.gpu
def test_gpu_wave_insertion_reproduces_golden():
    z = _wave_fixture()
    idx = g.Hnsw(g.HnswParams(max_nb_conn=12, ef=40, scale_modification=0.5), 192, np.uint32)
    idx.set_wave_max(24)
    idx.parallel_insert(z["base"], np.arange(260, dtype=np.uint64) * 2 + 1)
    _same_graph(idx.export_graph(), z)
