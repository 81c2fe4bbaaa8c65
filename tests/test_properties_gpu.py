"""Size-independent properties at (or near) the BASELINE.json sizes, where the CPU oracle is too
slow to be the checker: batch invariance of the sketcher, distances that follow the planted
family structure, self-distance zero, and HNSW answers that agree with brute force."""
import numpy as np
import pytest

import gsearch_b200 as g
from test_hnsw_gpu import tree_sigs

pytestmark = pytest.mark.gpu


def test_sketch_is_invariant_to_batch_composition_and_order():
    # configs[1] parameters (k=21, s=18000, prob) on 1 Mbp genomes: the result for a genome must
    # not depend on which other genomes share its batch, group, stream or slot
    files = [g.synth.dna_genome(i, 1_000_000, ncontigs=1 + i % 4) for i in range(40)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 18000))
    full, nb = sk.sketch_files(files)
    perm = np.random.default_rng(0).permutation(40)
    shuf, nb2 = sk.sketch_files([files[i] for i in perm])
    assert np.array_equal(shuf, full[perm]) and np.array_equal(nb2, nb[perm])
    for lo, hi in [(0, 1), (1, 4), (4, 13), (13, 40)]:
        part, _ = sk.sketch_files(files[lo:hi])
        assert np.array_equal(part, full[lo:hi]), (lo, hi)
    assert sk.retry_count == 0
    # no empty slot, and every signature element is a canonical 21-mer value (< 4^21)
    assert (full != 0).all() and int(full.max()) < 4**21


def test_distances_follow_the_planted_families():
    # the generator makes families of 16 genomes with substitution rates 0.1 % .. 10 %:
    # 1 - d estimates the weighted Jaccard, so mates are closer than strangers, in rate order
    files = [g.synth.dna_genome(i, 1_000_000) for i in range(32)]
    sig, _ = g.Sketcher(g.SeqSketcherParams(21, 18000)).sketch_files(files)
    d = g.DistHamming().matrix(sig, sig)
    assert (np.diag(d) == 0).all() and np.array_equal(d, d.T)
    for fam in (0, 1):
        blk = d[16 * fam:16 * fam + 16, 16 * fam:16 * fam + 16]
        off = d[16 * fam:16 * fam + 16, 16 * (1 - fam):16 * (1 - fam) + 16]
        assert blk[0, 1:].max() < 0.999 and off.min() > 0.999   # strangers share (almost) no 21-mer
        # members with a lower substitution rate are closer to the root than those with a higher one
        assert blk[0, 1] < blk[0, 5] or blk[0, 7] < blk[0, 5]


def test_optdens_and_super_agree_when_every_bin_is_hit():
    # SuperMinHash restricted to its first level IS the one-permutation bin minimum
    files = [g.synth.dna_genome(i, 600_000) for i in range(4)]
    a, _ = g.Sketcher(g.SeqSketcherParams(21, 4096, g.ALGO_OPTDENS)).sketch_files(files)
    b, _ = g.Sketcher(g.SeqSketcherParams(21, 4096, g.ALGO_SUPER)).sketch_files(files)
    assert a.dtype == b.dtype == np.float32 and np.array_equal(a, b)
    assert (a < 1.0).all() and (a >= 0.0).all()


def test_hnsw_at_reference_parameters_agrees_with_brute_force():
    # configs[2] parameters (s=18000 u64, n=128, ef=1600) on a 1500-point index built on device
    rng = np.random.default_rng(21)
    S, n = 18000, 1500
    sigs = tree_sigs(rng, n, S, np.uint64, keep=0.9)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), S, np.uint64)
    idx.parallel_insert(sigs, np.arange(n, dtype=np.uint64))
    assert idx.get_nb_point() == n
    q = sigs[::50]
    out, cnt, neval = idx.search_raw(q, 50, 1600)
    d = g.DistHamming().matrix(q, sigs)
    assert (cnt == 50).all()
    for i in range(len(q)):
        assert out["d_id"][i, 0] == 50 * i and out["distance"][i, 0] == 0.0
        assert (np.diff(out["distance"][i]) >= 0).all()
        # the distances reported are the true distances of the ids reported
        assert np.array_equal(out["distance"][i], d[i, out["d_id"][i].astype(np.int64)])
        # with ef >= n the search visits everything reachable: exact top-50 up to ties
        assert out["distance"][i, -1] == np.sort(d[i])[49]
    assert (neval <= n + 300).all()


def test_baseline_config0_whole_pipeline_against_the_oracle(oracle):
    """BASELINE configs[0]: tohnsw on 32 synthetic 1 Mbp FASTA, k=16 s=2048 --algo prob (-n 128
    --ef 1600 as in the README): signatures, graph and answers all equal the CPU restatement."""
    files = [g.synth.dna_genome(i, 1_000_000) for i in range(32)]
    sig, nb = g.Sketcher(g.SeqSketcherParams(16, 2048)).sketch_files(files)
    want, wnb = oracle.sketch_files(files, 16, 2048, nthreads=8)
    assert sig.dtype == np.uint32 and sig.tobytes() == want.tobytes() and nb.tolist() == wnb.tolist()
    ids = np.arange(32, dtype=np.uint64)
    idx = g.Hnsw(g.HnswParams(max_nb_conn=128, ef=1600), 2048, np.uint32)
    idx.set_wave_max(296)                                # the library's default on a B200: two points per SM
    idx.parallel_insert(sig, ids)
    h = oracle.Hnsw(128, 1600, 2048, np.uint32)
    h.insert_waves(want, ids, 296)
    ga, gb = idx.export_graph(), h.export()
    assert ga["entry_point"] == gb["entry_point"]
    for k in ("levels", "ranks", "nbr_offsets", "nbr_index"):
        assert np.array_equal(ga[k], gb[k]), k
    got, gc, _ = idx.search_raw(sig, 10, 5000)           # ef_search = 5000, src/bin/gsearch.rs:893
    ref, rc, _ = h.search(want, 10, 5000, nthreads=4)
    assert gc.tolist() == rc.tolist() and got["d_id"].tolist() == ref["d_id"].tolist()
    assert got["distance"].tobytes() == ref["distance"].tobytes()
