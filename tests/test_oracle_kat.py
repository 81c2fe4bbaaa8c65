"""CPU tests: the oracle against published primitive vectors and against the independent
pure-Python restatement of the definitions (tests/_pyref.py)."""
import math

import numpy as np
import pytest

import _oracle as O
import _pyref as R
import ctypes as C


def test_splitmix64_published_vectors():
    # SplitMix64 (Vigna), seed 1234567: the vectors quoted in SURVEY.md 4 / A.3
    st = C.c_uint64(1234567)
    got = [O.L().gso_splitmix64_next(C.byref(st)) for _ in range(3)]
    assert got == [6457827717110365317, 3203168211198807973, 9817491932198370423]


def test_xoshiro256pp_published_vectors():
    # xoshiro256++ reference implementation, state {1,2,3,4}
    x = O.Xoshiro()
    x.s[:] = [1, 2, 3, 4]
    got = [O.L().gso_xoshiro_next_u64(C.byref(x)) for _ in range(5)]
    assert got == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205]


def test_seed_from_u64_and_draws_match_python():
    for seed in [0, 1, 42, 2**63 + 12345, 2**64 - 1]:
        x = O.Xoshiro()
        O.L().gso_xoshiro_seed_from_u64(C.byref(x), C.c_uint64(seed))
        r = R.Xoshiro(seed)
        assert list(x.s) == r.s
        for _ in range(4):
            assert O.L().gso_uniform_f64(C.byref(x)) == r.f64()
            assert O.L().gso_uniform_f32(C.byref(x)) == r.f32()
            assert O.L().gso_uniform_usize(C.byref(x), 18000) == r.usize(18000)
            assert O.L().gso_xoshiro_next_u32(C.byref(x)) == r.next() >> 32


@pytest.mark.parametrize("lam", [1.0, 3.0, math.log(18000 / 17999), math.log(2048 / 2047)])
def test_exp_restricted01_matches_python_and_theory(lam):
    e = O.Exp01()
    O.L().gso_exp01_init(C.byref(e), lam)
    pe = R.Exp01(lam)
    assert (e.c1, e.c2, e.c3) == (pe.c1, pe.c2, pe.c3)
    x = O.Xoshiro()
    O.L().gso_xoshiro_seed_from_u64(C.byref(x), C.c_uint64(7))
    r = R.Xoshiro(7)
    xs = []
    for _ in range(20000):
        a = O.L().gso_exp01_sample(C.byref(e), C.byref(x))
        assert a == pe.sample(r)
        assert 0.0 <= a < 1.0
        xs.append(a)
    # truncated exponential on [0,1): mean = 1/lam - 1/(e^lam - 1)
    mean = 1.0 / lam - 1.0 / math.expm1(lam)
    assert abs(np.mean(xs) - mean) < 4 * np.std(xs) / math.sqrt(len(xs))


FASTA_CASES = [
    b"",
    b">a\nACGT\n",
    b">a\nACGTNNNNacgtRYKM\nGGCC\n>b desc\nTTTT\n",
    b">a\r\nACGT\r\nAC\r\n>b\r\nGG\r\n",                     # CRLF
    b">a\nACGT",                                             # no trailing newline
    b">a\n\n\nAC\n\nGT\n",                                   # blank lines
    b">virus capsid protein\nACGTACGT\n>ok\nGGGG\n",         # capsid record dropped
    b">x\nAC>GT\n>y\n>z\nAAAA\n",                            # '>' inside a line; empty record
    b">only header",
    b">a\nNNNN\n>b\nACG\n",                                  # record that encodes to nothing
    b">capsid\n>capsid2\nACGT\n>c\ncapsidACGT\n",            # "capsid" inside a sequence line is data
]


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("data", FASTA_CASES)
def test_fasta_parse_dna_matches_python(data, block):
    got = [list(map(int, s)) for s in O.parse_fasta(data, 0, block)]
    assert got == R.parse_fasta(data, 0, block)


AA_CASES = [
    b">p1\nMKVLAA*\n>p2 capsid\nMMMM\n>p3\nACDXBZUacd*EFG\n",
    b">p\nMK\nVL\n",
    b"",
]


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("data", AA_CASES)
def test_fasta_parse_aa_matches_python(data, block):
    got = [list(map(int, s)) for s in O.parse_fasta(data, 1, block)]
    assert got == R.parse_fasta(data, 1, block)


def test_not_fasta_is_an_error():
    with pytest.raises(RuntimeError):
        O.parse_fasta(b"ACGT\n", 0, False)


def _rand_fasta(rng, nrec, L, alphabet="ACGT", noise="N"):
    out = []
    for r in range(nrec):
        seq = "".join(rng.choice(list(alphabet), L))
        if noise and L > 20:
            i = int(rng.integers(0, L - 5))
            seq = seq[:i] + noise * 3 + seq[i + 3:]
        out.append(f">rec{r}\n" + "\n".join(seq[i:i + 60] for i in range(0, L, 60)) + "\n")
    return "".join(out).encode()


@pytest.mark.parametrize("k", [1, 3, 11, 14, 15, 16, 21, 31])
def test_kmer_values_dna_match_python(k):
    rng = np.random.default_rng(k)
    data = _rand_fasta(rng, 3, 200)
    for block in (False, True):
        got = [int(v) for v in O.kmer_values(data, 0, k, block)]
        assert got == R.kmers(R.parse_fasta(data, 0, block), 0, k)


def test_canonical_kmer_is_strand_symmetric():
    comp = str.maketrans("ACGT", "TGCA")
    rng = np.random.default_rng(3)
    seq = "".join(rng.choice(list("ACGT"), 300))
    rc = seq.translate(comp)[::-1]
    a = O.kmer_values(f">a\n{seq}\n".encode(), 0, 21)
    b = O.kmer_values(f">a\n{rc}\n".encode(), 0, 21)
    assert sorted(a.tolist()) == sorted(b.tolist())


@pytest.mark.parametrize("k", [1, 6, 7, 12])
def test_kmer_values_aa_match_python(k):
    rng = np.random.default_rng(100 + k)
    data = _rand_fasta(rng, 4, 90, R.AA, "*")
    got = [int(v) for v in O.kmer_values(data, 1, k)]
    assert got == R.kmers(R.parse_fasta(data, 1), 1, k)


@pytest.mark.parametrize("k,m,val_bytes", [(16, 64, 4), (21, 64, 8), (11, 256, 4), (31, 32, 8)])
@pytest.mark.parametrize("identity", [False, True])
def test_probminhash3a_two_pass_equals_definition(k, m, val_bytes, identity):
    rng = np.random.default_rng(k * 1000 + m)
    # repeats on purpose: half of the genome is a copy of the other half, plus a tandem repeat
    half = "".join(rng.choice(list("ACGT"), 700))
    data = (">g\n" + half + half[:500] + "ACGTTGCA" * 40 + "\n").encode()
    vals = [int(v) for v in O.kmer_values(data, 0, k)]
    want, _ = R.probminhash3a_definition(vals, m, val_bytes, identity)
    sig, _ = O.sketch_files([data], k, m, spec_flags=1 if identity else 0)
    assert [int(v) for v in sig[0]] == want


def test_probminhash3a_tiny_input_fills_every_slot():
    # fewer k-mers than slots: the reference keeps generating points until all slots are filled
    data = b">g\nACGTACGGTCA\n"
    vals = [int(v) for v in O.kmer_values(data, 0, 5)]
    want, best = R.probminhash3a_definition(vals, 32, 4, hcut=80.0)
    sig, _ = O.sketch_files([data], 5, 32)
    assert [int(v) for v in sig[0]] == want
    assert all(b[0] < float("inf") for b in best)


def test_probminhash3a_empty_input_is_all_zero():
    sig, nb = O.sketch_files([b">g\nACG\n", b""], 16, 64)
    assert not sig.any() and nb.tolist() == [3, 0]


def test_probminhash_estimates_weighted_jaccard():
    # statistical KAT: E[1 - d_hamming] = Jp; here equal multiplicities so Jp = J
    rng = np.random.default_rng(11)
    a = "".join(rng.choice(list("ACGT"), 6000))
    b = a[:3000] + "".join(rng.choice(list("ACGT"), 3000))
    k, m = 16, 2048
    sig, _ = O.sketch_files([f">a\n{a}\n".encode(), f">b\n{b}\n".encode()], k, m)
    ka = set(O.kmer_values(f">a\n{a}\n".encode(), 0, k).tolist())
    kb = set(O.kmer_values(f">b\n{b}\n".encode(), 0, k).tolist())
    J = len(ka & kb) / len(ka | kb)
    est = 1.0 - O.hamming(sig[0], sig[1])
    assert abs(est - J) < 4 * math.sqrt(J * (1 - J) / m)
    assert O.hamming(sig[0], sig[0]) == 0.0


@pytest.mark.parametrize("f64_draw", [False, True])
@pytest.mark.parametrize("n,m", [(3000, 64), (40, 64), (1, 16)])
def test_optdens_equals_definition_including_densification(n, m, f64_draw):
    rng = np.random.default_rng(n + m)
    vals = [int(v) for v in rng.integers(0, 2**35, n)]
    got = O.optdens(vals, m, 2 if f64_draw else 0)
    want = R.optdens_definition(vals, m, f64_draw)
    assert got.tobytes() == want.tobytes()
    assert (got <= 1.0).all()


def test_optdens_estimates_jaccard_aa():
    rng = np.random.default_rng(5)
    a = "".join(rng.choice(list(R.AA), 8000))
    b = a[:4000] + "".join(rng.choice(list(R.AA), 4000))
    k, m = 7, 1024
    fa, fb = f">a\n{a}\n".encode(), f">b\n{b}\n".encode()
    sig, _ = O.sketch_files([fa, fb], k, m, algo=2, data_t=1)
    ka, kb = set(O.kmer_values(fa, 1, k).tolist()), set(O.kmer_values(fb, 1, k).tolist())
    J = len(ka & kb) / len(ka | kb)
    est = 1.0 - O.hamming(sig[0], sig[1])
    assert abs(est - J) < 4 * math.sqrt(J * (1 - J) / m)


def test_superminhash_is_order_free_and_estimates_jaccard():
    rng = np.random.default_rng(9)
    vals = [int(v) for v in rng.integers(0, 2**40, 5000)]
    m = 512
    a = O.superminhash(vals, m)
    b = O.superminhash(vals[::-1], m)
    assert a.tobytes() == b.tobytes()
    other = vals[:2500] + [int(v) for v in rng.integers(0, 2**40, 2500)]
    c = O.superminhash(other, m)
    J = len(set(vals) & set(other)) / len(set(vals) | set(other))
    assert abs((1.0 - O.hamming(a, c)) - J) < 4 * math.sqrt(J * (1 - J) / m)


@pytest.mark.parametrize("dt", [np.uint16, np.uint32, np.uint64, np.float32])
def test_hamming_matches_definition(dt):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 5, 1001).astype(dt)
    b = rng.integers(0, 5, 1001).astype(dt)
    assert np.float32(O.hamming(a, b)) == R.hamming(a, b)
    assert O.hamming(a, a) == 0.0


def test_sig_type_table():
    # src/dna/dnasketch.rs:500-515, src/aa/aasketch.rs:457-466
    assert [O.sig_type(k, 64, 0, 0) for k in (8, 14, 15, 16, 17, 21, 31)] == [0, 0, 1, 0, 1, 1, 1]
    assert [O.sig_type(k, 64, 0, 1) for k in (3, 6, 7, 12)] == [0, 0, 1, 1]
    assert O.sig_type(21, 64, 2, 0) == 2 and O.sig_type(7, 64, 2, 1) == 2


def _family_sigs(n, S, seed=0, fam=8):
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 2**40, (n, S)).astype(np.uint64)
    for f in range(0, n, fam):
        for j in range(1, min(fam, n - f)):
            keep = rng.random(S) < (0.9 - 0.8 * j / fam)
            base[f + j] = np.where(keep, base[f], base[f + j])
    return base


def _tree_sigs(n, S, seed=0):
    """signatures with graded distances: every point copies a random share of the slots of a
    random earlier point (a random recursive tree), then the order is shuffled"""
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 2**40, (n, S)).astype(np.uint64)
    for i in range(1, n):
        par = int(rng.integers(0, i))
        keep = rng.random(S) < rng.uniform(0.5, 0.95)
        base[i] = np.where(keep, base[par], base[i])
    return base[rng.permutation(n)]


def test_hnsw_wave_insert_of_one_is_sequential_insert(oracle):
    """gso_hnsw_insert_waves(wave_max=1) must rebuild exactly the graph of the sequential insert."""
    base = _family_sigs(400, 128)
    ids = np.arange(400, dtype=np.uint64) + 1000
    a = oracle.Hnsw(8, 32, 128, np.uint64)
    a.insert(base, ids)
    b = oracle.Hnsw(8, 32, 128, np.uint64)
    b.insert_waves(base, ids, 1)
    ga, gb = a.export(), b.export()
    assert ga["entry_point"] == gb["entry_point"]
    for k in ("levels", "ranks", "ids", "nbr_offsets", "nbr_index", "nbr_dist"):
        assert np.array_equal(ga[k], gb[k]), k


def test_hnsw_wave_insert_threads_build_the_same_graph(oracle):
    """gso_hnsw_insert_waves_mt (phase A of a wave spread over host threads: how bench.py's CPU
    reference arm builds its index) must produce the single-threaded graph bit for bit, and the
    same evaluation count."""
    base = _tree_sigs(600, 96, seed=11)
    ids = np.arange(600, dtype=np.uint64)
    a = oracle.Hnsw(8, 40, 96, np.uint64)
    a.insert_waves(base[:250], ids[:250], 37)    # (a call boundary ends a wave: same calls on both sides)
    a.insert_waves(base[250:], ids[250:], 37)
    b = oracle.Hnsw(8, 40, 96, np.uint64)
    b.insert_waves(base[:250], ids[:250], 37, nthreads=4)
    b.insert_waves(base[250:], ids[250:], 37, nthreads=3)   # second call: grown buffers, warm stamps
    ga, gb = a.export(), b.export()
    assert ga["entry_point"] == gb["entry_point"] and a.nb_eval() == b.nb_eval()
    for k in ("levels", "ranks", "ids", "nbr_offsets", "nbr_index", "nbr_dist"):
        assert np.array_equal(ga[k], gb[k]), k


def test_hnsw_wave_insert_keeps_recall(oracle):
    """Points of one wave do not see each other's lists, only each other's data; recall against
    brute force must stay at the level of the sequential build."""
    base = _tree_sigs(800, 128, seed=3)
    ids = np.arange(800, dtype=np.uint64)
    qi = np.arange(0, 800, 12)[:64]
    D = oracle.hamming_matrix(base[qi], base)
    k = 4
    recalls = []
    for wave in (1, 148):
        h = oracle.Hnsw(12, 48, 128, np.uint64)
        h.insert_waves(base, ids, wave)
        out, cnt, _ = h.search(base[qi], k, 96)
        hit = 0
        for i in range(64):
            kth = np.sort(D[i])[k - 1]
            hit += sum(1 for j in range(cnt[i]) if out["distance"][i][j] <= kth)
        recalls.append(hit / (64 * k))
    assert recalls[1] >= recalls[0] - 0.02 and recalls[1] > 0.97, recalls


@pytest.mark.parametrize("wave", [1, 7])
def test_hnsw_oracle_equals_independent_python_restatement(oracle, wave):
    """tests/_pyref.py restates HNSW (Rust BinaryHeap movements, search_layer, Malkov selection
    with extension, wave insertion) from the published definitions; the C oracle must build the
    same graph and return the same neighbours, ties included."""
    import _pyref as P
    rng = np.random.default_rng(5 + wave)
    n, S, M, ef_c = 90, 24, 4, 12
    base = rng.integers(1, 6, (n, S)).astype(np.uint32)          # tiny alphabet: distances tie massively
    for i in range(1, n):
        keep = rng.random(S) < 0.7
        base[i] = np.where(keep, base[int(rng.integers(0, i))], base[i])
    ids = np.arange(n, dtype=np.uint64) + 100
    h = oracle.Hnsw(M, ef_c, S, np.uint32, scale=1.0)
    h.insert_waves(base, ids, wave)
    p = P.PyHnsw(M, ef_c)
    p.insert_waves(list(base), list(ids), wave)
    gr = h.export()
    assert gr["entry_point"] == p.entry and gr["levels"].tolist() == p.level and gr["ranks"].tolist() == p.rank
    li = 0
    for pt in range(n):
        for l in range(p.level[pt] + 1):
            a, b = int(gr["nbr_offsets"][li]), int(gr["nbr_offsets"][li + 1])
            assert gr["nbr_index"][a:b].tolist() == [x for x, _ in p.nbrs[pt][l]], (pt, l)
            li += 1
    out, cnt, _ = h.search(base[:20], 5, 9)
    for i in range(20):
        want = p.search(base[i], 5, 9)
        assert out["d_id"][i, :cnt[i]].tolist() == [w[0] for w in want], i
        assert [float(x) for x in out["distance"][i, :cnt[i]]] == [w[1] for w in want]


def test_superminhash_early_stop_equals_definition(oracle):
    """the oracle's a_upper early stop and lazy permutation arrays only skip work: same bits as the
    full-permutation definition in tests/_pyref.py"""
    import _pyref as P
    rng = np.random.default_rng(2)
    for n, m in [(1, 8), (5, 16), (60, 16), (400, 32)]:
        vals = rng.integers(1, 2**40, n).astype(np.uint64)
        vals = np.concatenate([vals, vals[: n // 2]])            # repeated items are idempotent
        got = oracle.superminhash(vals, m)
        want = P.superminhash_definition([int(v) for v in vals], m)
        assert got.tobytes() == want.tobytes(), (n, m)


def test_fasta_parser_fuzz_against_python():
    """random byte soup biased towards the parser's special bytes ('>', newlines, the letters of
    "capsid", CR, lower case, non-alphabet): the C oracle and the line-based Python restatement
    must produce the same records in every mode"""
    from hypothesis import given, settings, strategies as st

    soup = st.lists(st.sampled_from([b">", b"\n", b"\r\n", b"capsid", b"c", b"a", b"A", b"C", b"G", b"T", b"N",
                                     b"acgt", b"MKV", b"*", b"X", b" ", b"ACGTACGTAC", b"p", b"s", b"i", b"d"]),
                    min_size=0, max_size=60)

    @settings(max_examples=300, deadline=None)
    @given(soup)
    def check(parts):
        data = b">" + b"".join(parts)          # a FASTA file starts with '>' (anything else is an error)
        for data_t in (0, 1):
            for block in (False, True):
                got = [list(map(int, s)) for s in O.parse_fasta(data, data_t, block)]
                assert got == R.parse_fasta(data, data_t, block), (data, data_t, block)

    check()


def test_fastq_parser_fuzz_against_python():
    """FASTQ (needletail: four-line records, first byte '@'): well-formed files with hostile line
    CONTENT (CR, lower case, N, '@' / '+' / '>' inside sequence and quality lines, "capsid" ids, empty
    sequences, missing final terminator) -- the C oracle and the Python restatement must agree"""
    from hypothesis import given, settings, strategies as st

    word = st.lists(st.sampled_from([b"A", b"C", b"G", b"T", b"N", b"acgt", b"@", b"+", b">", b" ", b"MKV", b"*",
                                     b"ACGTACGTAC", b"capsid", b"c"]), min_size=0, max_size=12).map(b"".join)
    rec = st.tuples(word, word, word, word, st.sampled_from([b"\n", b"\r\n"]))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(rec, min_size=1, max_size=8), st.sampled_from([b"", b"\n", b"\n\n", b"\r\n"]), st.booleans())
    def check(recs, tail, cut_last):
        parts = []
        for hid, seq, plus, qual, eol in recs:
            parts.append(b"@r" + hid.replace(b"\n", b"") + eol + seq + eol + b"+" + plus + eol + qual + eol)
        data = b"".join(parts)
        if cut_last:
            data = data.rstrip(b"\r\n")      # the last line may lack its terminator
        data += tail
        for data_t in (0, 1):
            for block in (False, True):
                got = [list(map(int, s)) for s in O.parse_fasta(data, data_t, block)]
                assert got == R.parse_fasta(data, data_t, block), (data, data_t, block)

    check()
    for bad in (b"@r\nACGT\nX\nIIII\n", b"@r\nACGT\n+\nIIII\nACGT\n", b"@r\nACGT\n"):
        with pytest.raises(RuntimeError):
            O.parse_fasta(bad, 0, False)


def test_expm1_spec_is_the_same_function_in_c_and_python_and_close_to_libm(oracle):
    """the sampler's rejection branch evaluates a FROZEN expm1 (degree-24 Horner polynomial, one
    rounded operation per step) so that the C oracle, the Python restatement and the CUDA kernel
    agree bit for bit; it must stay within 2 ulp of libm on [0, ln 2]"""
    import ctypes
    import math
    import _pyref as P
    L = oracle.L()
    L.gso_expm1_spec.restype = ctypes.c_double
    L.gso_expm1_spec.argtypes = [ctypes.c_double]
    rng = np.random.default_rng(8)
    for z in list(rng.random(20000) * math.log(2.0)) + [0.0, 1e-300, 5.6e-5, math.log(2.0)]:
        a = L.gso_expm1_spec(float(z))
        assert a == P.expm1_spec(float(z))
        assert abs(a - math.expm1(z)) <= 2 * math.ulp(math.expm1(z)) if z > 0 else a == 0.0


def test_superminhash2_and_revoptdens_equal_their_definitions(oracle):
    """SuperMinHash2 (per slot the fx hash of the winning item) and RevOptDens (reverse densification)
    as the oracle implements them -- early stop, lazy permutation, in-place rounds -- against the plain
    Python definitions; repeated items are idempotent; RevOptDens without empty bins IS OptDens"""
    import _pyref as P
    rng = np.random.default_rng(21)
    for n, m, kt32 in [(5, 8, True), (40, 16, False), (300, 32, True), (3, 64, False), (2000, 24, False)]:
        vals = rng.integers(1, 2**31 if kt32 else 2**40, n).astype(np.uint64)
        vals = np.concatenate([vals, vals[: n // 2]])
        got = oracle.superminhash2(vals, m, kt32)
        assert got.tolist() == P.superminhash2_definition([int(v) for v in vals], m, kt32), (n, m, kt32)
        got = oracle.revoptdens(vals, m)
        want = P.revoptdens_definition([int(v) for v in vals], m)
        assert got.tobytes() == want.tobytes(), (n, m)
        if n >= 300:  # every bin filled: nothing to densify
            assert got.tobytes() == oracle.optdens(vals, m).tobytes()
        assert (got <= 1.0).all()
    # Jaccard by equality: two sets sharing half of their items
    a = rng.integers(1, 2**40, 4000).astype(np.uint64)
    b = np.concatenate([a[:2000], rng.integers(1, 2**40, 2000).astype(np.uint64)])
    sa, sb = oracle.superminhash2(a, 2048, False), oracle.superminhash2(b, 2048, False)
    j = 2000 / 6000
    assert abs((sa == sb).mean() - j) < 4 * np.sqrt(j * (1 - j) / 2048)


def test_ln_spec_and_setsketch_equal_their_definitions(oracle):
    """--algo hll: the frozen logarithm is the same function in C and Python and within 1 ulp of libm;
    the oracle's SetSketch (early stop on the running register minimum) equals the plain definition"""
    import math, random
    rnd = random.Random(7)
    for _ in range(3000):
        x = math.exp(rnd.uniform(-45, 2))
        a = oracle.ln_spec(x)
        assert a == R.ln_spec(x)
        assert abs(a - math.log(x)) <= 1.01 * 2.0 ** -52 * max(abs(math.log(x)), 1e-300) + 1e-320
    for m, n in ((8, 40), (33, 70), (64, 500)):
        vals = [rnd.getrandbits(42) for _ in range(n)]
        vals += vals[:10]                                  # repeated items change nothing
        assert oracle.setsketch(vals, m).tolist() == R.setsketch_definition(vals, m)
    # similar sets share most registers, disjoint ones almost none (what DistHamming sees)
    base = [rnd.getrandbits(42) for _ in range(30000)]
    near = base[:27000] + [rnd.getrandbits(42) for _ in range(3000)]
    far = [rnd.getrandbits(42) for _ in range(30000)]
    s0, s1, s2 = (oracle.setsketch(v, 1024) for v in (base, near, far))
    assert (s0 != s1).mean() < 0.35 < 0.9 < (s0 != s2).mean()
