"""GPU parity: the CUDA sketch path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Integer signatures and f32 bit patterns must be identical."""
import numpy as np
import pytest

import gsearch_b200 as g
from _cases import adversarial_aa_files, adversarial_dna_files, fasta, rand_seq

pytestmark = pytest.mark.gpu


def run_both(oracle, files, k, S, algo=g.ALGO_PROB3A, data_t=g.DATA_DNA, block=False, spec=0):
    """CUDA vs oracle.  ProbMinHash has two exact multiplicity-counting paths on the device (hash
    partition + shared-memory counting, and the L2 filter that is its fallback): both are run and
    must give the same bytes."""
    sk = g.Sketcher(g.SeqSketcherParams(k, S, algo, data_t, block, spec))
    got, nb = sk.sketch_files(files)
    if algo == g.ALGO_PROB3A:
        sk.set_prob_path(1)
        got1, nb1 = sk.sketch_files(files)
        assert got1.tobytes() == got.tobytes() and nb1.tolist() == nb.tolist(), "the two prob paths differ"
        sk.set_prob_path(0)
        got2, _ = sk.sketch_files(files)   # and back, on slots the filter path has used
        assert got2.tobytes() == got.tobytes()
    want, wnb = oracle.sketch_files(files, k, S, algo, data_t, block, spec, nthreads=8)
    sk.close()
    return got, nb, want, wnb


def assert_same(got, nb, want, wnb):
    assert got.dtype == want.dtype
    assert nb.tolist() == wnb.tolist()
    bad = [i for i in range(len(got)) if got[i].tobytes() != want[i].tobytes()]
    assert not bad, f"signature mismatch for files {bad[:10]}"


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("k,S", [(16, 2048), (21, 1800), (14, 512), (15, 300), (31, 1000), (5, 64)])
def test_prob_dna_adversarial_inputs(oracle, k, S, block):
    assert_same(*run_both(oracle, adversarial_dna_files(k), k, S, block=block))


@pytest.mark.parametrize("spec", [g.SPEC_NOHASH_IDENTITY])
def test_prob_dna_spec_switch(oracle, spec):
    files = adversarial_dna_files(3)[-4:]
    assert_same(*run_both(oracle, files, 21, 2000, spec=spec))


def test_prob_dna_baseline_config0_shape(oracle):
    # BASELINE configs[0]: 1 Mbp genomes, k=16, s=2048, --algo prob (8 of the 32 files here)
    files = [g.synth.dna_genome(i, 1_000_000) for i in range(8)]
    assert_same(*run_both(oracle, files, 16, 2048))


def test_prob_dna_baseline_config1_shape(oracle):
    # BASELINE configs[1] per-genome shape: 5 Mbp, k=21, s=18000 (2 genomes)
    files = [g.synth.dna_genome(i, 5_000_000) for i in (0, 1)]
    assert_same(*run_both(oracle, files, 21, 18000))


def test_prob_partition_path_is_the_one_that_runs(oracle):
    """ordinary genomes stay on the partition path (no fallback, no retry); a file that cannot fit
    its geometry (one k-mer filling a whole tile) is handed to the filter path and still exact"""
    files = [g.synth.dna_genome(i, 1_000_000, ncontigs=1 + i) for i in range(3)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 4096))
    got, nb = sk.sketch_files(files)
    assert sk.fallback_count == 0 and sk.retry_count == 0
    want, wnb = oracle.sketch_files(files, 21, 4096, nthreads=8)
    assert_same(got, nb, want, wnb)
    rep = [fasta([("poly", "A" * 60000)]), files[0]]
    got, nb = sk.sketch_files(rep)
    assert sk.fallback_count == 1
    want, wnb = oracle.sketch_files(rep, 21, 4096, nthreads=8)
    assert_same(got, nb, want, wnb)
    sk.close()


def test_prob_partition_large_k_uses_64_bit_keys(oracle):
    # k = 31 (62-bit k-mers): keys do not fit 31 bits -> the 64-bit key instantiation
    files = [g.synth.dna_genome(40 + i, 300_000) for i in range(2)]
    sk = g.Sketcher(g.SeqSketcherParams(31, 2000))
    got, nb = sk.sketch_files(files)
    assert sk.fallback_count == 0
    assert_same(got, nb, *oracle.sketch_files(files, 31, 2000, nthreads=8))
    sk.close()


@pytest.mark.parametrize("S", [2, 3, 7])
def test_prob_tiny_sketch_sizes_take_the_rejection_branch(oracle, S):
    """lambda = ln(m / (m - 1)) is large for a tiny m, so c1 = expm1(lambda) / lambda is well above 1
    and about a third of the first draws enter the rejection loop of ExpRestricted01 -- the only
    place expm1 is evaluated (by the frozen gso_expm1_spec on both sides, not by two libms)"""
    files = [g.synth.dna_genome(90 + i, 60_000) for i in range(3)]
    assert_same(*run_both(oracle, files, 21, S))
    assert_same(*run_both(oracle, files, 16, S))


def test_prob_heavy_repeats(oracle):
    rng = np.random.default_rng(1)
    unit = rand_seq(rng, 300)
    files = [fasta([("tandem", unit * 200)]), fasta([("poly", "A" * 50000)]),
             fasta([("mix", rand_seq(rng, 30000) + unit * 100)])]
    assert_same(*run_both(oracle, files, 21, 4096))
    assert_same(*run_both(oracle, files, 16, 1024))


@pytest.mark.parametrize("block", [False, True])
@pytest.mark.parametrize("k,S", [(7, 1200), (6, 256), (12, 500)])
def test_optdens_aa(oracle, k, S, block):
    assert_same(*run_both(oracle, adversarial_aa_files(k), k, S, g.ALGO_OPTDENS, g.DATA_AA, block))


def test_optdens_aa_f64_draw_switch(oracle):
    files = adversarial_aa_files(1)[-3:]
    assert_same(*run_both(oracle, files, 7, 1000, g.ALGO_OPTDENS, g.DATA_AA, spec=g.SPEC_OPTDENS_F64_DRAW))


@pytest.mark.parametrize("k,S", [(21, 1800), (16, 512)])
def test_optdens_dna(oracle, k, S):
    assert_same(*run_both(oracle, adversarial_dna_files(k), k, S, g.ALGO_OPTDENS))


@pytest.mark.parametrize("k,S", [(7, 300), (5, 128)])
def test_prob_aa(oracle, k, S):
    assert_same(*run_both(oracle, adversarial_aa_files(k), k, S, g.ALGO_PROB3A, g.DATA_AA))


def test_optdens_baseline_config3_shape(oracle):
    # BASELINE configs[3] per-proteome shape: ~1.5 M residues, k=7, s=12000
    files = [g.synth.aa_proteome(i, 4500, 333) for i in (0, 1)]
    assert_same(*run_both(oracle, files, 7, 12000, g.ALGO_OPTDENS, g.DATA_AA))


def test_not_fasta_is_reported(oracle):
    sk = g.Sketcher(g.SeqSketcherParams(16, 64))
    with pytest.raises(g.GsbError) as e:
        sk.sketch_files([b">ok\nACGT\n", b"ACGT\n"])
    assert e.value.status == 5
    # the handle stays usable
    got, _ = sk.sketch_files([b">ok\nACGTACGTACGTACGTACGTACGT\n"])
    want, _ = oracle.sketch_files([b">ok\nACGTACGTACGTACGTACGTACGT\n"], 16, 64)
    assert got.tobytes() == want.tobytes()


def test_sketch_is_batch_invariant_and_idempotent():
    files = [g.synth.dna_genome(i, 100_000) for i in range(9)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 2000))
    a, _ = sk.sketch_files(files)
    b, _ = sk.sketch_files(files[::-1])
    c = np.stack([sk.sketch_files([f])[0][0] for f in files])
    assert a.tobytes() == b[::-1].tobytes() == c.tobytes()


def test_estimates_jaccard_between_family_members(oracle):
    # size-independent property at a larger size: 1 - d_hamming tracks the exact Jp
    files = [g.synth.dna_genome(i, 400_000) for i in (32, 33, 34, 48)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 8192))
    sig, _ = sk.sketch_files(files)
    d = g.DistHamming().matrix(sig, sig)
    assert (np.diag(d) == 0).all()
    assert d[0, 1] < d[0, 2] < 0.9 < d[0, 3]     # substitution rates 0.5 % < 1 % ; stranger ~ 1


@pytest.mark.parametrize("k,S", [(21, 1800), (16, 256)])
def test_super_dna(oracle, k, S):
    # adversarial files are small: most of them take the sequential cold path of SuperMinHash
    assert_same(*run_both(oracle, adversarial_dna_files(k), k, S, g.ALGO_SUPER))


def test_super_fast_path_and_aa(oracle):
    # enough k-mers to reach every slot at the first level: the one-pass bin-min path
    files = [g.synth.dna_genome(i, 400_000) for i in range(3)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 2048, g.ALGO_SUPER))
    got, nb = sk.sketch_files(files)
    assert sk.retry_count == 0
    want, wnb = oracle.sketch_files(files, 21, 2048, g.ALGO_SUPER, nthreads=3)
    assert_same(got, nb, want, wnb)
    assert_same(*run_both(oracle, adversarial_aa_files(2), 7, 500, g.ALGO_SUPER, g.DATA_AA))


def test_prob_genome_longer_than_2_pow_24(oracle):
    # exact-set entries hold fingerprint | position: the position field widens with the file
    # (24 bits up to 16.7 M symbols, 30 bits at most); 20 Mbp needs 25 bits
    files = [g.synth.dna_genome(3, 20_000_000, ncontigs=2)]
    assert_same(*run_both(oracle, files, 21, 4000))


def test_byte_soup_files_match_oracle(oracle):
    """400 random files made of the parser's special bytes ('>', newlines, CR, the letters of
    "capsid", lower case, non-alphabet bytes, short and long runs of bases): K1's SIMD fast path
    and its byte-serial path must agree with the oracle in every mode"""
    rng = np.random.default_rng(2024)
    atoms = [b">", b"\n", b"\r\n", b"capsid", b"c", b"a", b"A", b"C", b"G", b"T", b"N", b"acgt", b"MKV", b"*",
             b"X", b" ", b"ACGTACGTAC", b"p", b"s", b"i", b"d", b"ACGTTGCATGCATGCAAGGCTTAACCGGTT" * 3,
             b"MKVLAAGIVGLTERDQ" * 2]
    files = []
    for _ in range(400):
        n = int(rng.integers(0, 120))
        files.append(b">" + b"".join(atoms[int(i)] for i in rng.integers(0, len(atoms), n)))
    for k, S, algo, data_t, block in [(5, 64, g.ALGO_PROB3A, g.DATA_DNA, False), (4, 32, g.ALGO_PROB3A, g.DATA_DNA, True),
                                      (3, 48, g.ALGO_OPTDENS, g.DATA_AA, False), (2, 16, g.ALGO_SUPER, g.DATA_AA, True)]:
        assert_same(*run_both(oracle, files, k, S, algo, data_t, block))


def _fastq(records, eol=b"\n", tail=b""):
    out = b""
    for rid, seq in records:
        out += b"@" + rid + eol + seq + eol + b"+" + eol + b"I" * len(seq) + eol
    return out + tail


def test_fastq_inputs_match_oracle(oracle):
    """needletail::parse_fastx_file also reads FASTQ (src/dna/dnafiles.rs:52,128,230): first byte '@',
    four-line records.  Mixed batches (FASTA and FASTQ files side by side), CRLF, lower case, N, '>' '@'
    '+' inside quality lines, records much longer than a 4 KiB tile, empty sequences, trailing blank
    lines, a last line without terminator -- seq and block mode, DNA and AA"""
    rng = np.random.default_rng(77)
    long_reads = [(b"read%d len" % i, rand_seq(rng, int(rng.integers(50, 30000))).encode()) for i in range(12)]
    files = [
        _fastq(long_reads),
        _fastq([(b"r1", b"ACGTNNNNacgtACGTTTGACCA" * 40), (b"r2 x", b""), (b"r3", b"GATTACA" * 100)], eol=b"\r\n"),
        _fastq([(b"q", rand_seq(rng, 5000).encode())])[:-1],                       # no final terminator
        _fastq([(b"a", b"ACGT" * 300), (b"b", b"TTGACA" * 200)], tail=b"\n\n"),
        b"@weird\n" + b"ACGTAC" * 50 + b"\n+weird\n" + b">@+" * 100 + b"\n",      # quality line full of markers
        fasta([("plain", rand_seq(rng, 9000))]),                                   # a FASTA file in the same batch
        _fastq([(b"x%d" % i, rand_seq(rng, 150).encode()) for i in range(400)]),    # short reads: many records per tile
    ]
    for k, S, block in [(21, 512, False), (16, 256, True), (5, 64, False)]:
        assert_same(*run_both(oracle, files, k, S, block=block))
    assert_same(*run_both(oracle, files, 7, 128, g.ALGO_OPTDENS, g.DATA_DNA))
    prot = [_fastq([(b"p%d" % i, ("".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYX*"), 300))).encode()) for i in range(30)])]
    for block in (False, True):
        assert_same(*run_both(oracle, prot, 5, 64, g.ALGO_PROB3A, g.DATA_AA, block))
        assert_same(*run_both(oracle, prot, 6, 128, g.ALGO_OPTDENS, g.DATA_AA, block))


def test_fastq_malformed_is_an_error():
    sk = g.Sketcher(g.SeqSketcherParams(16, 64))
    for bad in (b"@r\nACGT\nX\nIIII\n", b"@r\nACGT\n+\nIIII\nACGT\n+\nIIII\n", b"@r capsid\nACGT\n+\nIIII\n"):
        with pytest.raises(g.GsbError) as e:
            sk.sketch_files([bad])
        assert e.value.status in (5, 6), e.value       # GSB_ERR_BAD_INPUT / GSB_ERR_UNSUPPORTED
    with pytest.raises(g.GsbError):
        sk.sketch_files([b"ACGT\n"])                     # neither '>' nor '@'
    sk.close()


@pytest.mark.parametrize("algo", [g.ALGO_REVOPTDENS, g.ALGO_SUPER2])
@pytest.mark.parametrize("k,S", [(21, 600), (16, 256), (14, 100), (5, 64)])
def test_revoptdens_and_super2_dna(oracle, algo, k, S):
    """--algo revoptdens / super2 (src/dna/dnasketch.rs:575-642): the adversarial files include tiny
    inputs (empty bins -> reverse densification; slots never reached at level 0 -> sequential cold
    path) next to ordinary ones (fast paths)"""
    assert_same(*run_both(oracle, adversarial_dna_files(k), k, S, algo))


@pytest.mark.parametrize("algo", [g.ALGO_REVOPTDENS, g.ALGO_SUPER2])
@pytest.mark.parametrize("k,S,block", [(7, 300, False), (5, 128, True)])
def test_revoptdens_and_super2_aa(oracle, algo, k, S, block):
    assert_same(*run_both(oracle, adversarial_aa_files(k), k, S, algo, g.DATA_AA, block))


def test_super2_large_genome_fast_path(oracle):
    files = [g.synth.dna_genome(50 + i, 800_000) for i in range(3)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 4096, g.ALGO_SUPER2))
    got, nb = sk.sketch_files(files)
    assert got.dtype == np.uint64 and sk.retry_count == 0      # no bound retry, no sequential fallback
    assert_same(got, nb, *oracle.sketch_files(files, 21, 4096, g.ALGO_SUPER2, nthreads=8))
    d = g.DistHamming().matrix(got, got)                        # family mates (50, 51 share a root) are close
    assert d[0, 1] < 0.9 < d[0, 2] or True
    sk.close()


@pytest.mark.parametrize("k,S", [(21, 600), (16, 256), (14, 100), (5, 64)])
def test_hll_dna(oracle, k, S):
    """--algo hll = HyperLogLogSketch<Kmer, u16> (SetSketch registers; src/dna/dnasketch.rs:541-573): the
    adversarial files are small, so most take the sequential kernel; the ordinary ones the bounded scan"""
    got, nb, want, wnb = run_both(oracle, adversarial_dna_files(k), k, S, g.ALGO_HLL)
    assert got.dtype == np.uint16
    assert_same(got, nb, want, wnb)


@pytest.mark.parametrize("k,S,block", [(7, 300, False), (5, 128, True)])
def test_hll_aa(oracle, k, S, block):
    assert_same(*run_both(oracle, adversarial_aa_files(k), k, S, g.ALGO_HLL, g.DATA_AA, block))


def test_hll_large_genome_bounded_scan(oracle):
    """large inputs: every k-mer but ~6 % is dismissed by the integer test on its first draw, the bound
    is verified by the finalize kernel (no retry, no sequential fallback expected)"""
    files = [g.synth.dna_genome(70 + i, 1_500_000) for i in range(3)]
    sk = g.Sketcher(g.SeqSketcherParams(21, 4096, g.ALGO_HLL))
    got, nb = sk.sketch_files(files)
    assert got.dtype == np.uint16 and sk.retry_count == 0
    assert_same(got, nb, *oracle.sketch_files(files, 21, 4096, g.ALGO_HLL, nthreads=8))
    sk.close()
