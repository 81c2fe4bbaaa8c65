"""CPU tests of the drop-in boundary: the library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gsearch_b200 as g
from gsearch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gsearch_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"GSB_API[^;(]*?\b(gsb_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 20
    L = C.CDLL(_lib.lib_path)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gsearch_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SYMBOLS) == syms


def test_version_and_device_count():
    assert "sm_100a" in g.version()
    assert g.device_count() >= 0


def test_parameter_validation_happens_before_device_probe():
    for bad in [g.SeqSketcherParams(32, 1000), g.SeqSketcherParams(0, 1000), g.SeqSketcherParams(21, 1),
                g.SeqSketcherParams(21, 70000), g.SeqSketcherParams(13, 100, data_t=g.DATA_AA)]:
        with pytest.raises(g.GsbError) as e:
            g.Sketcher(bad)
        assert e.value.status == 1
    with pytest.raises(g.GsbError) as e:
        g.Sketcher(g.SeqSketcherParams(21, 1000, algo=6))
    assert e.value.status == 1


@pytest.mark.skipif(g.device_count() > 0, reason="CPU-box behaviour")
def test_no_cpu_fallback():
    with pytest.raises(g.GsbError) as e:
        g.Sketcher(g.SeqSketcherParams(21, 1000))
    assert e.value.status == 2 and "no CPU path" in str(e.value)
    a = np.zeros((1, 64), dtype=np.uint32)
    with pytest.raises(g.GsbError) as e:
        g.DistHamming().matrix(a, a)
    assert e.value.status == 2
    with pytest.raises(g.GsbError) as e:
        g.Hnsw(g.HnswParams(), 64, np.uint32)
    assert e.value.status == 2


def test_sig_type_table_matches_oracle(oracle):
    for data_t, ks in ((g.DATA_DNA, range(1, 32)), (g.DATA_AA, range(1, 13))):
        for k in ks:
            for algo in (g.ALGO_PROB3A, g.ALGO_OPTDENS, g.ALGO_SUPER):
                assert g.SeqSketcherParams(k, 64, algo, data_t).sig_type() == oracle.sig_type(k, 64, algo, data_t)


def test_synthetic_generator_is_seeded_and_parses(oracle):
    a = g.synth.dna_genome(5, 20000, 3)
    assert a == g.synth.dna_genome(5, 20000, 3)
    assert a != g.synth.dna_genome(6, 20000, 3)
    seqs = oracle.parse_fasta(a, 0, False)
    assert len(seqs) == 3
    n = sum(len(s) for s in seqs)
    assert 20000 * 0.99 <= n + 20 <= 20000        # 0.1 % 'N'
    # family mates share most k-mers, strangers do not
    k0 = set(oracle.kmer_values(g.synth.dna_genome(16, 20000), 0, 16).tolist())
    k1 = set(oracle.kmer_values(g.synth.dna_genome(17, 20000), 0, 16).tolist())
    k2 = set(oracle.kmer_values(g.synth.dna_genome(40, 20000), 0, 16).tolist())
    assert len(k0 & k1) > 0.5 * len(k0) and len(k0 & k2) < 0.01 * len(k0)
    p = g.synth.aa_proteome(3, 20, 100)
    assert len(oracle.parse_fasta(p, 1, False)) == 20
