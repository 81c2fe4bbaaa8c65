"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/) and the seeded generator.

The reference ships no golden vectors and cannot be built here, so these fixtures
pin the oracle itself against regressions and give the GPU tests a second, committed target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _oracle as O  # noqa: E402
import gsearch_b200 as g  # noqa: E402

CASES = {
    # name: (files builder, k, S, algo, data_t, block)
    "prob_dna_k16_s256": (lambda: [g.synth.dna_genome(i, 30000, 1 + i) for i in range(3)], 16, 256, 0, 0, False),
    "prob_dna_k21_s512_block": (lambda: [g.synth.dna_genome(i, 30000, 2) for i in range(16, 19)], 21, 512, 0, 0, True),
    "optdens_aa_k7_s512": (lambda: [g.synth.aa_proteome(i, 40, 150) for i in range(3)], 7, 512, 2, 1, False),
    "optdens_dna_k21_s256": (lambda: [g.synth.dna_genome(i, 20000) for i in range(2)], 21, 256, 2, 0, False),
    "prob_aa_k6_s128": (lambda: [g.synth.aa_proteome(i, 30, 120) for i in range(2)], 6, 128, 0, 1, False),
    # SuperMinHash: the first file is small (sequential cold path on the device), the second reaches
    # every slot at the first level (bin-min fast path)
    "super_dna_k21_s256": (lambda: [g.synth.dna_genome(7, 1500), g.synth.dna_genome(8, 40000)], 21, 256, 1, 0, False),
    # round 2: RevOptDens (first file: empty bins -> reverse densification), SuperMinHash2 (u64 and u32
    # hashes; first file: sequential cold path), FASTQ input
    "revoptdens_dna_k21_s256": (lambda: [g.synth.dna_genome(9, 200), g.synth.dna_genome(10, 30000)], 21, 256, 3, 0, False),
    "super2_dna_k21_s256": (lambda: [g.synth.dna_genome(11, 1500), g.synth.dna_genome(12, 40000)], 21, 256, 4, 0, False),
    "super2_dna_k14_s128": (lambda: [g.synth.dna_genome(13, 900), g.synth.dna_genome(14, 25000)], 14, 128, 4, 0, True),
    "prob_fastq_k16_s256": (lambda: [fastq_of(g.synth.dna_genome(15, 20000, 3)), g.synth.dna_genome(16, 20000)], 16, 256, 0, 0, False),
    # SetSketch (--algo hll, u16 registers): first file small (sequential kernel), second through the bounded scan
    "hll_dna_k21_s256": (lambda: [g.synth.dna_genome(17, 1200), g.synth.dna_genome(18, 120000)], 21, 256, 5, 0, False),
}
ROUND2 = ["hll_dna_k21_s256"]  # (the other round-2 fixtures were written when this list named them)


def fastq_of(fasta_bytes):
    """the records of a FASTA file as four-line FASTQ records"""
    out = []
    for rec in fasta_bytes.split(b">")[1:]:
        head, _, body = rec.partition(b"\n")
        seq = body.replace(b"\n", b"")
        out.append(b"@" + head + b"\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n")
    return b"".join(out)


def wave_graph_fixture():
    """graph built by the oracle's wave insertion (wave_max = 24): pins the construction semantics
    the device builder follows"""
    rng = np.random.default_rng(321)
    base = rng.integers(1, 2**31, (260, 192)).astype(np.uint32)
    for i in range(1, 260):
        redraw = rng.random(192) >= 0.8
        base[i] = np.where(redraw, base[i], base[(i - 1) // 3])
    base = base[rng.permutation(260)]
    h = O.Hnsw(12, 40, 192, np.uint32, scale=0.5)
    h.insert_waves(base, np.arange(260, dtype=np.uint64) * 2 + 1, 24)
    gr = h.export()
    np.savez_compressed(os.path.join(HERE, "hnsw_wave_u32_s192.npz"), base=base, levels=gr["levels"],
                        ranks=gr["ranks"], ids=gr["ids"], nbr_offsets=gr["nbr_offsets"], nbr_index=gr["nbr_index"],
                        nbr_dist=gr["nbr_dist"], entry=np.array([gr["entry_point"]], dtype=np.uint64))
    print("wave graph fixture", len(gr["nbr_index"]), "links, entry", gr["entry_point"])


def main(only=None):
    if only == "round2":   # fixtures added in round 2: leaves the committed ones untouched
        for name in ROUND2:
            mk, k, S, algo, data_t, block = CASES[name]
            files = mk()
            sig, nb = O.sketch_files(files, k, S, algo, data_t, block)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), sig=sig, nb=nb,
                                meta=np.array([k, S, algo, data_t, int(block), len(files)]),
                                sha=np.array([hash_bytes(f) for f in files], dtype=np.uint64))
            print(name, sig.shape, sig.dtype, nb.tolist())
        return
    if only == "new":   # fixtures added after the first set: leaves the committed ones untouched
        name = "super_dna_k21_s256"
        mk, k, S, algo, data_t, block = CASES[name]
        files = mk()
        sig, nb = O.sketch_files(files, k, S, algo, data_t, block)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sig=sig, nb=nb,
                            meta=np.array([k, S, algo, data_t, int(block), len(files)]),
                            sha=np.array([hash_bytes(f) for f in files], dtype=np.uint64))
        wave_graph_fixture()
        return
    for name, (mk, k, S, algo, data_t, block) in CASES.items():
        files = mk()
        sig, nb = O.sketch_files(files, k, S, algo, data_t, block)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sig=sig, nb=nb,
                            meta=np.array([k, S, algo, data_t, int(block), len(files)]),
                            sha=np.array([hash_bytes(f) for f in files], dtype=np.uint64))
        print(name, sig.shape, sig.dtype, nb.tolist())
    # hamming + hnsw search fixture
    rng = np.random.default_rng(123)
    base = rng.integers(1, 2**40, (300, 256)).astype(np.uint64)
    for i in range(1, 300):
        redraw = rng.random(256) >= 0.85
        base[i] = np.where(redraw, base[i], base[(i - 1) // 2])
    h = O.Hnsw(16, 64, 256, np.uint64)
    h.insert(base, np.arange(300, dtype=np.uint64) + 7)
    q = base[::29]
    out, counts, neval = h.search(q, 6, 80)
    gr = h.export()
    np.savez_compressed(os.path.join(HERE, "hnsw_u64_s256.npz"), base=base, q=q, d_id=out["d_id"],
                        distance=out["distance"], layer=out["layer"], rank=out["rank"], counts=counts,
                        neval=neval, levels=gr["levels"], ranks=gr["ranks"], ids=gr["ids"],
                        nbr_offsets=gr["nbr_offsets"], nbr_index=gr["nbr_index"],
                        entry=np.array([gr["entry_point"]], dtype=np.uint64),
                        dist_q_base=O.hamming_matrix(q, base))
    print("hnsw fixture", counts.tolist(), neval.tolist())
    wave_graph_fixture()


def hash_bytes(b):
    """FNV-1a 64 of the generated input, so a generator change is detected"""
    h = 0xcbf29ce484222325
    for c in np.frombuffer(b, dtype=np.uint8)[::97]:
        h = ((h ^ int(c)) * 0x100000001b3) & ((1 << 64) - 1)
    return h


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
