"""Parameter structs of the reference path (src/utils/parameters.rs:33-42,139-147 and
kmerutils::sketcharg [U])."""
from dataclasses import dataclass

import numpy as np

ALGO_PROB3A, ALGO_SUPER, ALGO_OPTDENS, ALGO_REVOPTDENS, ALGO_SUPER2, ALGO_HLL = range(6)
DATA_DNA, DATA_AA = 0, 1
SIG_U32, SIG_U64, SIG_F32, SIG_U16 = range(4)
SPEC_NOHASH_IDENTITY, SPEC_OPTDENS_F64_DRAW = 1, 2

_ALGO_NAMES = {"prob": ALGO_PROB3A, "super": ALGO_SUPER, "optdens": ALGO_OPTDENS,
               "revoptdens": ALGO_REVOPTDENS, "super2": ALGO_SUPER2, "hll": ALGO_HLL}


def sig_dtype(sig_type):
    return {SIG_U32: np.uint32, SIG_U64: np.uint64, SIG_F32: np.float32, SIG_U16: np.uint16}[sig_type]


@dataclass
class SeqSketcherParams:
    """kmerutils::sketcharg::SeqSketcherParams{kmer_size, sketch_size, algo, data_t} plus the
    block flag of ProcessingParams (src/utils/parameters.rs:139-147)."""
    kmer_size: int
    sketch_size: int
    algo: int = ALGO_PROB3A
    data_t: int = DATA_DNA
    block_flag: bool = False
    spec_flags: int = 0

    @staticmethod
    def algo_from_cli(name):
        """--algo values accepted by src/bin/gsearch.rs:181-196."""
        return _ALGO_NAMES[name]

    def sig_type(self):
        """Sig type chosen by the reference dispatch tables (src/dna/dnasketch.rs:493-644,
        src/aa/aasketch.rs:449-552)."""
        if self.algo in (ALGO_PROB3A, ALGO_SUPER2):  # SuperHash2Sketch<Kmer, u32 | u64, Fx>: dnasketch.rs:575-599
            if self.data_t == DATA_DNA:
                return SIG_U32 if (self.kmer_size <= 14 or self.kmer_size == 16) else SIG_U64
            return SIG_U32 if self.kmer_size <= 6 else SIG_U64
        if self.algo == ALGO_HLL:  # HyperLogLogSketch<Kmer, u16>: dnasketch.rs:541-573
            return SIG_U16
        return SIG_F32


@dataclass
class HnswParams:
    """src/utils/parameters.rs:33-42 + the constants of src/dna/dnasketch.rs:139-160."""
    max_nb_conn: int = 128
    capacity: int = 1_500_000
    ef: int = 1600
    scale_modification: float = 1.0
    max_layer: int = 16
    extend_candidates: bool = True
    keep_pruned: bool = False
    level_seed: int = 0x5EED
