"""Seeded synthetic workloads for bench.py and the tests.

NOT part of the product: the generator is host code in its own library, datagen/libgsb_synth.so
(datagen/synth.cpp), so that a process that only generates data -- the CPU reference arm of
bench.py -- maps no product code."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(os.path.dirname(_HERE), "datagen", "libgsb_synth.so")
_L = None


def _lib():
    global _L
    if _L is None:
        if not os.path.exists(lib_path):
            raise RuntimeError(f"{lib_path} is not built; run `make -C datagen`")
        L = C.CDLL(lib_path)
        u64, u32, vp = C.c_uint64, C.c_uint32, C.c_void_p
        L.gsb_synth_max_bytes.restype, L.gsb_synth_max_bytes.argtypes = u64, [u64, u32]
        L.gsb_synth_dna_genome.restype, L.gsb_synth_dna_genome.argtypes = u64, [u64, u64, u32, vp, u64]
        L.gsb_synth_aa_proteome.restype, L.gsb_synth_aa_proteome.argtypes = u64, [u64, u32, u32, vp, u64]
        L.gsb_synth_signatures.restype, L.gsb_synth_signatures.argtypes = C.c_int, [vp, u64, u32, u32, u64, u64]
        L.gsb_synth_queries.restype = C.c_int
        L.gsb_synth_queries.argtypes = [vp, u32, vp, u64, u32, u32, u64, C.c_double, vp]
        _L = L
    return _L


def dna_genome(index, length, ncontigs=1):
    L = _lib()
    cap = L.gsb_synth_max_bytes(length, ncontigs)
    buf = np.empty(cap, dtype=np.uint8)
    n = L.gsb_synth_dna_genome(index, length, ncontigs, C.c_void_p(buf.ctypes.data), cap)
    return buf[:n].tobytes()


def dna_genome_into(index, length, ncontigs, ptr, cap):
    """write straight into a caller buffer (e.g. pinned memory); returns bytes written"""
    return _lib().gsb_synth_dna_genome(index, length, ncontigs, C.c_void_p(ptr), cap)


def aa_proteome(index, nprot, mean_len=333):
    L = _lib()
    cap = L.gsb_synth_max_bytes(nprot * (mean_len + mean_len // 2 + 2), nprot)
    buf = np.empty(cap, dtype=np.uint8)
    n = L.gsb_synth_aa_proteome(index, nprot, mean_len, C.c_void_p(buf.ctypes.data), cap)
    return buf[:n].tobytes()


def aa_proteome_into(index, nprot, mean_len, ptr, cap):
    return _lib().gsb_synth_aa_proteome(index, nprot, mean_len, C.c_void_p(ptr), cap)


def max_bytes(length, nrecords=1):
    return _lib().gsb_synth_max_bytes(length, nrecords)


def signatures(n, S, dtype=np.uint64, seed=1234, first=0, out=None):
    """rows [first, first+n) of the synthetic signature database (random recursive tree with graded
    distances); deterministic in (seed, row, slot) whatever the thread count"""
    dtype = np.dtype(dtype)
    if out is None:
        out = np.empty((n, S), dtype=dtype)
    assert out.dtype == dtype and out.shape == (n, S) and out.flags.c_contiguous
    rc = _lib().gsb_synth_signatures(C.c_void_p(out.ctypes.data), n, S, dtype.itemsize, seed, first)
    if rc:
        raise ValueError("gsb_synth_signatures: bad argument")
    return out


def queries(nq, db, seed=99, noise=0.1):
    """nq mutated members of `db`: (queries, picked row of each)"""
    db = np.ascontiguousarray(db)
    out = np.empty((nq, db.shape[1]), dtype=db.dtype)
    picked = np.empty(nq, dtype=np.uint64)
    rc = _lib().gsb_synth_queries(C.c_void_p(out.ctypes.data), nq, C.c_void_p(db.ctypes.data), db.shape[0],
                                  db.shape[1], db.dtype.itemsize, seed, noise, C.c_void_p(picked.ctypes.data))
    if rc:
        raise ValueError("gsb_synth_queries: bad argument")
    return out, picked
