"""Host mirror of anndists::dist::DistHamming (trait Distance<T>::eval) [U], used at
src/dna/dnasketch.rs:139 and src/bin/bindash.rs:94-95."""
import ctypes as C

import numpy as np

from . import _lib
from .params import SIG_F32, SIG_U16, SIG_U32, SIG_U64

_TYPES = {np.dtype(np.uint32): SIG_U32, np.dtype(np.uint64): SIG_U64, np.dtype(np.float32): SIG_F32,
          np.dtype(np.uint16): SIG_U16}


class DistHamming:
    def __init__(self, device: int = 0):
        self.device = device

    def eval(self, va, vb) -> float:
        """eval(&[T], &[T]) -> f32"""
        va = np.ascontiguousarray(va)
        vb = np.ascontiguousarray(vb, dtype=va.dtype)
        return float(self.matrix(va[None, :], vb[None, :])[0, 0])

    def batch(self, q, cands):
        """one query against n candidates (the neighbour-expansion shape)"""
        q = np.ascontiguousarray(q)
        return self.matrix(q[None, :], cands)[0]

    def matrix(self, queries, cands):
        """nq x n distances (the all-pairs shape of src/bin/bindash.rs:93-164)"""
        queries = np.ascontiguousarray(queries)
        cands = np.ascontiguousarray(cands, dtype=queries.dtype)
        nq, S = queries.shape
        n = cands.shape[0]
        if n and cands.shape[1] != S:
            raise ValueError("signature lengths differ")
        out = np.zeros((nq, n), dtype=np.float32)
        st = _TYPES[queries.dtype]
        _lib.check(_lib.lib().gsb_hamming_matrix(C.c_void_p(queries.ctypes.data), nq,
                                                 C.c_void_p(cands.ctypes.data), n, S, st,
                                                 C.c_void_p(out.ctypes.data), self.device))
        return out
