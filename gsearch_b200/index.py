"""Host mirror of hnsw_rs::Hnsw<Sig, DistHamming> [U] as the reference uses it:
Hnsw::new (src/dna/dnasketch.rs:139), parallel_insert (:435), parallel_search
(src/dna/dnarequest.rs:353), get_nb_point, file_dump / load (src/utils/dumpload.rs:31)."""
import ctypes as C
import os
from collections import namedtuple

import numpy as np

from . import _lib
from .distance import _TYPES
from .params import HnswParams

Neighbour = namedtuple("Neighbour", "d_id distance p_id")

NEIGHBOUR_DTYPE = np.dtype({"names": ["d_id", "distance", "layer", "rank"],
                            "formats": [np.uint64, np.float32, np.uint8, np.int32],
                            "offsets": [0, 8, 12, 16], "itemsize": 24})  # = sizeof(gsb_neighbour)


class Hnsw:
    def __init__(self, params: HnswParams, sketch_size: int, dtype, device: int = 0):
        self.params = params
        self.dtype = np.dtype(dtype)
        self.sketch_size = sketch_size
        cp = _lib.IndexParams(params.max_nb_conn, params.capacity, params.max_layer, params.ef,
                              params.scale_modification, _TYPES[self.dtype], sketch_size,
                              1 if params.extend_candidates else 0, 1 if params.keep_pruned else 0,
                              params.level_seed)
        h = C.c_void_p()
        _lib.check(_lib.lib().gsb_index_create(C.byref(cp), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().gsb_index_destroy(self._h)
            self._h = None

    __del__ = close

    def get_nb_point(self):
        return _lib.lib().gsb_index_nb_point(self._h)

    def parallel_insert(self, sigs, ids):
        sigs = np.ascontiguousarray(sigs, dtype=self.dtype)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_index_insert_batch(self._h, C.c_void_p(sigs.ctypes.data),
                                                     C.c_void_p(ids.ctypes.data), len(ids)))

    def insert_device(self, d_sigs_ptr, ids):
        """parallel_insert with the signatures already in device memory (raw pointer, n x S)"""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_index_insert_batch_dev(self._h, C.c_void_p(d_sigs_ptr),
                                                         C.c_void_p(ids.ctypes.data), len(ids)))

    def insert_sharded(self, comm, sigs_or_ptr, ids):
        """parallel_insert over the GPUs of `comm` (gsb_index_insert_batch_sharded): every rank calls
        it with the same signatures (numpy array, or a raw device pointer to the all-gathered matrix)
        and ids; every replica ends with the graph one GPU builds with the same wave_max"""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        if isinstance(sigs_or_ptr, np.ndarray):
            keep = np.ascontiguousarray(sigs_or_ptr, dtype=self.dtype)
            ptr = keep.ctypes.data
        else:
            ptr = int(sigs_or_ptr)
        _lib.check(_lib.lib().gsb_index_insert_batch_sharded(self._h, comm._h, C.c_void_p(ptr),
                                                             C.c_void_p(ids.ctypes.data), len(ids)))

    def set_wave_max(self, wave_max):
        _lib.check(_lib.lib().gsb_index_set_wave_max(self._h, int(wave_max)))

    def export_graph(self):
        """graph image (same keys as the oracle's export): levels, ranks, ids, CSR lists, entry"""
        n = self.get_nb_point()
        tl, tn = C.c_uint64(), C.c_uint64()
        _lib.check(_lib.lib().gsb_index_graph_sizes(self._h, C.byref(tl), C.byref(tn)))
        levels = np.zeros(max(n, 1), dtype=np.uint8)
        ranks = np.zeros(max(n, 1), dtype=np.uint32)
        ids = np.zeros(max(n, 1), dtype=np.uint64)
        off = np.zeros(tl.value + 1, dtype=np.uint64)
        idx = np.zeros(max(tn.value, 1), dtype=np.uint32)
        dist = np.zeros(max(tn.value, 1), dtype=np.float32)
        entry = C.c_uint64()
        p = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(_lib.lib().gsb_index_export_graph(self._h, p(levels), p(ranks), p(ids), p(off), p(idx), p(dist),
                                                     C.byref(entry)))
        return dict(levels=levels[:n], ranks=ranks[:n], ids=ids[:n], nbr_offsets=off, nbr_index=idx[:tn.value],
                    nbr_dist=dist[:tn.value], entry_point=entry.value)

    def load_graph(self, sigs, ids, levels, ranks, nbr_offsets, nbr_index, entry_point, nbr_dist=None):
        sigs = np.ascontiguousarray(sigs, dtype=self.dtype)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        levels = np.ascontiguousarray(levels, dtype=np.uint8)
        ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
        nbr_offsets = np.ascontiguousarray(nbr_offsets, dtype=np.uint64)
        nbr_index = np.ascontiguousarray(nbr_index, dtype=np.uint32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        if nbr_dist is not None:
            nbr_dist = np.ascontiguousarray(nbr_dist, dtype=np.float32)
        _lib.check(_lib.lib().gsb_index_load_graph(self._h, p(sigs), p(ids), len(ids), p(levels), p(ranks),
                                                   p(nbr_offsets), p(nbr_index),
                                                   p(nbr_dist) if nbr_dist is not None else C.c_void_p(0),
                                                   int(entry_point)))

    def search_raw(self, queries, knbn, ef):
        """-> (structured neighbour array nq x knbn, counts, nb_eval)"""
        queries = np.ascontiguousarray(queries, dtype=self.dtype)
        nq = queries.shape[0]
        out = np.zeros((nq, knbn), dtype=NEIGHBOUR_DTYPE)
        counts = np.zeros(nq, dtype=np.uint32)
        neval = np.zeros(nq, dtype=np.uint64)
        p = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(_lib.lib().gsb_index_search_batch(self._h, p(queries), nq, knbn, ef, p(out), p(counts),
                                                     p(neval)))
        return out, counts, neval

    def search_pointers(self, queries_ptr, nq, knbn, ef, out_ptr, counts_ptr, nb_eval_ptr=0):
        """gsb_index_search_batch with raw HOST pointers (e.g. pinned buffers: the copies then run at link
        speed): queries nq x S, out nq x knbn gsb_neighbour (24 B each), counts nq x u32, nb_eval nq x u64"""
        _lib.check(_lib.lib().gsb_index_search_batch(self._h, C.c_void_p(queries_ptr), nq, knbn, ef,
                                                     C.c_void_p(out_ptr), C.c_void_p(counts_ptr),
                                                     C.c_void_p(nb_eval_ptr) if nb_eval_ptr else C.c_void_p(0)))

    def search_device(self, d_queries_ptr, nq, knbn, ef, d_out_ptr, d_counts_ptr, d_nb_eval_ptr=0):
        """gsb_index_search_batch_dev: queries and results in device memory (raw pointers); out is
        nq x knbn gsb_neighbour (24 B each), counts nq x u32, nb_eval nq x u64 (optional)"""
        _lib.check(_lib.lib().gsb_index_search_batch_dev(self._h, C.c_void_p(d_queries_ptr), nq, knbn, ef,
                                                         C.c_void_p(d_out_ptr), C.c_void_p(d_counts_ptr),
                                                         C.c_void_p(d_nb_eval_ptr) if d_nb_eval_ptr else C.c_void_p(0)))

    def parallel_search(self, queries, knbn, ef):
        """parallel_search(&[Vec<Sig>], knbn, ef) -> Vec<Vec<Neighbour>>"""
        out, counts, _ = self.search_raw(queries, knbn, ef)
        res = []
        for i in range(out.shape[0]):
            res.append([Neighbour(int(r["d_id"]), float(r["distance"]), (int(r["layer"]), int(r["rank"])))
                        for r in out[i, :counts[i]]])
        return res

    def export_signatures(self):
        """the signatures in insertion order (n x S), copied to host memory"""
        n = self.get_nb_point()
        out = np.zeros((n, self.sketch_size), dtype=self.dtype)
        if n:
            _lib.check(_lib.lib().gsb_index_export_signatures(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def file_dump(self, directory, basename="hnswdump"):
        """hnsw.file_dump (src/utils/dumpload.rs:31).  Native layout by default; GSB_DUMP_HNSWIO=1 writes
        the hnsw_rs hnswio-style layout of gsearch_b200/hnswio.py (byte compatibility unverified)"""
        if os.environ.get("GSB_DUMP_HNSWIO", "0") not in ("", "0"):
            from . import hnswio
            hnswio.dump(str(directory), basename, self.export_graph(), self.export_signatures(),
                        self.params.max_nb_conn, self.params.ef)
            return
        _lib.check(_lib.lib().gsb_index_dump(self._h, str(directory).encode(), basename.encode()))

    def load(self, directory, basename="hnswdump"):
        """HnswIo::load_hnsw; the layout is recognised by its magic number"""
        from . import hnswio
        if hnswio.is_hnswio(str(directory), basename):
            im = hnswio.load(str(directory), basename)
            if np.dtype(im["dtype"]) != np.dtype(self.dtype) or im["sigs"].shape[1] != self.sketch_size:
                raise ValueError("dump holds %s signatures of size %d" % (im["dtype"], im["sigs"].shape[1]))
            self.load_graph(im["sigs"], im["ids"], im["levels"], im["ranks"], im["nbr_offsets"], im["nbr_index"],
                            im["entry_point"], im["nbr_dist"])
            return
        _lib.check(_lib.lib().gsb_index_load(self._h, str(directory).encode(), basename.encode()))
