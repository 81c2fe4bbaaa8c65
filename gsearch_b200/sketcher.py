"""Host mirror of kmerutils::SeqSketcherT / SeqSketcherAAT as the reference drives them
(src/dna/dnasketch.rs:336,357; src/aa/aasketch.rs:313,329): FASTA files in, one signature per
file out.  All arithmetic runs in libgsearch_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from .params import SeqSketcherParams, sig_dtype


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Sketcher:
    def __init__(self, params: SeqSketcherParams, device: int = 0):
        self.params = params
        cp = _lib.SketchParams(params.kmer_size, params.sketch_size, params.algo, params.data_t,
                               1 if params.block_flag else 0, params.spec_flags)
        h = C.c_void_p()
        _lib.check(_lib.lib().gsb_sketcher_create(C.byref(cp), device, C.byref(h)))
        self._h = h
        self.device = device
        self.sig_type = _lib.lib().gsb_sketcher_sig_type(h)
        self.dtype = sig_dtype(self.sig_type)
        self.elem_size = _lib.lib().gsb_sketcher_elem_size(h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().gsb_sketcher_destroy(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def concat(files):
        """list of bytes-like -> (uint8 array, uint64 offsets)"""
        offs = np.zeros(len(files) + 1, dtype=np.uint64)
        for i, f in enumerate(files):
            offs[i + 1] = offs[i] + len(f)
        buf = np.empty(int(offs[-1]), dtype=np.uint8)
        for i, f in enumerate(files):
            buf[int(offs[i]):int(offs[i + 1])] = np.frombuffer(bytes(f), dtype=np.uint8)
        return buf, offs

    def sketch_files(self, files):
        """sketch_compressedkmer[_seqs] over a batch of FASTA files (bytes) -> (sigs, nb_bases)."""
        buf, offs = self.concat(files)
        return self.sketch_buffer(buf, offs)

    def sketch_buffer(self, buf, offsets):
        """Host buffers (numpy uint8 / uint64 offsets) -> (n x S signatures, n encoded lengths)."""
        n = len(offsets) - 1
        S = self.params.sketch_size
        sig = np.zeros((n, S), dtype=self.dtype)
        nb = np.zeros(n, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_sketch_fasta_batch(self._h, _ptr(buf), _ptr(offsets), n, _ptr(sig), _ptr(nb)))
        return sig, nb

    def sketch_buffer_to_device(self, buf, offsets, d_sig_ptr, d_nb_ptr=0):
        """host FASTA buffers in, signatures (n x S) written at the raw DEVICE pointer d_sig_ptr -- e.g.
        this rank's slice of the matrix that is all-gathered before HNSW insertion"""
        n = len(offsets) - 1
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_sketch_fasta_batch_to_dev(self._h, _ptr(buf), _ptr(offsets), n,
                                                            C.c_void_p(d_sig_ptr), C.c_void_p(d_nb_ptr)))

    def sketch_pointers(self, bytes_ptr, offsets, n, sig_ptr, nb_ptr=0):
        """Raw HOST pointers (e.g. torch pinned tensors): the e2e path of bench.py."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_sketch_fasta_batch(self._h, C.c_void_p(bytes_ptr), _ptr(offsets), n,
                                                     C.c_void_p(sig_ptr), C.c_void_p(nb_ptr)))

    def sketch_device(self, d_bytes_ptr, offsets, n, d_sig_ptr, d_nb_ptr=0, stream=0):
        """Raw DEVICE pointers; offsets is a host array of n+1 values."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(_lib.lib().gsb_sketch_fasta_batch_dev(self._h, C.c_void_p(d_bytes_ptr), _ptr(offsets), n,
                                                         C.c_void_p(d_sig_ptr), C.c_void_p(d_nb_ptr),
                                                         C.c_void_p(stream)))

    @property
    def launch_count(self):
        return _lib.lib().gsb_sketcher_launch_count(self._h)

    @property
    def retry_count(self):
        return _lib.lib().gsb_sketcher_retry_count(self._h)

    @property
    def fallback_count(self):
        """genomes the partition path handed to the general filter path (ProbMinHash only)"""
        return _lib.lib().gsb_sketcher_fallback_count(self._h)

    def set_prob_path(self, path):
        """0 = partition path (filter path as fallback), 1 = filter path only"""
        _lib.check(_lib.lib().gsb_sketcher_set_prob_path(self._h, int(path)))

    def enable_timing(self, on=True):
        _lib.lib().gsb_sketcher_enable_timing(self._h, 1 if on else 0)

    def kernel_times(self):
        """-> {family: (total ms, timed spans)} measured with CUDA events on the launch stream"""
        ms = np.zeros(8, dtype=np.float64)
        n = np.zeros(8, dtype=np.uint64)
        _lib.lib().gsb_sketcher_kernel_times(self._h, _ptr(ms), _ptr(n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(
            ("k1_pack", "k2_scan", "k3_slots", "reset", "k2_mark", "k2_classify", "k2_exact", "k1_summary"))}
