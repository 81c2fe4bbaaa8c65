"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference path.
TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's CPU arms."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.run(["make", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)
    return SO


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("kmer_size", "sketch_size", "algo", "data_t", "block_flag", "spec_flags")]


class Xoshiro(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class Exp01(C.Structure):
    _fields_ = [("lam", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("c3", C.c_double)]


class Seqs(C.Structure):
    _fields_ = [("codes", C.POINTER(C.c_uint8)), ("seq_off", C.POINTER(C.c_uint64)), ("nseq", C.c_uint64),
                ("nb_raw", C.c_uint64)]


NEIGHBOUR_DTYPE = np.dtype({"names": ["d_id", "distance", "layer", "rank"],
                            "formats": [np.uint64, np.float32, np.uint8, np.int32],
                            "offsets": [0, 8, 12, 16], "itemsize": 24})  # = sizeof(gsb_neighbour)

_L = None


def L():
    global _L
    if _L is None:
        build()
        lib = C.CDLL(SO)
        lib.gso_splitmix64_next.restype = C.c_uint64
        lib.gso_xoshiro_next_u64.restype = C.c_uint64
        lib.gso_xoshiro_next_u32.restype = C.c_uint32
        lib.gso_uniform_f64.restype = C.c_double
        lib.gso_uniform_f32.restype = C.c_float
        lib.gso_uniform_usize.restype = C.c_uint64
        lib.gso_uniform_usize.argtypes = [C.c_void_p, C.c_uint64]
        lib.gso_exp01_init.argtypes = [C.c_void_p, C.c_double]
        lib.gso_exp01_sample.restype = C.c_double
        lib.gso_hamming.restype = C.c_float
        lib.gso_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.gso_hnsw_new.restype = C.c_void_p
        lib.gso_hnsw_new.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        lib.gso_hnsw_free.argtypes = [C.c_void_p]
        lib.gso_hnsw_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.gso_hnsw_insert_waves.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        lib.gso_hnsw_insert_waves_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int]
        lib.gso_hnsw_wave_size.argtypes = [C.c_uint64, C.c_uint32]
        lib.gso_hnsw_wave_size.restype = C.c_uint32
        lib.gso_hnsw_nb_point.restype = C.c_uint64
        lib.gso_hnsw_nb_point.argtypes = [C.c_void_p]
        lib.gso_hnsw_nb_eval.restype = C.c_uint64
        lib.gso_hnsw_nb_eval.argtypes = [C.c_void_p]
        lib.gso_hnsw_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.gso_hnsw_total_lists.restype = C.c_uint64
        lib.gso_hnsw_total_lists.argtypes = [C.c_void_p]
        lib.gso_hnsw_total_nbrs.restype = C.c_uint64
        lib.gso_hnsw_total_nbrs.argtypes = [C.c_void_p]
        lib.gso_hnsw_export.argtypes = [C.c_void_p] * 8
        lib.gso_hnsw_import.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.gso_sketch_fasta_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                               C.c_void_p, C.c_int]
        lib.gso_hamming_matrix.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_void_p, C.c_int]
        lib.gso_parse_fasta.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.gso_kmer_values.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.gso_count_kmers.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gso_probminhash3a.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_void_p, C.c_void_p]
        lib.gso_optdens.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.gso_superminhash.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        lib.gso_seqs_free.argtypes = [C.c_void_p]
        _L = lib
    return _L


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


SIG_DTYPES = {0: np.uint32, 1: np.uint64, 2: np.float32, 3: np.uint16}


def sig_type(kmer_size, sketch_size, algo, data_t):
    p = Params(kmer_size, sketch_size, algo, data_t, 0, 0)
    return L().gso_sig_type(C.byref(p))


def sketch_files(files, kmer_size, sketch_size, algo=0, data_t=0, block_flag=False, spec_flags=0, nthreads=1):
    """-> (sigs n x S, nb_bases) ; raises on a parse error"""
    offs = np.zeros(len(files) + 1, dtype=np.uint64)
    for i, f in enumerate(files):
        offs[i + 1] = offs[i] + len(f)
    buf = np.frombuffer(b"".join(bytes(f) for f in files) + b"\0", dtype=np.uint8).copy()
    return sketch_buffer(buf, offs, kmer_size, sketch_size, algo, data_t, block_flag, spec_flags, nthreads)


def sketch_buffer(buf, offs, kmer_size, sketch_size, algo=0, data_t=0, block_flag=False, spec_flags=0,
                  nthreads=1):
    p = Params(kmer_size, sketch_size, algo, data_t, 1 if block_flag else 0, spec_flags)
    n = len(offs) - 1
    dt = SIG_DTYPES[L().gso_sig_type(C.byref(p))]
    sig = np.zeros((n, sketch_size), dtype=dt)
    nb = np.zeros(n, dtype=np.uint64)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    rc = L().gso_sketch_fasta_batch(C.byref(p), _p(buf), _p(offs), n, _p(sig), _p(nb), nthreads)
    if rc:
        raise RuntimeError(f"oracle status {rc}")
    return sig, nb


def parse_fasta(data, data_t=0, block_flag=False):
    """-> list of code arrays (one per sequence)"""
    buf = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8).copy()
    s = Seqs()
    rc = L().gso_parse_fasta(_p(buf), len(data), data_t, 1 if block_flag else 0, C.byref(s))
    if rc:
        raise RuntimeError(f"oracle status {rc}")
    out = []
    for i in range(s.nseq):
        a, b = s.seq_off[i], s.seq_off[i + 1]
        out.append(np.ctypeslib.as_array(s.codes, shape=(max(b, 1),))[a:b].copy())
    L().gso_seqs_free(C.byref(s))
    return out


def kmer_values(data, data_t, k, block_flag=False):
    buf = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8).copy()
    s = Seqs()
    rc = L().gso_parse_fasta(_p(buf), len(data), data_t, 1 if block_flag else 0, C.byref(s))
    if rc:
        raise RuntimeError(f"oracle status {rc}")
    vals = C.POINTER(C.c_uint64)()
    n = C.c_uint64()
    L().gso_kmer_values(C.byref(s), data_t, k, C.byref(vals), C.byref(n))
    out = np.ctypeslib.as_array(vals, shape=(max(n.value, 1),))[:n.value].copy()
    L().gso_seqs_free(C.byref(s))
    C.CDLL(None).free(vals)
    return out


def probminhash3a(keys, weights, m, val_bytes=8, spec_flags=0):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    sig = np.zeros(m, dtype=np.uint64)
    hmin = np.zeros(m, dtype=np.float64)
    rc = L().gso_probminhash3a(_p(keys), _p(w), len(keys), m, val_bytes, spec_flags, _p(sig), _p(hmin))
    assert rc == 0
    return sig, hmin


def optdens(vals, m, spec_flags=0):
    vals = np.ascontiguousarray(vals, dtype=np.uint64)
    sig = np.zeros(m, dtype=np.float32)
    assert L().gso_optdens(_p(vals), len(vals), m, spec_flags, _p(sig)) == 0
    return sig


def revoptdens(vals, m, spec_flags=0):
    vals = np.ascontiguousarray(vals, dtype=np.uint64)
    sig = np.zeros(m, dtype=np.float32)
    assert L().gso_revoptdens(_p(vals), len(vals), m, spec_flags, _p(sig)) == 0
    return sig


def superminhash2(vals, m, kt32):
    vals = np.ascontiguousarray(vals, dtype=np.uint64)
    sig = np.zeros(m, dtype=np.uint64)
    assert L().gso_superminhash2(_p(vals), len(vals), m, 1 if kt32 else 0, _p(sig)) == 0
    return sig


def setsketch(vals, m):
    vals = np.ascontiguousarray(vals, dtype=np.uint64)
    sig = np.zeros(m, dtype=np.uint16)
    assert L().gso_setsketch(_p(vals), len(vals), m, _p(sig)) == 0
    return sig


def ln_spec(x):
    L().gso_ln_spec.restype = C.c_double
    L().gso_ln_spec.argtypes = [C.c_double]
    return L().gso_ln_spec(float(x))


def superminhash(vals, m):
    vals = np.ascontiguousarray(vals, dtype=np.uint64)
    sig = np.zeros(m, dtype=np.float32)
    assert L().gso_superminhash(_p(vals), len(vals), m, _p(sig)) == 0
    return sig


_ST = {np.dtype(np.uint32): 0, np.dtype(np.uint64): 1, np.dtype(np.float32): 2, np.dtype(np.uint16): 3}


def hamming(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    return float(L().gso_hamming(_p(a), _p(b), a.shape[0], _ST[a.dtype]))


def hamming_matrix(q, c, nthreads=1):
    q = np.ascontiguousarray(q)
    c = np.ascontiguousarray(c, dtype=q.dtype)
    out = np.zeros((q.shape[0], c.shape[0]), dtype=np.float32)
    L().gso_hamming_matrix(_p(q), q.shape[0], _p(c), c.shape[0], q.shape[1], _ST[q.dtype], _p(out), nthreads)
    return out


class Hnsw:
    def __init__(self, M, ef_c, S, dtype, capacity=1_500_000, max_layer=16, scale=1.0, extend=True,
                 keep_pruned=False, seed=0x5EED):
        self.dtype = np.dtype(dtype)
        self.S = S
        self.h = L().gso_hnsw_new(M, capacity, max_layer, ef_c, scale, _ST[self.dtype], S, int(extend),
                                  int(keep_pruned), seed)

    def __del__(self):
        if getattr(self, "h", None):
            L().gso_hnsw_free(self.h)
            self.h = None

    def insert(self, sigs, ids):
        sigs = np.ascontiguousarray(sigs, dtype=self.dtype)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        rc = L().gso_hnsw_insert(self.h, _p(sigs), _p(ids), len(ids))
        assert rc == 0, rc

    def insert_waves(self, sigs, ids, wave_max, nthreads=1):
        """deterministic wave insertion; nthreads > 1 spreads phase A of each wave over host threads
        (same graph)"""
        sigs = np.ascontiguousarray(sigs, dtype=self.dtype)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        if nthreads > 1:
            rc = L().gso_hnsw_insert_waves_mt(self.h, _p(sigs), _p(ids), len(ids), wave_max, nthreads)
        else:
            rc = L().gso_hnsw_insert_waves(self.h, _p(sigs), _p(ids), len(ids), wave_max)
        assert rc == 0, rc

    def import_graph(self, sigs, gr):
        """load a graph image (dict as returned by export / gsearch_b200.Hnsw.export_graph)"""
        sigs = np.ascontiguousarray(sigs, dtype=self.dtype)
        dt = dict(ids=np.uint64, levels=np.uint8, ranks=np.uint32, nbr_offsets=np.uint64, nbr_index=np.uint32,
                  nbr_dist=np.float32)
        a = {k: np.ascontiguousarray(gr[k], dtype=t) for k, t in dt.items()}  # kept alive over the call
        rc = L().gso_hnsw_import(self.h, _p(sigs), _p(a["ids"]), len(a["ids"]), _p(a["levels"]), _p(a["ranks"]),
                                 _p(a["nbr_offsets"]), _p(a["nbr_index"]), _p(a["nbr_dist"]),
                                 int(gr["entry_point"]))
        assert rc == 0, rc

    def nb_point(self):
        return L().gso_hnsw_nb_point(self.h)

    def nb_eval(self):
        return L().gso_hnsw_nb_eval(self.h)

    def search(self, queries, knbn, ef, nthreads=1):
        queries = np.ascontiguousarray(queries, dtype=self.dtype)
        nq = queries.shape[0]
        out = np.zeros((nq, knbn), dtype=NEIGHBOUR_DTYPE)
        counts = np.zeros(nq, dtype=np.uint32)
        neval = np.zeros(nq, dtype=np.uint64)
        L().gso_hnsw_search_batch(self.h, _p(queries), nq, knbn, ef, _p(out), _p(counts), _p(neval), nthreads)
        return out, counts, neval

    def export(self):
        n = self.nb_point()
        tl = L().gso_hnsw_total_lists(self.h)
        tn = L().gso_hnsw_total_nbrs(self.h)
        levels = np.zeros(max(n, 1), dtype=np.uint8)
        ranks = np.zeros(max(n, 1), dtype=np.uint32)
        ids = np.zeros(max(n, 1), dtype=np.uint64)
        off = np.zeros(tl + 1, dtype=np.uint64)
        idx = np.zeros(max(tn, 1), dtype=np.uint32)
        dist = np.zeros(max(tn, 1), dtype=np.float32)
        entry = C.c_uint64()
        L().gso_hnsw_export(self.h, _p(levels), _p(ranks), _p(ids), _p(off), _p(idx), _p(dist), C.byref(entry))
        return dict(levels=levels[:n], ranks=ranks[:n], ids=ids[:n], nbr_offsets=off, nbr_index=idx[:tn],
                    nbr_dist=dist[:tn], entry_point=entry.value)
