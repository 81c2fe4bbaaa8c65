"""`gsearch` command line over the B200 path: the reference's tohnsw / add / request sub-commands
with the same flags, file names and text formats (src/bin/gsearch.rs:417-587).  The side files
(seqdict.json, parameters.json, processing_state.json, gsearch.neighbors.txt) are byte-compatible
with the reference's; hnswdump.hnsw.{graph,data} keep the reference's NAMES but this library's own
layout, so a database is served by the binary that built it.  Host side only walks directories, inflates files and writes
JSON / text; sketching, HNSW construction and search all run in libgsearch_b200.so on the GPU.

  gsearch [--pio N] [--nbthreads N] tohnsw -d DIR -k K -s S -n NBNG [--ef EF]
          [--scale_modify_f F] --algo prob|super|super2|optdens|revoptdens|hll [--aa] [--block]
  gsearch add -b DBDIR -n NEWDIR
  gsearch request -b DBDIR -r QUERYDIR -n NBANSWERS
  gsearch bindash -q QUERY_LIST -r REFERENCE_LIST [-k 16] [-s 2048] [-d 0|1] [-o OUT]   (src/bin/bindash.rs)
  gsearch reformat KMER MODEL gsearch.neighbors.txt OUT.tsv                              (src/bin/reformat.rs)

Outputs, as in the reference: tohnsw writes hnswdump.hnsw.graph, hnswdump.hnsw.data, seqdict.json,
parameters.json and processing_state.json into the CURRENT directory (src/dna/dnasketch.rs:152-156),
add writes them back into DBDIR, request writes gsearch.neighbors.txt (src/dna/dnarequest.rs:85-87) and, in
sequence mode, gsearch.matches (src/matcher.rs:233-277) into the current directory."""
import argparse
import bz2
import gzip
import json
import lzma
import math
import os
import sys
import time

import numpy as np

DNA_SUFFIXES = ("fna.gz", "fa.gz", "fa.xz", "fna.xz", "fasta.xz", "fa.bz2", "fna.bz2", "fasta.bz2", "fasta.gz",
                "fna", "fa", "fasta")                      # src/utils/files.rs:116-137
AA_SUFFIXES = ("faa.gz", "faa", "faa.xz", "faa.bz2")       # src/utils/files.rs:140-146
ALGO_JSON = {"prob": "PROB3A", "super": "SUPER", "optdens": "OPTDENS", "revoptdens": "REVOPTDENS",
             "super2": "SUPER2", "hll": "HLL"}
EF_SEARCH = 5000                                           # src/bin/gsearch.rs:893
ANSWER_THRESHOLD = 0.99                                    # src/dna/dnarequest.rs:83


def is_fasta_file(path, aa):
    return path.endswith(AA_SUFFIXES if aa else DNA_SUFFIXES)


def walk_fasta(directory, aa):
    """process_dir: recursive walk, files in directory-listing order (src/utils/files.rs:151-217)"""
    out = []
    for root, dirs, files in os.walk(directory):
        dirs.sort()
        for f in sorted(files):
            p = os.path.join(root, f)
            if is_fasta_file(p, aa):
                out.append(p)
    return out


def read_inflated(path):
    """needletail opens gz / bz2 / xz transparently (src/dna/dnafiles.rs:52)"""
    if path.endswith(".gz"):
        return gzip.open(path, "rb").read()
    if path.endswith(".bz2"):
        return bz2.open(path, "rb").read()
    if path.endswith(".xz"):
        return lzma.open(path, "rb").read()
    return open(path, "rb").read()


# ---------------------------------------------------------------- on-disk side files
def dump_parameters(dirpath, p):
    doc = {"hnsw": {"capacity": p["capacity"], "ef": p["ef"], "max_nb_conn": p["nbng"],
                    "scale_modification": p["scale"]},
           "sketch": {"kmer_size": p["kmer"], "sketch_size": p["sketch"], "algo": ALGO_JSON[p["algo"]],
                      "data_t": "AA" if p["aa"] else "DNA"},
           "block_flag": bool(p["block"])}
    with open(os.path.join(dirpath, "parameters.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))


def reload_parameters(dirpath):
    doc = json.load(open(os.path.join(dirpath, "parameters.json")))
    inv = {v: k for k, v in ALGO_JSON.items()}
    return {"capacity": doc["hnsw"]["capacity"], "ef": doc["hnsw"]["ef"], "nbng": doc["hnsw"]["max_nb_conn"],
            "scale": doc["hnsw"]["scale_modification"], "kmer": doc["sketch"]["kmer_size"],
            "sketch": doc["sketch"]["sketch_size"], "algo": inv[doc["sketch"]["algo"]],
            "aa": doc["sketch"]["data_t"] == "AA", "block": doc["block_flag"]}


def dump_seqdict(dirpath, items):
    """SeqDict::dump: serde objects back to back, no separator (src/utils/idsketch.rs:164-197)"""
    with open(os.path.join(dirpath, "seqdict.json"), "w") as f:
        for path, fasta_id, length in items:
            json.dump({"id": {"path": path, "fasta_id": fasta_id}, "len": int(length)}, f, separators=(",", ":"))


def reload_seqdict(dirpath):
    """stream reload of concatenated JSON objects (src/utils/idsketch.rs:201-253)"""
    txt = open(os.path.join(dirpath, "seqdict.json")).read()
    dec, pos, items = json.JSONDecoder(), 0, []
    while pos < len(txt):
        obj, pos = dec.raw_decode(txt, pos)
        items.append((obj["id"]["path"], obj["id"]["fasta_id"], obj["len"]))
        while pos < len(txt) and txt[pos].isspace():
            pos += 1
    return items


def dump_state(dirpath, nb_seq, nb_file, elapsed):
    with open(os.path.join(dirpath, "processing_state.json"), "w") as f:
        json.dump({"nb_seq": nb_seq, "nb_file": nb_file, "elapsed_t": float(elapsed)}, f, separators=(",", ":"))


def format_answers(rank, qitem, neighbours, seqdict, threshold=ANSWER_THRESHOLD):
    """ReqAnswer::dump, byte for byte (src/answer.rs:35-76)"""
    qpath, qfid, qlen = qitem
    if not any(d <= threshold for _, d in neighbours):
        return ""
    out = [f"\n{rank}\t{qpath}\tfasta_id:\t{qfid}\tlength:\t{qlen}"]
    for d_id, d in neighbours:
        if d < threshold:
            path, fid, length = seqdict[d_id]
            mant, exp = f"{d:.5E}".split("E")            # Rust {:.5E}: no sign, no padding in the exponent
            out.append(f"\nquery_id:\t{qpath}\tdistance:\t{mant}E{int(exp)}\tanswer_fasta_path\t{path}\t"
                       f"{fid} \t answer_seq_len:\t {length}")
    return "".join(out)


# ---------------------------------------------------------------- GPU pipeline
def read_all(paths, nbthreads):
    """inflate a batch of files on a thread pool (zlib / bz2 / lzma release the GIL); the reference
    does the same with rayon in process_dir_parallel (src/utils/files.rs:258-341)"""
    if nbthreads <= 1 or len(paths) <= 1:
        return [read_inflated(f) for f in paths]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=nbthreads) as ex:
        return list(ex.map(read_inflated, paths))


def algo_id(g, name):
    """--algo names as the reference spells them (src/bin/gsearch.rs:181-196); the library rejects the
    ones that are not built with GSB_ERR_UNSUPPORTED and says so"""
    return g.SeqSketcherParams.algo_from_cli(name)


class DeviceSignatures:
    """signatures of a list of files, sketched on one GPU and LEFT on it (n x S, row i = file i):
    the files are read and inflated by host threads into a pinned staging buffer, the library copies
    them to the device under its kernels, and the result never visits host memory -- it feeds
    gsb_index_insert_batch_dev / gsb_index_search_batch_dev or the all-gather directly"""

    def __init__(self, g, files, p, pio, device=0, nbthreads=0, rows=None):
        from .comm import DeviceBuffer, PinnedBuffer
        self.g, self.files, self.device = g, files, device
        self.sk = g.Sketcher(g.SeqSketcherParams(p["kmer"], p["sketch"], algo_id(g, p["algo"]),
                                                 g.DATA_AA if p["aa"] else g.DATA_DNA, bool(p["block"])), device=device)
        self.dtype = self.sk.dtype
        self.row = p["sketch"] * self.sk.elem_size
        self.rows = max(len(files), 1) if rows is None else max(rows, 1)
        self.d_sig = DeviceBuffer(self.rows * self.row, device)
        self.d_nb = DeviceBuffer(self.rows * 8, device)
        self.d_nb.upload(np.zeros(self.rows, dtype=np.uint64))
        batch = max(1, min(pio, 256))
        pinned = None
        threads = nbthreads or (os.cpu_count() or 1)
        for b in range(0, len(files), batch):
            blobs = read_all(files[b:b + batch], threads)
            total = sum(len(x) for x in blobs)
            if pinned is None or pinned.nbytes < total:
                if pinned is not None:
                    pinned.free()
                pinned = PinnedBuffer(int(total * 1.25) + 4096)
            offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
            pos = 0
            for i, x in enumerate(blobs):
                pinned.array[pos:pos + len(x)] = np.frombuffer(x, dtype=np.uint8)
                pos += len(x)
                offs[i + 1] = pos
            self.sk.sketch_buffer_to_device(pinned.array, offs, self.d_sig.ptr + b * self.row, self.d_nb.ptr + b * 8)
        if pinned is not None:
            pinned.free()

    def nb_bases(self, n=None):
        return self.d_nb.download(np.uint64, self.rows)[:len(self.files) if n is None else n]

    def host_signatures(self):
        n = len(self.files)
        S = self.row // np.dtype(self.dtype).itemsize
        return self.d_sig.download(self.dtype, n * S).reshape(n, S)


def item_of(path, nb, block):
    # ItemDict ids: seq mode keeps an empty fasta_id (src/dna/dnafiles.rs:86-93), block mode the
    # literal "-total-sequence" (src/dna/dnafiles.rs:268-272)
    return (path, "-total-sequence" if block else "", int(nb))


def open_index(g, p, dtype, device=0):
    return g.Hnsw(g.HnswParams(max_nb_conn=min(255, p["nbng"]), capacity=p["capacity"], ef=p["ef"],
                               scale_modification=p["scale"]), p["sketch"], dtype, device=device)


def dumpall(idx, dirpath, seqdict, p, nb_file, elapsed):
    """dumpall (src/utils/dumpload.rs:15-62)"""
    idx.file_dump(dirpath, "hnswdump")
    dump_seqdict(dirpath, seqdict)
    dump_parameters(dirpath, p)
    dump_state(dirpath, len(seqdict), nb_file, elapsed)


WAVE_PER_GPU = 296   # one insertion wave = two points per SM of every GPU (capped by the library)
WAVE_CAP = 1024       # (every point of a wave meets the earlier ones by exact distance: O(W^2) per wave)


def build_worker(rank, world, uid, files, p, pio, nbthreads, load_dir, out_dir, first_id, result_q):
    """one process per GPU: sketch this rank's files (file i -> rank i mod world), all-gather the
    signatures in file order, insert them (sharded by point inside every wave) and, on rank 0, dump"""
    import gsearch_b200 as g
    from .comm import Comm, DeviceBuffer, shard_rows
    t0 = time.time()
    n = len(files)
    per = shard_rows(n, world)
    local = DeviceSignatures(g, files[rank::world], p, pio, device=rank, nbthreads=nbthreads, rows=per)
    t1 = time.time()
    comm = None
    if world > 1:
        comm = Comm(uid, world, rank, rank)
        d_tmp = DeviceBuffer(world * per * local.row, rank)
        d_all = DeviceBuffer(n * local.row, rank)
        comm.all_gather_rows(local.d_sig.ptr, per, local.row, n, d_tmp.ptr, d_all.ptr)
        d_nb_all = DeviceBuffer(n * 8, rank)
        comm.all_gather_rows(local.d_nb.ptr, per, 8, n, d_tmp.ptr, d_nb_all.ptr)
        nb = d_nb_all.download(np.uint64, n)
        d_tmp.free()
        local.d_sig.free()
    else:
        d_all, nb = local.d_sig, local.nb_bases()
    idx = open_index(g, p, local.dtype, device=rank)
    if load_dir is not None:
        idx.load(load_dir, "hnswdump")
        assert idx.get_nb_point() == first_id                       # src/dna/dnasketch.rs:438
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    if world > 1:
        idx.set_wave_max(min(WAVE_CAP, WAVE_PER_GPU * world))
        idx.insert_sharded(comm, d_all.ptr, ids)
    else:
        idx.insert_device(d_all.ptr, ids)
    t2 = time.time()
    if rank == 0:
        idx.file_dump(out_dir, "hnswdump")
        result_q.put(([item_of(f, nb[i], p["block"]) for i, f in enumerate(files)], t1 - t0, t2 - t1))
    if comm is not None:
        comm.close()


def run_build(a, files, p, load_dir, out_dir, first_id):
    """-> (items of the new files, sketch seconds, insertion seconds); hnswdump.* written to out_dir"""
    world = max(1, a.gpus)
    if world == 1:
        import queue
        q = queue.Queue()
        build_worker(0, 1, None, files, p, a.pio, a.nbthreads, load_dir, out_dir, first_id, q)
        return q.get()
    import multiprocessing as mp
    import gsearch_b200 as g
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid = g.comm.unique_id()
    procs = [ctx.Process(target=build_worker, args=(r, world, uid, files, p, a.pio, a.nbthreads, load_dir, out_dir,
                                                    first_id, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    items, ts, ti = q.get()
    for pr in procs:
        pr.join()
        if pr.exitcode != 0:
            raise SystemExit(f"a GPU worker failed (exit code {pr.exitcode})")
    return items, ts, ti


def cmd_tohnsw(a):
    p = {"capacity": 1_500_000, "ef": a.ef, "nbng": a.nbng & 0xFF, "scale": a.scale_modify_f, "kmer": a.kmer,
         "sketch": a.sketch, "algo": a.algo, "aa": a.aa, "block": a.block}   # `nbng as u8`, gsearch.rs:268
    files = walk_fasta(a.dir, p["aa"])
    if not files:
        raise SystemExit(f"no fasta file found in {a.dir}")
    t0 = time.time()
    items, ts, ti = run_build(a, files, p, None, ".", 0)
    dump_seqdict(".", items)
    dump_parameters(".", p)
    dump_state(".", len(items), len(items), time.time() - t0)
    print(f"tohnsw: {len(items)} files on {max(1, a.gpus)} GPU(s), sketch {ts:.2f} s, hnsw insertion {ti:.2f} s")


def cmd_add(a):
    p = reload_parameters(a.hnsw)
    seqdict = reload_seqdict(a.hnsw)
    files = walk_fasta(a.new, p["aa"])
    if not files:
        raise SystemExit(f"no fasta file found in {a.new}")
    t0 = time.time()
    items, ts, ti = run_build(a, files, p, a.hnsw, a.hnsw, len(seqdict))
    seqdict += items
    dump_seqdict(a.hnsw, seqdict)
    dump_parameters(a.hnsw, p)
    dump_state(a.hnsw, len(seqdict), len(seqdict), time.time() - t0)
    print(f"add: {len(items)} new files, database now holds {len(seqdict)}")


def request_worker(rank, world, files, p, pio, nbthreads, dbdir, knbn, result_q):
    """replicated index, query j -> rank j mod world (src/dna/dnarequest.rs:353: one query = one task);
    query signatures go from the sketcher to the search kernel without leaving the device"""
    import gsearch_b200 as g
    from .comm import DeviceBuffer
    from .index import NEIGHBOUR_DTYPE
    mine = files[rank::world]
    q = DeviceSignatures(g, mine, p, pio, device=rank, nbthreads=nbthreads)
    idx = open_index(g, p, q.dtype, device=rank)
    idx.load(dbdir, "hnswdump")
    nq = len(mine)
    out = np.zeros((nq, knbn), dtype=NEIGHBOUR_DTYPE)
    cnt = np.zeros(nq, dtype=np.uint32)
    if nq:
        d_out = DeviceBuffer(nq * knbn * NEIGHBOUR_DTYPE.itemsize, rank)
        d_cnt = DeviceBuffer(nq * 4, rank)
        idx.search_device(q.d_sig.ptr, nq, knbn, EF_SEARCH, d_out.ptr, d_cnt.ptr)
        out = d_out.download(np.uint8, nq * knbn * NEIGHBOUR_DTYPE.itemsize).view(NEIGHBOUR_DTYPE).reshape(nq, knbn)
        cnt = d_cnt.download(np.uint32, nq)
    nb = q.nb_bases()
    result_q.put((rank, [item_of(f, nb[i], p["block"]) for i, f in enumerate(mine)],
                  out["d_id"].copy(), out["distance"].copy(), cnt))


def rust_exp(x, digits):
    """Rust's {:.<digits>E}: mantissa with `digits` decimals, 'E', exponent without sign padding"""
    mant, exp = f"{x:.{digits}E}".split("E")
    return f"{mant}E{int(exp)}"


def format_matches(queries, threshold=ANSWER_THRESHOLD):
    """Matcher::analyze (src/matcher.rs:233-277), the `gsearch.matches` file of sequence mode: per request
    genome the (at most five) database genomes of smallest merit, merit = product of the distances below
    the threshold of the sequences matched in that genome (MatchList::compute_merit_wl, :86-94), as f32.
    `queries` = [(request path, [(target path, distance), ...]), ...]; the reference walks a HashMap
    (arbitrary order), here the request genomes come in the order of their first appearance."""
    per_request = {}
    for qpath, nbrs in queries:
        targets = per_request.setdefault(qpath, {})
        for tpath, d in nbrs:
            merit = targets.get(tpath, 1.0)
            if np.float32(d) < np.float32(threshold):
                merit *= float(np.float32(d))
            targets[tpath] = merit
    out = []
    for qpath, targets in per_request.items():
        ranked = sorted(((t, float(np.float32(m))) for t, m in targets.items()), key=lambda tm: tm[1])
        out.append(f"\n\n request genome : {qpath}")
        for t, m in ranked[:5]:
            out.append(f"\n\t matched genome {t}  merit : {rust_exp(m, 3)}")
    return "".join(out)


def rust_f64(x):
    """Rust's `{}` for an f64: shortest digits that round-trip, never an exponent"""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    return np.format_float_positional(x, trim="-")


def cmd_reformat(a):
    """src/bin/reformat.rs: gsearch.neighbors.txt -> a table with an ANI column.  Rows are the lines that
    start with `query_id:`; columns 1, 3, 5, 7 of the tab-split line (file names reduced to their last
    component; column 7 is whatever the answer line holds there, kept byte for byte like the reference
    does); ANI model 1 (Poisson): (1 + ln(2J / (1 + J)) / k) * 100, model 2 (binomial):
    (2J / (1 + J))^(1/k) * 100 with J = 1 - distance; sorted by query name, then distance."""
    rows = []
    with open(a.input_file) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line.startswith("query_id:"):
                continue
            parts = line.split("\t")
            d = float(parts[3])
            j = 1.0 - d
            frac = j * 2.0 / (j + 1.0)
            if a.model == 1:
                ani = rust_f64((1.0 + (math.log(frac) if frac > 0 else float("-inf")) / a.kmer) * 100.0)
            elif a.model == 2:
                ani = rust_f64(math.pow(frac, 1.0 / a.kmer) * 100.0)
            else:
                ani = "Invalid Model"
            rows.append((os.path.basename(parts[1]), d, f"{os.path.basename(parts[1])}\t{rust_f64(d)}\t"
                         f"{os.path.basename(parts[5])}\t{parts[7]}\t{ani}"))
    rows.sort(key=lambda r: (r[0], r[1]))
    with open(a.output_file, "w") as f:
        f.write("Query_Name\tDistance\tNeighbor_Fasta_name\tNeighbor_Seq_Len\tANI\n")
        for r in rows:
            f.write(r[2] + "\n")


def cmd_request(a):
    p = reload_parameters(a.hnsw)
    seqdict = reload_seqdict(a.hnsw)
    files = walk_fasta(a.query, p["aa"])
    if not files:
        raise SystemExit(f"no fasta file found in {a.query}")
    world = max(1, a.gpus)
    if world == 1:
        import queue
        q = queue.Queue()
        request_worker(0, 1, files, p, a.pio, a.nbthreads, a.hnsw, a.nbanswers, q)
        parts = [q.get()]
    else:
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=request_worker, args=(r, world, files, p, a.pio, a.nbthreads, a.hnsw,
                                                          a.nbanswers, q)) for r in range(world)]
        for pr in procs:
            pr.start()
        parts = [q.get() for _ in procs]
        for pr in procs:
            pr.join()
            if pr.exitcode != 0:
                raise SystemExit(f"a GPU worker failed (exit code {pr.exitcode})")
    by_rank = {r: (items, ids, dd, cnt) for r, items, ids, dd, cnt in parts}
    matches = []
    with open("gsearch.neighbors.txt", "w") as f:
        for j in range(len(files)):   # query order = file order, whatever the number of GPUs
            items, ids, dd, cnt = by_rank[j % world]
            t = j // world
            nbrs = [(int(ids[t, i]), float(dd[t, i])) for i in range(cnt[t])]
            f.write(format_answers(j, items[t], nbrs, seqdict))
            matches.append((items[t][0], [(seqdict[d_id][0], d) for d_id, d in nbrs]))
    if not p["block"]:   # sequence mode: seq_matcher.analyze() (src/bin/gsearch.rs:926-929)
        with open("gsearch.matches", "w") as f:
            f.write(format_matches(matches))
    print(f"request: {len(files)} queries answered in gsearch.neighbors.txt")


def cmd_bindash(a):
    """bindash-rs (src/bin/bindash.rs): OptDens / RevOptDens sketches of two genome lists and ALL
    query x reference distances.  Sketches stay on the device; the distance matrix is one tiled
    Hamming kernel (K6) instead of |Q| x |R| DistHamming::eval calls (bindash.rs:93-164); the ANI-style
    transformation 1 - (2J / (1 + J))^(1/k) and the text output are the reference's (:92-99, :101-164).
    Unlike the reference's k <= 14 branch (bindash.rs:349-354) k-mers are always canonical here."""
    import ctypes as C
    import gsearch_b200 as g
    from . import _lib
    from .comm import DeviceBuffer
    read_list = lambda p: [ln.strip() for ln in open(p) if ln.strip()]
    qfiles, rfiles = read_list(a.query_list), read_list(a.reference_list)
    p = {"kmer": a.kmer_size, "sketch": a.sketch_size, "algo": "revoptdens" if a.densification == 1 else "optdens",
         "aa": False, "block": False}
    q = DeviceSignatures(g, qfiles, p, a.pio, nbthreads=a.threads or a.nbthreads)
    r = DeviceSignatures(g, rfiles, p, a.pio, nbthreads=a.threads or a.nbthreads) if rfiles != qfiles else q
    nq, nr = len(qfiles), len(rfiles)
    d_out = DeviceBuffer(max(1, nq * nr) * 4)
    _lib.check(_lib.lib().gsb_hamming_matrix_dev(C.c_void_p(q.d_sig.ptr), nq, C.c_void_p(r.d_sig.ptr), nr,
                                                 a.sketch_size, g.SIG_F32, C.c_void_p(d_out.ptr), C.c_void_p(0)))
    ham = d_out.download(np.float32, nq * nr).reshape(nq, nr)   # (the copy synchronises with the kernel)
    j = np.float32(1.0) - ham
    frac = (np.float32(2.0) * j) / (np.float32(1.0) + j)
    dist = 1.0 - np.power(frac, np.float32(1.0 / a.kmer_size), dtype=np.float32).astype(np.float64)
    out = open(a.output, "w") if a.output else sys.stdout
    out.write("Query\tReference\tDistance\n")
    for i, qp in enumerate(qfiles):
        for k, rp in enumerate(rfiles):
            d = 0.0 if os.path.basename(qp) == os.path.basename(rp) else dist[i, k]
            out.write(f"{qp}\t{rp}\t{d:.6f}\n")
    if a.output:
        out.close()


def build_parser():
    ap = argparse.ArgumentParser(prog="gsearch", description="GSearch sketch-and-search path on B200")
    ap.add_argument("--pio", type=int, default=64, help="files read and sketched together")
    ap.add_argument("--nbthreads", type=int, default=0, help="host threads that read and inflate files (0 = all cores)")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this node (the only extension of the reference's "
                                                       "command line): one process per GPU, NCCL all-gather of the "
                                                       "signatures, HNSW insertion sharded inside every wave")
    sub = ap.add_subparsers(dest="cmd", required=True)
    t = sub.add_parser("tohnsw")
    t.add_argument("-d", "--dir", required=True)
    t.add_argument("-k", "--kmer", type=int, required=True)
    t.add_argument("-s", "--sketch", type=int, required=True)
    t.add_argument("-n", "--nbng", type=int, required=True)
    t.add_argument("--ef", type=int, default=400)
    t.add_argument("--scale_modify_f", type=float, default=1.0)
    t.add_argument("--algo", required=True, choices=sorted(ALGO_JSON))
    t.add_argument("--aa", action="store_true")
    t.add_argument("--block", action="store_true")
    t.set_defaults(fn=cmd_tohnsw)
    ad = sub.add_parser("add")
    ad.add_argument("-b", "--hnsw", required=True)
    ad.add_argument("-n", "--new", required=True)
    ad.set_defaults(fn=cmd_add)
    r = sub.add_parser("request")
    r.add_argument("-b", "--hnsw", required=True)
    r.add_argument("-n", "--nbanswers", type=int, required=True)
    r.add_argument("-r", "--query", required=True)
    r.set_defaults(fn=cmd_request)
    rf = sub.add_parser("reformat", help="gsearch.neighbors.txt -> table with ANI (src/bin/reformat.rs)")
    rf.add_argument("kmer", type=int)
    rf.add_argument("model", type=int)
    rf.add_argument("input_file")
    rf.add_argument("output_file")
    rf.set_defaults(fn=cmd_reformat)
    b = sub.add_parser("bindash", help="all-pairs OptDens / RevOptDens distances (src/bin/bindash.rs)")
    b.add_argument("-q", "--query_list", required=True)
    b.add_argument("-r", "--reference_list", required=True)
    b.add_argument("-k", "--kmer_size", type=int, default=16)
    b.add_argument("-s", "--sketch_size", type=int, default=2048)
    b.add_argument("-d", "--densification", type=int, default=0, choices=[0, 1])
    b.add_argument("-t", "--threads", type=int, default=0)
    b.add_argument("-o", "--output", default=None)
    b.set_defaults(fn=cmd_bindash)
    return ap


def main(argv=None):
    a = build_parser().parse_args(argv)
    if a.cmd == "tohnsw" and not (0.2 <= a.scale_modify_f <= 1.0):
        raise SystemExit("scale_modify_f must be in [0.2, 1]")          # src/bin/gsearch.rs:230-240
    a.fn(a)


if __name__ == "__main__":
    main()
