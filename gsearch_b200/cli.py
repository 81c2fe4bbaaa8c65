"""`gsearch` command line over the B200 path: the reference's tohnsw / add / request sub-commands
with the same flags, file names and text formats (src/bin/gsearch.rs:417-587), so that a user of
the reference can switch binaries.  Host side only walks directories, inflates files and writes
JSON / text; sketching, HNSW construction and search all run in libgsearch_b200.so on the GPU.

  gsearch [--pio N] [--nbthreads N] tohnsw -d DIR -k K -s S -n NBNG [--ef EF]
          [--scale_modify_f F] --algo prob|super|optdens [--aa] [--block]
  gsearch add -b DBDIR -n NEWDIR
  gsearch request -b DBDIR -r QUERYDIR -n NBANSWERS

Outputs, as in the reference: tohnsw writes hnswdump.hnsw.graph, hnswdump.hnsw.data, seqdict.json,
parameters.json and processing_state.json into the CURRENT directory (src/dna/dnasketch.rs:152-156),
add writes them back into DBDIR, request writes gsearch.neighbors.txt into the current directory
(src/dna/dnarequest.rs:85-87)."""
import argparse
import bz2
import gzip
import json
import lzma
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


def sketch_directory(g, directory, p, pio, device=0, nbthreads=0):
    files = walk_fasta(directory, p["aa"])
    if not files:
        raise SystemExit(f"no fasta file found in {directory}")
    algo = g.SeqSketcherParams.algo_from_name(p["algo"]) if hasattr(g.SeqSketcherParams, "algo_from_name") else \
        {"prob": g.ALGO_PROB3A, "super": g.ALGO_SUPER, "optdens": g.ALGO_OPTDENS}[p["algo"]]
    sk = g.Sketcher(g.SeqSketcherParams(p["kmer"], p["sketch"], algo, g.DATA_AA if p["aa"] else g.DATA_DNA,
                                        bool(p["block"])), device=device)
    sigs, items = [], []
    batch = max(1, min(pio, 256))
    for b in range(0, len(files), batch):
        chunk = files[b:b + batch]
        sig, nb = sk.sketch_files(read_all(chunk, nbthreads or (os.cpu_count() or 1)))
        sigs.append(sig)
        # ItemDict ids: seq mode keeps an empty fasta_id (src/dna/dnafiles.rs:86-93), block mode the
        # literal "-total-sequence" (src/dna/dnafiles.rs:268-272)
        fid = "-total-sequence" if p["block"] else ""
        items += [(f, fid, int(n)) for f, n in zip(chunk, nb)]
    return np.concatenate(sigs), items, sk.dtype


def open_index(g, p, dtype, device=0):
    return g.Hnsw(g.HnswParams(max_nb_conn=min(255, p["nbng"]), capacity=p["capacity"], ef=p["ef"],
                               scale_modification=p["scale"]), p["sketch"], dtype, device=device)


def dumpall(idx, dirpath, seqdict, p, nb_file, elapsed):
    """dumpall (src/utils/dumpload.rs:15-62)"""
    idx.file_dump(dirpath, "hnswdump")
    dump_seqdict(dirpath, seqdict)
    dump_parameters(dirpath, p)
    dump_state(dirpath, len(seqdict), nb_file, elapsed)


def cmd_tohnsw(a):
    import gsearch_b200 as g
    p = {"capacity": 1_500_000, "ef": a.ef, "nbng": a.nbng & 0xFF, "scale": a.scale_modify_f, "kmer": a.kmer,
         "sketch": a.sketch, "algo": a.algo, "aa": a.aa, "block": a.block}   # `nbng as u8`, gsearch.rs:268
    t0 = time.time()
    sig, items, dtype = sketch_directory(g, a.dir, p, a.pio, nbthreads=a.nbthreads)
    t1 = time.time()
    idx = open_index(g, p, dtype)
    idx.parallel_insert(sig, np.arange(len(items), dtype=np.uint64))
    t2 = time.time()
    dumpall(idx, ".", items, p, len(items), t2 - t0)
    print(f"tohnsw: {len(items)} files, sketch {t1 - t0:.2f} s, hnsw insertion {t2 - t1:.2f} s")


def cmd_add(a):
    import gsearch_b200 as g
    p = reload_parameters(a.hnsw)
    seqdict = reload_seqdict(a.hnsw)
    t0 = time.time()
    sig, items, dtype = sketch_directory(g, a.new, p, a.pio, nbthreads=a.nbthreads)
    idx = open_index(g, p, dtype)
    idx.load(a.hnsw, "hnswdump")
    assert idx.get_nb_point() == len(seqdict)                       # src/dna/dnasketch.rs:438
    idx.parallel_insert(sig, np.arange(len(seqdict), len(seqdict) + len(items), dtype=np.uint64))
    seqdict += items
    dumpall(idx, a.hnsw, seqdict, p, len(seqdict), time.time() - t0)
    print(f"add: {len(items)} new files, database now holds {len(seqdict)}")


def cmd_request(a):
    import gsearch_b200 as g
    p = reload_parameters(a.hnsw)
    seqdict = reload_seqdict(a.hnsw)
    sig, items, dtype = sketch_directory(g, a.query, p, a.pio, nbthreads=a.nbthreads)
    idx = open_index(g, p, dtype)
    idx.load(a.hnsw, "hnswdump")
    out, counts, _ = idx.search_raw(sig, a.nbanswers, EF_SEARCH)
    with open("gsearch.neighbors.txt", "w") as f:
        for r, item in enumerate(items):
            nbrs = [(int(out["d_id"][r, j]), float(out["distance"][r, j])) for j in range(counts[r])]
            f.write(format_answers(r, item, nbrs, seqdict))
    print(f"request: {len(items)} queries answered in gsearch.neighbors.txt")


def build_parser():
    ap = argparse.ArgumentParser(prog="gsearch", description="GSearch sketch-and-search path on B200")
    ap.add_argument("--pio", type=int, default=64, help="files read and sketched together")
    ap.add_argument("--nbthreads", type=int, default=0, help="host threads that read and inflate files (0 = all cores)")
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
    return ap


def main(argv=None):
    a = build_parser().parse_args(argv)
    if a.cmd == "tohnsw" and not (0.2 <= a.scale_modify_f <= 1.0):
        raise SystemExit("scale_modify_f must be in [0.2, 1]")          # src/bin/gsearch.rs:230-240
    a.fn(a)


if __name__ == "__main__":
    main()
