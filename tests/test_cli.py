"""The gsearch command line: flags, file names and text formats of the reference
(src/bin/gsearch.rs:417-587, src/utils/{files,idsketch,parameters}.rs, src/answer.rs).  CPU tests
cover the host-side formats; the GPU test runs tohnsw -> request -> add end to end."""
import gzip
import json
import os

import numpy as np
import pytest

import gsearch_b200 as g
from gsearch_b200 import cli


def test_suffix_rules_and_walk(tmp_path):
    for name in ["a.fna", "b.fa.gz", "c.fasta.xz", "d.faa", "e.txt", "sub/f.fna.bz2", "sub/g.faa.gz"]:
        p = tmp_path / name
        p.parent.mkdir(exist_ok=True)
        p.write_bytes(b"")
    dna = [os.path.relpath(p, tmp_path) for p in cli.walk_fasta(str(tmp_path), aa=False)]
    aa = [os.path.relpath(p, tmp_path) for p in cli.walk_fasta(str(tmp_path), aa=True)]
    assert dna == ["a.fna", "b.fa.gz", "c.fasta.xz", "sub/f.fna.bz2"]
    assert aa == ["d.faa", "sub/g.faa.gz"]


def test_side_files_round_trip(tmp_path):
    p = {"capacity": 1_500_000, "ef": 1600, "nbng": 128, "scale": 0.25, "kmer": 21, "sketch": 18000,
         "algo": "prob", "aa": False, "block": False}
    cli.dump_parameters(tmp_path, p)
    doc = json.load(open(tmp_path / "parameters.json"))
    assert doc == {"hnsw": {"capacity": 1500000, "ef": 1600, "max_nb_conn": 128, "scale_modification": 0.25},
                   "sketch": {"kmer_size": 21, "sketch_size": 18000, "algo": "PROB3A", "data_t": "DNA"},
                   "block_flag": False}
    assert cli.reload_parameters(tmp_path) == p
    items = [("/db/GCA_1.fna.gz", "", 5000123), ("/db/x y.fa", "", 7)]
    cli.dump_seqdict(tmp_path, items)
    raw = open(tmp_path / "seqdict.json").read()
    assert raw.startswith('{"id":{"path":"/db/GCA_1.fna.gz","fasta_id":""},"len":5000123}{"id":')   # no separator
    assert cli.reload_seqdict(tmp_path) == items
    cli.dump_state(tmp_path, 2, 2, 1.5)
    assert json.load(open(tmp_path / "processing_state.json")) == {"nb_seq": 2, "nb_file": 2, "elapsed_t": 1.5}


def test_answer_rows_follow_answer_rs():
    seqdict = [("/db/a.fna", "", 100), ("/db/b.fna", "", 200)]
    txt = cli.format_answers(3, ("/q/x.fna", "", 42), [(1, 0.0), (0, 0.25), (1, 0.995)], seqdict)
    rows = txt.split("\n")
    assert rows[0] == ""                                   # every row starts with a newline
    assert rows[1] == "3\t/q/x.fna\tfasta_id:\t\tlength:\t42"
    assert rows[2] == "query_id:\t/q/x.fna\tdistance:\t0.00000E0\tanswer_fasta_path\t/db/b.fna\t \t answer_seq_len:\t 200"
    assert rows[3] == "query_id:\t/q/x.fna\tdistance:\t2.50000E-1\tanswer_fasta_path\t/db/a.fna\t \t answer_seq_len:\t 100"
    assert len(rows) == 4                                  # 0.995 >= 0.99 is not printed
    assert cli.format_answers(0, ("/q/y.fna", "", 1), [(0, 0.999)], seqdict) == ""


def test_parser_mirrors_reference_flags():
    a = cli.build_parser().parse_args("--pio 2000 --nbthreads 24 tohnsw -d db -k 21 -s 18000 -n 128 --ef 1600 "
                                      "--algo prob --scale_modify_f 0.25".split())
    assert (a.pio, a.dir, a.kmer, a.sketch, a.nbng, a.ef, a.algo, a.scale_modify_f, a.aa, a.block) == \
        (2000, "db", 21, 18000, 128, 1600, "prob", 0.25, False, False)
    a = cli.build_parser().parse_args("tohnsw -d db -k 7 -s 12000 -n 128 --algo optdens --aa --block".split())
    assert a.ef == 400 and a.aa and a.block                 # --ef defaults to 400 (gsearch.rs:218)
    a = cli.build_parser().parse_args("request -b dbdir -r qdir -n 50".split())
    assert (a.hnsw, a.query, a.nbanswers) == ("dbdir", "qdir", 50)
    a = cli.build_parser().parse_args("add -b dbdir -n newdir".split())
    assert (a.hnsw, a.new) == ("dbdir", "newdir")


def test_matches_file_and_reformat(tmp_path):
    """gsearch.matches (Matcher::analyze, src/matcher.rs:233-277) and `reformat` (src/bin/reformat.rs):
    text formats, merit = product of the distances below the threshold, ANI models, the column-7 quirk"""
    import math
    seqdict = [("/db/x.fna", "x", 1000), ("/db/y.fna", "y", 2000), ("/db/z.fna", "z", 3000)]
    q = ("/q/a.fna", "a", 49950)
    nbrs = [(0, 0.0), (1, 0.9453125), (2, 0.995)]
    txt = cli.format_matches([(q[0], [(seqdict[i][0], d) for i, d in nbrs]), ("/q/b.fna", [("/db/y.fna", 0.5)])])
    assert txt == ("\n\n request genome : /q/a.fna\n\t matched genome /db/x.fna  merit : 0.000E0"
                   "\n\t matched genome /db/y.fna  merit : 9.453E-1\n\t matched genome /db/z.fna  merit : 1.000E0"
                   "\n\n request genome : /q/b.fna\n\t matched genome /db/y.fna  merit : 5.000E-1")
    many = [(f"/db/g{i}.fna", 0.1 + 0.1 * i) for i in range(8)]
    assert cli.format_matches([("/q/c.fna", many)]).count("matched genome") == 5       # at most five per request
    # reformat reads what format_answers wrote
    (tmp_path / "n.txt").write_text(cli.format_answers(0, q, nbrs, seqdict) + cli.format_answers(1, ("/q/0.fna", "0", 5), [(1, 0.25)], seqdict))
    for model in (1, 2, 3):
        cli.main(f"reformat 16 {model} {tmp_path / 'n.txt'} {tmp_path / 'o.tsv'}".split())
        rows = (tmp_path / "o.tsv").read_text().splitlines()
        assert rows[0] == "Query_Name\tDistance\tNeighbor_Fasta_name\tNeighbor_Seq_Len\tANI"
        assert [r.split("\t")[0] for r in rows[1:]] == ["0.fna", "a.fna", "a.fna"]       # sorted by query, then distance
        name, d, nb, col7, ani = rows[3].split("\t")
        assert (name, d, nb, col7) == ("a.fna", "0.945312", "y.fna", " answer_seq_len:")  # (0.995 is above the threshold)
        j = 1.0 - 0.945312
        frac = j * 2.0 / (j + 1.0)
        want = {1: (1.0 + math.log(frac) / 16) * 100.0, 2: math.pow(frac, 1.0 / 16) * 100.0}.get(model)
        assert ani == ("Invalid Model" if want is None else cli.rust_f64(want))
        assert rows[2].split("\t")[1] == "0" and rows[2].split("\t")[4] == ("100" if model in (1, 2) else "Invalid Model")
    assert cli.rust_f64(0.00005) == "0.00005" and cli.rust_f64(float("nan")) == "NaN" and cli.rust_f64(1e21) == "1000000000000000000000"


@pytest.mark.gpu
def test_tohnsw_request_add_end_to_end(tmp_path, monkeypatch):
    db, new, qd, work = tmp_path / "db", tmp_path / "new", tmp_path / "q", tmp_path / "work"
    for d in (db, new, qd, work):
        d.mkdir()
    for i in range(24):
        data = g.synth.dna_genome(i, 60_000)
        if i % 3 == 0:
            (db / f"g{i:03d}.fna.gz").write_bytes(gzip.compress(data))
        else:
            (db / f"g{i:03d}.fna").write_bytes(data)
    for i in range(24, 28):
        (new / f"g{i:03d}.fa").write_bytes(g.synth.dna_genome(i, 60_000))
    (qd / "q0.fna").write_bytes(g.synth.dna_genome(5, 60_000))      # identical to a database genome
    (qd / "q1.fna").write_bytes(g.synth.dna_genome(25, 60_000))     # only in the database after `add`
    monkeypatch.chdir(work)
    cli.main("tohnsw -d {} -k 16 -s 512 -n 16 --ef 64 --algo prob".format(db).split())
    for f in ("hnswdump.hnsw.graph", "hnswdump.hnsw.data", "seqdict.json", "parameters.json",
              "processing_state.json"):
        assert (work / f).exists(), f
    assert len(cli.reload_seqdict(work)) == 24
    cli.main("request -b {} -r {} -n 5".format(work, qd).split())
    txt = open(work / "gsearch.neighbors.txt").read()
    assert "distance:\t0.00000E0\tanswer_fasta_path\t{}".format(db / "g005.fna") in txt
    assert str(new / "g025.fa") not in txt
    mtxt = open(work / "gsearch.matches").read()
    assert "\n\n request genome : {}\n\t matched genome {}  merit : 0.000E0".format(qd / "q0.fna", db / "g005.fna") in mtxt
    cli.main("add -b {} -n {}".format(work, new).split())
    assert len(cli.reload_seqdict(work)) == 28
    assert json.load(open(work / "processing_state.json"))["nb_seq"] == 28
    cli.main("request -b {} -r {} -n 5".format(work, qd).split())
    txt = open(work / "gsearch.neighbors.txt").read()
    assert "distance:\t0.00000E0\tanswer_fasta_path\t{}".format(new / "g025.fa") in txt


@pytest.mark.gpu
def test_tohnsw_request_aa_optdens(tmp_path, monkeypatch):
    db, qd, work = tmp_path / "db", tmp_path / "q", tmp_path / "work"
    for d in (db, qd, work):
        d.mkdir()
    for i in range(12):
        (db / f"p{i:03d}.faa").write_bytes(g.synth.aa_proteome(i, 60, 200))
    (qd / "q.faa.gz").write_bytes(gzip.compress(g.synth.aa_proteome(7, 60, 200)))
    (qd / "ignored.fna").write_bytes(g.synth.dna_genome(1, 5000))      # not an AA suffix
    monkeypatch.chdir(work)
    cli.main("tohnsw -d {} -k 7 -s 400 -n 8 --ef 40 --algo optdens --aa --scale_modify_f 0.25".format(db).split())
    doc = json.load(open(work / "parameters.json"))
    assert doc["sketch"] == {"kmer_size": 7, "sketch_size": 400, "algo": "OPTDENS", "data_t": "AA"}
    assert doc["hnsw"]["scale_modification"] == 0.25
    cli.main("request -b {} -r {} -n 3".format(work, qd).split())
    txt = open(work / "gsearch.neighbors.txt").read()
    assert txt.count("query_id:") >= 1 and "ignored" not in txt
    assert "distance:\t0.00000E0\tanswer_fasta_path\t{}".format(db / "p007.faa") in txt


@pytest.mark.gpu
@pytest.mark.parametrize("algo,json_name", [("hll", "HLL"), ("super2", "SUPER2"), ("revoptdens", "REVOPTDENS")])
def test_tohnsw_request_other_algos_through_hnswio_files(tmp_path, monkeypatch, algo, json_name):
    """the remaining --algo values end to end, with the database written in the hnswio-style layout
    (GSB_DUMP_HNSWIO=1) and reloaded from it by `request`; one query is a FASTQ copy of a database genome"""
    from gsearch_b200 import hnswio
    db, qd, work = tmp_path / "db", tmp_path / "q", tmp_path / "work"
    for d in (db, qd, work):
        d.mkdir()
    for i in range(16):
        (db / f"g{i:03d}.fna").write_bytes(g.synth.dna_genome(100 + i, 50_000))
    recs = g.synth.dna_genome(107, 50_000).split(b">")[1:]
    fq = b"".join(b"@" + r.partition(b"\n")[0] + b"\n" + r.partition(b"\n")[2].replace(b"\n", b"") + b"\n+\n" +
                  b"I" * len(r.partition(b"\n")[2].replace(b"\n", b"")) + b"\n" for r in recs)
    (qd / "q.fna").write_bytes(fq)   # FASTQ content under a suffix the directory walk accepts (src/utils/files.rs:116-137)
    monkeypatch.chdir(work)
    monkeypatch.setenv("GSB_DUMP_HNSWIO", "1")
    cli.main("tohnsw -d {} -k 21 -s 384 -n 12 --ef 48 --algo {}".format(db, algo).split())
    assert hnswio.is_hnswio(str(work))
    assert json.load(open(work / "parameters.json"))["sketch"]["algo"] == json_name
    cli.main("request -b {} -r {} -n 4".format(work, qd).split())
    txt = open(work / "gsearch.neighbors.txt").read()
    assert "distance:\t0.00000E0\tanswer_fasta_path\t{}".format(db / "g007.fna") in txt


@pytest.mark.gpu
def test_bindash_all_pairs(tmp_path, oracle):
    """`gsearch bindash` (src/bin/bindash.rs): all query x reference distances from OptDens / RevOptDens
    sketches; rows `Query<TAB>Reference<TAB>Distance` with 1 - (2J/(1+J))^(1/k), 0 for equal basenames"""
    files = []
    for i in range(5):
        pth = tmp_path / f"g{i}.fna"
        pth.write_bytes(g.synth.dna_genome(16 + i, 40_000))
        files.append(str(pth))
    ql, rl, out = tmp_path / "q.txt", tmp_path / "r.txt", tmp_path / "out.tsv"
    ql.write_text("\n".join(files[:2]) + "\n")
    rl.write_text("\n".join(files) + "\n")
    for dens, algo in ((0, g.ALGO_OPTDENS), (1, g.ALGO_REVOPTDENS)):
        cli.main(f"bindash -q {ql} -r {rl} -k 16 -s 1024 -d {dens} -o {out}".split())
        rows = open(out).read().splitlines()
        assert rows[0] == "Query\tReference\tDistance" and len(rows) == 1 + 2 * 5
        sig, _ = oracle.sketch_files([open(f, "rb").read() for f in files], 16, 1024, algo)
        for row in rows[1:]:
            qp, rp, d = row.split("\t")
            i, k = files.index(qp), files.index(rp)
            if i == k:
                assert float(d) == 0.0
                continue
            j = np.float32(1.0) - np.float32(oracle.hamming(sig[i], sig[k]))
            want = 1.0 - float(np.power((np.float32(2.0) * j) / (np.float32(1.0) + j), np.float32(1.0 / 16), dtype=np.float32))
            assert abs(float(d) - want) < 1e-6
