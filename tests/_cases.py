"""Shared input builders for the parity tests."""
import numpy as np

import gsearch_b200 as g

AA = "ACDEFGHIKLMNPQRSTVWY"


def rand_seq(rng, L, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), L))


def fasta(records, width=80, crlf=False):
    nl = "\r\n" if crlf else "\n"
    out = []
    for hdr, seq in records:
        out.append(">" + hdr + nl)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + nl)
    return "".join(out).encode()


def adversarial_dna_files(seed=0):
    """FASTA files that exercise every rule of the K1 state machine, including events that
    straddle the 4 KiB tile and 16-byte thread boundaries."""
    rng = np.random.default_rng(seed)
    files = []
    files.append(b"")                                                    # empty file
    files.append(b">only a header")                                      # no sequence, no newline
    files.append(b">h\nACG\n")                                           # shorter than k
    files.append(fasta([("one", rand_seq(rng, 5000))]))
    files.append(fasta([("crlf", rand_seq(rng, 3000))], crlf=True))
    files.append(fasta([("noeol", rand_seq(rng, 777))]).rstrip(b"\n"))   # no trailing newline
    files.append(fasta([("oneline", rand_seq(rng, 20000))], width=10**9))  # single 20 kb line
    # many short records, some dropped ("capsid"), some empty, lower case, N runs, '>' inside lines
    recs = []
    for r in range(120):
        L = int(rng.integers(0, 300))
        s = rand_seq(rng, L, "ACGTacgtNRY")
        if r % 7 == 3:
            s = s[: L // 2] + ">" + s[L // 2:]
        hdr = f"rec{r} " + ("major capsid protein" if r % 5 == 2 else "contig")
        recs.append((hdr, s))
    files.append(fasta(recs, width=70))
    # a header longer than a tile, with "capsid" placed across the tile boundary
    for shift in range(4090, 4102, 3):
        hdr = "x" * (shift - 1 - 6) + "capsid" + "y" * 50
        files.append(fasta([(hdr, rand_seq(rng, 500)), ("keep", rand_seq(rng, 600))]))
    # '\n>' split across tile boundaries: records of sizes that sweep the boundary
    for pad in range(4085, 4100):
        body = rand_seq(rng, pad - 4)
        files.append((">a\n" + body + "\n>b\n" + rand_seq(rng, 300) + "\n").encode())
    # "capsid" text inside a sequence line is data, not a header
    files.append(b">c\nACGTcapsidACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n>d capsid\nACGTACGTACGTACGTACGTACGT\n")
    # synthetic genomes of the benchmark generator (families, repeats, N runs, lower case)
    for i in range(4):
        files.append(g.synth.dna_genome(i, 60000 + 1000 * i, ncontigs=1 + i))
    return files


def adversarial_aa_files(seed=0):
    rng = np.random.default_rng(seed)
    files = [b"", b">p\nMK\n"]
    recs = []
    for r in range(200):
        L = int(rng.integers(0, 400))
        s = rand_seq(rng, L, AA + "XBZ*acd")
        recs.append((f"prot{r} " + ("capsid" if r % 9 == 4 else "hypothetical"), s + "*"))
    files.append(fasta(recs, width=60))
    for i in range(3):
        files.append(g.synth.aa_proteome(i, 150, 200))
    return files
