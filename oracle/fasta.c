/*
 * fasta.c -- FASTA bytes -> encoded sequences (test infrastructure, see gso.h).
 *
 * Follows the in-tree file tasks of the reference:
 *   DNA seq mode   src/dna/dnafiles.rs:43-107  (process_file_by_sequence) and
 *                  :115-195 (process_buffer_by_sequence)
 *   DNA block mode src/dna/dnafiles.rs:200-276, :283-360 (..._in_one_block)
 *   AA  seq mode   src/aa/aafiles.rs:107-160,165-229 ; block mode :33-99
 * plus needletail's record rules [U, high]: the first byte of the file picks the format
 * (parse_fastx_file, src/dna/dnafiles.rs:52,128,230): '>' FASTA, '@' FASTQ.
 *   FASTA  a record starts at a '>' that is the first byte of a line; its id is the rest of that
 *          line; its sequence is every following line up to the next record start, with line
 *          terminators removed.
 *   FASTQ  needletail's reader handles four-line records only ("Only supports single-line FASTQ
 *          files" [U, medium-high]): '@' + id, the sequence on ONE line, a line starting with '+',
 *          the qualities on one line.  A line that should start with '@' or '+' and does not is
 *          an error (the reference exits: dnafiles.rs:54-57); trailing blank lines end the input.
 *
 * Encoding [U, high]:
 *   DNA  kmerutils Alphabet2b: A,C,G,T -> 0,1,2,3 ; lower case accepted ; every other
 *        byte is skipped by Sequence::encode_and_add (dnafiles.rs:41,70-72)
 *   AA   kmerutils aautils Alphabet "ACDEFGHIKLMNPQRSTVWY" -> 1..20 ; upper case only ;
 *        invalid letters filtered (aafiles.rs:11-28,192)
 */
#include "gso.h"

#include <stdlib.h>
#include <string.h>

static int8_t dna_code[256];
static int8_t aa_code[256];
static int tables_ready = 0;

/* built once at load time: the per-file driver parses from several threads */
__attribute__((constructor)) static void init_tables(void) {
    if (tables_ready) return;
    memset(dna_code, -1, sizeof dna_code);
    memset(aa_code, -1, sizeof aa_code);
    dna_code['A'] = dna_code['a'] = 0;
    dna_code['C'] = dna_code['c'] = 1;
    dna_code['G'] = dna_code['g'] = 2;
    dna_code['T'] = dna_code['t'] = 3;
    const char *aa = "ACDEFGHIKLMNPQRSTVWY";
    for (int i = 0; aa[i]; i++) aa_code[(unsigned char)aa[i]] = (int8_t)(i + 1);
    tables_ready = 1;
}

static int contains_capsid(const uint8_t *p, uint64_t n) {
    static const char pat[] = "capsid";
    if (n < 6) return 0;
    for (uint64_t i = 0; i + 6 <= n; i++)
        if (memcmp(p + i, pat, 6) == 0) return 1;
    return 0;
}

void gso_seqs_free(gso_seqs *s) {
    free(s->codes);
    free(s->seq_off);
    memset(s, 0, sizeof *s);
}

int gso_parse_fasta(const uint8_t *bytes, uint64_t len, uint32_t data_t, uint32_t block_flag,
                    gso_seqs *out) {
    init_tables();
    const int8_t *tab = (data_t == GSO_DATA_AA) ? aa_code : dna_code;
    memset(out, 0, sizeof *out);
    out->codes = (uint8_t *)malloc(len ? len : 1);
    uint64_t cap_seq = 16;
    out->seq_off = (uint64_t *)malloc((cap_seq + 1) * sizeof(uint64_t));
    if (!out->codes || !out->seq_off) return 4;
    out->seq_off[0] = 0;
    uint64_t ncodes = 0;
    if (len > 0 && bytes[0] == '@') {
        /* ---- FASTQ: records of exactly four lines */
        uint64_t i = 0;
        while (i < len) {
            uint64_t e[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}; /* [b, e) of the four lines, terminators excluded */
            int nl = 0;
            uint64_t j = i;
            {   /* only line terminators left: end of input */
                uint64_t q = i;
                while (q < len && (bytes[q] == '\n' || bytes[q] == '\r')) q++;
                if (q == len) break;
            }
            while (nl < 4 && j <= len) {
                b[nl] = j;
                while (j < len && bytes[j] != '\n') j++;
                e[nl] = j;
                if (e[nl] > b[nl] && bytes[e[nl] - 1] == '\r') e[nl]--;
                nl++;
                if (j >= len) break;
                j++; /* past the newline */
            }
            if (bytes[b[0]] != '@' || e[0] == b[0]) {
                gso_seqs_free(out);
                return 5; /* InvalidStart */
            }
            if (nl < 3 || e[2] == b[2] || bytes[b[2]] != '+') {
                gso_seqs_free(out);
                return 6; /* InvalidSeparator / truncated record */
            }
            int dropped = contains_capsid(bytes + b[0] + 1, e[0] - b[0] - 1);
            uint64_t start_codes = ncodes, raw = e[1] - b[1];
            if (!dropped)
                for (uint64_t q = b[1]; q < e[1]; q++)
                    if (tab[bytes[q]] >= 0) out->codes[ncodes++] = (uint8_t)tab[bytes[q]];
            if (!dropped) out->nb_raw += raw;
            if (!block_flag && !dropped && ncodes > start_codes) {
                if (out->nseq == cap_seq) {
                    cap_seq *= 2;
                    uint64_t *n2 = (uint64_t *)realloc(out->seq_off, (cap_seq + 1) * sizeof(uint64_t));
                    if (!n2) {
                        gso_seqs_free(out);
                        return 4;
                    }
                    out->seq_off = n2;
                }
                out->nseq++;
                out->seq_off[out->nseq] = ncodes;
            }
            i = j;
            if (nl < 4) break; /* the qualities of the last record may be missing a terminator (or the line) */
        }
        if (block_flag) {
            out->nseq = 1;
            out->seq_off[1] = ncodes;
        }
        return 0;
    }
    if (len > 0 && bytes[0] != '>') {
        gso_seqs_free(out);
        return 5; /* needletail: InvalidStart */
    }
    uint64_t i = 0;
    while (i < len) {
        /* i is at a record start ('>') */
        uint64_t h0 = i + 1, h1 = h0;
        while (h1 < len && bytes[h1] != '\n') h1++;
        int dropped = contains_capsid(bytes + h0, h1 - h0);
        uint64_t j = (h1 < len) ? h1 + 1 : len;
        uint64_t start_codes = ncodes, raw = 0;
        /* sequence lines up to the next '>' at a line start */
        while (j < len) {
            if (bytes[j] == '>' && bytes[j - 1] == '\n') break;
            uint8_t c = bytes[j];
            if (c != '\n' && c != '\r') raw++;
            if (!dropped && tab[c] >= 0) out->codes[ncodes++] = (uint8_t)tab[c];
            j++;
        }
        if (!dropped) out->nb_raw += raw;
        if (!block_flag) {
            /* one Sequence per record with raw len > 0, kept only if it encodes to > 0 */
            if (!dropped && ncodes > start_codes) {
                if (out->nseq == cap_seq) {
                    cap_seq *= 2;
                    uint64_t *n2 =
                        (uint64_t *)realloc(out->seq_off, (cap_seq + 1) * sizeof(uint64_t));
                    if (!n2) {
                        gso_seqs_free(out);
                        return 4;
                    }
                    out->seq_off = n2;
                }
                out->nseq++;
                out->seq_off[out->nseq] = ncodes;
            }
        }
        i = j;
    }
    if (block_flag) {
        /* the whole file is one sequence (possibly empty) */
        out->nseq = 1;
        out->seq_off[1] = ncodes;
    }
    return 0;
}
