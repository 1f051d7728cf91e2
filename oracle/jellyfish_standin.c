/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked or executed by the product path.
 *
 * Stand-in for the external Jellyfish 2.x binary the reference shells out to
 * (reference: src/kmercounting.sh:8,11, called from src/main.c:70).  Jellyfish is a
 * third-party dependency that is not vendored under the reference tree (version unpinned,
 * README.md:18) and is not installed in this image.  Its arithmetic for this call is exact
 * counting of every forward-strand K-mer that lies inside one FASTA record (no -C), so the
 * stand-in is unambiguous:
 *
 *   jellyfish count -m K -o OUTBIN [-c .. -s .. -t ..] SRC   -> writes sorted (kmer,count)
 *                                                                 records to OUTBIN (binary)
 *   jellyfish dump -c -t -o OUTTXT OUTBIN                     -> writes "KMER count\n" text and
 *                                                                 leaves OUTBIN in place
 *
 * Counting is sort based (LSD radix sort of 2-bit packed K-mers), single threaded.
 * Records shorter than K contribute nothing.  Non-ACGT symbols break the current window,
 * which is what Jellyfish does.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <zlib.h>

static void die(const char *m) { fprintf(stderr, "jellyfish-standin: %s\n", m); exit(1); }

static void radix_sort_u64(uint64_t *a, uint64_t *tmp, size_t n, int bits)
{
    /* only the low `bits` bits vary (keys are right aligned here) */
    for (int shift = 0; shift < bits; shift += 11) {
        size_t hist[2049];
        memset(hist, 0, sizeof hist);
        for (size_t i = 0; i < n; i++) hist[((a[i] >> shift) & 2047) + 1]++;
        for (int i = 0; i < 2048; i++) hist[i + 1] += hist[i];
        for (size_t i = 0; i < n; i++) tmp[hist[(a[i] >> shift) & 2047]++] = a[i];
        uint64_t *t = a; a = tmp; tmp = t;
    }
    /* result may sit in either buffer; caller handles via pass count parity */
}

static int do_count(int argc, char **argv)
{
    int K = 0; const char *out = NULL; const char *src = NULL;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-m") && i + 1 < argc) K = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if ((!strcmp(argv[i], "-c") || !strcmp(argv[i], "-s") || !strcmp(argv[i], "-t")) && i + 1 < argc) ++i;
        else if (argv[i][0] != '-') src = argv[i];
    }
    if (K < 1 || K > 32 || !out || !src) die("usage: count -m K -o OUT SRC");
    gzFile fp = gzopen(src, "r");
    if (!fp) die("cannot open source");
    size_t cap = 1 << 20, n = 0;
    uint64_t *keys = malloc(cap * sizeof *keys);
    if (!keys) die("oom");
    const uint64_t mask = (K == 32) ? ~(uint64_t)0 : (((uint64_t)1 << (2 * K)) - 1);
    static char buf[1 << 16];
    int got, in_header = 0, at_line_start = 1, fastq = 0, fq_state = 0;
    uint64_t cur = 0; int run = 0;
    while ((got = gzread(fp, buf, sizeof buf)) > 0) {
        for (int i = 0; i < got; i++) {
            char c = buf[i];
            if (c == '\n') { in_header = 0; at_line_start = 1; if (fastq && fq_state == 2) fq_state = 3; else if (fastq && fq_state == 3) fq_state = 0; continue; }
            if (c == '\r') continue;
            if (at_line_start) {
                at_line_start = 0;
                if (c == '>' ) { in_header = 1; run = 0; continue; }
                if (c == '@' && (fq_state == 0)) { fastq = 1; fq_state = 1; in_header = 1; run = 0; continue; }
                if (fastq && c == '+' && fq_state == 1) { fq_state = 2; in_header = 1; continue; }
            }
            if (in_header) continue;
            if (fastq && fq_state == 3) continue; /* quality line */
            int v;
            switch (c) { case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break;
                         case 'G': case 'g': v = 2; break; case 'T': case 't': v = 3; break; default: v = -1; }
            if (v < 0) { run = 0; continue; }
            cur = ((cur << 2) | (uint64_t)v) & mask;
            if (++run >= K) {
                if (n == cap) { cap *= 2; keys = realloc(keys, cap * sizeof *keys); if (!keys) die("oom"); }
                keys[n++] = cur;
            }
        }
    }
    gzclose(fp);
    uint64_t *tmp = malloc((n ? n : 1) * sizeof *tmp);
    if (!tmp) die("oom");
    int bits = 2 * K, passes = (bits + 10) / 11;
    radix_sort_u64(keys, tmp, n, bits);
    uint64_t *sorted = (passes & 1) ? tmp : keys;
    FILE *fo = fopen(out, "wb");
    if (!fo) die("cannot create output");
    uint64_t hdr[2] = { (uint64_t)K, 0 };
    fwrite(hdr, 8, 2, fo);
    uint64_t rec[2], distinct = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && sorted[j] == sorted[i]) j++;
        rec[0] = sorted[i]; rec[1] = j - i;
        fwrite(rec, 8, 2, fo);
        distinct++; i = j;
    }
    fclose(fo);
    free(keys); free(tmp);
    fprintf(stderr, "jellyfish-standin: %zu %d-mers, %llu distinct\n", n, K, (unsigned long long)distinct);
    return 0;
}

static int do_dump(int argc, char **argv)
{
    const char *out = NULL, *in = NULL;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (argv[i][0] != '-') in = argv[i];
    }
    if (!out || !in) die("usage: dump -c -t -o OUT IN");
    FILE *fi = fopen(in, "rb"), *fo = fopen(out, "w");
    if (!fi || !fo) die("dump: cannot open files");
    uint64_t hdr[2];
    if (fread(hdr, 8, 2, fi) != 2) die("dump: bad input");
    int K = (int)hdr[0];
    static char obuf[1 << 20];
    setvbuf(fo, obuf, _IOFBF, sizeof obuf);
    uint64_t rec[2]; char line[80];
    while (fread(rec, 8, 2, fi) == 2) {
        for (int i = 0; i < K; i++) line[i] = "ACGT"[(rec[0] >> (2 * (K - 1 - i))) & 3];
        int l = K + sprintf(line + K, "\t%llu\n", (unsigned long long)rec[1]);
        fwrite(line, 1, l, fo);
    }
    fclose(fi); fclose(fo);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) die("need a subcommand (count|dump)");
    if (!strcmp(argv[1], "count")) return do_count(argc, argv);
    if (!strcmp(argv[1], "dump")) return do_dump(argc, argv);
    die("unknown subcommand");
    return 1;
}
