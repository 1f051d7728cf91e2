/*
 * ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of what the reference deBWT pipeline outputs, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.  The product
 * (debwt_b200/, libdebwt_b200.so) never links, loads or executes this file.
 *
 * Parity pin: tests/test_oracle.py checks oracle_bwt() against the golden vector the reference
 * produced and against committed outputs of the compiled reference
 * (tests/golden/, oracle/_ref/deBWT -t 1).
 *
 * What is restated (reference file:line):
 *   - the text T = S1 # S2 # ... Sn $ and BWTLEN = sum(len) + n     src/collect#$.c:56-90
 *   - symbol order A<C<G<T<#<$, equal '#' compared through, '$' largest
 *                                                                   src/collect#$.c:253-311 (cmp)
 *   - BWT[row] = symbol before the row's suffix, '$' for suffix 0    src/generateSP.c:584-605
 *   - output packing: 32 symbols per u64, symbol j at bits 2*(31-(j&31)), '#'/'$' stored as T,
 *     '#' rows ascending, '$' row                                    src/insertCase3.c:84-95,115-131
 *   - k-mer sort / count-by-sort ({kmer,count} records)              src/mySort.c:61-77,98-195
 * The BWT itself is obtained here by direct suffix sorting (the closed form the survey verified the
 * reference against), NOT by the de Bruijn-branch route: an independent check of the CUDA path.
 * The stage-by-stage restatement of the de Bruijn-branch route lives in oracle/stages.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- LSD radix sort of (key, idx) pairs by key ------------------------------------------- */
typedef struct { uint64_t key; uint64_t idx; } pair_t;

static void radix_pairs(pair_t *a, pair_t *tmp, size_t n)
{
    for (int shift = 0; shift < 64; shift += 16) {
        size_t *hist = calloc(65537, sizeof(size_t));
        for (size_t i = 0; i < n; i++) hist[((a[i].key >> shift) & 0xffff) + 1]++;
        for (int i = 0; i < 65536; i++) hist[i + 1] += hist[i];
        for (size_t i = 0; i < n; i++) tmp[hist[(a[i].key >> shift) & 0xffff]++] = a[i];
        free(hist);
        pair_t *t = a; a = tmp; tmp = t;
    }
    /* 4 passes: result is back in the original `a` */
}

/* ---- suffix comparison: 8 symbols per step on the byte text ------------------------------ */
static const uint8_t *g_sym;
static size_t g_n;

static inline uint64_t load_be(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return __builtin_bswap64(v);
}

static int cmp_suffix_from(size_t a, size_t b)
{
    /* text is padded with >= 8 zero bytes; '$' (5) is unique so distinct suffixes always differ
       at or before it */
    for (;;) {
        uint64_t x = load_be(g_sym + a), y = load_be(g_sym + b);
        if (x != y) return x < y ? -1 : 1;
        a += 8; b += 8;
        if (a >= g_n || b >= g_n) return a > b ? -1 : (a < b ? 1 : 0);
    }
}

#define PRE 21  /* symbols folded into the 63-bit radix key (3 bits each) */

static int cmp_idx(const void *pa, const void *pb)
{
    size_t a = ((const pair_t *)pa)->idx, b = ((const pair_t *)pb)->idx;
    if (a == b) return 0;
    return cmp_suffix_from(a + PRE, b + PRE);
}

/*
 * sym: n symbol codes (0..3 bases, 4 '#', 5 '$'; '$' last).  bwt_out: n symbol codes in row order.
 * Returns 0, or -1 on allocation failure.
 */
int oracle_bwt(const uint8_t *sym, uint64_t n, uint8_t *bwt_out)
{
    uint8_t *pad = calloc(n + PRE + 16, 1);
    pair_t *a = malloc((n ? n : 1) * sizeof *a), *tmp = malloc((n ? n : 1) * sizeof *tmp);
    if (!pad || !a || !tmp) { free(pad); free(a); free(tmp); return -1; }
    memcpy(pad, sym, n);
    /* rolling 21-symbol key, 3 bits per symbol; symbols past the end read as 0 which cannot
       create ties between distinct suffixes because '$' is unique */
    uint64_t key = 0;
    for (int i = 0; i < PRE; i++) key = (key << 3) | pad[i];
    for (uint64_t i = 0; i < n; i++) {
        a[i].key = key; a[i].idx = i;
        key = ((key << 3) | pad[i + PRE]) & (((uint64_t)1 << (3 * PRE)) - 1);
    }
    radix_pairs(a, tmp, n);
    g_sym = pad; g_n = n;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && a[j].key == a[i].key) j++;
        if (j - i > 1) qsort(a + i, j - i, sizeof *a, cmp_idx);
        i = j;
    }
    for (uint64_t i = 0; i < n; i++) bwt_out[i] = a[i].idx ? sym[a[i].idx - 1] : sym[n - 1];
    free(pad); free(a); free(tmp);
    return 0;
}

/* Output packing (src/insertCase3.c:84-95,115-131).  words: ceil(n/32) u64 (caller zeroes),
   sharp: capacity >= number of '#' rows.  Returns the number of '#' rows. */
uint64_t oracle_pack_bwt(const uint8_t *bwt, uint64_t n, uint64_t *words, uint64_t *sharp, uint64_t *dollar)
{
    uint64_t ns = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t c = bwt[i];
        if (c == 4) sharp[ns++] = i;
        if (c == 5) *dollar = i;
        if (c > 3) c = 3;
        words[i >> 5] |= c << (2 * (31 - (i & 31)));
    }
    return ns;
}

/* 2-bit packing of the text (src/collect#$.c:78,82,87-90): separators as T plus 32 T of padding.
   words: ceil((n+32)/32) u64, caller zeroes. */
void oracle_pack_text(const uint8_t *sym, uint64_t n, uint64_t *words)
{
    for (uint64_t i = 0; i < n + 32; i++) {
        uint64_t c = (i < n) ? sym[i] : 3;
        if (c > 3) c = 3;
        words[i >> 5] |= c << (2 * (31 - (i & 31)));
    }
}

/* All in-record 32-mers as left-aligned u64 keys in text order (src/kmercounting.sh:8 semantics,
   src/mySort.c:61-75 packing).  Returns the number of keys written. */
uint64_t oracle_extract_keys(const uint8_t *sym, uint64_t n, uint64_t *keys)
{
    uint64_t cur = 0, run = 0, m = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (sym[i] > 3) { run = 0; continue; }
        cur = (cur << 2) | sym[i];
        if (++run >= 32) keys[m++] = cur;
    }
    return m;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}

/* the reference sorts with qsort + cmpKmer (src/mySort.c:203-238,338-345) */
void oracle_sort_keys(uint64_t *keys, uint64_t n) { qsort(keys, n, 8, cmp_u64); }

/* count-by-sort -> {kmer,count} records like kmerInfo (src/mySort.c:76-77,194); returns D */
uint64_t oracle_rle(const uint64_t *sorted, uint64_t n, uint64_t *kmers, uint64_t *counts)
{
    uint64_t d = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && sorted[j] == sorted[i]) j++;
        kmers[d] = sorted[i]; counts[d] = j - i; d++; i = j;
    }
    return d;
}

/* Unpacks the reference's output format back into BWT symbols (0..3, 4 '#', 5 '$'): the inverse of
   oracle_pack_bwt / src/insertCase3.c:84-95,115-131. */
void oracle_unpack_bwt(const uint64_t *words, uint64_t n, const uint64_t *sharp, uint64_t n_sharp, uint64_t dollar, uint8_t *out)
{
    for (uint64_t i = 0; i < n; i++) out[i] = (uint8_t)((words[i >> 5] >> (2 * (31 - (i & 31)))) & 3);
    for (uint64_t i = 0; i < n_sharp; i++) out[sharp[i]] = 4;
    out[dollar] = 5;
}

/*
 * LF-walk inversion (idea of the reference's unreachable developer check, src/LFsearch.c:49-166,
 * findSeg :167-235): rebuilds T from the BWT symbols in one N-cycle.  Size independent verifier.
 * bwt: n symbols (0..5).  text_out: n symbols.  Returns 0 when the walk closes after exactly n
 * steps, 1 otherwise.
 *
 * All '#' are one symbol, so LF is the ordinary LF mapping over the 6-letter alphabet.
 */
int oracle_invert_bwt(const uint8_t *bwt, uint64_t n, uint8_t *text_out)
{
    uint64_t cnt[7] = {0}, C[6];
    for (uint64_t i = 0; i < n; i++) cnt[bwt[i]]++;
    C[0] = 0;
    for (int c = 1; c < 6; c++) C[c] = C[c - 1] + cnt[c - 1];
    uint32_t *lf = malloc((n ? n : 1) * sizeof *lf);
    if (!lf) return -1;
    uint64_t seen[6] = {0};
    for (uint64_t i = 0; i < n; i++) { int c = bwt[i]; lf[i] = (uint32_t)(C[c] + seen[c]++); }
    /* row of the suffix "$..." is the last row (unique largest symbol); text[n-1] = '$' */
    uint64_t row = n - 1;
    int ok = 0;
    for (uint64_t k = 0; k < n; k++) {
        /* suffix at `row` starts at text position n-1-k; its preceding symbol is bwt[row] */
        uint64_t pos = n - 1 - k;
        text_out[pos] = (k == 0) ? 5 : text_out[pos];
        uint8_t c = bwt[row];
        if (pos > 0) text_out[pos - 1] = c;
        else if (c != 5) ok = 1;
        row = lf[row];
    }
    if (row != n - 1) ok = 1;
    free(lf);
    return ok;
}
