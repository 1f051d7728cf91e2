/*
 * Minimal FASTA/FASTQ record reader over zlib (plain or gzip input) for the deBWT-B200 host.
 *
 * The reference reads its input with klib's kseq.h (reference src/collect#$.c:26,34-37:
 * kseq_init / kseq_read / kseq_destroy over gzFile).  This is an independent implementation with
 * the same observable behaviour for this path: records in file order, sequence lines concatenated,
 * header text and FASTQ quality lines skipped, 64-bit record lengths (kseq_read returns int, which
 * silently truncates the reference at 2^31 bp -- SURVEY.md appendix C).
 */
#ifndef DEBWT_FASTX_H
#define DEBWT_FASTX_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
    gzFile fp;
    unsigned char* buf;
    int begin, end, eof;
    int last;                 /* look-ahead header character ('>' or '@'), 0 when none */
    char* seq;                /* current record's bases */
    uint64_t len, cap;
} fastx_t;

#define FASTX_BUF (1 << 22)

static inline fastx_t* fastx_open(const char* path) {
    gzFile fp = gzopen(path, "r");
    if (!fp) return NULL;
    gzbuffer(fp, 1 << 20);
    fastx_t* f = (fastx_t*)calloc(1, sizeof(fastx_t));
    f->fp = fp;
    f->buf = (unsigned char*)malloc(FASTX_BUF);
    return f;
}

static inline void fastx_close(fastx_t* f) {
    if (!f) return;
    gzclose(f->fp);
    free(f->buf);
    free(f->seq);
    free(f);
}

static inline int fastx_getc(fastx_t* f) {
    if (f->begin >= f->end) {
        if (f->eof) return -1;
        f->begin = 0;
        f->end = gzread(f->fp, f->buf, FASTX_BUF);
        if (f->end <= 0) { f->eof = 1; f->end = 0; return -1; }
    }
    return f->buf[f->begin++];
}

static inline void fastx_skip_line(fastx_t* f) {
    int c;
    while ((c = fastx_getc(f)) >= 0 && c != '\n') {}
}

static inline int fastx_push(fastx_t* f, const unsigned char* p, uint64_t n) {
    if (f->len + n + 1 > f->cap) {
        uint64_t cap = f->cap ? f->cap : (1 << 16);
        while (cap < f->len + n + 1) cap += cap >> 1;
        char* s = (char*)realloc(f->seq, cap);
        if (!s) return -1;
        f->seq = s; f->cap = cap;
    }
    memcpy(f->seq + f->len, p, n);
    f->len += n;
    return 0;
}

/* Reads the next record into f->seq / f->len.  Returns 1 on success, 0 at end of file,
   -1 on a malformed file or allocation failure. */
static inline int fastx_read(fastx_t* f) {
    int c;
    if (f->last == 0) {
        while ((c = fastx_getc(f)) >= 0 && c != '>' && c != '@') {}
        if (c < 0) return 0;
        f->last = c;
    }
    const int fastq = (f->last == '@');
    f->last = 0;
    f->len = 0;
    fastx_skip_line(f);                                  /* header */
    for (;;) {                                           /* sequence lines */
        c = fastx_getc(f);
        if (c < 0) break;
        if (c == '>' || (c == '@' && !fastq) || c == '+') { if (c != '+') f->last = c; break; }
        if (c == '@' && fastq) { f->last = c; break; }
        /* copy the rest of this line in bulk */
        f->begin--;
        for (;;) {
            unsigned char* p = f->buf + f->begin;
            unsigned char* nl = (unsigned char*)memchr(p, '\n', (size_t)(f->end - f->begin));
            uint64_t n = nl ? (uint64_t)(nl - p) : (uint64_t)(f->end - f->begin);
            uint64_t m = n;
            while (m && (p[m - 1] == '\r' || p[m - 1] == ' ')) m--;
            if (fastx_push(f, p, m)) return -1;
            f->begin += (int)n + (nl ? 1 : 0);
            if (nl) break;
            if (fastx_getc(f) < 0) break;                /* refill */
            f->begin--;
        }
    }
    if (f->seq) f->seq[f->len] = 0;
    if (fastq && c == '+') {                             /* skip '+' line and the quality string */
        fastx_skip_line(f);
        uint64_t q = 0;
        while (q < f->len && (c = fastx_getc(f)) >= 0) if (c != '\n' && c != '\r') q++;
        fastx_skip_line(f);
    }
    return 1;
}

/* Streaming variant: the bases of every record are handed to `bases` piece by piece (no per-record buffer, no second
   copy), `record_end` is called after the last piece of a record with the record's length.  Either callback stops
   the stream by returning non-zero.  Returns 0 at end of file, -1 on a malformed file, else the callback's value. */
typedef int (*fastx_bases_fn)(void* user, const unsigned char* p, uint64_t n);
typedef int (*fastx_record_fn)(void* user, uint64_t len);

static inline int fastx_stream(fastx_t* f, fastx_bases_fn bases, fastx_record_fn record_end, void* user) {
    int c, rc;
    for (;;) {
        if (f->last == 0) {
            while ((c = fastx_getc(f)) >= 0 && c != '>' && c != '@') {}
            if (c < 0) return 0;
            f->last = c;
        }
        const int fastq = (f->last == '@');
        f->last = 0;
        uint64_t len = 0;
        fastx_skip_line(f);                                  /* header */
        for (;;) {                                           /* sequence lines */
            c = fastx_getc(f);
            if (c < 0) break;
            if (c == '>' || (c == '@' && !fastq) || c == '+') { if (c != '+') f->last = c; break; }
            if (c == '@' && fastq) { f->last = c; break; }
            f->begin--;
            for (;;) {
                unsigned char* p = f->buf + f->begin;
                unsigned char* nl = (unsigned char*)memchr(p, '\n', (size_t)(f->end - f->begin));
                uint64_t n = nl ? (uint64_t)(nl - p) : (uint64_t)(f->end - f->begin);
                uint64_t m = n;
                while (m && (p[m - 1] == '\r' || p[m - 1] == ' ')) m--;
                if (m && (rc = bases(user, p, m)) != 0) return rc;
                len += m;
                f->begin += (int)n + (nl ? 1 : 0);
                if (nl) break;
                if (fastx_getc(f) < 0) break;                /* refill */
                f->begin--;
            }
        }
        if (fastq && c == '+') {                             /* skip '+' line and the quality string */
            fastx_skip_line(f);
            uint64_t q = 0;
            while (q < len && (c = fastx_getc(f)) >= 0) if (c != '\n' && c != '\r') q++;
            fastx_skip_line(f);
        }
        if ((rc = record_end(user, len)) != 0) return rc;
    }
}

#endif /* DEBWT_FASTX_H */
