/*
 * deBWT (B200) -- host program with the reference's command line.
 *
 *   deBWT -o <out> [-t <threads>] [-k <12..32>] [-j <jellyfish_root>] [-g <gpu>] <input.fa[.gz]>
 *
 * Same contract as the reference driver (reference src/main.c:25-53,175-186): options are
 * "flag value" pairs, the input file comes last, exit code 0 on success and 1 on any error with a
 * message on stderr; the three output files are the reference's (src/insertCase3.c:115-131):
 *   <out>    ceil(N/32) little-endian u64, 32 BWT symbols per word, '#'/'$' stored as T
 *   <out>.#  rows holding '#', ascending
 *   <out>.$  row holding '$'
 * Differences: -j is accepted and ignored (Jellyfish is replaced by on-GPU count-by-sort), -t is
 * accepted and ignored (the work runs on the GPU), -g picks the CUDA device, -n <seed> resolves IUPAC ambiguity codes
 * reproducibly (the reference needs a separate otherTool/transferN pass, which is time-seeded), no temp files.
 *
 * The reference reads the input twice through kseq (src/collect#$.c:37-48, 66-86) before anything else starts.  Here the
 * file is read ONCE and streamed: the reader writes the bases straight into the library's pinned staging window
 * (debwt_ingest_reserve / _commit), every full window is copied to the GPU and 2-bit packed while the next one is being
 * parsed, and the CUDA context is created on a second thread while the first megabytes are read.  The result comes back
 * into page-locked memory.  All computation happens in libdebwt_b200.so (include/debwt_b200.h); there is no CPU fallback.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../include/debwt_b200.h"
#include "fastx.h"

static void usage(void) {
    fprintf(stderr, "usage:\n");
    fprintf(stderr, "deBWT [options] reference\n");
    fprintf(stderr, "Please make sure your sequence don't contain any uncertain characters like 'N'\n");
    fprintf(stderr, "options:\n");
    fprintf(stderr, "-o: output bwt file(binary)\n");
    fprintf(stderr, "-t (optional): accepted for compatibility (the build runs on the GPU)\n");
    fprintf(stderr, "-k (optional): k-mer length (from 12 to 32, default 32)\n");
    fprintf(stderr, "-j (optional): accepted for compatibility, ignored (no Jellyfish needed)\n");
    fprintf(stderr, "-g (optional): CUDA device ordinal (default 0), or a comma-separated list (e.g. 0,1,2,3): the build is sharded over those GPUs\n");
    fprintf(stderr, "-n (optional): seed; IUPAC ambiguity codes are replaced by a pseudo-random compatible base (what otherTool/transferN does, reproducibly)\n");
    fprintf(stderr, "reference: sequence in fasta or fastq format (plain or gzip)\n");
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int write_file(const char* path, const void* data, size_t bytes) {
    FILE* f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot create %s!\n", path); return -1; }
    size_t w = bytes ? fwrite(data, 1, bytes, f) : 0;
    if (fclose(f) != 0 || w != bytes) { fprintf(stderr, "short write on %s\n", path); return -1; }
    return 0;
}

/* ---- context creation on a second thread: CUDA start-up (context, kernel images) overlaps reading the file ---- */
typedef struct { int gpu; debwt_ctx* ctx; int rc; char err[256]; volatile int done; } create_job;
static void* create_thread(void* p) {
    create_job* j = (create_job*)p;
    j->rc = debwt_create(&j->ctx, j->gpu);
    if (j->rc) { strncpy(j->err, debwt_last_error(), sizeof j->err - 1); j->err[sizeof j->err - 1] = 0; }
    __sync_synchronize();
    j->done = 1;
    return NULL;
}

/* ---- sink of the streaming reader: T = S1 # S2 # ... Sn $ written into the library's staging window ---- */
typedef struct {
    debwt_ctx* ctx;
    char* win;                 /* current window */
    uint64_t cap, used;        /* its capacity / bytes written and not yet committed */
    uint64_t n;                /* symbols of T emitted so far */
    uint64_t* seps;
    uint64_t nrec, sep_cap;
    int pending_sep;           /* the separator after the last record is written once we know whether it is '#' or '$' */
    int failed;
    /* until the context exists the symbols are parked in ordinary memory */
    create_job* job;
    uint64_t hint;
    char* park;
    uint64_t park_n, park_cap;
    int dry;                   /* no device: only the input checks run */
    int collect;               /* several GPUs: the whole text is collected in host memory, every GPU uploads its slice */
    int resolve;               /* -n: IUPAC policy */
    unsigned long long amb_seed;
} sink_t;

/* the context has come up: open the streaming input and push what was parked */
static int sink_attach(sink_t* s) {
    if (s->job->rc) { s->dry = 1; free(s->park); s->park = NULL; return 0; }
    s->ctx = s->job->ctx;
    if (s->resolve && debwt_set_ambiguity_policy(s->ctx, 1, s->amb_seed)) return -1;
    if (debwt_ingest_begin(s->ctx, s->hint)) return -1;
    if (s->park_n && debwt_ingest_append(s->ctx, s->park, s->park_n)) return -1;
    free(s->park);
    s->park = NULL;
    return 0;
}

static int sink_room(sink_t* s) {
    if (s->used && debwt_ingest_commit(s->ctx, s->used)) return -1;
    s->used = 0;
    if (debwt_ingest_reserve(s->ctx, &s->win, &s->cap)) return -1;
    return 0;
}

static int sink_put(sink_t* s, const unsigned char* p, uint64_t n) {
    if (s->dry) { s->n += n; return 0; }
    if (!s->ctx) {
        if (!s->collect && s->job->done) {
            if (sink_attach(s)) { s->failed = 1; return 2; }
            return sink_put(s, p, n);
        }
        if (s->park_n + n > s->park_cap) {
            uint64_t cap = s->park_cap ? s->park_cap : (64u << 20);
            while (cap < s->park_n + n) cap += cap >> 1;
            char* q = (char*)realloc(s->park, cap);
            if (!q) { fprintf(stderr, "out of host memory\n"); s->failed = 1; return 3; }
            s->park = q; s->park_cap = cap;
        }
        memcpy(s->park + s->park_n, p, n);
        s->park_n += n; s->n += n;
        return 0;
    }
    while (n) {
        if (s->used == s->cap && sink_room(s)) { s->failed = 1; return 2; }
        uint64_t m = s->cap - s->used;
        if (m > n) m = n;
        memcpy(s->win + s->used, p, m);
        s->used += m; s->n += m; p += m; n -= m;
    }
    return 0;
}

static int on_bases(void* u, const unsigned char* p, uint64_t n) {
    sink_t* s = (sink_t*)u;
    if (s->pending_sep) {
        const unsigned char c = '#';
        s->pending_sep = 0;
        if (sink_put(s, &c, 1)) return 2;
    }
    return sink_put(s, p, n);
}

static int on_record(void* u, uint64_t len) {
    sink_t* s = (sink_t*)u;
    if (len <= 32) { fprintf(stderr, "Length <= 32!\n"); return 3; }                  /* src/collect#$.c:41-45 */
    if (s->nrec == s->sep_cap) {
        s->sep_cap = s->sep_cap ? s->sep_cap * 2 : 64;
        s->seps = (uint64_t*)realloc(s->seps, s->sep_cap * 8);
        if (!s->seps) { fprintf(stderr, "out of host memory\n"); return 3; }
    }
    s->seps[s->nrec++] = s->n;             /* position the separator will take */
    s->pending_sep = 1;
    return 0;
}

int main(int argc, char* argv[]) {
    if (argc < 4 || (argc & 1) == 1) { usage(); return 1; }
    const char* source = argv[argc - 1];
    const char* obj = NULL;
    int k = 32, gpu = 0, resolve = 0, ngpu = 1;
    int gpus[16] = {0};
    unsigned long long amb_seed = 0;
    for (int i = 1; i < argc - 1; i += 2) {
        if (strcmp(argv[i], "-o") == 0) obj = argv[i + 1];
        else if (strcmp(argv[i], "-t") == 0) {
            if (atoi(argv[i + 1]) == 0) { fprintf(stderr, "thread number must be a number!\n"); return 1; }
        } else if (strcmp(argv[i], "-j") == 0) { /* ignored */
        } else if (strcmp(argv[i], "-g") == 0) {
            const char* q = argv[i + 1];
            ngpu = 0;
            while (*q && ngpu < 16) {
                gpus[ngpu++] = atoi(q);
                while (*q && *q != ',') q++;
                if (*q == ',') q++;
            }
            if (ngpu == 0) { usage(); return 1; }
            gpu = gpus[0];
        }
        else if (strcmp(argv[i], "-n") == 0) { resolve = 1; amb_seed = strtoull(argv[i + 1], NULL, 10); }
        else if (strcmp(argv[i], "-k") == 0) {
            k = atoi(argv[i + 1]);
            if (k < 12 || k > 32) { fprintf(stderr, "-k: k-mer length (from 12 to 32, default 32)\n"); return 1; }
        } else { usage(); return 1; }
    }
    if (!obj) { usage(); return 1; }
    FILE* probe = fopen(obj, "wb");                        /* src/main.c:55-58 */
    if (!probe) { fprintf(stderr, "cannot create %s!\n", obj); return 1; }
    fclose(probe);
    remove(obj);

    double t0 = now_s();
    if (ngpu > 1) {
        /* ---- several GPUs: debwt_build_multi, one thread per GPU, each uploads its own position slice ---- */
        if (resolve) { fprintf(stderr, "-n is not available with several GPUs\n"); return 1; }
        fastx_t* fxm = fastx_open(source);
        if (!fxm) { fprintf(stderr, "can not open ref file\n"); return 1; }
        sink_t m;
        memset(&m, 0, sizeof m);
        m.collect = 1;
        int mrc = fastx_stream(fxm, on_bases, on_record, &m);
        fastx_close(fxm);
        if (mrc == 3) return 1;
        if (mrc < 0) { fprintf(stderr, "malformed input file\n"); return 1; }
        if (m.failed) return 1;
        if (m.nrec == 0) { fprintf(stderr, "no sequence found in %s\n", source); return 1; }
        { const unsigned char c = '$'; if (sink_put(&m, &c, 1)) return 1; }
        double t1m = now_s();
        const uint64_t nwords = (m.n + 31) / 32, nsharp = m.nrec - 1;
        uint64_t* bwt = (uint64_t*)malloc((nwords ? nwords : 1) * 8);
        uint64_t* sharp = (uint64_t*)malloc((nsharp ? nsharp : 1) * 8);
        uint64_t dollar = 0;
        debwt_shard_stats ss;
        if (!bwt || !sharp) { fprintf(stderr, "out of host memory\n"); return 1; }
        if (debwt_build_multi(gpus, ngpu, m.park, m.n, m.seps, m.nrec, bwt, sharp, &dollar, &ss)) {
            fprintf(stderr, "deBWT: %s\n", debwt_last_error());
            return 1;
        }
        double t2m = now_s();
        size_t olm = strlen(obj);
        char* pm = (char*)malloc(olm + 3);
        memcpy(pm, obj, olm);
        pm[olm] = '.'; pm[olm + 2] = 0;
        if (write_file(obj, bwt, nwords * 8)) return 1;
        pm[olm + 1] = '#';
        if (write_file(pm, sharp, nsharp * 8)) return 1;
        pm[olm + 1] = '$';
        if (write_file(pm, &dollar, 8)) return 1;
        fprintf(stderr, "BWTLEN=%llu (%llu records), read %.3f s; %d GPUs: upload + build + download %.3f s (rank 0 device %.1f ms, sort %.1f ms); write %.3f s\n",
                (unsigned long long)m.n, (unsigned long long)m.nrec, t1m - t0, ngpu, t2m - t1m, ss.ms_total, ss.ms_sort, now_s() - t2m);
        fflush(NULL);
        _exit(0);
    }
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);          /* kernel images load with the context, on the second thread */
    create_job job;
    memset(&job, 0, sizeof job);
    job.gpu = gpu;
    pthread_t th;
    if (pthread_create(&th, NULL, create_thread, &job) != 0) { fprintf(stderr, "cannot start a thread\n"); return 1; }

    fastx_t* fx = fastx_open(source);
    if (!fx) { fprintf(stderr, "can not open ref file\n"); return 1; }
    /* an upper bound of N for a plain file is its size; a gzip file (magic 1f 8b) gives none */
    uint64_t hint = 0;
    {
        struct stat sb;
        FILE* f = fopen(source, "rb");
        unsigned char magic[2] = {0, 0};
        if (f) { if (fread(magic, 1, 2, f) != 2) magic[0] = 0; fclose(f); }
        if (stat(source, &sb) == 0 && !(magic[0] == 0x1f && magic[1] == 0x8b)) hint = (uint64_t)sb.st_size + 64;
    }
    sink_t s;
    memset(&s, 0, sizeof s);
    s.job = &job;
    s.hint = hint;
    s.resolve = resolve;
    s.amb_seed = amb_seed;
    int rc = fastx_stream(fx, on_bases, on_record, &s);
    fastx_close(fx);
    pthread_join(th, NULL);                               /* never leave while CUDA is still starting up */
    if (rc == 3) return 1;
    if (rc < 0) { fprintf(stderr, "malformed input file\n"); return 1; }
    if (rc == 0 && s.nrec == 0 && !s.failed) { fprintf(stderr, "no sequence found in %s\n", source); return 1; }
    if (job.rc) { fprintf(stderr, "deBWT: %s\n", job.err); return 1; }                /* input is fine, but there is no device */
    if (!s.ctx && !s.failed && sink_attach(&s)) s.failed = 1;                        /* a file shorter than the start-up */
    if (rc == 2 || s.failed) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    debwt_ctx* ctx = s.ctx;
    double t_ctx = now_s();
    {
        const unsigned char c = '$';                       /* the last separator is '$' (src/collect#$.c:83) */
        if (sink_put(&s, &c, 1)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    }
    if (s.used && debwt_ingest_commit(ctx, s.used)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    if (debwt_ingest_end(ctx, s.seps, s.nrec)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    double t1 = now_s();
    fprintf(stderr, "BWTLEN=%llu (%llu records), read + parse + upload + pack %.3f s (CUDA start-up overlapped; %llu MB parked meanwhile)\n",
            (unsigned long long)s.n, (unsigned long long)s.nrec, t1 - t0, (unsigned long long)(s.park_n >> 20));
    (void)t_ctx;

    if (debwt_build(ctx, k)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    uint64_t nsym = 0, nwords = 0, nsharp = 0, dollar = 0;
    if (debwt_result_sizes(ctx, &nsym, &nwords, &nsharp)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    uint64_t* bwt = (uint64_t*)debwt_host_alloc((nwords ? nwords : 1) * 8);     /* page-locked: the D2H copy runs at link speed */
    uint64_t* sharp = (uint64_t*)malloc((nsharp ? nsharp : 1) * 8);
    if (!bwt || !sharp) { fprintf(stderr, "out of host memory\n"); return 1; }
    if (debwt_result_copy(ctx, bwt, sharp, &dollar)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    debwt_stats st;
    debwt_get_stats(ctx, &st);
    double t2 = now_s();

    size_t ol = strlen(obj);
    char* p2 = (char*)malloc(ol + 3);
    memcpy(p2, obj, ol);
    p2[ol] = '.'; p2[ol + 2] = 0;
    if (write_file(obj, bwt, nwords * 8)) return 1;
    p2[ol + 1] = '#';
    if (write_file(p2, sharp, nsharp * 8)) return 1;
    p2[ol + 1] = '$';
    if (write_file(p2, &dollar, 8)) return 1;
    double t3 = now_s();
    fprintf(stderr, "GPU build + download %.3f s (device %.1f ms: sort %.1f ms, %u launches), write %.3f s; branch k-mers %llu, blue %llu, SP codes %llu\n",
            t2 - t1, st.ms_total, st.ms_sort, st.total_launches, t3 - t2, (unsigned long long)st.n_branch,
            (unsigned long long)st.n_blue, (unsigned long long)st.n_codes);
    /* the three files are written and closed: leave without tearing the 56 GB arena and the CUDA context down piece by
       piece (the reference ends with exit(0) right after its fwrite()s too, src/insertCase3.c:137) */
    fflush(NULL);
    _exit(0);
}
