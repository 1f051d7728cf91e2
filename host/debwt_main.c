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
 * accepted and ignored (the work runs on the GPU), -g picks the CUDA device, no temp files.
 * All computation happens in libdebwt_b200.so (include/debwt_b200.h); there is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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
    fprintf(stderr, "-g (optional): CUDA device ordinal (default 0)\n");
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

int main(int argc, char* argv[]) {
    if (argc < 4 || (argc & 1) == 1) { usage(); return 1; }
    const char* source = argv[argc - 1];
    const char* obj = NULL;
    int k = 32, gpu = 0;
    for (int i = 1; i < argc - 1; i += 2) {
        if (strcmp(argv[i], "-o") == 0) obj = argv[i + 1];
        else if (strcmp(argv[i], "-t") == 0) {
            if (atoi(argv[i + 1]) == 0) { fprintf(stderr, "thread number must be a number!\n"); return 1; }
        } else if (strcmp(argv[i], "-j") == 0) { /* ignored */
        } else if (strcmp(argv[i], "-g") == 0) gpu = atoi(argv[i + 1]);
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
    fastx_t* fx = fastx_open(source);
    if (!fx) { fprintf(stderr, "can not open ref file\n"); return 1; }
    /* T = S1 # S2 # ... Sn $ assembled in one host buffer (src/collect#$.c:56-90) */
    char* text = NULL;
    uint64_t n = 0, cap = 0, nrec = 0, sep_cap = 0;
    uint64_t* seps = NULL;
    int rc;
    while ((rc = fastx_read(fx)) == 1) {
        if (fx->len <= 32) { fprintf(stderr, "Length <= 32!\n"); return 1; }            /* src/collect#$.c:41-45 */
        if (n + fx->len + 1 > cap) {
            cap = cap ? cap : (1u << 20);
            while (cap < n + fx->len + 1) cap += cap >> 1;
            text = (char*)realloc(text, cap);
            if (!text) { fprintf(stderr, "out of host memory\n"); return 1; }
        }
        memcpy(text + n, fx->seq, fx->len);
        n += fx->len;
        if (nrec == sep_cap) { sep_cap = sep_cap ? sep_cap * 2 : 64; seps = (uint64_t*)realloc(seps, sep_cap * 8); }
        seps[nrec++] = n;
        text[n++] = '#';
    }
    fastx_close(fx);
    if (rc < 0) { fprintf(stderr, "malformed input file\n"); return 1; }
    if (nrec == 0) { fprintf(stderr, "no sequence found in %s\n", source); return 1; }
    text[n - 1] = '$';
    double t1 = now_s();
    fprintf(stderr, "BWTLEN=%llu (%llu records), read in %.3f s\n", (unsigned long long)n, (unsigned long long)nrec, t1 - t0);

    debwt_ctx* ctx = NULL;
    if (debwt_create(&ctx, gpu) || debwt_set_text(ctx, text, n, seps, nrec) || debwt_build(ctx, k)) {
        fprintf(stderr, "deBWT: %s\n", debwt_last_error());
        return 1;
    }
    free(text);
    uint64_t nsym = 0, nwords = 0, nsharp = 0, dollar = 0;
    if (debwt_result_sizes(ctx, &nsym, &nwords, &nsharp)) { fprintf(stderr, "deBWT: %s\n", debwt_last_error()); return 1; }
    uint64_t* bwt = (uint64_t*)malloc((nwords ? nwords : 1) * 8);
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
    fprintf(stderr, "GPU build %.3f s (device %.1f ms: sort %.1f ms, %u launches), write %.3f s; branch k-mers %llu, blue %llu, SP codes %llu\n",
            t2 - t1, st.ms_total, st.ms_sort, st.total_launches, t3 - t2, (unsigned long long)st.n_branch,
            (unsigned long long)st.n_blue, (unsigned long long)st.n_codes);
    debwt_destroy(ctx);
    return 0;
}
