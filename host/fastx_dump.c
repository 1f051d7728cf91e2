/* Test helper: prints one "<length> <fnv1a64 of the sequence>" line per record read by host/fastx.h, then "records <n> rc <rc>".
   With -s as first argument the records are read through the streaming interface (fastx_stream) instead of fastx_read. */
#include <stdio.h>
#include "fastx.h"

typedef struct { unsigned long long h, n; } acc_t;
static int on_bases(void* u, const unsigned char* p, uint64_t n) {
    acc_t* a = (acc_t*)u;
    for (uint64_t i = 0; i < n; i++) { a->h ^= p[i]; a->h *= 1099511628211ull; }
    return 0;
}
static int on_record(void* u, uint64_t len) {
    acc_t* a = (acc_t*)u;
    printf("%llu %llu\n", (unsigned long long)len, a->h);
    a->h = 1469598103934665603ull;
    a->n++;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const int stream = argc > 2 && strcmp(argv[1], "-s") == 0;
    fastx_t* f = fastx_open(argv[stream ? 2 : 1]);
    if (!f) return 3;
    int rc;
    unsigned long long n = 0;
    if (stream) {
        acc_t a = {1469598103934665603ull, 0};
        rc = fastx_stream(f, on_bases, on_record, &a);
        n = a.n;
    } else {
        while ((rc = fastx_read(f)) == 1) {
            unsigned long long h = 1469598103934665603ull;
            for (uint64_t i = 0; i < f->len; i++) { h ^= (unsigned char)f->seq[i]; h *= 1099511628211ull; }
            printf("%llu %llu\n", (unsigned long long)f->len, h);
            n++;
        }
    }
    fastx_close(f);
    printf("records %llu rc %d\n", n, rc);
    return rc < 0 ? 1 : 0;
}
