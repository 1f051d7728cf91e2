/* Test helper: prints "<n_records>" then one "<length> <fnv1a64 of the sequence>" line per record read by host/fastx.h. */
#include <stdio.h>
#include "fastx.h"
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    fastx_t* f = fastx_open(argv[1]);
    if (!f) return 3;
    int rc;
    unsigned long long n = 0;
    while ((rc = fastx_read(f)) == 1) {
        unsigned long long h = 1469598103934665603ull;
        for (uint64_t i = 0; i < f->len; i++) { h ^= (unsigned char)f->seq[i]; h *= 1099511628211ull; }
        printf("%llu %llu\n", (unsigned long long)f->len, h);
        n++;
    }
    fastx_close(f);
    printf("records %llu rc %d\n", n, rc);
    return rc < 0 ? 1 : 0;
}
