// Sentinel-window ("special") suffixes: the 32 suffixes per record that start within k symbols of a
// separator (reference src/collect#$.c:131-157 generateSpecialSA, :228-311 cmp, :348-602 seeKMER /
// divideKmer, src/INandOut.c:419-439 insertion after the equal T-padded k-mer).
#pragma once
#include "stages.cuh"

namespace debwt {

// true suffix order under A<C<G<T<#<$; equal '#' are compared through, '$' is largest
// (reference cmp, src/collect#$.c:253-311).  words: packed text, seps: R ascending separator positions.
__host__ __device__ inline bool special_less(const u64* __restrict__ words, const u64* __restrict__ seps, u64 R,
                                             u64 pa, u64 pb) {
    if (pa == pb) return false;
    u64 ra = lower_bound_u64(seps, 0, R, pa), rb = lower_bound_u64(seps, 0, R, pb);
    for (;;) {
        const u64 da = seps[ra] - pa, db = seps[rb] - pb;
        const u64 m = da < db ? da : db;
        for (u64 off = 0; off < m; off += 32) {
            u64 wa = text_window32(words, pa + off), wb = text_window32(words, pb + off);
            const u64 len = m - off;
            if (len < 32) { const u64 mask = ~(~0ull >> (2 * len)); wa &= mask; wb &= mask; }
            if (wa != wb) return wa < wb;
        }
        if (da != db) return da > db;            // the side that reaches its separator first is larger
        const bool a_end = (ra + 1 == R), b_end = (rb + 1 == R);
        if (a_end || b_end) return !a_end && b_end;
        pa = seps[ra] + 1; pb = seps[rb] + 1;
        ++ra; ++rb;
    }
}

// The same with the records of both positions known (suffix t = rec * 32 + j starts at seps[rec] - j): no record search.
__host__ __device__ inline bool special_less_rec(const u64* __restrict__ words, const u64* __restrict__ seps, u64 R,
                                                 u64 pa, u64 ra, u64 pb, u64 rb) {
    if (pa == pb) return false;
    for (;;) {
        const u64 da = seps[ra] - pa, db = seps[rb] - pb;
        const u64 m = da < db ? da : db;
        for (u64 off = 0; off < m; off += 32) {
            u64 wa = text_window32(words, pa + off), wb = text_window32(words, pb + off);
            const u64 len = m - off;
            if (len < 32) { const u64 mask = ~(~0ull >> (2 * len)); wa &= mask; wb &= mask; }
            if (wa != wb) return wa < wb;
        }
        if (da != db) return da > db;
        const bool a_end = (ra + 1 == R), b_end = (rb + 1 == R);
        if (a_end || b_end) return !a_end && b_end;
        pa = seps[ra] + 1; pb = seps[rb] + 1;
        ++ra; ++rb;
    }
}

// what the host needs to know about one special suffix (position = seps[rec] - j)
struct SpecialInfo {
    u64 w0;        // 32 symbols starting at the position
    u64 w1;        // 32 symbols starting right after its separator
    u64 ins;       // number of sorted keys <= its T-padded key (local to this device's key range)
    u32 rank;      // number of special suffixes smaller than it
    u8 prev;       // symbol before the position (always a base)
    u8 next;       // symbol 31 positions further (the next symbol of its 31-symbol window)
    u8 pad_[2];
};

// Ranks the m = 32 R special suffixes (suffix t = rec*32 + j) under special_less and gathers their windows and
// insertion points.  Up to kSpecialAllPairs suffixes: one block per suffix counts the smaller ones (one launch).
// Beyond: a bitonic sorting network over the suffix ids with the same comparator -- O(m log^2 m) comparisons, every
// stage fully parallel, the text never leaves the device (the reference sorts them with a single-threaded qsort,
// src/collect#$.c:157).
constexpr u64 kSpecialAllPairs = 16384;
int k_special_scan(const u64* words, const u64* d_seps, u64 n_rec, const u64* sorted, u64 n_keys, KeyIndex ki,
                   SpecialInfo* out, cudaStream_t st);

}  // namespace debwt
