// Small device helpers shared by the single-GPU stages (stages.cu) and the sharded stages (dist_kernels.cu).
#pragma once
#include "stages.cuh"

namespace debwt {

__device__ __forceinline__ u64 record_of(const u64* __restrict__ seps, u64 n_rec, u64 p) {
    return lower_bound_u64(seps, 0, n_rec, p);     // number of separators strictly before p
}


// first index of the group (k-mer = key >> 2) that sorted[i] belongs to
__device__ __forceinline__ u64 group_head(const u64* __restrict__ k, u64 i) {
    const u64 x = k[i] >> 2;
    u64 hi = i, step = 1, lo;
    for (;;) {
        if (step > hi) { lo = 0; break; }
        u64 j = hi - step;
        if ((k[j] >> 2) == x) { hi = j; step <<= 1; }
        else { lo = j + 1; break; }
    }
    return lower_bound_u64(k, lo, hi, x << 2);
}

// one past the last index of the group that sorted[i] belongs to
__device__ __forceinline__ u64 group_end(const u64* __restrict__ k, u64 n, u64 i) {
    const u64 x = k[i] >> 2;
    u64 lo = i, step = 1, hi;
    for (;;) {
        u64 j = lo + step;
        if (j >= n) { hi = n; break; }
        if ((k[j] >> 2) == x) { lo = j; step <<= 1; }
        else { hi = j; break; }
    }
    return upper_bound_u64(k, lo, hi, (x << 2) | 3ull);
}


// b: branch id (hash table in hmode 1: the slot index instead); f: the entry's flags (multi_in << 1 | multi_out)
__device__ __forceinline__ bool branch_lookup(const BranchTable& bt, u64 x /* k-mer << 2, low bits 0 */, u64& b, u32& f) {
    // presence bitmap first: small enough to stay in L2, it rejects most positions (branch k-mers are rare)
    // before the table -- which does not fit in L2 at human scale -- is touched
    const u64 fi = bt.filter_of(x);
    if (!((bt.filter()[fi >> 5] >> (fi & 31)) & 1u)) return false;
    if (bt.hslots) {
        const u64 mask = (1ull << bt.hbits) - 1;
        for (u64 h = bt.hash_of(x);; h = (h + 1) & mask) {
            const ulonglong2 v = __ldg(bt.hslots + h);
            if (v.x == 0) return false;
            if ((v.x & ~3ull) == x) { b = bt.hmode ? h : v.y; f = (u32)(v.x & 3ull); return true; }
        }
    }
    const u64 t = x >> (64 - bt.bits);
    u64 lo = bt.bidx[t], hi = bt.bidx[t + 1];
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        u64 v = bt.kmer[mid] & ~3ull;
        if (v < x) lo = mid + 1; else hi = mid;
    }
    if (lo < bt.n_branch) {
        const u64 v = bt.kmer[lo];
        if ((v & ~3ull) == x) { b = lo; f = (u32)(v & 3ull); return true; }
    }
    return false;
}


__device__ __forceinline__ u64 sp_index_of(const u32* __restrict__ mo_bits, const u32* __restrict__ word_prefix, u64 p) {
    const u32 bits = mo_bits[p >> 5];
    return (u64)word_prefix[p >> 5] + __popc(bits & ((1u << (p & 31)) - 1u));
}


__device__ __forceinline__ void bwt_or(u64* __restrict__ bwt, u64 row, u32 code) {
    atomicOr(bwt + (row >> 5), (u64)code << (2 * (31 - (row & 31))));
}


}  // namespace debwt
