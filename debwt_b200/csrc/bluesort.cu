// K10: segmented sort of the blue entries of every multi-in k-mer by their branch-code strings.
//
// Replaces sortBlue / multiQuickSort / myQsort / cmpSP (reference src/sortBlue.c:10-280).  An entry is
// (spIndex << 4) | prev; entries of one k-mer are ordered by the code string that starts at spIndex,
// codes compared 32 at a time (one u64), '#' > T and equal '#' compared through, '$' largest
// (src/sortBlue.c:109-173).
//
// Every entry first caches the first 32 codes of its string (one u64, fetched once); almost all
// comparisons are decided on the cached word without touching memory, the rest continue in the code
// array from code 32 on.  Segments are binned by size:
//   <= 32   : one warp, entries in registers, rank by counting;
//   <= 128  : one warp, entries + cached words in shared memory, same-direction bitonic network;
//   larger  : round-based refinement (MSD style) driven by work lists: a work item is a run of entries that
//             agree on their first `depth` codes.  Items above 512 entries are cut by a sample-sort split on their code
//             word at `depth` (up to 256 buckets, equality buckets advance 32 codes; a window with a separator code is
//             bucketed by an order-preserving word, order_word()); items of at most 512 entries are sorted in shared
//             memory by a packed network (one u64 = 27 codes | index, two network stages per pass) whose comparisons
//             touch no memory; equal-word runs that still hold two prev symbols are finished by direct comparisons
//             (<= 32 entries) or become the next round's items, so long common prefixes are walked once per
//             entry instead of once per comparison and the work shrinks geometrically.  Items that hold a separator
//             code take a comparator network (in shared memory, or over HBM with 4096-entry blocks staged per merge level
//             when a whole long item ties on such windows).
// The network uses virtual +inf padding (all compare-exchanges point the same way), so no segment
// needs scratch for padding.  Segments whose prev symbols are all equal are skipped
// (src/sortBlue.c:192-219): any order gives the same BWT.
#include "stages.cuh"

#include <cstdio>
#include <cstdlib>

namespace debwt {

namespace {

constexpr int TPB = 256;
constexpr int WARPS = TPB / 32;
constexpr int MID_SEG = 128;
constexpr int BIG_TPB = 512;
constexpr int SMEM_SEG = 4096;     // == CHUNK: largest segment sorted entirely in shared memory

__device__ __forceinline__ u32 fetch_sep(const SpView& v, u64 s) {
    const u32* __restrict__ sep = v.sep;
    const u64 i = s >> 5;
    const u32 sh = (u32)(s & 31);
    const u32 lo = sep[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (sep[i + 1] << (32 - sh));
}

// separator flags of the 32 codes from s on, for the sites that fetch one window per entry: summary first
__device__ __forceinline__ u32 sep_window(const SpView& v, u64 s) {
    if (v.sep_sum) {
        const u64 i = s >> 5, j0 = i >> 5, j1 = (i + 1) >> 5;
        const u32 b0 = (__ldg(v.sep_sum + (j0 >> 5)) >> (j0 & 31)) & 1u, b1 = (__ldg(v.sep_sum + (j1 >> 5)) >> (j1 & 31)) & 1u;
        if (!(b0 | b1)) return 0u;
    }
    return fetch_sep(v, s);
}

// Code word of the window at s as an integer that sorts like the string among windows of one item: a window whose first
// separator code is its t-th code is larger than every window that agrees on the t codes before it and has a base there, so its
// word is filled with ones from code t on and `np` tells it apart from a plain word with the same value (plain first).
// Windows with equal (word, np) are not ordered by this: plain ones tie on 32 codes, the others need the comparator.
__device__ __forceinline__ u64 order_word(const SpView& v, u64 s, bool& np) {
    u64 wd = text_window32(v.codes, s);
    const u32 fs = sep_window(v, s);
    np = fs != 0;
    if (np) wd |= ~0ull >> (2 * (__ffs(fs) - 1));
    return wd;
}

// strict "string at sa < string at sb" starting the comparison `skip` codes in; sa != sb
__device__ __forceinline__ bool sp_less_from(const SpView& v, u64 sa, u64 sb, u32 skip) {
    sa += skip;
    sb += skip;
    for (;;) {
        const u64 ca = text_window32(v.codes, sa), cb = text_window32(v.codes, sb);
        const u32 fa = fetch_sep(v, sa), fb = fetch_sep(v, sb);
        if ((fa | fb) == 0) {
            if (ca != cb) return ca < cb;
        } else {
            for (int t = 0; t < 32; ++t) {
                u32 xa = (u32)(ca >> (2 * (31 - t))) & 3u, xb = (u32)(cb >> (2 * (31 - t))) & 3u;
                if ((fa >> t) & 1u) xa = (sa + t == v.dollar_index) ? 5u : 4u;
                if ((fb >> t) & 1u) xb = (sb + t == v.dollar_index) ? 5u : 4u;
                if (xa != xb) return xa < xb;
            }
        }
        sa += 32;
        sb += 32;
        if (sa >= v.n_codes || sb >= v.n_codes) return sa > sb;   // unreachable: the '$' code is unique and last
    }
}

// cached first word of an entry's string; `plain` = no separator code among its first 32 codes
struct Cached {
    u64 word;
    bool plain;
};
__device__ __forceinline__ Cached cache_of(const SpView& v, u64 entry) {
    const u64 s = entry >> 4;
    Cached c;
    c.word = text_window32(v.codes, s);
    c.plain = sep_window(v, s) == 0;
    return c;
}

// a cached word is only trusted when both sides are free of separator codes
__device__ __forceinline__ bool entry_less(const SpView& v, u64 ea, u64 wa, bool pa, u64 eb, u64 wb, bool pb) {
    if (pa && pb) {
        if (wa != wb) return wa < wb;
        return sp_less_from(v, ea >> 4, eb >> 4, 32);
    }
    return sp_less_from(v, ea >> 4, eb >> 4, 0);
}

// summary of the separator bitmap: bit j = some word of sep[32 j, 32 j + 32) is non-zero
__global__ void __launch_bounds__(TPB) sep_summary_kernel(const u32* __restrict__ sep, u64 nwords, u32* __restrict__ sum) {
    const u64 i = (u64)blockIdx.x * TPB + threadIdx.x;
    if (i < nwords && sep[i]) atomicOr(sum + (i >> 10), 1u << ((i >> 5) & 31));
}

// ---------------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) bin_segments_kernel(BranchTable bt, u32* __restrict__ counts /*[4]*/,
                                                          u32* __restrict__ small, u32* __restrict__ mid,
                                                          u32* __restrict__ block, u32* __restrict__ huge) {
    const u64 b = (u64)blockIdx.x * TPB + threadIdx.x;
    if (b >= bt.n_branch) return;
    if (!(bt.kmer[b] & 2ull)) return;
    const u32 len = bt.blue[b + 1] - bt.blue[b];
    if (len <= 1) return;
    if (len <= 32) small[atomicAdd(counts + 0, 1u)] = (u32)b;
    else if (len <= MID_SEG) mid[atomicAdd(counts + 1, 1u)] = (u32)b;
    else if (len <= SMEM_SEG) block[atomicAdd(counts + 2, 1u)] = (u32)b;
    else huge[atomicAdd(counts + 3, 1u)] = (u32)b;
}

// ---------------------------------------------------------------------------------------------
// <= 32: one warp, registers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) sort_small_kernel(u64* __restrict__ blue, BranchTable bt, SpView sp,
                                                        const u32* __restrict__ list, const u32* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const u32 n = *count;
    const u32 nwarps = gridDim.x * WARPS;
    for (u32 idx = (blockIdx.x * TPB + threadIdx.x) >> 5; idx < n; idx += nwarps) {
        const u32 b = list[idx];
        const u32 off = bt.blue[b];
        const u32 len = bt.blue[b + 1] - off;
        const bool act = (u32)lane < len;
        const u64 e = act ? blue[(u64)off + lane] : 0;
        const u32 c0 = __shfl_sync(0xffffffffu, (u32)(e & 15ull), 0);
        if (__all_sync(0xffffffffu, !act || (u32)(e & 15ull) == c0)) continue;
        Cached c{0, false};
        if (act) c = cache_of(sp, e);
        u32 rank = 0;
        for (u32 j = 0; j < len; ++j) {
            __syncwarp();
            const u64 ej = __shfl_sync(0xffffffffu, e, j);
            const u64 wj = __shfl_sync(0xffffffffu, c.word, j);
            const bool pj = __shfl_sync(0xffffffffu, (int)c.plain, j);
            if (act && j != (u32)lane && entry_less(sp, ej, wj, pj, e, c.word, c.plain)) ++rank;
        }
        __syncwarp();
        if (act) blue[(u64)off + rank] = e;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// same-direction bitonic network over (entry, cached word, plain flag); `nthreads` cooperating threads
// ---------------------------------------------------------------------------------------------
template <typename Sync>
__device__ __forceinline__ void bitonic_cached(u64* ent, u64* wrd, u8* pln, u64 len, const SpView& sp, u32 tid, u32 nthreads,
                                               Sync sync) {
    u64 P = 1;
    while (P < len) P <<= 1;
    auto cmpswap = [&](u64 i, u64 l) {
        const u64 ei = ent[i], el = ent[l], wi = wrd[i], wl = wrd[l];
        const bool pi = pln[i], pl = pln[l];
        if (entry_less(sp, el, wl, pl, ei, wi, pi)) {
            ent[i] = el; ent[l] = ei; wrd[i] = wl; wrd[l] = wi; pln[i] = pl; pln[l] = pi;
        }
    };
    for (u64 k = 2; k <= P; k <<= 1) {
        const u64 half = k >> 1;
        for (u64 t = tid; t < (P >> 1); t += nthreads) {                 // mirror step
            const u64 i = (t / half) * k + (t % half);
            const u64 l = i ^ (k - 1);
            if (l < len) cmpswap(i, l);
        }
        sync();
        for (u64 j = k >> 2; j > 0; j >>= 1) {                           // half cleaners
            for (u64 t = tid; t < (P >> 1); t += nthreads) {
                const u64 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u64 l = i | j;
                if (l < len) cmpswap(i, l);
            }
            sync();
        }
    }
}

// 33..128: one warp per segment, shared memory
__global__ void __launch_bounds__(TPB) sort_mid_kernel(u64* __restrict__ blue, BranchTable bt, SpView sp,
                                                      const u32* __restrict__ list, const u32* __restrict__ count) {
    __shared__ u64 s_ent[WARPS][MID_SEG];
    __shared__ u64 s_wrd[WARPS][MID_SEG];
    __shared__ u8 s_pln[WARPS][MID_SEG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n = *count;
    const u32 nwarps = gridDim.x * WARPS;
    for (u32 idx = (blockIdx.x * TPB + threadIdx.x) >> 5; idx < n; idx += nwarps) {
        const u32 b = list[idx];
        const u64 off = bt.blue[b];
        const u32 len = bt.blue[b + 1] - (u32)off;
        bool same = true;
        const u32 c0 = (u32)(blue[off] & 15ull);
        for (u32 t = lane; t < len; t += 32) {
            const u64 e = blue[off + t];
            const Cached c = cache_of(sp, e);
            s_ent[warp][t] = e; s_wrd[warp][t] = c.word; s_pln[warp][t] = c.plain;
            same &= (u32)(e & 15ull) == c0;
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, same)) continue;
        bitonic_cached(s_ent[warp], s_wrd[warp], s_pln[warp], len, sp, lane, 32, [] { __syncwarp(); });
        for (u32 t = lane; t < len; t += 32) blue[off + t] = s_ent[warp][t];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// > 128 entries: round-based refinement.  Round r sorts the segment by (tie group, 32 codes at depth 32 r):
// the comparisons of a round touch no memory beyond the staged arrays (one code word is fetched per
// entry per round), so near-identical suffixes -- repeat families whose copies agree for hundreds of
// codes -- cost one cheap network per 32 codes instead of a deep string walk inside every comparison.
// A tie group is a run of entries that agreed on every word so far; it keeps its place, is refined by the
// next word, and is finished as soon as it is a singleton or all of its prev symbols agree
// (src/sortBlue.c:192-219).  A word that contains a separator code ('#'/'$', at most R in the whole code
// string) cannot be compared as an integer: such a segment falls back to the comparator network.
// ---------------------------------------------------------------------------------------------
constexpr int CHUNK = 4096;                              // entries staged in shared memory at a time
constexpr size_t kChunkSmem = (size_t)CHUNK * 20;        // entry 8 + key 8 + tag 4 bytes

struct SegArrays {
    u64* ent;      // entries
    u64* key;      // current code word (or cached first word for the comparator fallback)
    u32* tag;      // tie group id (position of the group's first entry) / plain flag for the fallback
};

template <typename Less>
__device__ __forceinline__ void cmpswap3(const SegArrays& a, u32 i, u32 l, const Less& less) {
    const u64 ei = a.ent[i], el = a.ent[l], ki = a.key[i], kl = a.key[l];
    if constexpr (Less::kUsesTag) {
        const u32 ti = a.tag[i], tl = a.tag[l];
        if (less(el, kl, tl, ei, ki, ti)) {
            a.ent[i] = el; a.ent[l] = ei; a.key[i] = kl; a.key[l] = ki; a.tag[i] = tl; a.tag[l] = ti;
        }
    } else {                                            // tags carry nothing the order depends on: leave them
        if (less(el, kl, 0u, ei, ki, 0u)) { a.ent[i] = el; a.ent[l] = ei; a.key[i] = kl; a.key[l] = ki; }
    }
}

// all stages of the network whose partner distance is <= span/2, on `len` entries held in `a`
// (span = power of two >= len for a full sort, or CHUNK for the tail stages of a longer merge level)
template <typename Less>
__device__ __forceinline__ void net_local(const SegArrays& a, u32 len, u32 k_first, u32 k_last, bool tail_only, const Less& less) {
    for (u32 k = k_first; k <= k_last; k <<= 1) {
        const u32 half = k >> 1;
        if (!tail_only) {
            for (u32 t = threadIdx.x; t < (k_last >> 1); t += blockDim.x) {               // mirror step
                const u32 i = ((t & ~(half - 1)) << 1) | (t & (half - 1));
                const u32 l = i ^ (k - 1);
                if (l < len) cmpswap3(a, i, l, less);
            }
            __syncthreads();
        }
        for (u32 j = tail_only ? half : (k >> 2); j > 0; j >>= 1) {                       // half cleaners
            for (u32 t = threadIdx.x; t < (k_last >> 1); t += blockDim.x) {
                const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u32 l = i | j;
                if (l < len) cmpswap3(a, i, l, less);
            }
            __syncthreads();
        }
        if (tail_only) break;
    }
}

__device__ __forceinline__ u32 pow2_at_least(u32 n) {
    u32 p = 1;
    while (p < n) p <<= 1;
    return p;
}

// full network over `len` entries in `g` (HBM when len > CH), staging through `s`
template <int CH, typename Less>
__device__ void net_sort(const SegArrays& g, const SegArrays& s, u32 len, bool in_smem, const Less& less) {
    if (in_smem) {                                  // the segment already sits in `s`
        net_local(s, len, 2, pow2_at_least(len), false, less);
        return;
    }
    const u32 P = pow2_at_least(len);
    for (u32 k = CH; k <= P; k <<= 1) {
        if (k > CH) {                            // stages with partner distance >= CH run over HBM
            const u32 half = k >> 1;
            for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const u32 i = ((t & ~(half - 1)) << 1) | (t & (half - 1));
                const u32 l = i ^ (k - 1);
                if (l < len) cmpswap3(g, i, l, less);
            }
            __syncthreads();
            for (u32 j = k >> 2; j >= (u32)CH; j >>= 1) {
                for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const u32 l = i | j;
                    if (l < len) cmpswap3(g, i, l, less);
                }
                __syncthreads();
            }
        }
        for (u32 c0 = 0; c0 < len; c0 += CH) {   // the remaining stages block by block in shared memory
            const u32 clen = (len - c0 < (u32)CH) ? len - c0 : (u32)CH;
            for (u32 t = threadIdx.x; t < clen; t += blockDim.x) {
                s.ent[t] = g.ent[c0 + t]; s.key[t] = g.key[c0 + t]; s.tag[t] = g.tag[c0 + t];
            }
            __syncthreads();
            if (k == (u32)CH) net_local(s, clen, 2, pow2_at_least(clen), false, less);
            else net_local(s, clen, CH, CH, true, less);
            for (u32 t = threadIdx.x; t < clen; t += blockDim.x) {
                g.ent[c0 + t] = s.ent[t]; g.key[c0 + t] = s.key[t]; g.tag[c0 + t] = s.tag[t];
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Packed network for items that sit in shared memory and hold no separator code: one u64 per entry =
// (the next ADV codes of its string, left-aligned) | (its index in the item), so a compare-exchange moves 8 bytes
// instead of an (entry, word) pair, all elements are distinct (the order is the stable one), and the entries are
// permuted once at the end.  (A (word, 16-bit index) pair network that keeps all 32 codes was measured too: the second
// array costs as many shared-memory instructions as it saves bytes -- no gain over the (entry, word) network.)
// Two consecutive stages of the bitonic network always act on closed groups of four elements (mirror k + half cleaner
// k/4; half cleaners 2j + j): a thread loads such a quad, does both stages in registers and stores it, which halves the
// shared-memory traffic and the barriers again.
// ---------------------------------------------------------------------------------------------
constexpr int ilog2_ceil(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }

template <int CH>
struct Packed {
    static constexpr int IB = ilog2_ceil(CH);                          // index bits
    static constexpr u32 ADV = (64 - IB) / 2;                          // codes compared per step: 27 (CH 512), 26 (CH 4096)
    static constexpr u64 WMASK = ~((1ull << (64 - 2 * (int)ADV)) - 1ull);
    static constexpr u64 IMASK = (1ull << IB) - 1ull;
};

__device__ __forceinline__ void cx64(u64& a, u64& b) {
    if (b < a) { const u64 t = a; a = b; b = t; }
}

// elements beyond len are virtual +inf: never stored, and every compare-exchange moves the minimum down
struct Quad {
    u32 q[4];
    u64 r[4];
    __device__ __forceinline__ void load(const u64* a, u32 len) {
#pragma unroll
        for (int m = 0; m < 4; ++m) r[m] = q[m] < len ? a[q[m]] : ~0ull;
    }
    __device__ __forceinline__ void store(u64* a, u32 len) const {
#pragma unroll
        for (int m = 0; m < 4; ++m) if (q[m] < len) a[q[m]] = r[m];
    }
};

__device__ void packed_sort(u64* a, u32 len) {
    const u32 P = pow2_at_least(len < 4 ? 4 : len);
    const u32 nq = P >> 2;
    // levels 2 and 4 on aligned quads: mirror 2, mirror 4, half cleaner 1
    for (u32 t = threadIdx.x; t < nq; t += blockDim.x) {
        Quad v;
#pragma unroll
        for (int m = 0; m < 4; ++m) v.q[m] = 4 * t + m;
        v.load(a, len);
        cx64(v.r[0], v.r[1]); cx64(v.r[2], v.r[3]);
        cx64(v.r[0], v.r[3]); cx64(v.r[1], v.r[2]);
        cx64(v.r[0], v.r[1]); cx64(v.r[2], v.r[3]);
        v.store(a, len);
    }
    __syncthreads();
    for (u32 k = 8; k <= P; k <<= 1) {
        const u32 kq = k >> 2;
        // mirror k + half cleaner k/4
        for (u32 t = threadIdx.x; t < nq; t += blockDim.x) {
            const u32 blk = (t / kq) * k, off = t & (kq - 1);
            Quad v;
            v.q[0] = blk + off; v.q[1] = blk + off + kq; v.q[2] = blk + k - 1 - off - kq; v.q[3] = blk + k - 1 - off;
            v.load(a, len);
            cx64(v.r[0], v.r[3]); cx64(v.r[1], v.r[2]);
            cx64(v.r[0], v.r[1]); cx64(v.r[2], v.r[3]);
            v.store(a, len);
        }
        __syncthreads();
        u32 j = k >> 3;                                  // next half cleaner
        for (; j >= 2; j >>= 2) {                        // half cleaners j and j/2 together
            const u32 jb = j >> 1;
            for (u32 t = threadIdx.x; t < nq; t += blockDim.x) {
                const u32 i = ((t & ~(jb - 1)) << 2) | (t & (jb - 1));
                Quad v;
                v.q[0] = i; v.q[1] = i + jb; v.q[2] = i + j; v.q[3] = i + j + jb;
                v.load(a, len);
                cx64(v.r[0], v.r[2]); cx64(v.r[1], v.r[3]);
                cx64(v.r[0], v.r[1]); cx64(v.r[2], v.r[3]);
                v.store(a, len);
            }
            __syncthreads();
        }
        if (j == 1) {                                    // one half cleaner left over
            for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const u32 i = 2 * t, l = i + 1;
                if (l < len) {
                    const u64 x = a[i], y = a[l];
                    if (y < x) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

struct LessKey {
    static constexpr bool kUsesTag = false;
    __device__ __forceinline__ bool operator()(u64, u64 ka, u32, u64, u64 kb, u32) const { return ka < kb; }
};
// comparator for items whose code words contain a separator: words at `depth` are cached in key (tag = plain)
struct LessFromDepth {
    static constexpr bool kUsesTag = true;
    SpView sp;
    u32 depth;
    __device__ __forceinline__ bool operator()(u64 ea, u64 ka, u32 ta, u64 eb, u64 kb, u32 tb) const {
        if (ta && tb) {
            if (ka != kb) return ka < kb;
            return sp_less_from(sp, ea >> 4, eb >> 4, depth + 32);
        }
        return sp_less_from(sp, ea >> 4, eb >> 4, depth);
    }
};

// A work item = a run of blue entries that agree on their first `depth` codes and still hold two different
// prev symbols.  One block sorts it by the next 32 codes; the runs of equal words that come out are
// finished (singletons, equal prev symbols), finished here by direct comparisons (<= 32 entries) or
// appended to the next round's list.  Work shrinks geometrically from round to round.
struct WorkItem {
    u64 off;
    u32 len, depth;
};

// Work lists of one round, by item size: `tiny` (<= TINY entries: 128-thread blocks, many per SM -- the per-item cost
// is barrier latency, so what counts is the number of items in flight), `small` (<= CHUNK: one 512-thread block,
// whole item in shared memory), `huge` (first cut into tiny/small ones by split_kernel).  Counters live in
// device memory: cnt[0] small, cnt[1] huge, cnt[2] tiny.
constexpr int TINY = 512;
constexpr int TINY_TPB = 64;
struct WorkLists {
    WorkItem* small;
    WorkItem* huge;
    WorkItem* tiny;
    u32* cnt;
    u32 cap_small, cap_huge, cap_tiny;
};

// SPLIT_ABOVE: items longer than this are cut by split_kernel first.  One big network over a 513..4096-entry item costs six
// times more per entry than the sample-sort split plus the small networks of its buckets (measured at 3.1 Gbp: 2.7 against
// 16 G entries/s), so everything above the tiny class is split.
constexpr u32 SPLIT_ABOVE = 512;       // == TINY
__device__ __forceinline__ void push_item(const WorkLists& l, const WorkItem& w) {
    if (w.len > SPLIT_ABOVE) {
        const u32 i = atomicAdd(l.cnt + 1, 1u);
        if (i < l.cap_huge) l.huge[i] = w;
    } else {
        const u32 i = atomicAdd(l.cnt + 2, 1u);
        if (i < l.cap_tiny) l.tiny[i] = w;
    }
}
// an item refine_kernel is to sort in this round (a bucket split_kernel has just cut, at most CHUNK entries)
__device__ __forceinline__ void push_for_network(const WorkLists& l, const WorkItem& w) {
    if (w.len > (u32)512) {
        const u32 i = atomicAdd(l.cnt + 0, 1u);
        if (i < l.cap_small) l.small[i] = w;
    } else {
        const u32 i = atomicAdd(l.cnt + 2, 1u);
        if (i < l.cap_tiny) l.tiny[i] = w;
    }
}

__global__ void __launch_bounds__(TPB) seed_items_kernel(BranchTable bt, const u32* __restrict__ list, u32 n, WorkLists cur) {
    const u32 i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const u32 b = list[i];
    WorkItem w;
    w.off = bt.blue[b];
    w.len = bt.blue[b + 1] - bt.blue[b];
    w.depth = 0;
    push_item(cur, w);
}

// Sample-sort partition of the huge items (one block per item): up to 128 buckets delimited by splitters drawn
// from the item's own code words at `depth` (8x oversampled), with a separate bucket for "equal to a splitter"
// so that heavy ties -- the copies of a repeat family that agree on all 32 codes -- never unbalance a bucket.
// Entries are moved bucket by bucket (through `scratch`), every bucket becomes an item of its own: ordinary
// buckets at the same depth (sorted by refine_kernel this round when they fit CHUNK, split again next round
// otherwise), equality buckets at depth + 32.  O(len) traffic per item instead of the O(len log^2 len) of a
// sorting network over HBM.  A window with a separator code is bucketed by order_word(): it lands where it belongs among
// the plain windows, and only the small item it ends up in takes refine_kernel's comparator path (a huge item that holds one
// such window used to go to the comparator network over HBM as a whole: 7 ms at 3.1 Gbp).
constexpr int SPLIT_MAX_BUCKETS = 256;
constexpr int SPLIT_OVERSAMPLE = 8;
constexpr int SPLIT_TARGET = 256;      // aimed-for bucket size (128 measured the same)

__global__ void __launch_bounds__(BIG_TPB) split_kernel(u64* __restrict__ blue, SpView sp, const WorkItem* __restrict__ items,
                                                       const u32* __restrict__ n_items_ptr, WorkLists cur, WorkLists next,
                                                       u64* __restrict__ scratch, u32* __restrict__ g_bucket) {
    __shared__ u64 s_sample[SPLIT_MAX_BUCKETS * SPLIT_OVERSAMPLE];
    __shared__ u64 s_split[SPLIT_MAX_BUCKETS];
    __shared__ u32 s_hist[2 * SPLIT_MAX_BUCKETS], s_start[2 * SPLIT_MAX_BUCKETS], s_pos[2 * SPLIT_MAX_BUCKETS];
    __shared__ int s_flag, s_mixed;
    const u32 n_items = *n_items_ptr;
    for (u32 idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const WorkItem it = items[idx];
        const u32 len = it.len, depth = it.depth;
        u64* ent = blue + it.off;
        u32 nb = (len + SPLIT_TARGET - 1) / SPLIT_TARGET;
        if (nb > (u32)SPLIT_MAX_BUCKETS) nb = SPLIT_MAX_BUCKETS;
        if (nb < 2) nb = 2;
        const u32 m = nb * SPLIT_OVERSAMPLE, P = pow2_at_least(m), n_split = nb - 1, n_buckets = 2 * nb - 1;
        if (threadIdx.x == 0) { s_flag = 0; s_mixed = 0; }
        for (u32 t = threadIdx.x; t < 2u * SPLIT_MAX_BUCKETS; t += blockDim.x) s_hist[t] = 0;
        // ---- sample, sort the sample, keep every SPLIT_OVERSAMPLE-th word as a splitter ----
        for (u32 t = threadIdx.x; t < P; t += blockDim.x) {
            u64 wd = ~0ull;
            if (t < m) {
                const u64 e = ent[(u64)t * len / m];
                bool np;
                wd = order_word(sp, (e >> 4) + depth, np);
            }
            s_sample[t] = wd;
        }
        __syncthreads();
        for (u32 k = 2; k <= P; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const u64 a = s_sample[i], b = s_sample[l];
                    if ((a > b) == ((i & k) == 0)) { s_sample[i] = b; s_sample[l] = a; }
                }
                __syncthreads();
            }
        }
        for (u32 t = threadIdx.x; t < n_split; t += blockDim.x) s_split[t] = s_sample[t * SPLIT_OVERSAMPLE + SPLIT_OVERSAMPLE - 1];
        __syncthreads();
        // ---- pass 1: bucket of every entry, histogram ----
        const u32 prev0 = (u32)(ent[0] & 15ull);
        for (u32 t = threadIdx.x; t < len; t += blockDim.x) {
            const u64 e = ent[t];
            if ((u32)(e & 15ull) != prev0) s_mixed = 1;
            bool np;
            const u64 wd = order_word(sp, (e >> 4) + depth, np);
            u32 lo = 0, hi = n_split;                       // first splitter >= wd
            while (lo < hi) {
                const u32 mid = (lo + hi) >> 1;
                if (s_split[mid] < wd) lo = mid + 1; else hi = mid;
            }
            // equal to a splitter: the plain windows tie on 32 codes (equality bucket, 32 codes deeper next time); a window with a
            // separator code sorts after them, i.e. first in the ordinary bucket that follows
            const u32 b = 2 * lo + ((lo < n_split && s_split[lo] == wd) ? (np ? 2u : 1u) : 0u);
            atomicAdd(&s_hist[b], 1u);
            g_bucket[it.off + t] = b;
        }
        __syncthreads();
        const bool mixed = s_mixed != 0;
        if (!mixed) { __syncthreads(); continue; }          // every prev symbol equal: nothing to order
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (u32 b = 0; b < n_buckets; ++b) {
                s_start[b] = run; s_pos[b] = run; run += s_hist[b];
                if (s_hist[b] == len && !(b & 1u)) s_flag = 1;      // no progress: windows with separator codes that all tie
            }
        }
        __syncthreads();
        if (s_flag) {                                       // the comparator network of refine_kernel orders them
            if (threadIdx.x == 0) {
                const u32 i = atomicAdd(cur.cnt + 0, 1u);
                if (i < cur.cap_small) cur.small[i] = it;
            }
            __syncthreads();
            continue;
        }
        // ---- pass 2: move the entries bucket by bucket ----
        for (u32 t = threadIdx.x; t < len; t += blockDim.x) {
            const u32 b = g_bucket[it.off + t];
            scratch[it.off + atomicAdd(&s_pos[b], 1u)] = ent[t];
        }
        __syncthreads();
        for (u32 t = threadIdx.x; t < len; t += blockDim.x) ent[t] = scratch[it.off + t];
        // ---- every bucket with two or more entries is an item ----
        for (u32 b = threadIdx.x; b < n_buckets; b += blockDim.x) {
            const u32 c = s_hist[b];
            if (c < 2) continue;
            WorkItem nw;
            nw.off = it.off + s_start[b]; nw.len = c; nw.depth = depth + ((b & 1u) ? 32u : 0u);
            // still long (a heavy tie, or one of 256 buckets of a huge item): cut again next round.  (Sending the long ties that fit a
            // block to refine_kernel's dominant-word peel instead measured 8 ms slower on the 10-haplotype collection.)
            if (c > SPLIT_ABOVE) push_item(next, nw);
            else push_for_network(cur, nw);
        }
        __syncthreads();
    }
}

// One warp puts `size` <= 32 entries that agree on their first `depth` codes into their final order: every lane
// fetches its word at `depth` once, the pairwise comparisons then run on registers and only walk the code strings when
// two words tie as well.
__device__ __forceinline__ void warp_rank_short(const SpView& sp, u64* ent, u32 size, u32 depth) {
    const u32 lane = threadIdx.x & 31;
    const bool mem = lane < size;
    const u64 e = mem ? ent[lane] : 0;
    u64 nw = 0;
    u32 np = 0;
    if (mem) {
        const u64 sidx = (e >> 4) + depth;
        nw = text_window32(sp.codes, sidx);
        np = (sep_window(sp, sidx) == 0 && sidx + 32 <= sp.n_codes) ? 1u : 0u;
    }
    u32 rank = 0;
    for (u32 j = 0; j < size; ++j) {
        __syncwarp();
        const u64 ej = __shfl_sync(0xffffffffu, e, j);
        const u64 wj = __shfl_sync(0xffffffffu, nw, j);
        const u32 pj = __shfl_sync(0xffffffffu, np, j);
        if (mem && j != lane) {
            bool less;
            if (pj && np) less = wj != nw ? wj < nw : sp_less_from(sp, ej >> 4, e >> 4, depth + 32);
            else less = sp_less_from(sp, ej >> 4, e >> 4, depth);
            if (less) ++rank;
        }
    }
    __syncwarp();
    if (mem) ent[rank] = e;
    __syncwarp();
}

template <int NT, int CH, int MINB>
__global__ void __launch_bounds__(NT, MINB) refine_kernel(u64* __restrict__ blue, SpView sp, const WorkItem* __restrict__ items,
                                                         const u32* __restrict__ n_items_ptr, WorkLists next,
                                                        u64* __restrict__ g_key, u32* __restrict__ g_tag) {
    const u32 n_items = *n_items_ptr;
    extern __shared__ __align__(16) unsigned char blk_smem[];
    SegArrays s;
    s.ent = reinterpret_cast<u64*>(blk_smem);
    s.key = s.ent + CH;
    s.tag = reinterpret_cast<u32*>(s.key + CH);
    __shared__ int s_flag, s_mixed;
    __shared__ u32 s_cnt[4];
    __shared__ u32 s_list[CH / 2];         // short unresolved runs of the current item: (size << 16) | head
    __shared__ u32 s_wm[32], s_ws[32];     // per-warp totals of the run scan (max head, prev-symbol changes)
    constexpr u32 MAX_PEEL = 64;           // dominant-word peels per item and launch
    for (u32 idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
        const WorkItem it = items[idx];
        const u32 len = it.len;
        const bool in_hbm = len > (u32)CH;
        SegArrays g;
        g.ent = blue + it.off;
        g.key = in_hbm ? g_key + it.off : s.key;
        g.tag = in_hbm ? g_tag + it.off : s.tag;
        const SegArrays& w = in_hbm ? g : s;
        if (!in_hbm) {
            for (u32 t = threadIdx.x; t < len; t += blockDim.x) s.ent[t] = g.ent[t];
        }
        // the part of the item still being refined: entries [lo, lo + vlen) agree on their first `depth` codes
        u32 lo = 0, vlen = len, depth = it.depth, peels = 0;
        for (;;) {
            if (threadIdx.x == 0) { s_flag = 0; s_mixed = 0; s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0; }
            __syncthreads();
            SegArrays v;
            v.ent = w.ent + lo; v.key = w.key + lo; v.tag = w.tag + lo;
            const u64 voff = it.off + lo;
            if (peels == MAX_PEEL) {                  // keep the launch bounded: the rest of the tie goes to the next round
                if (threadIdx.x == 0) { WorkItem nw; nw.off = voff; nw.len = vlen; nw.depth = depth; push_item(next, nw); }
                break;
            }
            // ---- fetch the code word of every entry at this depth ----
            const u32 prev0 = (u32)(v.ent[0] & 15ull);
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) {
                const u64 e = v.ent[t];
                if ((u32)(e & 15ull) != prev0) s_mixed = 1;
                const u64 sidx = (e >> 4) + depth;
                v.key[t] = text_window32(sp.codes, sidx);
                const bool plain = sep_window(sp, sidx) == 0 && sidx + 32 <= sp.n_codes;
                v.tag[t] = plain ? 1u : 0u;
                if (!plain) s_flag = 1;
            }
            __syncthreads();
            const bool fallback = s_flag != 0, mixed = s_mixed != 0;
            if (!mixed) break;         // every prev symbol equal: any order gives the same BWT (src/sortBlue.c:192-219)
            if (fallback) {
                LessFromDepth lf{sp, depth};
                SegArrays gv = g;
                gv.ent += lo; gv.key += lo; gv.tag += lo;
                net_sort<CH>(gv, in_hbm ? s : v, vlen, !in_hbm, lf);       // final order for this item
                break;
            }
            // ---- dominant word (the copies of a repeat that agree on these 32 codes as well): three-way partition
            //      around it instead of a sort; the equal part is refined further right here, 32 codes deeper ----
            if (!in_hbm && vlen > 32) {
                const u64 pivot = v.key[vlen >> 1];
                for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) {
                    const u64 kx = v.key[t];
                    const u32 cls = kx < pivot ? 0u : (kx == pivot ? 1u : 2u);
                    v.tag[t] = cls;
                    atomicAdd(&s_cnt[cls], 1u);
                }
                __syncthreads();
                const u32 n_lt = s_cnt[0], n_eq = s_cnt[1], n_gt = s_cnt[2];
                __syncthreads();
                if (2 * n_eq >= vlen) {
                    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = n_lt; s_cnt[2] = n_lt + n_eq; }
                    __syncthreads();
                    for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) v.key[atomicAdd(&s_cnt[v.tag[t]], 1u)] = v.ent[t];
                    __syncthreads();
                    for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) v.ent[t] = v.key[t];
                    __syncthreads();
                    // the two minorities: a warp orders a short one right away, a long one is an item of its own
                    if (threadIdx.x == 0) {
                        WorkItem nw;
                        nw.depth = depth;
                        if (n_lt > 32) { nw.off = voff; nw.len = n_lt; push_item(next, nw); }
                        if (n_gt > 32) { nw.off = voff + n_lt + n_eq; nw.len = n_gt; push_item(next, nw); }
                    }
                    if ((threadIdx.x >> 5) == 0 && n_lt >= 2 && n_lt <= 32) warp_rank_short(sp, v.ent, n_lt, depth);
                    if ((threadIdx.x >> 5) == 1 && n_gt >= 2 && n_gt <= 32) warp_rank_short(sp, v.ent + n_lt + n_eq, n_gt, depth);
                    lo += n_lt; vlen = n_eq; depth += 32; ++peels;
                    __syncthreads();
                    continue;
                }
            }
            if (!in_hbm) {
                // ---- packed network (see Packed / packed_sort): order by the next ADV codes, stable ----
                using PK = Packed<CH>;
                constexpr int IPT = CH / NT;
                constexpr u32 ADV = PK::ADV;
                const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) v.key[t] = (v.key[t] & PK::WMASK) | t;
                __syncthreads();
                packed_sort(v.key, vlen);
                u64 pe[IPT];
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const u32 t = threadIdx.x + q * NT;
                    pe[q] = t < vlen ? v.ent[v.key[t] & PK::IMASK] : 0;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const u32 t = threadIdx.x + q * NT;
                    if (t < vlen) v.ent[t] = pe[q];
                }
                __syncthreads();
                // ---- runs of equal words: one block scan gives every entry its run head (max) and the number of prev-symbol
                //      changes inside runs before it (sum); a run is unresolved when that number grows across it ----
                u32 hm[IPT], hs[IPT];
                u32 rm = 0, rs = 0;
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const u32 t = threadIdx.x * IPT + q;
                    if (t < vlen && t > 0) {
                        if ((v.key[t] ^ v.key[t - 1]) & PK::WMASK) rm = t;
                        else if ((v.ent[t] ^ v.ent[t - 1]) & 15ull) ++rs;
                    }
                    hm[q] = rm; hs[q] = rs;
                }
                u32 im = rm, is = rs;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 um = __shfl_up_sync(0xffffffffu, im, o), us = __shfl_up_sync(0xffffffffu, is, o);
                    if (lane >= (u32)o) { im = um > im ? um : im; is += us; }
                }
                if (lane == 31) { s_wm[warp] = im; s_ws[warp] = is; }
                u32 pm = __shfl_up_sync(0xffffffffu, im, 1), ps = __shfl_up_sync(0xffffffffu, is, 1);
                if (lane == 0) { pm = 0; ps = 0; }
                __syncthreads();
                for (u32 w2 = 0; w2 < warp; ++w2) { const u32 a = s_wm[w2]; pm = a > pm ? a : pm; ps += s_ws[w2]; }
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const u32 t = threadIdx.x * IPT + q;
                    hm[q] = hm[q] > pm ? hm[q] : pm;
                    if (t < vlen) v.tag[t] = hs[q] + ps;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const u32 t = threadIdx.x * IPT + q;
                    if (t >= vlen) continue;
                    if (t + 1 < vlen && !((v.key[t + 1] ^ v.key[t]) & PK::WMASK)) continue;      // not the last entry of its run
                    const u32 h = hm[q];
                    if (v.tag[t] == v.tag[h]) continue;                                           // one prev symbol: resolved
                    const u32 size = t + 1 - h;
                    if (size > 32) {
                        WorkItem nw;
                        nw.off = voff + h; nw.len = size; nw.depth = depth + ADV;
                        push_item(next, nw);
                    } else {
                        s_list[atomicAdd(&s_cnt[3], 1u)] = (size << 16) | h;
                    }
                }
                __syncthreads();
                const u32 n_list = s_cnt[3];
                for (u32 q = threadIdx.x >> 5; q < n_list; q += blockDim.x >> 5) {
                    const u32 x = s_list[q];
                    warp_rank_short(sp, v.ent + (x & 0xffffu), x >> 16, depth + ADV);
                }
                break;
            }
            {
                SegArrays gv = g;
                gv.ent += lo; gv.key += lo; gv.tag += lo;
                net_sort<CH>(gv, s, vlen, false, LessKey());
            }
            // ---- runs of equal words: head index of every entry (max-scan over "own index if head") ----
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) v.tag[t] = (t == 0 || v.key[t] != v.key[t - 1]) ? 1u : 0u;
            __syncthreads();
            u32* k32 = reinterpret_cast<u32*>(v.key);
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) k32[2 * t] = v.tag[t] ? t : 0u;
            __syncthreads();
            u32 ph = 0;
            for (u32 d = 1; d < vlen; d <<= 1) {
                for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) {
                    u32 x = k32[2 * t + ph];
                    if (t >= d) { const u32 o = k32[2 * (t - d) + ph]; x = o > x ? o : x; }
                    k32[2 * t + (ph ^ 1u)] = x;
                }
                __syncthreads();
                ph ^= 1u;
            }
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) v.tag[t] = k32[2 * t + ph];
            __syncthreads();
            // ---- which runs still hold two different prev symbols (flag at the run head) ----
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) k32[2 * t] = 0u;
            __syncthreads();
            for (u32 t = threadIdx.x + 1; t < vlen; t += blockDim.x) {
                const u32 h = v.tag[t];
                if (h == v.tag[t - 1] && ((v.ent[t] ^ v.ent[t - 1]) & 15ull)) k32[2 * h] = 1u;
            }
            __syncthreads();
            // ---- unresolved runs, reported by their last entry: long ones go to the next round, short ones are
            //      listed and ordered right away, one warp each, by direct comparisons ----
            for (u32 t = threadIdx.x; t < vlen; t += blockDim.x) {
                const u32 h = v.tag[t];
                if ((t + 1 == vlen || v.tag[t + 1] != h) && k32[2 * h]) {
                    const u32 size = t + 1 - h;
                    if (size > 32) {
                        WorkItem nw;
                        nw.off = voff + h; nw.len = size; nw.depth = depth + 32;
                        push_item(next, nw);
                    } else {
                        s_list[atomicAdd(&s_cnt[3], 1u)] = (size << 16) | h;
                    }
                }
            }
            __syncthreads();
            const u32 n_list = s_cnt[3];
            for (u32 q = threadIdx.x >> 5; q < n_list; q += blockDim.x >> 5) {
                const u32 x = s_list[q];
                warp_rank_short(sp, v.ent + (x & 0xffffu), x >> 16, depth + 32);
            }
            break;
        }
        __syncthreads();
        if (!in_hbm) {
            for (u32 t = threadIdx.x; t < len; t += blockDim.x) g.ent[t] = s.ent[t];
        }
        __syncthreads();
    }
}

}  // namespace

// d_work: 8 counter words + 4 lists of n_branch entries each (u32) => 4 * n_branch + 16 words
int k_sort_blue(u64* blue, BranchTable bt, SpView sp, u32* d_work, cudaStream_t st) {
    if (bt.n_blue == 0 || bt.n_branch == 0) return 0;
    u32* d_sum = nullptr;
    struct SumFree {                        // lives outside the build arena: hand it back on every exit path
        cudaStream_t st;
        u32** p;
        ~SumFree() { if (*p) cudaFreeAsync(*p, st); }
    } sum_guard{st, &d_sum};
    {
        const u64 nsw = sp.n_codes / 32 + 2, sum_words = (nsw >> 10) + 4;
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_sum), sum_words * 4, st));
        CUDA_TRY(cudaMemsetAsync(d_sum, 0, sum_words * 4, st));
        sep_summary_kernel<<<(unsigned)((nsw + TPB - 1) / TPB), TPB, 0, st>>>(sp.sep, nsw, d_sum);
        sp.sep_sum = d_sum;
    }
    u32* counts = d_work;
    u32* small = d_work + 8;
    u32* mid = small + bt.n_branch;
    u32* block = mid + bt.n_branch;
    u32* huge = block + bt.n_branch;
    CUDA_TRY(cudaMemsetAsync(counts, 0, 32, st));
    bin_segments_kernel<<<(unsigned)((bt.n_branch + TPB - 1) / TPB), TPB, 0, st>>>(bt, counts, small, mid, block, huge);
    u32 h[4];
    CUDA_TRY(cudaMemcpyAsync(h, counts, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int launched = 2;
    auto blocks_for = [](u32 warps) { u32 b = (warps + WARPS - 1) / WARPS; return b > 148u * 16u ? 148u * 16u : (b ? b : 1u); };
    if (h[0]) { sort_small_kernel<<<blocks_for(h[0]), TPB, 0, st>>>(blue, bt, sp, small, counts + 0); ++launched; }
    if (h[1]) { sort_mid_kernel<<<blocks_for(h[1]), TPB, 0, st>>>(blue, bt, sp, mid, counts + 1); ++launched; }
    const u32 n_big = h[2] + h[3];
    if (n_big) {
        // round-based refinement of the segments beyond one warp's shared-memory slice
        auto refine_small = refine_kernel<BIG_TPB, CHUNK, 2>;
        auto refine_tiny = refine_kernel<TINY_TPB, TINY, 16>;    // 8, 10 or 12 blocks per SM measured the same
        static bool attr_done[64] = {};        // per device: function attributes belong to the device's context
        int cur_dev = 0;
        cudaGetDevice(&cur_dev);
        if (!attr_done[cur_dev & 63]) {
            CUDA_TRY(cudaFuncSetAttribute(refine_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChunkSmem));
            attr_done[cur_dev & 63] = true;
        }
        const u64 cap_tiny64 = bt.n_blue / 16 + n_big + 16, cap_huge64 = bt.n_blue / SPLIT_ABOVE + n_big + 16;
        const u64 cap_small64 = bt.n_blue / TINY + cap_huge64 + 16;
        if (cap_tiny64 > 0xffffffffull) { set_error("internal: K10 work list too large"); return -1; }
        const u32 cap_tiny = (u32)cap_tiny64, cap_small = (u32)cap_small64, cap_huge = (u32)cap_huge64;
        const u64 per_set = (u64)cap_tiny + cap_small + cap_huge;
        WorkItem* lists = nullptr;
        u32* d_cnt = nullptr;
        u64* g_key = nullptr;
        struct AsyncFree {                  // the work lists live outside the build arena: hand them back on every exit path
            cudaStream_t st;
            void** p[3];
            ~AsyncFree() { for (void** q : p) if (*q) cudaFreeAsync(*q, st); }
        } guard{st, {reinterpret_cast<void**>(&lists), reinterpret_cast<void**>(&d_cnt), reinterpret_cast<void**>(&g_key)}};
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&lists), 2 * per_set * sizeof(WorkItem), st));
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_cnt), 64, st));
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&g_key), bt.n_blue * 12 + 64, st));
        u32* g_tag = g_key ? reinterpret_cast<u32*>(g_key + bt.n_blue) : nullptr;
        WorkLists cur{lists, lists + cap_small, lists + cap_small + cap_huge, d_cnt, cap_small, cap_huge, cap_tiny};
        WorkLists nxt{lists + per_set, lists + per_set + cap_small, lists + per_set + cap_small + cap_huge, d_cnt + 4,
                      cap_small, cap_huge, cap_tiny};
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 32, st));
        if (h[2]) seed_items_kernel<<<(h[2] + TPB - 1) / TPB, TPB, 0, st>>>(bt, block, h[2], cur);
        if (h[3]) seed_items_kernel<<<(h[3] + TPB - 1) / TPB, TPB, 0, st>>>(bt, huge, h[3], cur);
        launched += (h[2] ? 1 : 0) + (h[3] ? 1 : 0);
        u32 n_cur[3] = {1, 1, 1};                         // the exact counts are on the device
        constexpr int kMaxRounds = 100000;
        int round = 0;
        for (; (n_cur[0] || n_cur[1] || n_cur[2]) && round < kMaxRounds; ++round) {
            CUDA_TRY(cudaMemsetAsync(nxt.cnt, 0, 16, st));
            if (n_cur[1]) {
                const u32 grid = 148u * 2u;
                split_kernel<<<grid, BIG_TPB, 0, st>>>(blue, sp, cur.huge, cur.cnt + 1, cur, nxt, g_key, g_tag);
                ++launched;
            }
            refine_small<<<148u * 2u, BIG_TPB, kChunkSmem, st>>>(blue, sp, cur.small, cur.cnt + 0, nxt, g_key, g_tag);
            refine_tiny<<<148u * 16u, TINY_TPB, (size_t)TINY * 20, st>>>(blue, sp, cur.tiny, cur.cnt + 2, nxt, g_key, g_tag);
            launched += 2;
            u32 n_fin[8];
            CUDA_TRY(cudaMemcpyAsync(n_fin, d_cnt, 32, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const u32* c = cur.cnt == d_cnt ? n_fin : n_fin + 4;
            const u32* n = cur.cnt == d_cnt ? n_fin + 4 : n_fin;
            if (c[0] > cap_small || c[1] > cap_huge || c[2] > cap_tiny || n[0] > cap_small || n[1] > cap_huge || n[2] > cap_tiny) {
                set_error("internal: K10 work list overflow");
                return -1;
            }
            if (getenv("DEBWT_K10_DEBUG"))
                fprintf(stderr, "K10 round %d: small %u tiny %u (huge split %u) -> next small %u huge %u tiny %u\n", round, c[0], c[2],
                        n_cur[1], n[0], n[1], n[2]);
            n_cur[0] = n[0]; n_cur[1] = n[1]; n_cur[2] = n[2];
            WorkLists t = cur; cur = nxt; nxt = t;
        }
        if (n_cur[0] || n_cur[1] || n_cur[2]) {
            set_error("K10: segmented sort did not finish within the round limit (work items remain)");
            return -1;
        }
    }
    DEBWT_COUNT(launched);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace debwt
