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
//   <= 4096 : one thread block, shared memory, same network;
//   larger  : one thread block; every aligned 4096-entry block of the segment is staged in shared memory
//             once per merge level, only the stages with a longer partner distance run over HBM.
// The network uses virtual +inf padding (all compare-exchanges point the same way), so no segment
// needs scratch for padding.  Segments whose prev symbols are all equal are skipped
// (src/sortBlue.c:192-219): any order gives the same BWT.
#include "stages.cuh"

namespace debwt {

namespace {

constexpr int TPB = 256;
constexpr int WARPS = TPB / 32;
constexpr int MID_SEG = 128;
constexpr int BIG_TPB = 512;
constexpr int SMEM_SEG = 4096;     // == CHUNK: largest segment sorted entirely in shared memory

__device__ __forceinline__ u32 fetch_sep(const u32* __restrict__ sep, u64 s) {
    const u64 i = s >> 5;
    const u32 sh = (u32)(s & 31);
    const u32 lo = sep[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (sep[i + 1] << (32 - sh));
}

// strict "string at sa < string at sb" starting the comparison `skip` codes in; sa != sb
__device__ __forceinline__ bool sp_less_from(const SpView& v, u64 sa, u64 sb, u32 skip) {
    sa += skip;
    sb += skip;
    for (;;) {
        const u64 ca = text_window32(v.codes, sa), cb = text_window32(v.codes, sb);
        const u32 fa = fetch_sep(v.sep, sa), fb = fetch_sep(v.sep, sb);
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
    c.plain = fetch_sep(v.sep, s) == 0;
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

// One network stage over global arrays (mirror step when `mirror`, half cleaner otherwise)
__device__ __forceinline__ void global_stage(u64* ent, u64* wrd, u8* pln, u64 len, u64 P, u64 k, u64 j, bool mirror,
                                             const SpView& sp) {
    const u64 half = k >> 1;
    for (u64 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        u64 i, l;
        if (mirror) { i = (t / half) * k + (t % half); l = i ^ (k - 1); }
        else { i = ((t & ~(j - 1)) << 1) | (t & (j - 1)); l = i | j; }
        if (l < len) {
            const u64 ei = ent[i], el = ent[l], wi = wrd[i], wl = wrd[l];
            const bool pi = pln[i], pl = pln[l];
            if (entry_less(sp, el, wl, pl, ei, wi, pi)) {
                ent[i] = el; ent[l] = ei; wrd[i] = wl; wrd[l] = wi; pln[i] = pl; pln[l] = pi;
            }
        }
    }
    __syncthreads();
}

// 129..CHUNK entries: one block, everything in shared memory.  Larger segments: the network stages whose
// partner distance is below CHUNK/2 touch only one aligned CHUNK-sized block, so each block of the
// segment is loaded into shared memory once per merge level and finishes there; only the few stages
// with a longer partner distance run over HBM (20 instead of 136 for a 64 K segment).
constexpr int CHUNK = 4096;
constexpr size_t kChunkSmem = (size_t)CHUNK * 17;

__global__ void __launch_bounds__(BIG_TPB) sort_block_kernel(u64* __restrict__ blue, BranchTable bt, SpView sp,
                                                            const u32* __restrict__ list, const u32* __restrict__ count,
                                                            u64* __restrict__ g_wrd, u8* __restrict__ g_pln, bool in_hbm) {
    extern __shared__ __align__(16) unsigned char blk_smem[];
    u64* s_ent = reinterpret_cast<u64*>(blk_smem);
    u64* s_wrd = s_ent + CHUNK;
    u8* s_pln = reinterpret_cast<u8*>(s_wrd + CHUNK);
    __shared__ int s_same;
    const u32 n = *count;
    for (u32 idx = blockIdx.x; idx < n; idx += gridDim.x) {
        const u32 b = list[idx];
        const u64 off = bt.blue[b];
        const u64 len = bt.blue[b + 1] - off;
        if (threadIdx.x == 0) s_same = 1;
        __syncthreads();
        const u32 c0 = (u32)(blue[off] & 15ull);
        bool same = true;
        u64* ent = blue + off;
        u64* wrd = in_hbm ? g_wrd + off : s_wrd;
        u8* pln = in_hbm ? g_pln + off : s_pln;
        for (u64 t = threadIdx.x; t < len; t += blockDim.x) {
            const u64 e = ent[t];
            const Cached c = cache_of(sp, e);
            if (!in_hbm) s_ent[t] = e;
            wrd[t] = c.word; pln[t] = c.plain;
            same &= (u32)(e & 15ull) == c0;
        }
        if (!same) s_same = 0;
        __syncthreads();
        if (s_same) continue;
        if (!in_hbm) {
            bitonic_cached(s_ent, s_wrd, s_pln, len, sp, threadIdx.x, blockDim.x, [] { __syncthreads(); });
            for (u64 t = threadIdx.x; t < len; t += blockDim.x) ent[t] = s_ent[t];
            __syncthreads();
            continue;
        }
        u64 P = 1;
        while (P < len) P <<= 1;
        // level 0: sort every aligned CHUNK block in shared memory; later levels: long stages over HBM, then
        // the stages with partner distance < CHUNK/2 block by block in shared memory
        for (u64 k = CHUNK; k <= P; k <<= 1) {
            if (k > CHUNK) {
                global_stage(ent, wrd, pln, len, P, k, 0, true, sp);
                for (u64 j = k >> 2; j >= CHUNK; j >>= 1) global_stage(ent, wrd, pln, len, P, k, j, false, sp);
            }
            for (u64 c0b = 0; c0b < len; c0b += CHUNK) {
                const u64 clen = (len - c0b < CHUNK) ? len - c0b : CHUNK;
                for (u64 t = threadIdx.x; t < clen; t += blockDim.x) {
                    s_ent[t] = ent[c0b + t]; s_wrd[t] = wrd[c0b + t]; s_pln[t] = pln[c0b + t];
                }
                __syncthreads();
                if (k == CHUNK) {
                    bitonic_cached(s_ent, s_wrd, s_pln, clen, sp, threadIdx.x, blockDim.x, [] { __syncthreads(); });
                } else {
                    // half cleaners j = CHUNK/2 .. 1 of merge level k, restricted to this block
                    for (u64 j = CHUNK >> 1; j > 0; j >>= 1) {
                        for (u64 t = threadIdx.x; t < (CHUNK >> 1); t += blockDim.x) {
                            const u64 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const u64 l = i | j;
                            if (l < clen) {
                                const u64 ei = s_ent[i], el = s_ent[l], wi = s_wrd[i], wl = s_wrd[l];
                                const bool pi = s_pln[i], pl = s_pln[l];
                                if (entry_less(sp, el, wl, pl, ei, wi, pi)) {
                                    s_ent[i] = el; s_ent[l] = ei; s_wrd[i] = wl; s_wrd[l] = wi; s_pln[i] = pl; s_pln[l] = pi;
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                for (u64 t = threadIdx.x; t < clen; t += blockDim.x) {
                    ent[c0b + t] = s_ent[t]; wrd[c0b + t] = s_wrd[t]; pln[c0b + t] = s_pln[t];
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace

// d_work: 8 counter words + 4 lists of n_branch entries each (u32) => 4 * n_branch + 16 words
int k_sort_blue(u64* blue, BranchTable bt, SpView sp, u32* d_work, cudaStream_t st) {
    if (bt.n_blue == 0 || bt.n_branch == 0) return 0;
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
    int launched = 1;
    auto blocks_for = [](u32 warps) { u32 b = (warps + WARPS - 1) / WARPS; return b > 148u * 16u ? 148u * 16u : (b ? b : 1u); };
    if (h[0]) { sort_small_kernel<<<blocks_for(h[0]), TPB, 0, st>>>(blue, bt, sp, small, counts + 0); ++launched; }
    if (h[1]) { sort_mid_kernel<<<blocks_for(h[1]), TPB, 0, st>>>(blue, bt, sp, mid, counts + 1); ++launched; }
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(sort_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChunkSmem));
        attr_done = true;
    }
    if (h[2]) {
        sort_block_kernel<<<h[2] < 148u * 3u ? h[2] : 148u * 3u, BIG_TPB, kChunkSmem, st>>>(blue, bt, sp, block, counts + 2, nullptr, nullptr, false);
        ++launched;
    }
    if (h[3]) {
        u64* g_wrd = nullptr;
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&g_wrd), bt.n_blue * 9 + 64, st));
        u8* g_pln = reinterpret_cast<u8*>(g_wrd + bt.n_blue);
        sort_block_kernel<<<h[3] < 148u * 3u ? h[3] : 148u * 3u, BIG_TPB, kChunkSmem, st>>>(blue, bt, sp, huge, counts + 3, g_wrd, g_pln, true);
        CUDA_TRY(cudaFreeAsync(g_wrd, st));
        ++launched;
    }
    DEBWT_COUNT(launched);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace debwt
