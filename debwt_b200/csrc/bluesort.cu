// K10: segmented sort of the blue entries of every multi-in k-mer by their branch-code strings.
//
// Replaces sortBlue / multiQuickSort / myQsort / cmpSP (reference src/sortBlue.c:10-280).  An entry is
// (spIndex << 4) | prev; entries of one k-mer are ordered by the code string that starts at spIndex,
// codes compared 32 at a time (one u64), '#' > T and equal '#' compared through, '$' largest
// (src/sortBlue.c:109-173).
//
//   * segments of <= 32 entries (the overwhelming majority): one warp per segment, rank by counting;
//   * larger segments: one thread block per segment, same-direction bitonic network (virtual +inf
//     padding, so no scratch), in shared memory up to 2048 entries, in place in HBM beyond that.
#include "stages.cuh"

namespace debwt {

namespace {

constexpr int TPB = 256;
constexpr int BIG_TPB = 512;
constexpr int SMEM_SEG = 2048;

__device__ __forceinline__ u32 fetch_sep(const u32* __restrict__ sep, u64 s) {
    const u64 i = s >> 5;
    const u32 sh = (u32)(s & 31);
    const u32 lo = sep[i];
    if (sh == 0) return lo;
    return (lo >> sh) | (sep[i + 1] << (32 - sh));
}

// strict "string at sa < string at sb"; sa != sb
__device__ __forceinline__ bool sp_less(const SpView& v, u64 sa, u64 sb) {
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

__global__ void __launch_bounds__(TPB) sort_blue_small_kernel(u64* __restrict__ blue, BranchTable bt, SpView sp,
                                                             u32* __restrict__ big_list, u32* __restrict__ big_count) {
    const int lane = threadIdx.x & 31;
    const u64 nwarps = (u64)gridDim.x * (TPB / 32);
    for (u64 b = ((u64)blockIdx.x * TPB + threadIdx.x) >> 5; b < bt.n_branch; b += nwarps) {
        if (!(bt.kmer[b] & 2ull)) continue;
        const u32 off = bt.blue[b];
        const u32 len = bt.blue[b + 1] - off;
        if (len <= 1) continue;
        if (len > 32) {
            if (lane == 0) big_list[atomicAdd(big_count, 1u)] = (u32)b;
            continue;
        }
        const u64 e = (u32)lane < len ? blue[(u64)off + lane] : 0;
        const u64 s = e >> 4;
        // every prev symbol equal -> any order gives the same BWT (src/sortBlue.c:192-219)
        const u32 c0 = __shfl_sync(0xffffffffu, (u32)(e & 15ull), 0);
        if (__all_sync(0xffffffffu, (u32)lane >= len || (u32)(e & 15ull) == c0)) continue;
        u32 rank = 0;
        for (u32 j = 0; j < len; ++j) {
            __syncwarp();
            const u64 sj = __shfl_sync(0xffffffffu, s, j);
            if ((u32)lane < len && j != (u32)lane && sp_less(sp, sj, s)) ++rank;
        }
        __syncwarp();
        if ((u32)lane < len) blue[(u64)off + rank] = e;
        __syncwarp();
    }
}

template <typename Ptr>
__device__ __forceinline__ void cmpswap(Ptr a, u64 i, u64 l, const SpView& sp) {
    const u64 x = a[i], y = a[l];
    if (sp_less(sp, y >> 4, x >> 4)) { a[i] = y; a[l] = x; }
}

template <typename Ptr>
__device__ void bitonic_same_direction(Ptr a, u64 len, const SpView& sp) {
    u64 P = 1;
    while (P < len) P <<= 1;
    for (u64 k = 2; k <= P; k <<= 1) {
        const u64 half = k >> 1;
        for (u64 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {      // mirror step
            const u64 i = (t / half) * k + (t % half);
            const u64 l = i ^ (k - 1);
            if (l < len) cmpswap(a, i, l, sp);
        }
        __syncthreads();
        for (u64 j = k >> 2; j > 0; j >>= 1) {                          // half cleaners
            for (u64 t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const u64 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u64 l = i | j;
                if (l < len) cmpswap(a, i, l, sp);
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(BIG_TPB) sort_blue_big_kernel(u64* __restrict__ blue, BranchTable bt, SpView sp,
                                                               const u32* __restrict__ big_list,
                                                               const u32* __restrict__ big_count) {
    __shared__ u64 s_seg[SMEM_SEG];
    const u32 nbig = *big_count;
    for (u32 idx = blockIdx.x; idx < nbig; idx += gridDim.x) {
        const u32 b = big_list[idx];
        const u64 off = bt.blue[b];
        const u64 len = bt.blue[b + 1] - off;
        u64* a = blue + off;
        if (len <= SMEM_SEG) {
            for (u64 t = threadIdx.x; t < len; t += blockDim.x) s_seg[t] = a[t];
            __syncthreads();
            bitonic_same_direction(s_seg, len, sp);
            for (u64 t = threadIdx.x; t < len; t += blockDim.x) a[t] = s_seg[t];
            __syncthreads();
        } else {
            bitonic_same_direction(a, len, sp);
        }
    }
}

}  // namespace

int k_sort_blue(u64* blue, BranchTable bt, SpView sp, u32* d_work, cudaStream_t st) {
    if (bt.n_blue == 0 || bt.n_branch == 0) return 0;
    u32* big_count = d_work;
    u32* big_list = d_work + 4;
    CUDA_TRY(cudaMemsetAsync(big_count, 0, 16, st));
    u64 blocks = (bt.n_branch + (TPB / 32) - 1) / (TPB / 32);
    if (blocks > 148 * 32) blocks = 148 * 32;
    sort_blue_small_kernel<<<(unsigned)blocks, TPB, 0, st>>>(blue, bt, sp, big_list, big_count);
    CUDA_TRY(cudaGetLastError());
    sort_blue_big_kernel<<<148 * 2, BIG_TPB, 0, st>>>(blue, bt, sp, big_list, big_count);
    DEBWT_COUNT(2);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace debwt
