// Synthetic-genome generators on the device, bit-identical to debwt_b200/synth.py (splitmix64,
// 32 bases per 64-bit output taken from the top bits down).  Bench / test plumbing: a 3.1 Gbp human-scale workload is
// generated in HBM in milliseconds instead of minutes of numpy on every rank.  Not part of the BWT path.
#include "../../include/debwt_b200.h"
#include "common.cuh"

namespace debwt {

namespace {

constexpr int TPB = 256;

__host__ __device__ __forceinline__ u64 splitmix_at(u64 seed, u64 index1) {      // output number index1 (1-based) of the stream
    u64 z = seed + index1 * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u32 code_at(u64 seed, u64 p) {                         // synth.random_codes(seed, ...)[p]
    return (u32)(splitmix_at(seed, (p >> 5) + 1) >> (2 * (31 - (p & 31)))) & 3u;
}
__device__ __forceinline__ u8 base_of(u32 code) { return (u8)((0x54474341u >> (8 * code)) & 255u); }   // "ACGT"
__device__ __forceinline__ u32 code_of(u8 c) {
    const u32 u = c & 0xDFu;
    const u32 x = (u >> 1) & 3u;
    return x ^ (x >> 1);
}

__global__ void __launch_bounds__(TPB) random_bases_kernel(u8* __restrict__ out, u64 n, u64 seed) {
    const u64 w = (u64)blockIdx.x * TPB + threadIdx.x;
    const u64 base = w * 32;
    if (base >= n) return;
    const u64 z = splitmix_at(seed, w + 1);
    const u64 lim = n - base < 32 ? n - base : 32;
    for (u64 j = 0; j < lim; ++j) out[base + j] = base_of((u32)(z >> (2 * (31 - j))) & 3u);
}

// synth._insert_family: later copies overwrite earlier ones => every position takes the LAST copy that covers it
__global__ void __launch_bounds__(TPB) family_owner_kernel(u32* __restrict__ owner, u64 n, u64 seed, u64 copies, u64 len) {
    const u64 idx = (u64)blockIdx.x * TPB + threadIdx.x;
    if (idx >= copies * len) return;
    const u64 i = idx / len, p = idx - i * len;
    const u64 o = splitmix_at(seed + (1ull << 33), i + 1) % (n - len);
    atomicMax(owner + o + p, (u32)(i + 1));
}

__global__ void __launch_bounds__(TPB) family_apply_kernel(u8* __restrict__ seq, u32* __restrict__ owner, u64 n, u64 seed, u64 len,
                                                          u64 thr) {
    const u64 pos = (u64)blockIdx.x * TPB + threadIdx.x;
    if (pos >= n) return;
    const u32 w = owner[pos];
    if (!w) return;
    owner[pos] = 0;                                                            // ready for the next family
    const u64 i = w - 1;
    const u64 o = splitmix_at(seed + (1ull << 33), i + 1) % (n - len);
    const u64 p = pos - o;
    u32 code = code_at(seed, p);
    if (thr) {                                                                 // synth._mutate_many, row seed = i + seed + 2^34
        const u64 r = splitmix_at(i + seed + (1ull << 34), p + 1);
        if ((r >> 11) < thr) code = (code + (u32)((r & 0x7FFull) % 3ull) + 1u) & 3u;
    }
    seq[pos] = base_of(code);
}

__global__ void __launch_bounds__(TPB) mutate_kernel(const u8* __restrict__ in, u8* __restrict__ out, u64 n, u64 seed, u64 thr) {
    const u64 pos = (u64)blockIdx.x * TPB + threadIdx.x;
    if (pos >= n) return;
    u8 c = in[pos];
    const u64 r = splitmix_at(seed, pos + 1);                                  // synth._mutate
    if ((r >> 11) < thr) c = base_of((code_of(c) + (u32)((r & 0x7FFull) % 3ull) + 1u) & 3u);
    out[pos] = c;
}

int on_device(int device) {
    if (debwt_device_count() <= device || device < 0) {
        set_error("no such CUDA device (this library has no CPU fallback)");
        return -1;
    }
    CUDA_TRY(cudaSetDevice(device));
    return 0;
}

inline unsigned grid_for(u64 work) { return (unsigned)((work + TPB - 1) / TPB); }

}  // namespace
}  // namespace debwt

using namespace debwt;

extern "C" {

int debwt_synth_random_bases(int device, void* d_out, uint64_t n, uint64_t seed) {
    if (on_device(device)) return -1;
    if (n == 0) return 0;
    random_bases_kernel<<<grid_for((n + 31) / 32), TPB>>>(static_cast<u8*>(d_out), n, seed);
    DEBWT_COUNT(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

int debwt_synth_insert_family(int device, void* d_seq, uint64_t n, uint64_t seed, uint64_t copies, uint64_t length, uint64_t thr,
                              void* d_owner_zeroed) {
    if (on_device(device)) return -1;
    if (copies == 0 || n <= length) return 0;                                  // synth._insert_family's early return
    if (copies >= 0xFFFFFFFFull) { set_error("too many copies"); return -1; }
    u32* owner = static_cast<u32*>(d_owner_zeroed);
    const u64 work = copies * length;
    if (work / TPB >= 0x7FFFFFFFull) { set_error("family too large for one launch"); return -1; }
    family_owner_kernel<<<grid_for(work), TPB>>>(owner, n, seed, copies, length);
    family_apply_kernel<<<grid_for(n), TPB>>>(static_cast<u8*>(d_seq), owner, n, seed, length, thr);
    DEBWT_COUNT(2);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

int debwt_synth_mutate(int device, const void* d_in, void* d_out, uint64_t n, uint64_t seed, uint64_t thr) {
    if (on_device(device)) return -1;
    if (n == 0) return 0;
    mutate_kernel<<<grid_for(n), TPB>>>(static_cast<const u8*>(d_in), static_cast<u8*>(d_out), n, seed, thr);
    DEBWT_COUNT(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
