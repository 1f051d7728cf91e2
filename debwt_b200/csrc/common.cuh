// Shared device/host helpers for the deBWT-B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

namespace debwt {

constexpr int KMER = 32;        // counted (k+1)-mer length == one 64-bit key   (reference src/main.c:43)
constexpr int KNODE = 31;       // de Bruijn node length k                      (reference src/main.c:44)

// group mask bits (one u16 per sorted key)
constexpr u32 GM_IN_SEP = 1u << 4;     // bits 0..3: in-base seen; bit 4: predecessor is '#'/'$'
constexpr u32 GM_OUT_SHIFT = 8;        // bits 8..11: out-base seen
constexpr u32 GM_OUT_TAIL = 1u << 12;  // the k-mer also occurs right before a separator

void set_error(const std::string& msg);
extern unsigned g_launches;        // kernel launches since the last debwt_build() started
extern unsigned long long g_launches_total;   // kernel launches in this process
#define DEBWT_COUNT(n) (debwt::g_launches += (n), debwt::g_launches_total += (n))

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            char _b[512];                                                                  \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                     cudaGetErrorString(_e));                                              \
            debwt::set_error(_b);                                                          \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

__host__ __device__ __forceinline__ bool gm_multi_in(u32 m) {
    u32 b = m & 15u;
    return (b & (b - 1)) != 0 || (m & GM_IN_SEP);
}
__host__ __device__ __forceinline__ bool gm_multi_out(u32 m) {
    u32 b = (m >> GM_OUT_SHIFT) & 15u;
    return (b & (b - 1)) != 0 || (m & GM_OUT_TAIL);
}

// ---- packed 2-bit text access: 32 symbols per u64, symbol j at bits 2*(31-(j&31)) -------------
// 32 symbols starting at symbol position p (needs words[p>>5] and words[(p>>5)+1] readable)
__host__ __device__ __forceinline__ u64 text_window32(const u64* __restrict__ w, u64 p) {
    u64 i = p >> 5;
    u32 s = (u32)(p & 31) * 2;
    u64 a = w[i];
    if (s == 0) return a;
    return (a << s) | (w[i + 1] >> (64 - s));
}
__host__ __device__ __forceinline__ u32 text_symbol(const u64* __restrict__ w, u64 p) {
    return (u32)(w[p >> 5] >> (2 * (31 - (p & 31)))) & 3u;
}

// ---- loads the compiler must keep together and in program order (memory-level parallelism by hand) ----
__device__ __forceinline__ u32 ld_nc_u32(const u32* p) {
    u32 v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ ulonglong2 ld_nc_u64x2(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

// ---- streaming loads / stores --------------------------------------------------------------
__device__ __forceinline__ u64 ld_stream(const u64* p) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(u64* p, u64 v) {
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(v));
}
__device__ __forceinline__ u64 ld_volatile(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 16-bit atomic OR through the enclosing aligned 32-bit word
__device__ __forceinline__ void atomic_or_u16(u16* base, u64 idx, u32 bits) {
    u32* w = reinterpret_cast<u32*>(base) + (idx >> 1);
    atomicOr(w, bits << ((idx & 1) * 16));
}

// ---- block-wide exclusive scan (u32), THREADS a multiple of 32, <= 1024 -----------------------
template <int THREADS>
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total, u32* smem /* >= 33 u32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();            // protect smem reuse across calls
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = (lane < THREADS / 32) ? smem[lane] : 0;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < THREADS / 32) smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    if (total) *total = smem[32];
    return smem[warp] + inc - v;
}

// lower_bound / upper_bound on a sorted u64 array, range [lo, hi)
__host__ __device__ __forceinline__ u64 lower_bound_u64(const u64* __restrict__ a, u64 lo, u64 hi, u64 key) {
    while (lo < hi) {
        u64 mid = lo + ((hi - lo) >> 1);
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__host__ __device__ __forceinline__ u64 upper_bound_u64(const u64* __restrict__ a, u64 lo, u64 hi, u64 key) {
    while (lo < hi) {
        u64 mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace debwt
