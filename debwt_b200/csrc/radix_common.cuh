// Device helpers shared by the two onesweep kernels (radix_sort.cu: register-staged tiles; radix_sort_tma.cu:
// TMA bulk-copy staged tiles).
#pragma once
#include "radix_sort.cuh"

namespace debwt {
namespace radix {

constexpr int RADIX = 256;
constexpr int PASSES = 8;
constexpr u64 LB_VALUE_MASK = (1ull << 56) - 1;
constexpr u64 LB_EPOCH_MASK = 63ull << 56;
constexpr u64 LB_AGG = 1ull << 62;
constexpr u64 LB_INCL = 2ull << 62;

template <int PASS>
__device__ __forceinline__ u32 digit_of(u64 key) {
    const u32 w = PASS < 4 ? (u32)key : (u32)(key >> 32);
    constexpr int s = 8 * (PASS & 3);
    return s == 24 ? (w >> 24) : ((w >> s) & 255u);
}

// lanes of the warp that hold the same 8-bit digit: one ballot per digit bit
__device__ __forceinline__ u32 match_digit(u32 d) {
    u32 peers;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 v, t;\n\t"
        "and.b32 t, %1, 1;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; mov.b32 %0, v;\n\t"
        "and.b32 t, %1, 2;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 4;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 8;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 16;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 32;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 64;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "and.b32 t, %1, 128; setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; @!p not.b32 v, v; and.b32 %0, %0, v;\n\t"
        "}"
        : "=&r"(peers)
        : "r"(d));
    return peers;
}

// Decoupled look-back: exclusive prefix of this digit over all earlier tiles.  `lb` = this tile's look-back word of
// this digit, predecessors at lb - i * RADIX.  The predecessors are fetched BATCH at a time: the loads of one batch
// are independent, so a run of count-only predecessors costs one memory round trip per batch instead of per tile.
// Tiles are claimed at R per microsecond and a round trip takes L: the walk only stays short when R * L / BATCH is
// well below 1 (it is about 1 for BATCH = 4 on B200, which lets walks grow to the number of tiles in flight).
template <int BATCH>
__device__ __forceinline__ void lookback_fetch(const u64* p, u32 left, u64 (&v)[BATCH]) {
#pragma unroll
    for (int i = 0; i < BATCH; ++i) v[i] = ((u32)i < left) ? ld_volatile(p - (size_t)i * RADIX) : 0;
}

// SLEEP: nanoseconds to back off when a round made no progress (0 = spin).  PRE: the caller fetched the first batch
// itself (lookback_fetch(lb - RADIX, tile, pre)) some time ago -- e.g. before the shared-memory reorder, so that the
// first round trip is hidden behind it.
template <int BATCH, int SLEEP = 0, bool PRE = false>
__device__ __forceinline__ u64 lookback_exclusive(const u64* lb, u32 tile, u64 epoch, const u64* pre = nullptr) {
    u64 excl = 0;
    u32 left = tile;                       // predecessors not yet consumed
    const u64* p = lb - RADIX;             // nearest unconsumed predecessor
    u32 spins = 0;
    bool done = false;
    bool first = PRE;
    while (!done) {
        u64 v[BATCH];
        if (first) {
#pragma unroll
            for (int i = 0; i < BATCH; ++i) v[i] = pre[i];
            first = false;
        } else {
            lookback_fetch<BATCH>(p, left, v);
        }
        int used = 0;
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            if (done || used != i) continue;                       // stop at the first unpublished entry
            if ((u32)i >= left) continue;
            const u64 x = v[i];
            if ((x & LB_EPOCH_MASK) != epoch || (x >> 62) == 0) continue;
            excl += x & LB_VALUE_MASK;
            used = i + 1;
            if ((x >> 62) == 2) done = true;
        }
        p -= (size_t)used * RADIX;
        left -= used;
        if (used == 0) {
            if (SLEEP) __nanosleep(SLEEP);
            if (++spins > (1u << 24)) __trap();                    // never hang the device on a bug
        }
    }
    return excl;
}

// ---- named barriers (a subset of the CTA's warps) ------------------------------------------------
template <int ID, int NT>
__device__ __forceinline__ void bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory"); }
template <int ID, int NT>
__device__ __forceinline__ void bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NT) : "memory"); }

// block_exclusive_scan (common.cuh) over the first NT threads of the CTA, synchronised on named barrier ID
template <int NT, int ID>
__device__ __forceinline__ u32 group_exclusive_scan(u32 v, u32* smem /* >= 33 u32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    bar_sync<ID, NT>();
    if (lane == 31) smem[warp] = inc;
    bar_sync<ID, NT>();
    if (warp == 0) {
        u32 w = (lane < NT / 32) ? smem[lane] : 0;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NT / 32) smem[lane] = winc - w;
    }
    bar_sync<ID, NT>();
    return smem[warp] + inc - v;
}

// Look-back for DPT digits at once (digit j of this thread = dig0 + 32 j): the loads of all digits and of BATCH
// predecessors each are in flight together.  lb_row = look-back row of this tile.
template <int DPT, int BATCH>
__device__ __forceinline__ void lookback_multi(const u64* lb_row, int dig0, u32 tile, u64 epoch, u64 (&excl)[DPT]) {
    u32 used_tot[DPT];
    bool done[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) { excl[j] = 0; used_tot[j] = 0; done[j] = false; }
    u32 spins = 0;
    for (;;) {
        bool all = true;
#pragma unroll
        for (int j = 0; j < DPT; ++j) all = all && done[j];
        if (all) break;
        u64 v[DPT][BATCH];
#pragma unroll
        for (int j = 0; j < DPT; ++j)
#pragma unroll
            for (int i = 0; i < BATCH; ++i)
                v[j][i] = (!done[j] && used_tot[j] + (u32)i < tile)
                              ? ld_volatile(lb_row + dig0 + 32 * j - (size_t)(used_tot[j] + (u32)i + 1u) * RADIX)
                              : 0;
        bool progress = false;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            if (done[j]) continue;
            const u32 left = tile - used_tot[j];
            int used = 0;
            bool fin = false;
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                if (fin || used != i) continue;
                if ((u32)i >= left) continue;
                const u64 x = v[j][i];
                if ((x & LB_EPOCH_MASK) != epoch || (x >> 62) == 0) continue;
                excl[j] += x & LB_VALUE_MASK;
                used = i + 1;
                if ((x >> 62) == 2) fin = true;
            }
            used_tot[j] += (u32)used;
            done[j] = fin;
            progress = progress || used != 0;
        }
        if (!progress) {
            __nanosleep(64);
            if (++spins > (1u << 22)) __trap();                   // never hang the device on a bug
        }
    }
}

}  // namespace radix

// one digit pass with the TMA-staged persistent kernel (radix_sort_tma.cu); cfg >= TMA_CFG_BASE
constexpr int TMA_CFG_BASE = 32;
int tma_config_tile(int cfg);
int launch_tma_sweep(int cfg, const u64* in, u64* out, u64 n, int pass, const SortWorkspace& ws, cudaStream_t st);

}  // namespace debwt
