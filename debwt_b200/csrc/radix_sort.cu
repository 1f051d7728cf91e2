// K3: LSD radix sort of 64-bit (k+1)-mer keys -- the named roofline phase.
//
// Replaces the reference's bucket-by-12-bases + per-bucket qsort (reference src/mySort.c:98-176,
// multiDistri :371-401, multiThreadSort :203-238, cmpKmer :338-345).
//
// Design (sm_100a, HBM bound, no tensor cores):
//   * one histogram sweep builds all eight 256-bin digit histograms in shared memory (8 B/key read);
//   * eight "onesweep" passes, each reading every key once and writing it once (16 B/key):
//       - a tile of THREADS*ITEMS keys is loaded warp-striped (coalesced 256 B per warp load);
//       - keys are ranked inside the warp by warp-ballot matching (8 votes per key, one per digit bit)
//         against a per-warp shared-memory digit histogram;
//       - per-digit tile totals are chained across tiles by decoupled look-back (one 64-bit word
//         carries status + epoch + value, so no fences and no reset between passes); the look-back
//         runs after the tile has been reordered in shared memory, so its latency is hidden;
//       - the tile is written out bin by bin, so every warp store covers contiguous addresses.
//   B200 has ~2.4x the HBM bandwidth of H100 with about the same SM count, so the pass is
//   instruction-issue bound unless the per-key instruction count is kept near 2: the digit position
//   is a template parameter, the ballot sequence is hand-written, destinations are 32-bit indices.
//   Algorithmic traffic: 8 + 8*16 = 136 B/key.
#include "radix_common.cuh"

namespace debwt {

namespace {

using namespace radix;

template <int THREADS>
__global__ void __launch_bounds__(THREADS) radix_hist_kernel(const u64* __restrict__ keys, u64 n,
                                                            u64* __restrict__ ghist) {
    __shared__ u32 sh[PASSES * RADIX];
    for (int i = threadIdx.x; i < PASSES * RADIX; i += THREADS) sh[i] = 0;
    __syncthreads();
    const u64 nvec = n >> 1;
    const ulonglong2* kv = reinterpret_cast<const ulonglong2*>(keys);
    const u64 stride = (u64)gridDim.x * THREADS;
    u64 i = (u64)blockIdx.x * THREADS + threadIdx.x;
    auto acc = [&](u64 k) {
        const u32 lo = (u32)k, hi = (u32)(k >> 32);
        atomicAdd(&sh[0 * RADIX + (lo & 255u)], 1u);
        atomicAdd(&sh[1 * RADIX + ((lo >> 8) & 255u)], 1u);
        atomicAdd(&sh[2 * RADIX + ((lo >> 16) & 255u)], 1u);
        atomicAdd(&sh[3 * RADIX + (lo >> 24)], 1u);
        atomicAdd(&sh[4 * RADIX + (hi & 255u)], 1u);
        atomicAdd(&sh[5 * RADIX + ((hi >> 8) & 255u)], 1u);
        atomicAdd(&sh[6 * RADIX + ((hi >> 16) & 255u)], 1u);
        atomicAdd(&sh[7 * RADIX + (hi >> 24)], 1u);
    };
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        ulonglong2 a = __ldg(kv + i), b = __ldg(kv + i + stride), c = __ldg(kv + i + 2 * stride),
                   d = __ldg(kv + i + 3 * stride);
        acc(a.x); acc(a.y); acc(b.x); acc(b.y); acc(c.x); acc(c.y); acc(d.x); acc(d.y);
    }
    for (; i < nvec; i += stride) {
        ulonglong2 a = __ldg(kv + i);
        acc(a.x); acc(a.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc(keys[n - 1]);
    __syncthreads();
    for (int j = threadIdx.x; j < PASSES * RADIX; j += THREADS) {
        u32 c = sh[j];
        if (c) atomicAdd(&ghist[j], (u64)c);
    }
}

// counts -> exclusive bases, in place; skip[p] = 1 when one bin of pass p holds every key
__global__ void __launch_bounds__(RADIX) radix_scan_kernel(u64* __restrict__ ghist, u64 n, u32* __restrict__ skip) {
    __shared__ u64 s[RADIX];
    for (int p = 0; p < PASSES; ++p) {
        u64 c = ghist[p * RADIX + threadIdx.x];
        s[threadIdx.x] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 run = 0;
            u32 sk = 0;
            for (int d = 0; d < RADIX; ++d) {
                u64 v = s[d];
                if (v == n) sk = 1;
                s[d] = run;
                run += v;
            }
            skip[p] = sk;
        }
        __syncthreads();
        ghist[p * RADIX + threadIdx.x] = s[threadIdx.x];
        __syncthreads();
    }
}

template <int THREADS, int ITEMS>
struct SweepSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr size_t bytes = (size_t)TILE * 8 + (size_t)WARPS * RADIX * 4 + RADIX * 4 + RADIX * 4 + 40 * 4 +
                                    (size_t)(THREADS / RADIX) * RADIX * 4;
};

// PERSIST = false: one CTA per tile.  PERSIST = true: a resident CTA loops over tiles and issues the loads
// of its next tile right after the current one has been reordered into shared memory, so the load latency
// and the look-back wait of tile t overlap the global loads of tile t+1 (the key registers are free then).
// RANK = 1: the digit group's first lane bumps the warp histogram with one shared-memory atomic (ATOMS) and hands the old
// count to its peers by shuffle, instead of a warp-wide load + a leader store around two warp barriers.
template <int THREADS, int ITEMS, int MIN_BLOCKS, int PASS, bool PERSIST, int LB, int SLEEP, bool PRE, int RANK = 0>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_kernel(const u64* __restrict__ in, u64* __restrict__ out, u32 n, u32 ntiles, const u64* __restrict__ gbase,
                u64* __restrict__ lookback, u32* __restrict__ tile_counter, u64 epoch, u32* __restrict__ kidx, int kshift) {
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    static_assert(THREADS >= RADIX, "one thread per digit is assumed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* s_keys = reinterpret_cast<u64*>(smem_raw);
    u32* s_whist = reinterpret_cast<u32*>(s_keys + TILE);
    u32* s_binoff = s_whist + WARPS * RADIX;
    u32* s_goff = s_binoff + RADIX;
    u32* s_scan = s_goff + RADIX;   // 33 words + tile ids
    u32* s_part = s_scan + 40;      // [GROUPS][RADIX] partial digit totals
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32* wh = s_whist + warp * RADIX;
    const u32 lt = lanemask_lt();
    constexpr int GROUPS = THREADS / RADIX;               // 256 -> 1, 384 -> 1, 512 -> 2, 1024 -> 4
    constexpr int WPG = WARPS / GROUPS;                   // warps per group
    static_assert(WPG * GROUPS == WARPS, "warps must divide evenly over the digit groups");
    const int dg = tid & (RADIX - 1), grp = tid >> 8;

    u64 key[ITEMS];
    auto load_tile = [&](u32 t) {
        const u32 first = t * (u32)TILE + warp * (ITEMS * 32) + lane;
        const u64* src = in + first;
        if (t + 1 < ntiles) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) key[j] = ld_stream(src + j * 32);
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) key[j] = (first + j * 32 < n) ? ld_stream(src + j * 32) : ~0ull;
        }
    };

    if (tid == 0) s_scan[34] = atomicAdd(tile_counter, 1u);
    __syncthreads();
    u32 tile = s_scan[34];
    if (tile >= ntiles) return;
    load_tile(tile);

    for (;;) {
        for (int i = tid; i < WARPS * RADIX; i += THREADS) s_whist[i] = 0;
        if (PERSIST && tid == 0) s_scan[35] = atomicAdd(tile_counter, 1u);      // claim the next tile early
        __syncthreads();
        const u32 next_tile = PERSIST ? s_scan[35] : ntiles;
        const u32 tile_base = tile * (u32)TILE;
        const u32 remain = n - tile_base;
        const int valid = remain < (u32)TILE ? (int)remain : TILE;

        // ---- rank inside the warp (stable: item order = memory order); two 16-bit ranks per register ----
        u32 rank2[(ITEMS + 1) / 2];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 d = digit_of<PASS>(key[j]);
            const u32 peers = match_digit(d);
            const u32 below = __popc(peers & lt);
            u32 pre;
            if constexpr (RANK == 1) {
                pre = 0;
                if (below == 0) pre = atomicAdd(wh + d, (u32)__popc(peers));
                pre = __shfl_sync(0xffffffffu, pre, __ffs(peers) - 1);
            } else {
                pre = wh[d];
                __syncwarp();
                if (below == 0) wh[d] = pre + __popc(peers);
                __syncwarp();
            }
            const u32 r = pre + below;
            if (j & 1) rank2[j >> 1] |= r << 16; else rank2[j >> 1] = r;
        }
        __syncthreads();

        // ---- per digit: exclusive offsets across warps (GROUPS thread groups share the warps), tile total ----
        u32 part = 0;
        if (grp < GROUPS) {
#pragma unroll
            for (int w = 0; w < WPG; ++w) {
                u32* q = s_whist + (grp * WPG + w) * RADIX + dg;
                const u32 c = *q;
                *q = part;
                part += c;
            }
            if (GROUPS > 1) s_part[grp * RADIX + dg] = part;
        }
        u32 count = part, before = 0;
        if (GROUPS > 1) {
            __syncthreads();
            count = 0;
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const u32 v = s_part[g * RADIX + dg];
                if (g < grp) before += v;
                count += v;
            }
        }
        const u32 off = block_exclusive_scan<THREADS>(tid < RADIX ? count : 0u, nullptr, s_scan);
        u64 vcount = count;
        u64* lb = lookback + (u64)tile * RADIX + tid;
        if (tid < RADIX) {
            s_binoff[tid] = off;
            if (tid == RADIX - 1) vcount -= (u64)(TILE - valid);   // padding keys sit at the end of the last bin
            // publish this tile's count right away; the prefix is resolved after the shared-memory reorder
            st_volatile(lb, (tile == 0 ? LB_INCL : LB_AGG) | epoch | vcount);
        }
        __syncthreads();
        // fold the bin offset into the per-warp offsets: one shared-memory lookup per key in the reorder
        if (grp < GROUPS) {
            const u32 base = s_binoff[dg] + before;
#pragma unroll
            for (int w = 0; w < WPG; ++w) s_whist[(grp * WPG + w) * RADIX + dg] += base;
        }
        __syncthreads();

        // ---- reorder through shared memory (PRE: the first look-back round trip is in flight meanwhile) ----
        u64 pre[LB];
        if (PRE && tid < RADIX && tile != 0) lookback_fetch<LB>(lb - RADIX, tile, pre);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const u32 d = digit_of<PASS>(key[j]);
            const u32 r = (j & 1) ? (rank2[j >> 1] >> 16) : (rank2[j >> 1] & 0xffffu);
            s_keys[wh[d] + r] = key[j];
        }

        // ---- the key registers are free: start fetching the next tile ----
        if (PERSIST && next_tile < ntiles) load_tile(next_tile);

        // ---- decoupled look-back (one thread per digit) ----
        if (tid < RADIX) {
            u64 excl = 0;
            if (tile != 0) {
                excl = lookback_exclusive<LB, SLEEP, PRE>(lb, tile, epoch, pre);
                st_volatile(lb, LB_INCL | epoch | (excl + vcount));
            }
            s_goff[tid] = (u32)(gbase[tid] + excl) - s_binoff[tid];
        }
        __syncthreads();

        // ---- write out: consecutive threads -> consecutive addresses inside a bin ----
        // The last pass leaves the keys in their final order, and the keys of one digit sit in that order inside the
        // tile: a key whose top bits differ from its left neighbour's opens a bucket of the direct key index here
        // (KeyIndex, stages.cuh); the minimum over the tiles is the bucket's first position.  Saves the sweep over
        // the sorted keys that would otherwise build the index.
        auto index_key = [&](u32 i, u64 k, u32 g) {
            if constexpr (PASS == PASSES - 1) {
                if (kidx != nullptr) {
                    const u64 b = k >> kshift;
                    if (i == 0 || (s_keys[i - 1] >> kshift) != b) atomicMin(kidx + b, g);
                }
            }
        };
        if (valid == TILE) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const u32 i = tid + j * THREADS;
                const u64 k = s_keys[i];
                const u32 g = s_goff[digit_of<PASS>(k)] + i;
                out[g] = k;
                index_key(i, k, g);
            }
        } else {
            for (u32 i = tid; i < (u32)valid; i += THREADS) {
                const u64 k = s_keys[i];
                const u32 g = s_goff[digit_of<PASS>(k)] + i;
                out[g] = k;
                index_key(i, k, g);
            }
        }
        if (!PERSIST || next_tile >= ntiles) break;
        tile = next_tile;
        __syncthreads();                       // s_keys / s_goff / s_whist are reused by the next tile
    }
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, int PASS, bool PERSIST, int LB, int SLEEP, bool PRE, int RANK>
int launch_sweep_pass(const u64* in, u64* out, u64 n, const SortWorkspace& ws, cudaStream_t st) {
    using S = SweepSmem<THREADS, ITEMS>;
    auto kern = onesweep_kernel<THREADS, ITEMS, MIN_BLOCKS, PASS, PERSIST, LB, SLEEP, PRE, RANK>;
    static bool attr_done[64] = {};            // per device: function attributes belong to the device's context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (!attr_done[cur_dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
        attr_done[cur_dev & 63] = true;
    }
    const u64 ntiles = (n + S::TILE - 1) / S::TILE;
    u64 grid = ntiles;
    if (PERSIST) {
        static int resident = 0;
        if (!resident) {
            int dev = 0, sms = 148, per_sm = MIN_BLOCKS;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, S::bytes);
            resident = sms * (per_sm > 0 ? per_sm : 1);
        }
        if (grid > (u64)resident) grid = resident;
    }
    kern<<<(unsigned)grid, THREADS, S::bytes, st>>>(in, out, (u32)n, (u32)ntiles, ws.hist + PASS * RADIX, ws.lookback,
                                                      ws.tile_counter + PASS, (u64)(PASS + 1) << 56,
                                                      PASS == PASSES - 1 ? ws.key_index : nullptr, 64 - ws.key_index_bits);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


template <int THREADS, int ITEMS, int MIN_BLOCKS, bool PERSIST = false, int LB = 4, int SLEEP = 0, bool PRE = false, int RANK = 0>
int launch_sweep(const u64* in, u64* out, u64 n, int pass, const SortWorkspace& ws, cudaStream_t st) {
    switch (pass) {
        case 0: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 0, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 1: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 1, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 2: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 2, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 3: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 3, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 4: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 4, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 5: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 5, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        case 6: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 6, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
        default: return launch_sweep_pass<THREADS, ITEMS, MIN_BLOCKS, 7, PERSIST, LB, SLEEP, PRE, RANK>(in, out, n, ws, st);
    }
}

}  // namespace

int sort_config_tile(int cfg) {
    if (cfg >= TMA_CFG_BASE) return tma_config_tile(cfg);
    switch (cfg) {

        case 1: return 512 * 16;
        case 2: return 256 * 24;
        case 3: return 384 * 16;
        case 4: return 512 * 12;
        case 5: return 1024 * 8;
        case 6: return 512 * 16;
        case 8: return 384 * 16;
        case 9: return 512 * 20;
        case 10: case 11: return 320 * 16;
        case 12: return 384 * 14;
        case 13: return 448 * 12;
        case 14: return 288 * 16;
        case 15: return 384 * 18;
        case 16: return 384 * 16;
        case 17: return 512 * 16;
        case 18: return 256 * 16;
        case 19: return 384 * 18;
        default: return 256 * 16;
    }
}

size_t sort_workspace_bytes(u64 n, int cfg) {
    const u64 tile = (u64)sort_config_tile(cfg);
    const u64 ntiles = (n + tile - 1) / tile + 1;
    return PASSES * RADIX * 8 + 64 + ntiles * RADIX * 8 + ntiles * 64;
}

int sort_workspace_bind(SortWorkspace& ws, void* mem, u64 n, int cfg) {
    ws.cfg = cfg;
    ws.hist = reinterpret_cast<u64*>(mem);
    ws.tile_counter = reinterpret_cast<u32*>(ws.hist + PASSES * RADIX);
    ws.skip = ws.tile_counter + 8;
    ws.lookback = ws.hist + PASSES * RADIX + 8;
    const u64 tile = (u64)sort_config_tile(cfg);
    ws.ntiles = (n + tile - 1) / tile;
    ws.lookback_par = ws.lookback + (ws.ntiles + 1) * RADIX;
    return 0;
}

// Sorts `n` keys (n < 2^32 per call: one device's share).  `a` holds the input; `b` is scratch of the
// same size.  Returns in *result which of the two buffers holds the sorted keys (passes whose digit
// is constant are skipped).
int radix_sort_clear(const SortWorkspace& ws, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(ws.hist, 0, PASSES * RADIX * 8 + 64 + (ws.ntiles + 1) * RADIX * 8 + (ws.ntiles + 1) * 64, st));
    return 0;
}

int radix_sort_u64(u64* a, u64* b, u64 n, const SortWorkspace& ws, cudaStream_t st, u64** result, bool hist_ready) {
    *result = a;
    if (n <= 1) return 0;
    if (n >= (1ull << 32) - (1u << 16)) {
        set_error("radix_sort_u64: at most 2^32 - 65536 keys per device call");
        return -1;
    }
    if (!hist_ready) {
        if (radix_sort_clear(ws, st)) return -1;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        u64 want = (n / 2 + 511) / 512;
        unsigned hgrid = (unsigned)(want < (u64)sms * 8 ? (want ? want : 1) : (u64)sms * 8);
        radix_hist_kernel<512><<<hgrid, 512, 0, st>>>(a, n, ws.hist);
        CUDA_TRY(cudaGetLastError());
        DEBWT_COUNT(1);
    }
    radix_scan_kernel<<<1, RADIX, 0, st>>>(ws.hist, n, ws.skip);
    DEBWT_COUNT(1);
    CUDA_TRY(cudaGetLastError());
    u32 skip[PASSES];
    CUDA_TRY(cudaMemcpyAsync(skip, ws.skip, sizeof skip, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    u64 *src = a, *dst = b;
    int sweeps = 0;
    if (ws.ev_sweep_begin) CUDA_TRY(cudaEventRecord(ws.ev_sweep_begin, st));
    for (int p = ws.first_pass; p < PASSES; ++p) {
        if (skip[p]) continue;
        int rc;
        if (ws.cfg >= TMA_CFG_BASE) rc = launch_tma_sweep(ws.cfg, src, dst, n, p, ws, st);
        else switch (ws.cfg) {
            case 1: rc = launch_sweep<512, 16, 2>(src, dst, n, p, ws, st); break;
            case 2: rc = launch_sweep<256, 24, 2>(src, dst, n, p, ws, st); break;
            case 3: rc = launch_sweep<384, 16, 2>(src, dst, n, p, ws, st); break;
            case 4: rc = launch_sweep<512, 12, 2>(src, dst, n, p, ws, st); break;
            case 5: rc = launch_sweep<1024, 8, 1>(src, dst, n, p, ws, st); break;
            case 6: rc = launch_sweep<512, 16, 2, true>(src, dst, n, p, ws, st); break;
            case 7: rc = launch_sweep<256, 16, 4>(src, dst, n, p, ws, st); break;
            case 8: rc = launch_sweep<384, 16, 3>(src, dst, n, p, ws, st); break;
            case 9: rc = launch_sweep<512, 20, 2>(src, dst, n, p, ws, st); break;
            case 10: rc = launch_sweep<320, 16, 3>(src, dst, n, p, ws, st); break;
            case 11: rc = launch_sweep<320, 16, 4>(src, dst, n, p, ws, st); break;
            case 12: rc = launch_sweep<384, 14, 3>(src, dst, n, p, ws, st); break;
            case 13: rc = launch_sweep<448, 12, 3>(src, dst, n, p, ws, st); break;
            case 14: rc = launch_sweep<288, 16, 4>(src, dst, n, p, ws, st); break;
            case 15: rc = launch_sweep<384, 18, 3>(src, dst, n, p, ws, st); break;
            case 16: rc = launch_sweep<384, 16, 3, false, 4, 0, false, 1>(src, dst, n, p, ws, st); break;
            case 17: rc = launch_sweep<512, 16, 2, false, 4, 0, false, 1>(src, dst, n, p, ws, st); break;
            case 18: rc = launch_sweep<256, 16, 4, false, 4, 0, false, 1>(src, dst, n, p, ws, st); break;
            case 19: rc = launch_sweep<384, 18, 3, false, 4, 0, false, 1>(src, dst, n, p, ws, st); break;
            default: rc = launch_sweep<256, 16, 3>(src, dst, n, p, ws, st); break;
        }
        if (rc) return rc;
        DEBWT_COUNT(1);
        ++sweeps;
        u64* t = src; src = dst; dst = t;
    }
    if (ws.ev_sweep_end) CUDA_TRY(cudaEventRecord(ws.ev_sweep_end, st));
    if (ws.sweeps_out) *ws.sweeps_out = sweeps;
    if (ws.key_index_done) *ws.key_index_done = ws.key_index != nullptr && !skip[PASSES - 1];
    *result = src;
    return 0;
}

}  // namespace debwt
