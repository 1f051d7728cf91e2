// K3, bulk-store variant of the onesweep digit pass (reference src/mySort.c:98-176 is what the sort replaces).
//
// Same algorithm as onesweep_kernel in radix_sort.cu (warp-ballot ranking against per-warp shared-memory digit
// histograms, decoupled look-back across tiles).  What differs is the write-out -- the Blackwell way:
//   * every digit's run of the reordered tile leaves shared memory as ONE bulk async copy (cp.async.bulk
//     global <- shared::cta, SASS UBLKCP.G.S), issued by the thread that owns the digit: no per-key LDS / LDS / STG in
//     the write-out (5.5 of the ~24 LSU wavefronts per 32 keys that bound the register-staged kernel on B200), the
//     copy engine moves the bytes, and no barrier follows the look-back: a digit thread stores its run as soon as
//     its own prefix is known, the other warps are done after the reorder.
//   * Bulk copies need 16-byte aligned addresses on both sides and keys are 8 bytes, so a run must start in shared
//     memory on a slot of the same PARITY as its first global index -- which the count look-back only knows after the
//     reorder (it is deferred behind it on purpose: started earlier, the walks grow to the number of tiles in
//     flight, profiles/r02_sort_experiments.md).  The parities are therefore chained separately: per tile one 64-byte
//     row of eight words (32 digit-count parities each, same status|epoch|value format), one warp load fetches one
//     word of 32 predecessors, a XOR reduction gives the parity of every exclusive prefix in one or two round trips
//     BEFORE the reorder.  One slack slot per digit absorbs the shift; at most one head and one tail key per run
//     are stored with plain 8-byte stores.
// Algorithmic traffic per launch: 16 B/key (8 read + 8 written).
#include "radix_common.cuh"

namespace debwt {

namespace {

using namespace radix;

// ---- async proxy (bulk copy engine) helpers -------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int THREADS, int ITEMS>
struct BulkSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int SLOTS = TILE + RADIX;          // one slack slot per digit (parity of the run start)
    static_assert(SLOTS < 65536, "run start and count are packed into 16 bits each");
    static constexpr size_t off_hist = (size_t)SLOTS * 8;
    static constexpr size_t off_run = off_hist + (size_t)WARPS * RADIX * 4;    // [256] valid count << 16 | first slot
    static constexpr size_t off_g = off_run + RADIX * 4;                       // [256] global index of the run's first key
    static constexpr size_t off_scan = off_g + RADIX * 4;                      // 40 scan words + tile id
    static constexpr size_t bytes = off_scan + 48 * 4;
};

template <int THREADS, int ITEMS, int MIN_BLOCKS, int PASS, int LB>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_bulk_kernel(const u64* __restrict__ in, u64* __restrict__ out, u32 n, u32 ntiles, const u64* __restrict__ gbase,
                     u64* __restrict__ lookback, u64* __restrict__ lookback_par, u32* __restrict__ tile_counter, u64 epoch,
                     u32* __restrict__ kidx, int kshift) {
    using S = BulkSmem<THREADS, ITEMS>;
    constexpr int WARPS = S::WARPS, TILE = S::TILE, SLOTS = S::SLOTS;
    static_assert(THREADS >= RADIX, "one thread per digit is assumed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* s_keys = reinterpret_cast<u64*>(smem_raw);                           // [SLOTS]
    u32* s_whist = reinterpret_cast<u32*>(smem_raw + S::off_hist);            // [WARPS][RADIX]
    u32* s_run = reinterpret_cast<u32*>(smem_raw + S::off_run);
    u32* s_g = reinterpret_cast<u32*>(smem_raw + S::off_g);
    u32* s_scan = reinterpret_cast<u32*>(smem_raw + S::off_scan);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32* wh = s_whist + warp * RADIX;
    const u32 lt = lanemask_lt();
    constexpr bool MARK = PASS == PASSES - 1;           // the last pass may also mark the direct key index (radix_sort.cu)

    if (tid == 0) s_scan[40] = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) s_whist[i] = 0;
    __syncthreads();
    const u32 tile = s_scan[40];
    if (tile >= ntiles) return;
    const u32 tile_base = tile * (u32)TILE;
    const u32 remain = n - tile_base;
    const int valid = remain < (u32)TILE ? (int)remain : TILE;

    u64 key[ITEMS];
    {
        const u32 first = tile_base + warp * (ITEMS * 32) + lane;
        const u64* src = in + first;
        if (valid == TILE) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) key[j] = ld_stream(src + j * 32);
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) key[j] = (first + j * 32 < n) ? ld_stream(src + j * 32) : ~0ull;
        }
    }

    // ---- rank inside the warp (stable: item order = memory order); two 16-bit ranks per register ----
    u32 rank2[(ITEMS + 1) / 2];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 d = digit_of<PASS>(key[j]);
        const u32 peers = match_digit(d);
        const u32 pre = wh[d];
        __syncwarp();
        const u32 below = __popc(peers & lt);
        if (below == 0) wh[d] = pre + __popc(peers);
        __syncwarp();
        const u32 r = pre + below;
        if (j & 1) rank2[j >> 1] |= r << 16; else rank2[j >> 1] = r;
    }
    __syncthreads();

    // ---- per digit: tile total, publish count and count parity ----
    u32 count = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) count += s_whist[w * RADIX + tid];
    }
    u32 vcount = count;
    if (tid == RADIX - 1) vcount -= (u32)(TILE - valid);                        // padding keys sit at the end of the last bin
    u64* lb = lookback + (u64)tile * RADIX + tid;
    u64* par_row = lookback_par + (u64)tile * 8;
    u32 own_par = 0;
    if (tid < RADIX) {
        st_volatile(lb, (tile == 0 ? LB_INCL : LB_AGG) | epoch | (u64)vcount);
        own_par = __ballot_sync(0xffffffffu, vcount & 1u);                      // warps 0..7 are complete
        if (lane == 0) st_volatile(par_row + warp, (tile == 0 ? LB_INCL : LB_AGG) | epoch | (u64)own_par);
    }
    const u32 q = block_exclusive_scan<THREADS>(tid < RADIX ? count + 1u : 0u, nullptr, s_scan);

    // ---- parity of every exclusive prefix: warp w chains word w (digits 32w..32w+31) over the predecessors ----
    u32 spos = 0, gpar = 0;
    if (tid < RADIX) {
        u32 excl_par = 0;
        if (tile != 0) {
            u32 left = tile;
            const u64* p = lookback_par + (u64)(tile - 1) * 8 + warp;
            u32 spins = 0;
            for (;;) {
                const u64 x = ((u32)lane < left) ? ld_volatile(p - (size_t)lane * 8) : 0;
                const bool ready = (x & LB_EPOCH_MASK) == epoch && (x >> 62) != 0;
                const u32 m_ready = __ballot_sync(0xffffffffu, ready);
                const u32 m_incl = __ballot_sync(0xffffffffu, ready && (x >> 62) == 2);
                const u32 nready = (~m_ready) ? (u32)__ffs(~m_ready) - 1u : 32u;   // leading published predecessors
                const u32 first_incl = m_incl ? (u32)__ffs(m_incl) - 1u : 32u;
                const bool fin = first_incl < nready;
                const u32 take = fin ? first_incl + 1u : nready;
                excl_par ^= __reduce_xor_sync(0xffffffffu, (u32)lane < take ? (u32)x : 0u);
                if (fin) break;
                if (take == 0) {
                    __nanosleep(32);
                    if (++spins > (1u << 22)) __trap();                           // never hang the device on a bug
                }
                p -= (size_t)take * 8;
                left -= take;
            }
            if (lane == 0) st_volatile(par_row + warp, LB_INCL | epoch | (u64)(excl_par ^ own_par));
        }
        gpar = ((u32)gbase[tid] ^ (excl_par >> lane)) & 1u;                       // parity of the run's first global index
        spos = q + ((q ^ gpar) & 1u);
        u32 run = spos;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const u32 c = s_whist[w * RADIX + tid]; s_whist[w * RADIX + tid] = run; run += c; }
    }
    __syncthreads();

    // ---- reorder through shared memory ----
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const u32 d = digit_of<PASS>(key[j]);
        const u32 r = (j & 1) ? (rank2[j >> 1] >> 16) : (rank2[j >> 1] & 0xffffu);
        s_keys[wh[d] + r] = key[j];
    }
    fence_proxy_async();                                                       // generic-proxy writes -> visible to the bulk copies
    __syncthreads();
    if (tid >= RADIX && !(MARK && kidx != nullptr)) return;                    // only the digit threads are left with work

    // ---- decoupled look-back (count), then the digit's run leaves as one bulk copy ----
    if (tid < RADIX) {
        u64 excl = 0;
        if (tile != 0) {
            excl = lookback_exclusive<LB, 0, false>(lb, tile, epoch);
            st_volatile(lb, LB_INCL | epoch | (excl + vcount));
        }
        u32 g = (u32)(gbase[tid] + excl);
        if ((g & 1u) != gpar) __trap();                                        // the two chains disagree: never store misaligned
        if (MARK && kidx != nullptr) { s_run[tid] = (vcount << 16) | spos; s_g[tid] = g; }
        u32 c = vcount, s = spos;
        if (c) {
            if (g & 1u) { out[g] = s_keys[s]; ++s; ++g; --c; }
            const u32 body = c & ~1u;
            if (body) bulk_store(out + g, s_keys + s, body * 8u);
            if (c & 1u) out[g + body] = s_keys[s + body];
        }
        bulk_commit();
    }
    if (MARK && kidx != nullptr) {
        // The last pass leaves the keys in final order: a key whose top bits differ from its left neighbour's opens
        // a bucket of the direct key index (KeyIndex, stages.cuh); the minimum over the tiles is its first position.
        __syncthreads();
        for (u32 s = tid; s < (u32)SLOTS; s += THREADS) {
            const u64 k = s_keys[s];
            const u32 d = digit_of<PASS>(k);
            const u32 run = s_run[d];
            const u32 off = s - (run & 0xffffu);
            if (off < (run >> 16)) {                                           // unsigned: also rejects s < first slot
                const u64 bk = k >> kshift;
                if (off == 0 || (s_keys[s - 1] >> kshift) != bk) atomicMin(kidx + bk, s_g[d] + off);
            }
        }
    }
    if (tid < RADIX) bulk_wait_read();                                         // shared memory must outlive the copies' reads
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, int PASS, int LB>
int launch_bulk_pass(const u64* in, u64* out, u64 n, const SortWorkspace& ws, cudaStream_t st) {
    using S = BulkSmem<THREADS, ITEMS>;
    auto kern = onesweep_bulk_kernel<THREADS, ITEMS, MIN_BLOCKS, PASS, LB>;
    static bool attr_done[64] = {};             // per device: function attributes belong to the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
        attr_done[dev & 63] = true;
    }
    const u64 ntiles = (n + S::TILE - 1) / S::TILE;
    kern<<<(unsigned)ntiles, THREADS, S::bytes, st>>>(in, out, (u32)n, (u32)ntiles, ws.hist + PASS * RADIX, ws.lookback,
                                                        ws.lookback_par, ws.tile_counter + PASS, (u64)(PASS + 1) << 56,
                                                        PASS == PASSES - 1 ? ws.key_index : nullptr, 64 - ws.key_index_bits);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int THREADS, int ITEMS, int MIN_BLOCKS, int LB>
int launch_bulk(const u64* in, u64* out, u64 n, int pass, const SortWorkspace& ws, cudaStream_t st) {
    switch (pass) {
        case 0: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 0, LB>(in, out, n, ws, st);
        case 1: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 1, LB>(in, out, n, ws, st);
        case 2: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 2, LB>(in, out, n, ws, st);
        case 3: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 3, LB>(in, out, n, ws, st);
        case 4: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 4, LB>(in, out, n, ws, st);
        case 5: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 5, LB>(in, out, n, ws, st);
        case 6: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 6, LB>(in, out, n, ws, st);
        default: return launch_bulk_pass<THREADS, ITEMS, MIN_BLOCKS, 7, LB>(in, out, n, ws, st);
    }
}

}  // namespace

// cfg = TMA_CFG_BASE + shape: (threads, keys per thread, CTAs per SM, count look-back batch)
#define DEBWT_BULK_SHAPES(X) \
    X(0, 384, 16, 3, 4)     \
    X(1, 512, 16, 2, 4)

int tma_config_tile(int cfg) {
    switch (cfg - TMA_CFG_BASE) {
#define X(id, T, I, B, LB) case id: return T * I;
        DEBWT_BULK_SHAPES(X)
#undef X
        default: return 384 * 16;
    }
}

int launch_tma_sweep(int cfg, const u64* in, u64* out, u64 n, int pass, const SortWorkspace& ws, cudaStream_t st) {
    switch (cfg - TMA_CFG_BASE) {
#define X(id, T, I, B, LB) case id: return launch_bulk<T, I, B, LB>(in, out, n, pass, ws, st);
        DEBWT_BULK_SHAPES(X)
#undef X
        default: return launch_bulk<384, 16, 3, 4>(in, out, n, pass, ws, st);
    }
}

}  // namespace debwt
