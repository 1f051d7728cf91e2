// FM-index tables over the finished BWT and the at-scale verifier built on them.
//
// The reference builds these in its unreachable "developer mode": occ checkpoints every 32 rows
// (reference src/insertCase3.c:139-208: occ[(N >> 5) + 1][4], rows holding '#'/'$' are stored as T but not counted),
// the C-array of cumulative base counts (src/collect#$.c:92-100), and an LF walk that inverts the BWT
// (src/LFsearch.c:49-166, findSeg :167-235: '#' rows map to the last R rows in rank order, :131).  Here:
//   * occ + C are produced on the device in the reference's layout (one sweep over the packed BWT);
//   * LF(r) for every row in one kernel;
//   * the reference walks the LF permutation one row at a time (N dependent steps); on the GPU the walk is list ranking by
//     pointer jumping (Wyllie): ceil(log2 N) rounds of one random 8-byte gather per row, after which every row knows
//     its distance to the end of the walk, i.e. its text position.  The BWT is the BWT of T iff the permutation is ONE
//     cycle and the symbol of row r equals T[position(r) - 1] for every r -- checked against the caller's ASCII text.
#include <algorithm>
#include <string>
#include <vector>

#include "ctx.cuh"

namespace debwt {

namespace {

constexpr int TPB = 256;
constexpr u32 NIL = 0xFFFFFFFFu;
inline unsigned grid_for(u64 work, u64 per_block = TPB) { return (unsigned)((work + per_block - 1) / per_block); }

// bit j (0..31) of the result set <=> symbol j of the packed word (bits 2*(31-j)) equals `code`
__device__ __forceinline__ u32 symbol_mask(u64 w, u32 code) {
    const u64 x = w ^ (code & 2u ? 0ull : 0xAAAAAAAAAAAAAAAAull) ^ (code & 1u ? 0ull : 0x5555555555555555ull);
    u64 m = x & (x >> 1) & 0x5555555555555555ull;          // bit 2*(31-j) set where both bits match
    // compress the even bits: symbol j sits at bit 2*(31-j) -> bit (31-j); then reverse so that symbol j is bit j
    m = (m | (m >> 1)) & 0x3333333333333333ull;
    m = (m | (m >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    m = (m | (m >> 4)) & 0x00FF00FF00FF00FFull;
    m = (m | (m >> 8)) & 0x0000FFFF0000FFFFull;
    m = (m | (m >> 16)) & 0x00000000FFFFFFFFull;
    return __brev((u32)m);
}
__device__ __forceinline__ u32 valid_rows(u64 n, u64 w) {                 // rows of word w that exist
    const u64 left = n - w * 32;
    return left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
}

__global__ void __launch_bounds__(TPB) spec_bits_kernel(u32* __restrict__ bits, const u64* __restrict__ rows, u64 m) {
    const u64 t = (u64)blockIdx.x * TPB + threadIdx.x;
    if (t < m) atomicOr(bits + (rows[t] >> 5), 1u << (rows[t] & 31));
}

// per block of TPB words: totals of A, C, G, T (specials excluded)
__global__ void __launch_bounds__(TPB) occ_count_kernel(const u64* __restrict__ bwt, const u32* __restrict__ spec, u64 n, u64 nwords,
                                                       u32* __restrict__ block_tot) {
    __shared__ u32 sm[40];
    const u64 w = (u64)blockIdx.x * TPB + threadIdx.x;
    u32 ac = 0, gt = 0;
    if (w < nwords) {
        const u64 x = bwt[w];
        const u32 ok = valid_rows(n, w) & ~spec[w];
        ac = __popc(symbol_mask(x, 0) & ok) | (__popc(symbol_mask(x, 1) & ok) << 16);
        gt = __popc(symbol_mask(x, 2) & ok) | (__popc(symbol_mask(x, 3) & ok) << 16);
    }
    u32 t_ac, t_gt;
    block_exclusive_scan<TPB>(ac, &t_ac, sm);
    block_exclusive_scan<TPB>(gt, &t_gt, sm);
    if (threadIdx.x == 0) {
        u32* o = block_tot + (u64)blockIdx.x * 4;
        o[0] = t_ac & 0xffffu; o[1] = t_ac >> 16; o[2] = t_gt & 0xffffu; o[3] = t_gt >> 16;
    }
}

// exclusive scan over the block totals (one block, sequential over chunks) -> 64-bit bases; totals[4] at the end
__global__ void __launch_bounds__(1024) occ_scan_kernel(const u32* __restrict__ block_tot, u64 nb, u64* __restrict__ block_base,
                                                       u64* __restrict__ totals) {
    __shared__ u32 sm[40];
    __shared__ u64 carry[4];
    if (threadIdx.x < 4) carry[threadIdx.x] = 0;
    __syncthreads();
    for (u64 base = 0; base < nb; base += 1024) {
        const u64 i = base + threadIdx.x;
        for (int c = 0; c < 4; ++c) {
            const u32 v = i < nb ? block_tot[i * 4 + c] : 0;
            u32 total;
            const u32 ex = block_exclusive_scan<1024>(v, &total, sm);
            if (i < nb) block_base[i * 4 + c] = carry[c] + ex;
            __syncthreads();
            if (threadIdx.x == 0) carry[c] += total;
            __syncthreads();
        }
    }
    if (threadIdx.x < 4) totals[threadIdx.x] = carry[threadIdx.x];
}

// occ[w][c] = number of symbol c in rows [0, 32 w)   (reference layout, src/insertCase3.c:141-142)
__global__ void __launch_bounds__(TPB) occ_write_kernel(const u64* __restrict__ bwt, const u32* __restrict__ spec, u64 n, u64 nwords,
                                                       const u64* __restrict__ block_base, u64* __restrict__ occ, u64 occ_len) {
    __shared__ u32 sm[40];
    const u64 w = (u64)blockIdx.x * TPB + threadIdx.x;
    u32 ac = 0, gt = 0;
    if (w < nwords) {
        const u64 x = bwt[w];
        const u32 ok = valid_rows(n, w) & ~spec[w];
        ac = __popc(symbol_mask(x, 0) & ok) | (__popc(symbol_mask(x, 1) & ok) << 16);
        gt = __popc(symbol_mask(x, 2) & ok) | (__popc(symbol_mask(x, 3) & ok) << 16);
    }
    const u32 e_ac = block_exclusive_scan<TPB>(ac, nullptr, sm);
    const u32 e_gt = block_exclusive_scan<TPB>(gt, nullptr, sm);
    if (w < occ_len) {
        const u64* b = block_base + (u64)blockIdx.x * 4;
        ulonglong2* o = reinterpret_cast<ulonglong2*>(occ + w * 4);
        o[0] = make_ulonglong2(b[0] + (e_ac & 0xffffu), b[1] + (e_ac >> 16));
        o[1] = make_ulonglong2(b[2] + (e_gt & 0xffffu), b[3] + (e_gt >> 16));
    }
    // the checkpoint one past the last word (when N is a multiple of 32 times ... the reference allocates (N >> 5) + 1)
    if (w + 1 == nwords && w + 1 < occ_len) {
        const u64* b = block_base + (u64)blockIdx.x * 4;
        ulonglong2* o = reinterpret_cast<ulonglong2*>(occ + (w + 1) * 4);
        o[0] = make_ulonglong2(b[0] + (e_ac & 0xffffu) + (ac & 0xffffu), b[1] + (e_ac >> 16) + (ac >> 16));
        o[1] = make_ulonglong2(b[2] + (e_gt & 0xffffu) + (gt & 0xffffu), b[3] + (e_gt >> 16) + (gt >> 16));
    }
}

struct FmView {
    const u64* bwt;
    const u32* spec;
    const u64* occ;
    const u64* sharp;      // sorted rows holding '#'
    u64 n_sharp;
    u64 dollar_row;
    u64 n;
    u64 C[6];
};

// number of rows < r holding base c (0..3)   (findSeg, src/LFsearch.c:167-235, without the -1 / +ACGT[type])
__device__ __forceinline__ u64 occ_before(const FmView& f, u32 c, u64 r) {
    const u64 w = r >> 5;
    const u32 below = (1u << (r & 31)) - 1u;
    u64 v = f.occ[w * 4 + c];
    if (below) v += __popc(symbol_mask(f.bwt[w], c) & ~f.spec[w] & below);
    return v;
}

// symbol of row r: 0..3, 4 = '#', 5 = '$'
__device__ __forceinline__ u32 row_symbol(const FmView& f, u64 r) {
    if ((f.spec[r >> 5] >> (r & 31)) & 1u) return r == f.dollar_row ? 5u : 4u;
    return (u32)(f.bwt[r >> 5] >> (2 * (31 - (r & 31)))) & 3u;
}

__device__ __forceinline__ u64 lf_of(const FmView& f, u64 r) {
    const u32 c = row_symbol(f, r);
    if (c == 5u) return f.n - 1;                                                  // '$' is the largest suffix
    if (c == 4u) return f.C[4] + lower_bound_u64(f.sharp, 0, f.n_sharp, r);       // src/LFsearch.c:131
    return f.C[c] + occ_before(f, c, r);
}

// list node of row r: successor (LF) in the high half, distance to it in the low half; the row whose LF is the start
// row (the row of suffix 0, which holds '$') ends the list
__global__ void __launch_bounds__(TPB) lf_init_kernel(FmView f, u64* __restrict__ node) {
    const u64 r = (u64)blockIdx.x * TPB + threadIdx.x;
    if (r >= f.n) return;
    const u64 s = lf_of(f, r);
    node[r] = s == f.dollar_row ? ((u64)NIL << 32) : ((s << 32) | 1ull);
}

__global__ void __launch_bounds__(TPB) jump_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 n, u32* __restrict__ live) {
    const u64 r = (u64)blockIdx.x * TPB + threadIdx.x;
    if (r >= n) return;
    u64 a = in[r];
    const u32 s = (u32)(a >> 32);
    bool more = false;
    if (s != NIL) {
        const u64 b = in[s];
        a = (b & 0xFFFFFFFF00000000ull) | (u32)((u32)a + (u32)b);
        more = (u32)(b >> 32) != NIL;
    }
    out[r] = a;
    if (__any_sync(0xffffffffu, more) && (threadIdx.x & 31) == 0) *live = 1u;
}

__device__ __forceinline__ u32 ascii_symbol(u8 c) {
    if (c == '#') return 4u;
    if (c == '$') return 5u;
    const u32 u = c & 0xDFu;
    const u32 x = (u >> 1) & 3u;
    return x ^ (x >> 1);
}

// after the jumps: the low half of node[r] is the number of LF steps from r to the end of the walk = position(r) - 1
__global__ void __launch_bounds__(TPB) lf_check_kernel(FmView f, const u64* __restrict__ node, const u8* __restrict__ text,
                                                      unsigned long long* __restrict__ bad) {
    const u64 r = (u64)blockIdx.x * TPB + threadIdx.x;
    if (r >= f.n) return;
    const u64 a = node[r];
    bool ok = (u32)(a >> 32) == NIL;                                              // reached the end: r is on the one cycle
    if (ok) ok = row_symbol(f, r) == ascii_symbol(text[(u32)a]);
    if (!ok) atomicAdd(bad, 1ull);
}

// sequential LF walk from the row of suffix 0 (it holds '$'): step i must show T[N-1-i]; one thread, `steps` dependent
// steps -- the reference's own developer check (src/LFsearch.c:49-166), for texts beyond the list-ranking verifier's 2^32
__global__ void lf_walk_kernel(FmView f, const u8* __restrict__ tail, u64 steps, unsigned long long* __restrict__ bad) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u64 r = f.dollar_row;
    unsigned long long nb = 0;
    for (u64 i = 0; i < steps; ++i) {
        if (row_symbol(f, r) != ascii_symbol(tail[steps - 1 - i])) ++nb;
        r = lf_of(f, r);
    }
    *bad = nb;
}

// backward search of one pattern per thread (count only)
__global__ void __launch_bounds__(TPB) fm_count_kernel(FmView f, const u8* __restrict__ pats, const u64* __restrict__ offs, u64 m,
                                                      u64* __restrict__ counts) {
    const u64 t = (u64)blockIdx.x * TPB + threadIdx.x;
    if (t >= m) return;
    u64 sp = 0, ep = f.n;                                                         // rows [sp, ep)
    for (u64 i = offs[t + 1]; i > offs[t] && sp < ep; --i) {
        const u32 c = ascii_symbol(pats[i - 1]);
        if (c > 3u) { sp = ep = 0; break; }
        sp = f.C[c] + occ_before(f, c, sp);
        ep = f.C[c] + (ep == f.n ? f.C[c + 1] - f.C[c] : occ_before(f, c, ep));
    }
    counts[t] = ep > sp ? ep - sp : 0;
}

FmView view_of(const debwt_ctx* c) {
    FmView f;
    f.bwt = c->d_bwt; f.spec = c->d_spec_bits; f.occ = c->d_occ; f.sharp = c->d_sharp_sorted;
    f.n_sharp = c->n_rec - 1; f.dollar_row = c->dollar_row; f.n = c->n;
    for (int i = 0; i < 6; ++i) f.C[i] = c->c_array[i];
    return f;
}

}  // namespace

void drop_index(debwt_ctx* c) {
    if (c->d_occ) cudaFree(c->d_occ);
    if (c->d_spec_bits) cudaFree(c->d_spec_bits);
    if (c->d_sharp_sorted) cudaFree(c->d_sharp_sorted);
    c->d_occ = nullptr; c->d_spec_bits = nullptr; c->d_sharp_sorted = nullptr;
    c->indexed = false;
}

}  // namespace debwt

using namespace debwt;

#define FAIL(msg)              \
    do {                       \
        debwt::set_error(msg); \
        return -1;             \
    } while (0)

extern "C" {

// sharp: the rows holding '#', ascending
static int index_build_with(debwt_ctx* c, std::vector<u64> sharp, u64 dollar_row) {
    cudaStream_t st = c->st;
    const u64 n = c->n, nwords = c->n_words, occ_len = (n >> 5) + 1;
    const u64 cnt = sharp.size();
    c->dollar_row = dollar_row;
    sharp.push_back(dollar_row);                                                  // all special rows: cnt + 1 entries
    drop_index(c);
    const u64 nb = (nwords + TPB - 1) / TPB;
    u32* d_tot = nullptr;
    u64 *d_base = nullptr, *d_totals = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_occ), (occ_len + 1) * 32));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_spec_bits), (nwords + 1) * 4));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_sharp_sorted), (cnt + 1) * 8));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_tot), nb * 16 + 16));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_base), nb * 32 + 64));
    d_totals = d_base + nb * 4;
    CUDA_TRY(cudaMemcpyAsync(c->d_sharp_sorted, sharp.data(), (cnt + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(c->d_spec_bits, 0, (nwords + 1) * 4, st));
    spec_bits_kernel<<<grid_for(cnt + 1), TPB, 0, st>>>(c->d_spec_bits, c->d_sharp_sorted, cnt + 1);
    occ_count_kernel<<<(unsigned)nb, TPB, 0, st>>>(c->d_bwt, c->d_spec_bits, n, nwords, d_tot);
    occ_scan_kernel<<<1, 1024, 0, st>>>(d_tot, nb, d_base, d_totals);
    occ_write_kernel<<<(unsigned)nb, TPB, 0, st>>>(c->d_bwt, c->d_spec_bits, n, nwords, d_base, c->d_occ, occ_len);
    DEBWT_COUNT(4);
    u64 tot[4];
    CUDA_TRY(cudaMemcpyAsync(tot, d_totals, 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    cudaFree(d_tot);
    cudaFree(d_base);
    if (tot[0] + tot[1] + tot[2] + tot[3] + c->n_rec != n) FAIL("internal: base counts of the BWT do not add up to N");
    c->c_array[0] = 0;                                                            // src/collect#$.c:92-100
    c->c_array[1] = tot[0];
    c->c_array[2] = tot[0] + tot[1];
    c->c_array[3] = tot[0] + tot[1] + tot[2];
    c->c_array[4] = tot[0] + tot[1] + tot[2] + tot[3];                            // '#' rows
    c->c_array[5] = n - 1;                                                        // '$'
    c->indexed = true;
    return 0;
}

int debwt_index_build(debwt_ctx* c) {
    if (!c || !c->built) FAIL("no result: call debwt_build first");
    if (c->indexed) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->st;
    // separator rows: '#' rows sorted (src/insertCase3.c:84-95 writes them ascending), '$' row
    std::vector<u64> sharp(c->n_rec + 1);
    u32 cnt = 0;
    u64 dollar = 0;
    CUDA_TRY(cudaMemcpyAsync(&cnt, c->d_sharp_count, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&dollar, c->d_dollar, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(sharp.data(), c->d_sharp, (c->n_rec + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (cnt != c->n_rec - 1) FAIL("internal: wrong number of '#' rows");
    sharp.resize(cnt);
    std::sort(sharp.begin(), sharp.end());
    return index_build_with(c, std::move(sharp), dollar);
}

int debwt_index_sizes(const debwt_ctx* c, uint64_t* occ_rows) {
    if (!c || !c->indexed) FAIL("no index: call debwt_index_build first");
    if (occ_rows) *occ_rows = (c->n >> 5) + 1;
    return 0;
}

int debwt_index_copy(debwt_ctx* c, uint64_t* occ, uint64_t* c_array) {
    if (!c || !c->indexed) FAIL("no index: call debwt_index_build first");
    CUDA_TRY(cudaSetDevice(c->device));
    if (occ) CUDA_TRY(cudaMemcpyAsync(occ, c->d_occ, ((c->n >> 5) + 1) * 32, cudaMemcpyDeviceToHost, c->st));
    CUDA_TRY(cudaStreamSynchronize(c->st));
    if (c_array) for (int i = 0; i < 6; ++i) c_array[i] = c->c_array[i];
    return 0;
}

int debwt_index_count(debwt_ctx* c, const char* patterns, const uint64_t* offsets, uint64_t n_patterns, uint64_t* counts) {
    if (!c || !c->indexed) FAIL("no index: call debwt_index_build first");
    if (n_patterns == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    u8* d_p = nullptr; u64 *d_o = nullptr, *d_c = nullptr;
    const u64 bytes = offsets[n_patterns];
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_p), bytes + 16));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_o), (n_patterns + 1) * 8));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_c), n_patterns * 8));
    CUDA_TRY(cudaMemcpyAsync(d_p, patterns, bytes, cudaMemcpyHostToDevice, c->st));
    CUDA_TRY(cudaMemcpyAsync(d_o, offsets, (n_patterns + 1) * 8, cudaMemcpyHostToDevice, c->st));
    fm_count_kernel<<<grid_for(n_patterns), TPB, 0, c->st>>>(view_of(c), d_p, d_o, n_patterns, d_c);
    DEBWT_COUNT(1);
    CUDA_TRY(cudaMemcpyAsync(counts, d_c, n_patterns * 8, cudaMemcpyDeviceToHost, c->st));
    CUDA_TRY(cudaStreamSynchronize(c->st));
    CUDA_TRY(cudaGetLastError());
    cudaFree(d_p); cudaFree(d_o); cudaFree(d_c);
    return 0;
}

static int verify_indexed(debwt_ctx* c, const void* d_text, uint64_t* n_bad_out, float* ms_out) {
    if (c->n >= 0xFFFFFFFFull) FAIL("verify: N must be below 2^32 - 1");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->st;
    const u64 n = c->n;
    u64 *d_a = nullptr, *d_b = nullptr;
    u32* d_live = nullptr;
    unsigned long long* d_bad = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d_a), n * 8) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&d_b), n * 8) != cudaSuccess) {
        cudaGetLastError();
        if (d_a) cudaFree(d_a);
        FAIL("verify: not enough device memory for the list-ranking buffers (16 bytes per symbol)");
    }
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_live), 16));
    d_bad = reinterpret_cast<unsigned long long*>(d_live + 2);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, st));
    const FmView f = view_of(c);
    lf_init_kernel<<<grid_for(n), TPB, 0, st>>>(f, d_a);
    DEBWT_COUNT(1);
    int rounds = 0;
    for (; rounds < 34; ++rounds) {
        u32 live = 0;
        CUDA_TRY(cudaMemsetAsync(d_live, 0, 4, st));
        jump_kernel<<<grid_for(n), TPB, 0, st>>>(d_a, d_b, n, d_live);
        DEBWT_COUNT(1);
        CUDA_TRY(cudaMemcpyAsync(&live, d_live, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        std::swap(d_a, d_b);
        if (!live) break;
    }
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, 8, st));
    lf_check_kernel<<<grid_for(n), TPB, 0, st>>>(f, d_a, static_cast<const u8*>(d_text), d_bad);
    DEBWT_COUNT(1);
    unsigned long long bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_live);
    if (n_bad_out) *n_bad_out = bad;
    if (ms_out) *ms_out = ms;
    return 0;
}

int debwt_verify_text_device(debwt_ctx* c, const void* d_text, uint64_t n_symbols, uint64_t* n_bad_out, float* ms_out) {
    if (!c || !c->built) FAIL("no result: call debwt_build first");
    if (n_symbols != c->n) FAIL("verify: text length differs from the build's");
    if (debwt_index_build(c)) return -1;
    return verify_indexed(c, d_text, n_bad_out, ms_out);
}

int debwt_verify_bwt_device(int device, const uint64_t* d_bwt_words, uint64_t n_symbols, const uint64_t* sharp_rows, uint64_t n_sharp,
                            uint64_t dollar_row, const void* d_text, uint64_t* n_bad_out, float* ms_out) {
    if (!d_bwt_words || !d_text || (n_sharp && !sharp_rows)) FAIL("null argument");
    if (debwt_device_count() <= device || device < 0) FAIL("no such CUDA device (this library has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));
    debwt_ctx tmp;                                        // a view of the caller's buffers, nothing owned but the index
    tmp.device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&tmp.st, cudaStreamNonBlocking));
    tmp.n = n_symbols; tmp.n_rec = n_sharp + 1; tmp.n_words = (n_symbols + 31) / 32;
    tmp.d_bwt = reinterpret_cast<u64*>(const_cast<uint64_t*>(d_bwt_words));
    tmp.built = true;
    std::vector<u64> sharp(sharp_rows, sharp_rows + n_sharp);
    std::sort(sharp.begin(), sharp.end());
    CUDA_TRY(cudaDeviceSynchronize());                    // the caller's buffers may have been produced on another stream
    int rc = index_build_with(&tmp, std::move(sharp), dollar_row);
    if (!rc) rc = verify_indexed(&tmp, d_text, n_bad_out, ms_out);
    drop_index(&tmp);
    cudaStreamDestroy(tmp.st);
    return rc;
}

int debwt_verify_walk_device(int device, const uint64_t* d_bwt_words, uint64_t n_symbols, const uint64_t* sharp_rows, uint64_t n_sharp,
                             uint64_t dollar_row, const void* d_text_tail, uint64_t steps, uint64_t* n_bad_out, uint64_t* c_array_out) {
    if (!d_bwt_words || !d_text_tail || (n_sharp && !sharp_rows)) FAIL("null argument");
    if (steps > n_symbols) FAIL("verify: more steps than symbols");
    if (debwt_device_count() <= device || device < 0) FAIL("no such CUDA device (this library has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(device));
    debwt_ctx tmp;
    tmp.device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&tmp.st, cudaStreamNonBlocking));
    tmp.n = n_symbols; tmp.n_rec = n_sharp + 1; tmp.n_words = (n_symbols + 31) / 32;
    tmp.d_bwt = reinterpret_cast<u64*>(const_cast<uint64_t*>(d_bwt_words));
    tmp.built = true;
    std::vector<u64> sharp(sharp_rows, sharp_rows + n_sharp);
    std::sort(sharp.begin(), sharp.end());
    CUDA_TRY(cudaDeviceSynchronize());
    int rc = index_build_with(&tmp, std::move(sharp), dollar_row);           // also checks that the base counts add up to N
    unsigned long long* d_bad = nullptr;
    if (!rc && cudaMalloc(reinterpret_cast<void**>(&d_bad), 8) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); rc = -1; }
    if (!rc) {
        lf_walk_kernel<<<1, 32, 0, tmp.st>>>(view_of(&tmp), static_cast<const u8*>(d_text_tail), steps, d_bad);
        DEBWT_COUNT(1);
        unsigned long long bad = 0;
        if (cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, tmp.st) != cudaSuccess || cudaStreamSynchronize(tmp.st) != cudaSuccess) {
            set_error(std::string("verify walk: ") + cudaGetErrorString(cudaGetLastError()));
            rc = -1;
        }
        if (n_bad_out) *n_bad_out = bad;
        if (c_array_out) for (int i = 0; i < 6; ++i) c_array_out[i] = tmp.c_array[i];
    }
    if (d_bad) cudaFree(d_bad);
    drop_index(&tmp);
    cudaStreamDestroy(tmp.st);
    return rc;
}

int debwt_verify_text(debwt_ctx* c, const char* text, uint64_t n_symbols, uint64_t* n_bad_out, float* ms_out) {
    if (!c || !text) FAIL("null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    u8* d_t = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_t), n_symbols + 16));
    CUDA_TRY(cudaMemcpy(d_t, text, n_symbols, cudaMemcpyHostToDevice));
    const int rc = debwt_verify_text_device(c, d_t, n_symbols, n_bad_out, ms_out);
    cudaFree(d_t);
    return rc;
}

}  // extern "C"
