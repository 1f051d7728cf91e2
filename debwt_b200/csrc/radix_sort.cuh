// K3 interface: LSD radix sort of 64-bit keys (see radix_sort.cu).
#pragma once
#include "common.cuh"

namespace debwt {

struct SortWorkspace {
    int cfg = 0;
    u64* hist = nullptr;          // [8][256] counts, then exclusive bases
    u32* tile_counter = nullptr;  // [8] dynamic tile ids, one per pass
    u32* skip = nullptr;          // [8] pass has a constant digit
    u64* lookback = nullptr;      // [ntiles][256] decoupled look-back words
    u64* lookback_par = nullptr;  // [ntiles][8] look-back words of the count PARITIES (bulk-store kernel, radix_sort_tma.cu)
    u64 ntiles = 0;
    cudaEvent_t ev_sweep_begin = nullptr, ev_sweep_end = nullptr;   // optional: bracket the scatter passes
    int* sweeps_out = nullptr;                                        // optional: number of passes launched
    // optional: the last pass also marks the direct key index (KeyIndex::idx, preset to 0xFFFFFFFF, `key_index_bits`
    // top bits); *key_index_done tells whether it ran (a constant top digit skips the pass)
    u32* key_index = nullptr;
    int key_index_bits = 0;
    bool* key_index_done = nullptr;
    // only the passes >= first_pass run (the caller knows the low 8 * first_pass bits need no ordering, e.g. K9's
    // (branch id, position) keys, which only have to be grouped by branch id)
    int first_pass = 0;
};

int sort_config_tile(int cfg);
size_t sort_workspace_bytes(u64 n, int cfg);
int sort_workspace_bind(SortWorkspace& ws, void* mem, u64 n, int cfg);
// hist_ready: the caller zeroed the workspace with radix_sort_clear() and filled ws.hist ([8][256] digit counts of
// exactly these n keys, pass p = bits 8p..8p+7) itself, e.g. from the text (k_text_digit_hist); the key sweep that
// would count them is skipped.
int radix_sort_clear(const SortWorkspace& ws, cudaStream_t st);
int radix_sort_u64(u64* a, u64* b, u64 n, const SortWorkspace& ws, cudaStream_t st, u64** result, bool hist_ready = false);

}  // namespace debwt
