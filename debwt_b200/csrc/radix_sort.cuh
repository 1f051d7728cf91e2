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
    u64 ntiles = 0;
    cudaEvent_t ev_sweep_begin = nullptr, ev_sweep_end = nullptr;   // optional: bracket the scatter passes
    int* sweeps_out = nullptr;                                        // optional: number of passes launched
};

int sort_config_tile(int cfg);
size_t sort_workspace_bytes(u64 n, int cfg);
int sort_workspace_bind(SortWorkspace& ws, void* mem, u64 n, int cfg);
int radix_sort_u64(u64* a, u64* b, u64 n, const SortWorkspace& ws, cudaStream_t st, u64** result);

}  // namespace debwt
