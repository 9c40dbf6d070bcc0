// topk_list.cuh -- sorted (distance asc, index asc) top-k lists kept in shared memory, maintained
// by one warp.  The order is the total order of src_legacy/tests/test_helpers.rs:116-120.
#pragma once
#include <stdint.h>

__device__ __forceinline__ bool topk_key_less(double da, uint32_t ia, double db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}

// All 32 lanes call with identical (d, j).  ld/li: list of current length c, capacity k <= 128.
// Inserts at its sorted position, dropping the last entry when the list is full.
__device__ __forceinline__ void warp_list_insert(double* ld, uint32_t* li, uint32_t& c, uint32_t k, double d, uint32_t j,
                                                 int lane) {
    uint32_t p = 0;
    for (uint32_t base = 0; base < c; base += 32) {
        uint32_t t = base + lane;
        bool less = t < c && topk_key_less(ld[t], li[t], d, j);
        p += __popc(__ballot_sync(0xffffffffu, less));
    }
    if (p >= k) return;
    const uint32_t last = c < k ? c : k - 1;  // entries [p, last) move up by one
    double rd[4];
    uint32_t ri[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        uint32_t t = p + lane + 32 * u;
        if (t < last) { rd[u] = ld[t]; ri[u] = li[t]; }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        uint32_t t = p + lane + 32 * u;
        if (t < last) { ld[t + 1] = rd[u]; li[t + 1] = ri[u]; }
    }
    if (lane == 0) { ld[p] = d; li[p] = j; }
    __syncwarp();
    c = last + 1;
}
