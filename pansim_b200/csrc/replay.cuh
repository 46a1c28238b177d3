// replay.cuh -- K6: apply an explicit, ordered list of core writes with the
// reference's store semantics (`pop[[row, site]] = value`, later entries win;
// population.rs:537 and :745) -- bit-exact and deterministic on a parallel
// machine.
//
// Two passes over the event list with an open-addressing hash table keyed by
// cell: pass 1 records, per cell, the highest event index that targets it
// (atomicMax); pass 2 lets exactly that event store its value. Distinct cells
// that share a 32-bit word are combined with atomicAnd/atomicOr on disjoint bit
// pairs, which commute.
#pragma once
#include "common.cuh"

namespace pansim {

struct CoreWriteList {
    const uint32_t *row, *site;   // site = global core coordinate
    const uint8_t *value;         // one-hot
    size_t n;
    uint32_t index_base;          // order offset (mutations first, then HR)
};

__device__ __forceinline__ uint64_t hash64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__global__ void replay_insert_kernel(CoreWriteList ev, uint64_t site_begin, uint64_t site_end,
                                     uint64_t local_sites, unsigned long long *keys, uint32_t *vals,
                                     uint64_t mask)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ev.n) return;
    const uint64_t site = ev.site[k];
    if (site < site_begin || site >= site_end) return;
    const unsigned long long cell = (unsigned long long)ev.row[k] * local_sites + (site - site_begin) + 1ull;
    uint64_t slot = hash64(cell) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[slot], 0ull, cell);
        if (prev == 0ull || prev == cell) {
            atomicMax(&vals[slot], ev.index_base + (uint32_t)k + 1u);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void replay_apply_kernel(CoreWriteList ev, uint64_t site_begin, uint64_t site_end,
                                    uint64_t local_sites, const unsigned long long *keys,
                                    const uint32_t *vals, uint64_t mask, uint8_t *state, uint64_t row_stride,
                                    int *bad)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ev.n) return;
    const uint64_t site = ev.site[k];
    if (site < site_begin || site >= site_end) return;
    const uint64_t ls = site - site_begin;
    const unsigned long long cell = (unsigned long long)ev.row[k] * local_sites + ls + 1ull;
    uint64_t slot = hash64(cell) & mask;
    while (keys[slot] != cell) slot = (slot + 1) & mask;
    if (vals[slot] != ev.index_base + (uint32_t)k + 1u) return;      // a later event owns this cell
    const uint32_t b = ev.value[k];
    if (!(b == 1 || b == 2 || b == 4 || b == 8)) { *bad = 1; return; }
    const uint32_t code = (b >> 1) - (b >> 3);
    uint32_t *word = reinterpret_cast<uint32_t *>(state + (uint64_t)ev.row[k] * row_stride) + (ls >> 4);
    const uint32_t sh = (uint32_t)(ls & 15u) * 2u;
    atomicAnd(word, ~(3u << sh));
    atomicOr(word, code << sh);
}

}  // namespace pansim
