// core_hr.cuh -- K4b: homologous recombination of the core genome as a sparse
// pass over the gathered + mutated rows (population.rs:544-751, core branch).
//
// Reference process: every row acts as donor, emits n ~ Poisson(lambda_HR)
// events, each with a uniform recipient != donor and a uniform locus, carrying
// the donor's allele at that locus read BEFORE any event is applied (snapshot,
// population.rs:693-695); events are then applied in a shuffled donor order,
// later writes win (:745). Seen from the recipient (Poisson superposition /
// splitting): an 8192-site region of row r receives K ~ Poisson(8192 *
// lambda_HR / L) events, each at a uniform site of the region from a donor
// uniform on the other N-1 rows; the order among the events of one cell is the
// draw order.
//
// Two kernels, because every donor read must see the pre-recombination state:
//   hr_collect_kernel  one warp per (region, row), one event per lane. Reads
//                      the donor's and the recipient's cell from the state
//                      (read-only here) and appends (word, shift, allele XOR)
//                      for every cell whose LAST event changes it. Same-cell
//                      events are resolved here: __match_any_sync inside a
//                      window of 32 events, a per-warp 8192-bit claim map over
//                      windows visited last to first.
//   hr_apply_kernel    atomicXor of the collected deltas (cells are distinct,
//                      so the order of application is irrelevant).
// Items are ordered region-major so that concurrently running warps read donor
// cells of the same few column regions (L2 locality).
#pragma once
#include "common.cuh"

namespace pansim {

constexpr uint32_t STREAM_CORE_HR_COUNT = 7;          // counter word 3 = 7 << 16 (| 0x8000 | i for extra count draws)
constexpr uint32_t HR_EVENT_W0 = 0x01000000u;         // counter word 3 of event e = HR_EVENT_W0 + e
constexpr int HR_WARPS = 8;

struct HrArgs {
    uint32_t *state;          // packed core rows (gathered + mutated), words
    uint32_t n_rows, n_regions, region0;
    uint64_t row_stride_words;
    uint64_t site_limit;
    uint2 key;
    uint32_t gen;
    const uint32_t *tab;      // device image [256 guide][size thresholds] of Poisson(region mean / nsub)
    uint32_t nsub, kmax;
    unsigned long long *list; // entries: word index << 7 | shift << 2 | allele xor
    uint32_t *count;          // entries appended by this launch
    uint32_t *count_other;    // zeroed here for the next generation
    uint32_t cap;
    int *err_flag;
    // optional event dump
    uint32_t *dump_counters;  // [1] = HR events
    uint32_t dump_cap;
    uint32_t *d_hr_rec, *d_hr_locus, *d_hr_donor, *d_hr_seq;
    uint8_t *d_hr_value;
};

template <bool DUMP>
__global__ void __launch_bounds__(HR_WARPS * 32) hr_collect_kernel(const HrArgs a)
{
    __shared__ uint32_t claim_all[HR_WARPS][REGION_SITES / 32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.count_other = 0u;
    const uint64_t item = (uint64_t)blockIdx.x * HR_WARPS + warp;
    if (item >= (uint64_t)a.n_rows * a.n_regions) return;
    const uint32_t reg = (uint32_t)(item / a.n_rows), row = (uint32_t)(item % a.n_rows);
    const uint32_t greg = a.region0 + reg;

    const uint4 cctr = make_ctr(greg, row, a.gen, STREAM_CORE_HR_COUNT);
    const uint32_t first = philox4x32_10(cctr, a.key).x;
    const uint32_t K = stream_count(cctr, a.key, first, a.tab, a.nsub, a.kmax);     // warp-uniform
    if (K == 0u) return;

    const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
    const uint64_t rem_sites = a.site_limit - reg_site0;
    const uint32_t lim = rem_sites < REGION_SITES ? (uint32_t)rem_sites : REGION_SITES;
    const uint32_t n_other = a.n_rows - 1u;
    const uint64_t reg_word0 = (uint64_t)reg * REGION_WORDS;
    const uint64_t own_word0 = (uint64_t)row * a.row_stride_words + reg_word0;

    uint32_t *claim = claim_all[warp];
    const bool multi = K > 32u;
    if (multi) {
#pragma unroll
        for (int i = 0; i < (int)(REGION_SITES / 32 / 32); i++) claim[i * 32 + lane] = 0u;
        __syncwarp();
    }

#pragma unroll 1
    for (int base = (int)((K - 1u) & ~31u); base >= 0; base -= 32) {
        const uint32_t e = (uint32_t)base + lane;
        uint32_t pos = 0, donor = 0;
        bool valid = false;
        if (e < K) {
            const uint4 r = philox4x32_10(make_uint4(greg, row, a.gen, HR_EVENT_W0 + e), a.key);
            pos = r.x >> 19;                                                   // uniform site of the region
            donor = (uint32_t)__umul64hi(((uint64_t)r.y << 32) | r.z, (uint64_t)n_other);   // bias <= N / 2^64
            donor += donor >= row ? 1u : 0u;                                   // population.rs:616-619
            valid = pos < lim;                                                 // ragged last region: thinned away
        }
        // the last event of a cell wins (population.rs:745): highest lane of this window ...
        const uint32_t same = __match_any_sync(0xffffffffu, valid ? pos : (0x80000000u | lane));
        bool keep = valid && ((same >> lane) >> 1) == 0u;
        // ... unless a later window (visited earlier) already claimed the cell
        if (multi && keep) {
            const uint32_t bit = 1u << (pos & 31u);
            keep = (atomicOr(&claim[pos >> 5], bit) & bit) == 0u;
        }
        uint32_t delta = 0, sh = (pos & 15u) * 2u;
        if (keep || (DUMP && valid)) {
            const uint32_t dw = __ldg(a.state + (uint64_t)donor * a.row_stride_words + reg_word0 + (pos >> 4));
            const uint32_t v = (dw >> sh) & 3u;
            if (keep) delta = ((__ldg(a.state + own_word0 + (pos >> 4)) >> sh) & 3u) ^ v;
            if (DUMP) {
                const uint32_t slot = atomicAdd(&a.dump_counters[1], 1u);
                if (slot < a.dump_cap) {
                    a.d_hr_rec[slot] = row;
                    a.d_hr_locus[slot] = (uint32_t)(reg_site0 + pos);
                    a.d_hr_donor[slot] = donor;
                    a.d_hr_seq[slot] = e;
                    a.d_hr_value[slot] = (uint8_t)(1u << v);
                }
            }
        }
        const uint32_t emit = __ballot_sync(0xffffffffu, delta != 0u);
        if (emit) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(a.count, (uint32_t)__popc(emit));
            at = __shfl_sync(0xffffffffu, at, 0) + (uint32_t)__popc(emit & ((1u << lane) - 1u));
            if (delta) {
                if (at < a.cap)
                    a.list[at] = ((unsigned long long)(own_word0 + (pos >> 4)) << 7) | (sh << 2) | delta;
                else
                    *a.err_flag = 2;
            }
        }
        if (multi) __syncwarp();
    }
}

__global__ void __launch_bounds__(256) hr_apply_kernel(uint32_t *state, const unsigned long long *list,
                                                       const uint32_t *count, uint32_t cap)
{
    const uint32_t n = min(*count, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long e = list[i];
        atomicXor(state + (e >> 7), (uint32_t)(e & 3u) << (uint32_t)((e >> 2) & 31u));
    }
}

}  // namespace pansim
