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
// Two steps, because every donor read must see the pre-recombination state:
//   collect  one warp per (region, 4 rows), one event per lane. Reads the
//            donor's and the recipient's cell from the state (read-only here)
//            and stores (site, allele XOR) for every cell whose LAST event
//            changes it, into the fixed slot array of that (region, row) item
//            (u16 entries; an item that overflows its slots appends to a small
//            global list). Same-cell events are resolved here:
//            __match_any_sync inside a window of 32 events, a per-warp
//            8192-bit claim map over windows visited last to first.
//   apply    atomicXor of the collected deltas (cells are distinct, so the
//            order of application is irrelevant).
// The steps run as two kernels after core_mut_kernel whenever a reader needs the
// materialised state (distances, downloads, replay steps, the event dump). In the
// default generate-mode loop the same events are instead recomputed and applied
// by the NEXT generation's core_mut_kernel (core_mut.cuh, hr_window_fetch/apply).
#pragma once
#include "common.cuh"

namespace pansim {

constexpr uint32_t STREAM_CORE_HR_EXTRA = 8;          // extra count draws of (region, row) for means above the table range
constexpr uint32_t HR_EVENT_W0 = 0x01000000u;         // counter word 3 of event e = HR_EVENT_W0 + e
// Philox call of event e of item (region, row): x -> site, (y, z) -> donor; word w of the call of
// event 0 is the uniform the item's event COUNT is drawn from (every lane of a warp makes the call
// of event `lane` anyway, so the count costs one shuffle instead of a Philox call of its own).
constexpr int HR_WARPS = 8;
constexpr uint32_t HR_ROWS_PER_TASK = 4;
constexpr uint32_t HR_CLAIM_WORDS = REGION_SITES / 32;   // per warp
constexpr uint32_t HR_APPLY_ITEMS = 32;                  // items per apply task (one warp)

struct HrArgs {
    uint32_t *state;          // packed core rows (gathered + mutated), words
    uint32_t n_rows, n_regions, region0;
    uint32_t lemire_t;        // 2^32 mod (n_rows - 1): rejection threshold of the donor draw
    uint64_t row_stride_words;
    uint64_t site_limit;
    uint2 key;
    uint32_t gen;
    const uint32_t *tab;      // device thresholds T[j] = round(CDF(j) * 2^32) of Poisson(region mean / nsub), padded with 0xFFFFFFFF
    uint32_t tab_words, nsub, kmax;
    uint16_t *slots;          // [items][slot_cap]: site-in-region | xor << 13; item = region * n_rows + row
    uint16_t *counts;         // [items]
    uint32_t slot_cap;
    unsigned long long *ovf;  // overflow entries: word index << 7 | shift << 2 | xor
    uint32_t *ovf_count;      // [0] used by this generation
    uint32_t *ovf_count_other;// zeroed for the next one
    uint32_t ovf_cap;
    int *err_flag;
    // optional event dump
    uint32_t *dump_counters;  // [1] = HR events
    uint32_t dump_cap;
    uint32_t *d_hr_rec, *d_hr_locus, *d_hr_donor, *d_hr_seq;
    uint8_t *d_hr_value;
};

static inline size_t hr_collect_smem_bytes(uint32_t tab_words)
{
    return (size_t)tab_words * 4 + (size_t)HR_WARPS * HR_CLAIM_WORDS * 4;
}

// ---- event count of one (region, row) item: warp-collective, exact CDF inversion ----------------
// k = #{j : T[j] <= u} capped at kmax; lane-parallel compare + ballot over the threshold table
__device__ __forceinline__ uint32_t hr_count_from_uniform(const uint32_t *thr, uint32_t size, uint32_t kmax, uint32_t u,
                                                          uint32_t lane)
{
    // size is a multiple of 32 (host pads with 0xFFFFFFFF; those only compare <= u for u = 2^32 - 1,
    // which the cap at kmax absorbs)
    uint32_t k = 0;
#pragma unroll 1
    for (uint32_t j0 = 0; j0 < size; j0 += 32) k += (uint32_t)__popc(__ballot_sync(0xffffffffu, thr[j0 + lane] <= u));
    return min(k, kmax);
}

// means above the table range: nsub - 1 more draws from dedicated count calls (exact by additivity)
__device__ __noinline__ uint32_t hr_count_extra(uint32_t greg, uint32_t row, uint32_t gen, uint2 key, const uint32_t *thr,
                                                uint32_t size, uint32_t nsub, uint32_t kmax, uint32_t lane)
{
    uint32_t k = 0;
    for (uint32_t s = 1; s < nsub; s++) {
        uint4 c = make_ctr(greg, row, gen, STREAM_CORE_HR_EXTRA);
        c.w |= 0x8000u | ((s - 1) >> 2);
        const uint4 r = philox_core(c, key);
        const uint32_t sel = (s - 1) & 3u;
        const uint32_t u = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
        k += hr_count_from_uniform(thr, size, kmax, u, lane);
    }
    return k;
}

// event e of item (global region greg, row): site of the region and donor row (population.rs:616-619)
struct HrEvent { uint32_t pos, donor, w; };
template <typename Key>      // uint2 key or precomputed PhiloxKeys
__device__ __forceinline__ HrEvent hr_event(uint32_t greg, uint32_t row, uint32_t gen, const Key &key, uint32_t e, uint32_t n_other,
                                            uint32_t lemire_t)      // lemire_t = 2^32 mod n_other (host)
{
    const uint4 r = philox_core(make_uint4(greg, row, gen, HR_EVENT_W0 + e), key);
    HrEvent ev;
    ev.pos = r.x >> 19;                                                                   // uniform site of the region
    // uniform on the other N - 1 rows: multiply-high of a 32-bit word with Lemire's rejection test; a
    // rejected word (probability < N / 2^32) is replaced by the next one (residual bias < (N / 2^32)^2)
    uint64_t m = (uint64_t)r.y * n_other;
    if ((uint32_t)m < n_other) {
        if ((uint32_t)m < lemire_t) m = (uint64_t)r.z * n_other;
    }
    ev.donor = (uint32_t)(m >> 32);
    ev.donor += ev.donor >= row ? 1u : 0u;                                                // population.rs:616-619
    ev.w = r.w;
    return ev;
}

// one (region, row) item; warp-collective
template <bool DUMP>
__device__ __forceinline__ void hr_collect_item(const HrArgs &a, uint32_t *claim, uint32_t reg, uint32_t row,
                                                uint32_t K, uint32_t lane)
{
    const uint64_t item = (uint64_t)reg * a.n_rows + row;
    if (K == 0u) {
        if (lane == 0) a.counts[item] = 0;
        return;
    }
    const uint32_t greg = a.region0 + reg;
    const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
    const uint64_t rem_sites = a.site_limit - reg_site0;
    const uint32_t lim = rem_sites < REGION_SITES ? (uint32_t)rem_sites : REGION_SITES;
    const uint32_t n_other = a.n_rows - 1u;
    const uint64_t reg_word0 = (uint64_t)reg * REGION_WORDS;
    const uint64_t own_word0 = (uint64_t)row * a.row_stride_words + reg_word0;
    uint16_t *slots = a.slots + item * a.slot_cap;

    const bool multi = K > 32u;
    if (multi) {
#pragma unroll
        for (int i = 0; i < (int)(HR_CLAIM_WORDS / 32); i++) claim[i * 32 + lane] = 0u;
        __syncwarp();
    }
    uint32_t n_emit = 0;
#pragma unroll 1
    for (int base = (int)((K - 1u) & ~31u); base >= 0; base -= 32) {
        const uint32_t e = (uint32_t)base + lane;
        uint32_t pos = 0, donor = 0;
        bool valid = false;
        if (e < K) {
            const HrEvent ev = hr_event(greg, row, a.gen, a.key, e, n_other, a.lemire_t);
            pos = ev.pos; donor = ev.donor;
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
        uint32_t delta = 0;
        const uint32_t sh = (pos & 15u) * 2u;
        if (keep || (DUMP && valid)) {
            // L2 loads: in the fused kernel these rows were written by other SMs of the same launch
            const uint32_t dw = __ldcg(a.state + (uint64_t)donor * a.row_stride_words + reg_word0 + (pos >> 4));
            const uint32_t v = (dw >> sh) & 3u;
            if (keep) delta = ((__ldcg(a.state + own_word0 + (pos >> 4)) >> sh) & 3u) ^ v;
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
        if (delta) {
            const uint32_t at = n_emit + (uint32_t)__popc(emit & ((1u << lane) - 1u));
            if (at < a.slot_cap) {
                slots[at] = (uint16_t)(pos | (delta << 13));
            } else {
                const uint32_t o = atomicAdd(a.ovf_count, 1u);
                if (o < a.ovf_cap)
                    a.ovf[o] = ((unsigned long long)(own_word0 + (pos >> 4)) << 7) | (sh << 2) | delta;
                else
                    *a.err_flag = 2;
            }
        }
        n_emit += (uint32_t)__popc(emit);
        if (multi) __syncwarp();
    }
    if (lane == 0) a.counts[item] = (uint16_t)min(n_emit, a.slot_cap);
}

// task = (region, rows 4q .. 4q+3). When every row has at most 32 events (one window each) the four
// windows are processed together: all Philox calls, then all loads, then all stores, so one L2/DRAM
// round trip serves the whole task.
template <bool DUMP>
__device__ __forceinline__ void hr_collect_task(const HrArgs &a, const uint32_t *tab, uint32_t *claim, uint32_t reg,
                                                uint32_t q, uint32_t lane)
{
    constexpr int R = (int)HR_ROWS_PER_TASK;
    const uint32_t greg = a.region0 + reg;
    const uint32_t n_other = a.n_rows - 1u;
    uint32_t K[R], pos[R], donor[R];
    uint32_t kmaxi = 0;
#pragma unroll
    for (int i = 0; i < R; i++) {
        const uint32_t row = q * R + i;
        K[i] = 0; pos[i] = 0; donor[i] = 0;
        if (row < a.n_rows) {
            const HrEvent ev = hr_event(greg, row, a.gen, a.key, lane, n_other, a.lemire_t);      // first window: event `lane`
            pos[i] = ev.pos; donor[i] = ev.donor;
            K[i] = hr_count_from_uniform(tab, a.tab_words, a.kmax, __shfl_sync(0xffffffffu, ev.w, 0), lane);   // warp-uniform
            if (a.nsub > 1) K[i] += hr_count_extra(greg, row, a.gen, a.key, tab, a.tab_words, a.nsub, a.kmax, lane);
        }
        kmaxi = max(kmaxi, K[i]);
    }
    if (kmaxi > 32u) {
#pragma unroll 1
        for (int i = 0; i < R; i++)
            if (q * R + i < a.n_rows) hr_collect_item<DUMP>(a, claim, reg, q * R + i, K[i], lane);
        return;
    }

    const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
    const uint64_t rem_sites = a.site_limit - reg_site0;
    const uint32_t lim = rem_sites < REGION_SITES ? (uint32_t)rem_sites : REGION_SITES;
    const uint32_t *col = a.state + (uint64_t)reg * REGION_WORDS;
    bool valid[R], keep[R];
#pragma unroll
    for (int i = 0; i < R; i++) valid[i] = lane < K[i] && pos[i] < lim;
    uint32_t dw[R], ow[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        const uint32_t same = __match_any_sync(0xffffffffu, valid[i] ? pos[i] : (0x80000000u | lane));
        keep[i] = valid[i] && ((same >> lane) >> 1) == 0u;
        dw[i] = 0; ow[i] = 0;
        if (keep[i] || (DUMP && valid[i])) dw[i] = __ldcg(col + (uint64_t)donor[i] * a.row_stride_words + (pos[i] >> 4));
        if (keep[i]) ow[i] = __ldcg(col + (uint64_t)(q * R + i) * a.row_stride_words + (pos[i] >> 4));
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
        const uint32_t row = q * R + i;
        if (row >= a.n_rows) break;
        const uint64_t item = (uint64_t)reg * a.n_rows + row;
        const uint32_t sh = (pos[i] & 15u) * 2u;
        const uint32_t v = (dw[i] >> sh) & 3u;
        const uint32_t delta = keep[i] ? (((ow[i] >> sh) & 3u) ^ v) : 0u;
        if (DUMP && valid[i]) {
            const uint32_t slot = atomicAdd(&a.dump_counters[1], 1u);
            if (slot < a.dump_cap) {
                a.d_hr_rec[slot] = row;
                a.d_hr_locus[slot] = (uint32_t)(reg_site0 + pos[i]);
                a.d_hr_donor[slot] = donor[i];
                a.d_hr_seq[slot] = lane;
                a.d_hr_value[slot] = (uint8_t)(1u << v);
            }
        }
        const uint32_t emit = __ballot_sync(0xffffffffu, delta != 0u);
        if (delta) a.slots[item * a.slot_cap + (uint32_t)__popc(emit & ((1u << lane) - 1u))] = (uint16_t)(pos[i] | (delta << 13));
        if (lane == 0) a.counts[item] = (uint16_t)__popc(emit);
    }
}

// apply the slots of items [item0, item0 + n) (n <= 32): one item per lane, its slot line read
// with 16-byte loads (slot_cap is a multiple of 32 entries = 64 bytes)
__device__ __forceinline__ void hr_apply_task(const HrArgs &a, uint64_t item0, uint32_t n, uint32_t lane)
{
    if (lane >= n) return;
    const uint64_t item = item0 + lane;
    const uint32_t cnt = __ldcg(a.counts + item);
    if (cnt == 0u) return;
    const uint32_t reg = (uint32_t)(item / a.n_rows), row = (uint32_t)(item % a.n_rows);
    uint32_t *own = a.state + (uint64_t)row * a.row_stride_words + (uint64_t)reg * REGION_WORDS;
    const uint4 *line = reinterpret_cast<const uint4 *>(a.slots + item * a.slot_cap);
#pragma unroll 1
    for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (e0 + 8u * j < cnt) ? __ldcg(line + (e0 >> 3) + j) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t en = (w[t >> 1] >> (16 * (t & 1))) & 0xFFFFu;
                const uint32_t pos = en & 0x1FFFu;
                if (e0 + 8u * j + t < cnt) atomicXor(own + (pos >> 4), (en >> 13) << ((pos & 15u) * 2u));
            }
        }
    }
}

__device__ __forceinline__ void hr_apply_overflow(const HrArgs &a, uint32_t first, uint32_t stride)
{
    const uint32_t n = min(*a.ovf_count, a.ovf_cap);
    for (uint32_t i = first; i < n; i += stride) {
        const unsigned long long e = __ldcg(a.ovf + i);
        atomicXor(a.state + (e >> 7), (uint32_t)(e & 3u) << (uint32_t)((e >> 2) & 31u));
    }
}

// ---- stand-alone launches (after core_mut_kernel has finished) ----------------------------

// grid = ceil(tasks / HR_WARPS), tasks = n_regions * ceil(n_rows / 4), task = region-major
template <bool DUMP>
__global__ void __launch_bounds__(HR_WARPS * 32) hr_collect_kernel(const HrArgs a)
{
    extern __shared__ uint32_t hr_smem[];
    uint32_t *tab = hr_smem;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *claim = hr_smem + a.tab_words + warp * HR_CLAIM_WORDS;
    for (uint32_t i = threadIdx.x; i < a.tab_words; i += blockDim.x) tab[i] = a.tab[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.ovf_count_other = 0u;
    __syncthreads();
    const uint32_t qn = (a.n_rows + HR_ROWS_PER_TASK - 1) / HR_ROWS_PER_TASK;
    const uint64_t task = (uint64_t)blockIdx.x * HR_WARPS + warp;
    if (task >= (uint64_t)a.n_regions * qn) return;
    hr_collect_task<DUMP>(a, tab, claim, (uint32_t)(task / qn), (uint32_t)(task % qn), lane);
}

// grid = ceil(items / (32 * HR_WARPS)); block 0 also drains the overflow list
__global__ void __launch_bounds__(HR_WARPS * 32) hr_apply_kernel(const HrArgs a)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t items = (uint64_t)a.n_rows * a.n_regions;
    const uint64_t item0 = ((uint64_t)blockIdx.x * HR_WARPS + warp) * HR_APPLY_ITEMS;
    if (item0 < items) hr_apply_task(a, item0, (uint32_t)min((uint64_t)HR_APPLY_ITEMS, items - item0), lane);
    hr_apply_overflow(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

}  // namespace pansim
