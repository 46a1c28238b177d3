// core_gen.cuh -- K4: one launch per generation for the core genome.
//
// The gather+SNP pass (core_mut.cuh) and the recombination pass (core_hr.cuh)
// are tied by one dependency: a recombination event reads its donor cell from
// the finished (gathered + mutated) rows, and donor and recipient cell share
// the locus (population.rs:693-695, 745). So recombination of a COLUMN BLOCK
// (a few 8192-site regions, all rows, ~20 MB) only needs the gather+SNP pass of
// that block. This kernel runs the three kinds of work as CTA roles of one
// grid, ordered by a ticket each CTA draws when it starts:
//
//   CTA of phase p:  gather+SNP items of block p, then collect tasks of block p-lag_c,
//                    then apply tasks of block p-lag_a
//
// Before its collect tasks a CTA waits (acquire spin on a counter) until every
// gather+SNP CTA of that block has published its rows; before its apply tasks,
// for the block's collect step. All waits are on CTAs with LOWER tickets, which
// have already started and depend only on still lower ones, so the grid cannot
// deadlock; the lags are a full wave of CTAs, so in steady state nothing spins. The
// recombination reads and atomics hit rows that were written microseconds
// earlier and are still in L2, and they run while the TMA pipelines of the next
// blocks stream: the latency-bound sparse work hides under the bandwidth-bound
// dense work.
//
// Results are identical to core_mut_kernel followed by hr_collect_kernel and
// hr_apply_kernel (same device functions, same counters).
#pragma once
#include "core_hr.cuh"
#include "core_mut.cuh"

namespace pansim {

struct GenSched {
    uint32_t n_blocks;        // column blocks
    uint32_t blk_regs;        // regions per full block
    uint32_t lag_collect;     // phases between gather+SNP of a block and its collect step (>= 1)
    uint32_t lag_apply;       // phases between gather+SNP of a block and its apply step (> lag_collect)
    uint32_t ctas;            // CTAs per phase (what a full block needs for its gather+SNP items)
    uint32_t ctas_overflow;   // trailing CTAs that drain the overflow list
    uint32_t *ctl;            // [0] ticket, [1 .. B] gather+SNP done, [1+B .. 2B] collect done (CTAs per block)
    uint32_t *ctl_other;      // control block of the next launch (zeroed here)
    uint32_t ctl_words;
};

static inline uint32_t gen_grid(const GenSched &s)
{
    return (s.n_blocks + s.lag_apply) * s.ctas + s.ctas_overflow;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// CTA-wide: wait until *ctr >= target (thread 0 spins), then everyone may read what the signallers wrote
__device__ __forceinline__ void cta_wait_counter(const uint32_t *ctr, uint32_t target)
{
    if (threadIdx.x == 0) {
        while (ld_acquire_gpu(ctr) < target) __nanosleep(100);
    }
    __syncthreads();
}

// CTA-wide: everything this CTA wrote (generic and bulk-async proxy) becomes visible, then *ctr += 1
__device__ __forceinline__ void cta_signal_counter(uint32_t *ctr)
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
    }
}

// Every CTA of phase p does, in this order: its share of the gather+SNP items of block p, its
// share of the collect tasks of block p - lag_collect, its share of the apply tasks of block
// p - lag_apply. The lags are chosen on the host so that the blocks waited for were finished a
// full wave of CTAs ago: the waits are there for correctness, in steady state they do not spin.
template <bool DUMP>
__global__ void __launch_bounds__(CM_THREADS, 4) core_gen_kernel(const CoreMutArgs a, const HrArgs h, const GenSched s)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint32_t s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&s.ctl[0], 1u);
    __syncthreads();
    const uint32_t ticket = s_ticket;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (ticket == 0) {
        for (uint32_t i = threadIdx.x; i < s.ctl_words; i += CM_THREADS) s.ctl_other[i] = 0u;
        if (threadIdx.x == 0) *h.ovf_count_other = 0u;
    }
    uint32_t *mut_done = s.ctl + 1, *collect_done = s.ctl + 1 + s.n_blocks;
    const uint32_t n_phases = s.n_blocks + s.lag_apply;
    const uint32_t phase = ticket / s.ctas, idx = ticket % s.ctas;
    const uint32_t qn = (a.n_rows + HR_ROWS_PER_TASK - 1) / HR_ROWS_PER_TASK;
    auto regs_of = [&](uint32_t b) { return min(s.blk_regs, a.n_regions - b * s.blk_regs); };

    if (phase >= n_phases) {
        // ---- overflow list: after the collect step of every block ----
        for (uint32_t b = 0; b < s.n_blocks; b++) cta_wait_counter(&collect_done[b], s.ctas);
        hr_apply_overflow(h, (ticket - n_phases * s.ctas) * CM_THREADS + threadIdx.x, s.ctas_overflow * CM_THREADS);
        return;
    }

    // ---- gather + SNP of column block `phase` ----
    if (phase < s.n_blocks) {
        const uint32_t regs = regs_of(phase);
        const uint32_t items = a.n_rows * regs;
        const uint32_t cta_items = CM_WARPS * a.items_per_warp;
        const uint32_t begin = idx * cta_items;
        if (begin < items) {
            const MutSmem m = mut_smem_carve(smem_dyn, a.mut_size);
            mut_cta_setup<true>(a, m);
            mut_cta_items<true, DUMP>(a, m, begin, min(items, begin + cta_items), regs, phase * s.blk_regs);
        }
        cta_signal_counter(&mut_done[phase]);       // its __syncthreads also releases the stage buffers
    }

    // ---- recombination, collect step of column block `phase - lag_collect` ----
    if (phase >= s.lag_collect && phase - s.lag_collect < s.n_blocks) {
        const uint32_t b = phase - s.lag_collect;
        const uint32_t tasks = regs_of(b) * qn;
        const uint32_t share = (tasks + s.ctas - 1) / s.ctas;
        const uint32_t t0 = idx * share, t1 = min(tasks, t0 + share);
        if (t0 < t1) {
            uint32_t *tab = reinterpret_cast<uint32_t *>(smem_dyn);
            uint32_t *claim = tab + h.tab_words + warp * HR_CLAIM_WORDS;
            for (uint32_t i = threadIdx.x; i < h.tab_words; i += CM_THREADS) tab[i] = h.tab[i];
            cta_wait_counter(&mut_done[b], s.ctas);                  // also orders the table copy
#pragma unroll 1
            for (uint32_t task = t0 + warp; task < t1; task += HR_WARPS)
                hr_collect_task<DUMP>(h, tab, claim, b * s.blk_regs + task / qn, task % qn, lane);
        }
        cta_signal_counter(&collect_done[b]);
    }

    // ---- recombination, apply step of column block `phase - lag_apply` ----
    if (phase >= s.lag_apply) {
        const uint32_t b = phase - s.lag_apply;
        const uint32_t items = a.n_rows * regs_of(b);
        const uint32_t tasks = (items + HR_APPLY_ITEMS - 1) / HR_APPLY_ITEMS;
        const uint32_t share = (tasks + s.ctas - 1) / s.ctas;
        const uint32_t t0 = idx * share, t1 = min(tasks, t0 + share);
        if (t0 < t1) {
            cta_wait_counter(&collect_done[b], s.ctas);
            const uint64_t blk_item0 = (uint64_t)b * s.blk_regs * a.n_rows;     // items are region-major
#pragma unroll 1
            for (uint32_t task = t0 + warp; task < t1; task += HR_WARPS) {
                const uint32_t i0 = task * HR_APPLY_ITEMS;
                hr_apply_task(h, blk_item0 + i0, min(HR_APPLY_ITEMS, items - i0), lane);
            }
        }
    }
}

}  // namespace pansim
