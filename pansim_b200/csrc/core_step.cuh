// core_step.cuh -- K4: the fused core-genome generation step.
//
// One pass over the 2-bit packed core alignment does, per output row i:
//   gather-by-parent   next[i,:] = pop[parents[i],:]      (population.rs:450-465)
//   SNP mutation       population.rs:512-539
//   homologous recomb. population.rs:544-751 (core branch)
// Every output cell is a pure function of (old state, parents, Philox key), so
// there are no races, no atomics, and the result is independent of grid size
// and of how the columns are sharded over GPUs.
//
// Exact thinning used instead of the reference's per-row event loops
// (Poisson splitting):
//   * per row the reference draws n ~ Poisson(lambda) SNP events at uniform
//     sites, each writing U{C,G,T} (core_vec[1 >> value] is always core_vec[0],
//     population.rs:531). Restricted to a 256-site block that is a
//     Poisson(256*lambda/L) number of events at uniform positions in the block.
//   * HR: every donor d != r emits Poisson(lambda_HR/((N-1)L)) events onto cell
//     (r,l); summed over donors the block receives Poisson(256*lambda_HR/L)
//     events, each with a donor uniform on the other N-1 rows, carrying the
//     donor's post-mutation (pre-recombination) allele at the same locus
//     (snapshot semantics, population.rs:693-695). The snapshot value is
//     recomputed from the donor's parent row plus the donor's own (counter-
//     based, hence reproducible) mutation events. Later events overwrite
//     earlier ones (population.rs:745).
//
// Data movement: warp-private TMA pipelines. Each warp owns CS_STAGES 2 KiB
// shared-memory buffers; lane 0 issues cp.async.bulk global->shared for the
// parent's region (mbarrier complete_tx), all lanes apply their events to their
// own bank-conflict-free words, then lane 0 issues cp.async.bulk shared->global
// into the child's row. No CTA-wide barrier in the steady state.
#pragma once
#include "common.cuh"

namespace pansim {

constexpr int CS_WARPS = 8;
constexpr int CS_STAGES = 4;
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr uint32_t POISSON_TABLE_MAX = 1024;

struct CoreStepArgs {
    const uint8_t *old_state;
    uint8_t *new_state;
    const uint32_t *parents;
    uint32_t n_rows;
    uint32_t n_regions;       // regions per (local) row
    uint64_t row_stride;      // bytes
    uint32_t region0;         // global index of local region 0
    uint64_t site_limit;      // global site index one past the last valid site of this shard
    uint2 key;
    uint32_t gen;
    const uint32_t *mut_thr;  // device tables (copied to shared memory)
    uint32_t mut_size, mut_nsub, mut_kmax;
    const uint32_t *hr_thr;
    uint32_t hr_size, hr_nsub, hr_kmax;
    // optional event dump (parity instrumentation)
    uint32_t *dump_counters;  // [0] = SNP events, [1] = HR events
    uint32_t dump_cap;
    uint32_t *d_mut_row, *d_mut_site, *d_mut_seq;
    uint8_t *d_mut_allele;
    uint32_t *d_hr_rec, *d_hr_locus, *d_hr_donor, *d_hr_seq;
    uint8_t *d_hr_value;
};

static inline size_t core_step_smem_bytes(uint32_t mut_size, uint32_t hr_size)
{
    return (size_t)CS_WARPS * CS_STAGES * REGION_BYTES + (size_t)CS_WARPS * CS_STAGES * sizeof(uint64_t) +
           (size_t)(mut_size + hr_size) * sizeof(uint32_t);
}

// number of SNP events for one (site block, row) and the stream positioned after it
__device__ __forceinline__ uint32_t draw_count(BitStream &bs, const uint32_t *thr, uint32_t size,
                                               uint32_t nsub, uint32_t kmax)
{
    uint32_t k = 0;
    for (uint32_t s = 0; s < nsub; s++) k += poisson_from_uniform(thr, size, kmax, bs.take32());
    return k;
}

// one SNP event: position in the 256-site block (8 bits) and a new allele
// uniform on {C,G,T} = codes {1,2,3}: 2-bit draws, rejecting 0.
__device__ __forceinline__ void draw_snp(BitStream &bs, uint32_t &pos, uint32_t &allele)
{
    const uint32_t x = bs.take(10);
    pos = x & 255u;
    allele = x >> 8;
    while (allele == 0) allele = bs.take(2);
}

template <bool RNG, bool DUMP>
__global__ void __launch_bounds__(CS_THREADS) core_step_kernel(const CoreStepArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *bars_all = reinterpret_cast<uint64_t *>(smem_raw + (size_t)CS_WARPS * CS_STAGES * REGION_BYTES);
    uint32_t *tab_mut = reinterpret_cast<uint32_t *>(bars_all + CS_WARPS * CS_STAGES);
    uint32_t *tab_hr = tab_mut + a.mut_size;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *stages = smem_raw + (size_t)warp * CS_STAGES * REGION_BYTES;
    uint64_t *bars = bars_all + warp * CS_STAGES;

    if (RNG) {
        for (uint32_t i = threadIdx.x; i < a.mut_size; i += CS_THREADS) tab_mut[i] = a.mut_thr[i];
        for (uint32_t i = threadIdx.x; i < a.hr_size; i += CS_THREADS) tab_hr[i] = a.hr_thr[i];
    }
    if (lane == 0) {
        for (int s = 0; s < CS_STAGES; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t total = a.n_rows * a.n_regions;
    const uint32_t n_warps = gridDim.x * CS_WARPS;
    const uint32_t gw = blockIdx.x * CS_WARPS + warp;
    if (gw >= total) return;
    const uint32_t n_my = (total - gw + n_warps - 1) / n_warps;

    auto issue_load = [&](uint32_t j) {      // lane 0 only
        const uint32_t t = gw + j * n_warps;
        const uint32_t row = t / a.n_regions, reg = t - row * a.n_regions;
        const uint8_t *src = a.old_state + (uint64_t)a.parents[row] * a.row_stride + (uint64_t)reg * REGION_BYTES;
        const uint32_t s = j % CS_STAGES;
        mbar_arrive_expect_tx(&bars[s], REGION_BYTES);
        bulk_g2s(stages + s * REGION_BYTES, src, REGION_BYTES, &bars[s]);
    };

    if (lane == 0) {
        const uint32_t pre = n_my < (uint32_t)(CS_STAGES - 1) ? n_my : (uint32_t)(CS_STAGES - 1);
        for (uint32_t j = 0; j < pre; j++) issue_load(j);
    }

    for (uint32_t j = 0; j < n_my; j++) {
        const uint32_t t = gw + j * n_warps;
        const uint32_t row = t / a.n_regions, reg = t - row * a.n_regions;
        const uint32_t s = j % CS_STAGES;
        uint32_t *sw = reinterpret_cast<uint32_t *>(stages + s * REGION_BYTES);
        mbar_wait(&bars[s], (j / CS_STAGES) & 1u);

        if (RNG) {
            const uint32_t greg = a.region0 + reg;
            const uint32_t block_id = greg * 32u + lane;
            const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
            const uint64_t rem = a.site_limit - reg_site0;
            const uint32_t lim = rem >= REGION_SITES ? REGION_SITES : (uint32_t)rem;

            // ---- SNP mutation (population.rs:512-539) ----
            if (a.mut_nsub) {
                BitStream bs(make_ctr(block_id, row, a.gen, STREAM_CORE_MUT), a.key);
                const uint32_t k = draw_count(bs, tab_mut, a.mut_size, a.mut_nsub, a.mut_kmax);
                for (uint32_t e = 0; e < k; e++) {
                    uint32_t pos, al;
                    draw_snp(bs, pos, al);
                    const uint32_t widx = ((pos >> 4) << 5) + lane;
                    const uint32_t sir = widx * 16u + (pos & 15u);
                    if (sir < lim) {
                        const uint32_t sh = (pos & 15u) * 2u;
                        uint32_t w = sw[widx];
                        w = (w & ~(3u << sh)) | (al << sh);
                        sw[widx] = w;
                        if (DUMP) {
                            const uint32_t slot = atomicAdd(&a.dump_counters[0], 1u);
                            if (slot < a.dump_cap) {
                                a.d_mut_row[slot] = row;
                                a.d_mut_site[slot] = (uint32_t)(reg_site0 + sir);
                                a.d_mut_seq[slot] = e;
                                a.d_mut_allele[slot] = (uint8_t)(1u << al);
                            }
                        }
                    }
                }
            }

            // ---- homologous recombination (population.rs:544-751, core) ----
            if (a.hr_nsub) {
                BitStream hs(make_ctr(block_id, row, a.gen, STREAM_CORE_HR), a.key);
                const uint32_t kh = draw_count(hs, tab_hr, a.hr_size, a.hr_nsub, a.hr_kmax);
                for (uint32_t e = 0; e < kh; e++) {
                    const uint32_t pos = hs.take(8);
                    uint32_t d = hs.below(a.n_rows - 1u);      // population.rs:584, 616-619
                    d += (d >= row) ? 1u : 0u;
                    const uint32_t widx = ((pos >> 4) << 5) + lane;
                    const uint32_t sir = widx * 16u + (pos & 15u);
                    if (sir >= lim) continue;
                    const uint32_t sh = (pos & 15u) * 2u;
                    // donor's allele after gather, before mutation
                    const uint32_t *dsrc = reinterpret_cast<const uint32_t *>(
                        a.old_state + (uint64_t)a.parents[d] * a.row_stride + (uint64_t)reg * REGION_BYTES);
                    uint32_t val = (__ldg(dsrc + widx) >> sh) & 3u;
                    // donor's own SNP events in this block (same counter => same events)
                    if (a.mut_nsub) {
                        BitStream ds(make_ctr(block_id, d, a.gen, STREAM_CORE_MUT), a.key);
                        const uint32_t kd = draw_count(ds, tab_mut, a.mut_size, a.mut_nsub, a.mut_kmax);
                        for (uint32_t e2 = 0; e2 < kd; e2++) {
                            uint32_t p2, a2;
                            draw_snp(ds, p2, a2);
                            if (p2 == pos) val = a2;
                        }
                    }
                    uint32_t w = sw[widx];
                    w = (w & ~(3u << sh)) | (val << sh);
                    sw[widx] = w;
                    if (DUMP) {
                        const uint32_t slot = atomicAdd(&a.dump_counters[1], 1u);
                        if (slot < a.dump_cap) {
                            a.d_hr_rec[slot] = row;
                            a.d_hr_locus[slot] = (uint32_t)(reg_site0 + sir);
                            a.d_hr_donor[slot] = d;
                            a.d_hr_seq[slot] = e;
                            a.d_hr_value[slot] = (uint8_t)(1u << val);
                        }
                    }
                }
            }
            fence_proxy_async();     // generic-proxy writes -> visible to the bulk store
        }
        __syncwarp();

        if (lane == 0) {
            uint8_t *dst = a.new_state + (uint64_t)row * a.row_stride + (uint64_t)reg * REGION_BYTES;
            bulk_s2g(dst, sw, REGION_BYTES);
            bulk_commit();
            const uint32_t jn = j + CS_STAGES - 1;
            if (jn < n_my) {
                bulk_wait_read<1>();      // the store that last used stage (j-1)%S has left smem
                issue_load(jn);
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait<0>();
}

}  // namespace pansim
