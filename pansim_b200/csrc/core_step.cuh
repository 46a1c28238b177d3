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
// (Poisson splitting / thinning):
//   * per row the reference draws n ~ Poisson(lambda) SNP events at uniform
//     sites, each writing U{C,G,T} (core_vec[1 >> value] is always core_vec[0],
//     population.rs:531). Restricted to a 256-site block that is a
//     Poisson(256*lambda/L) number of events at uniform positions (one byte of
//     Philox output each); the alleles are successive base-3 digits of a
//     64-bit Philox word W (digit = floor(3W / 2^64), W <- 3W mod 2^64), i.e.
//     uniform on codes {1,2,3} = {C,G,T} up to 3^20/2^64 = 2e-10 for the 20
//     digits taken per word, with no rejection loop.
//   * HR: every donor d != r emits Poisson(lambda_HR/((N-1)L)) events onto cell
//     (r,l); summed over donors the block receives Poisson(256*lambda_HR/L)
//     events, each with a donor uniform on the other N-1 rows, carrying the
//     donor's post-mutation (pre-recombination) allele at the same locus
//     (snapshot semantics, population.rs:693-695). The snapshot value is
//     recomputed from the donor's parent row plus the donor's own (counter-
//     based, hence reproducible) SNP slots. Later events overwrite earlier ones
//     (population.rs:745).
//
// Control flow is warp-uniform: every lane makes the same two Philox calls for
// its first 20 SNP events (a lane that drew more handles the excess in a short
// tail), positions are extracted with compile-time byte selects, the snapshot
// probe compares four position bytes per instruction, and the HR events of all
// 32 lanes are compacted into a
// shared-memory queue so that the expensive snapshot recomputation runs with
// all lanes busy.
//
// Data movement: warp-private TMA pipelines. Each warp owns CS_STAGES 2 KiB
// shared-memory buffers; lane 0 issues cp.async.bulk global->shared for the
// parent's region (mbarrier complete_tx), all lanes apply their events to their
// own bank-conflict-free words, then lane 0 issues cp.async.bulk shared->global
// into the child's row. No CTA-wide barrier in the steady state.
#pragma once
#include <utility>
#include "common.cuh"

namespace pansim {

constexpr int CS_WARPS = 8;
constexpr int CS_STAGES = 3;
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr int HRQ_CAP = 64;                       // HR queue entries per warp
constexpr uint32_t POISSON_TABLE_MAX = 1024;
constexpr uint32_t SNP_GROUP = 20;                // SNP events per pair of Philox calls

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
    const uint32_t *mut_tab;  // device image [256 guide][mut_size thresholds] (copied to shared memory)
    uint32_t mut_size, mut_nsub, mut_kmax;
    const uint32_t *hr_tab;
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
           (size_t)CS_WARPS * HRQ_CAP * sizeof(uint2) +
           (size_t)(2 * GUIDE_ENTRIES + mut_size + hr_size) * sizeof(uint32_t);
}

// Poisson count of a (block,row) stream. Draw 0 uses `first` (word x of Philox
// call 0); a mean above the table range adds draws from dedicated count calls
// (exact by additivity).
__device__ __forceinline__ uint32_t stream_count(uint4 ctr, uint2 key, uint32_t first, const uint32_t *tab,
                                                 uint32_t nsub, uint32_t kmax)
{
    uint32_t k = poisson_from_uniform(tab, kmax, first);
    for (uint32_t s = 1; s < nsub; s++) {
        uint4 c = ctr;
        c.w |= 0x8000u | ((s - 1) >> 2);
        const uint4 r = philox4x32_10(c, key);
        const uint32_t sel = (s - 1) & 3u;
        const uint32_t u = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
        k += poisson_from_uniform(tab, kmax, u);
    }
    return k;
}

// next base-3 digit of the 64-bit fraction hi:lo (digit = floor(3F), F <- frac(3F))
__device__ __forceinline__ uint32_t next_trit(uint32_t &lo, uint32_t &hi)
{
    const uint64_t a = (uint64_t)lo * 3u;
    const uint64_t b = (uint64_t)hi * 3u + (uint32_t)(a >> 32);
    lo = (uint32_t)a;
    hi = (uint32_t)b;
    return (uint32_t)(b >> 32);
}

// SNP stream layout of (block,row), group g = events 20g .. 20g+19:
//   call 2g   : x = Poisson count word (g == 0 only), y:z = allele digit source, w = positions 0..3
//   call 2g+1 : x,y,z,w = positions 4..19
struct SnpGroup {
    uint32_t tlo, thi;
    uint32_t p[5];
};

__device__ __forceinline__ SnpGroup snp_group(uint4 ctr, uint2 key, uint32_t g, const uint4 *first)
{
    uint4 c0 = ctr, c1 = ctr;
    c0.w += 2u * g;
    c1.w += 2u * g + 1u;
    const uint4 a = first ? *first : philox4x32_10(c0, key);
    const uint4 b = philox4x32_10(c1, key);
    SnpGroup r;
    r.tlo = a.y; r.thi = a.z;
    r.p[0] = a.w; r.p[1] = b.x; r.p[2] = b.y; r.p[3] = b.z; r.p[4] = b.w;
    return r;
}

template <bool DUMP>
struct MutApply {
    uint32_t lane_base;       // shared-space byte address of this lane's word 0
    uint32_t lane, pos_lim, k, row;
    uint64_t reg_site0;
    const CoreStepArgs *a;

    __device__ __forceinline__ void slot(uint32_t pos, uint32_t digit, uint32_t idx) const
    {
        if (idx < k && pos < pos_lim) {
            const uint32_t addr = lane_base + ((pos & 0xF0u) << 3);      // word (pos>>4)*32 + lane
            const uint32_t one = 1u << ((pos & 15u) * 2u);
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
            w = (w & ~(one * 3u)) | (one * digit + one);                 // code = digit + 1 in {C,G,T}
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(w) : "memory");
            if (DUMP) {
                const uint32_t s = atomicAdd(&a->dump_counters[0], 1u);
                if (s < a->dump_cap) {
                    const uint32_t sir = (((pos >> 4) << 5) + lane) * 16u + (pos & 15u);
                    a->d_mut_row[s] = row;
                    a->d_mut_site[s] = (uint32_t)(reg_site0 + sir);
                    a->d_mut_seq[s] = idx;
                    a->d_mut_allele[s] = (uint8_t)(2u << digit);
                }
            }
        }
    }

    __device__ __forceinline__ void group(SnpGroup g, uint32_t base) const
    {
#pragma unroll
        for (int j = 0; j < 5; j++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t pos = (g.p[j] >> (8 * b)) & 255u;
                const uint32_t digit = next_trit(g.tlo, g.thi);
                slot(pos, digit, base + 4 * j + b);
            }
        }
    }
};

// index (0..19) of the last of the first `n` position bytes equal to `pos`, or -1
__device__ __forceinline__ int find_last_pos(const SnpGroup &g, uint32_t pos, uint32_t n)
{
    const uint32_t rep = pos * 0x01010101u;
    int last = -1;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint32_t t = g.p[j] ^ rep;
        // exact per-byte zero test: bit 7 of each byte of z set iff that byte of t is 0
        uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
        const int nv = (int)n - 4 * j;                    // valid bytes in this word
        if (nv < 4) z &= nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u);
        if (z) last = 4 * j + ((31 - __clz(z)) >> 3);
    }
    return last;
}

// The SNP allele (code 1..3) that row d's own mutation writes at position `pos`
// of site block `block`, or 0 if none of its events hits that position.
__device__ __forceinline__ uint32_t snp_probe(uint32_t block, uint32_t d, uint32_t pos, const CoreStepArgs &a,
                                              const uint32_t *tab_mut)
{
    const uint4 ctr = make_ctr(block, d, a.gen, STREAM_CORE_MUT);
    const uint4 g0 = philox4x32_10(ctr, a.key);
    const uint32_t kd = stream_count(ctr, a.key, g0.x, tab_mut, a.mut_nsub, a.mut_kmax);
    SnpGroup g = snp_group(ctr, a.key, 0, &g0);
    int last = find_last_pos(g, pos, kd);
    uint32_t lo = g.tlo, hi = g.thi;
    for (uint32_t grp = 1; grp * SNP_GROUP < kd; grp++) {          // rare: more than 20 events
        const SnpGroup h = snp_group(ctr, a.key, grp, nullptr);
        const int l2 = find_last_pos(h, pos, kd - grp * SNP_GROUP);
        if (l2 >= 0) { last = l2; lo = h.tlo; hi = h.thi; }
    }
    if (last < 0) return 0u;
    uint32_t digit = 0;
    for (int i = 0; i <= last; i++) digit = next_trit(lo, hi);
    return digit + 1u;
}

// HR event e of a (block,row) stream: position and donor. Call c = e/2 of the
// HR stream carries events 2c and 2c+1; call 0 additionally carries the count.
//   call 0:  x = count word, y = donor rnd (e0), w = donor rnd (e1), z = positions
//   call c:  x = donor rnd (e 2c), y = donor rnd (e 2c+1), z = positions
__device__ __forceinline__ void hr_event(const uint4 g, bool first_call, uint32_t which, uint32_t n_other,
                                         uint4 hctr, uint2 key, uint32_t e, uint32_t &pos, uint32_t &donor_raw)
{
    const uint32_t rnd = first_call ? (which ? g.w : g.y) : (which ? g.y : g.x);
    pos = (g.z >> (8u * which)) & 255u;
    // Lemire multiply-shift, exact up to one retry (a fresh dedicated Philox word);
    // residual bias <= (n/2^32)^2
    uint64_t m = (uint64_t)rnd * n_other;
    const uint32_t l = (uint32_t)m;
    if (l < n_other) {
        const uint32_t t = (0u - n_other) % n_other;
        if (l < t) {
            uint4 c = hctr;
            c.w |= 0x4000u;
            c.y ^= 0x80000000u;
            c.x += e * 0x9E3779B9u;
            m = (uint64_t)philox4x32_10(c, key).x * n_other;
        }
    }
    donor_raw = (uint32_t)(m >> 32);
}

template <bool RNG, bool DUMP>
__global__ void __launch_bounds__(CS_THREADS, 3) core_step_kernel(const CoreStepArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *bars_all = reinterpret_cast<uint64_t *>(smem_raw + (size_t)CS_WARPS * CS_STAGES * REGION_BYTES);
    uint2 *hrq_all = reinterpret_cast<uint2 *>(bars_all + CS_WARPS * CS_STAGES);
    uint32_t *tab_mut = reinterpret_cast<uint32_t *>(hrq_all + CS_WARPS * HRQ_CAP);
    uint32_t *tab_hr = tab_mut + GUIDE_ENTRIES + a.mut_size;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *stages = smem_raw + (size_t)warp * CS_STAGES * REGION_BYTES;
    uint64_t *bars = bars_all + warp * CS_STAGES;
    uint2 *hrq = hrq_all + warp * HRQ_CAP;

    if (RNG) {
        for (uint32_t i = threadIdx.x; i < GUIDE_ENTRIES + a.mut_size; i += CS_THREADS) tab_mut[i] = a.mut_tab[i];
        for (uint32_t i = threadIdx.x; i < GUIDE_ENTRIES + a.hr_size; i += CS_THREADS) tab_hr[i] = a.hr_tab[i];
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
    // item t = gw + j*n_warps  ->  (row, reg), advanced incrementally
    const uint32_t d_row = n_warps / a.n_regions, d_reg = n_warps % a.n_regions;

    // the load side runs CS_STAGES-1 items ahead with its own (row, reg) cursor
    uint32_t l_row = gw / a.n_regions, l_reg = gw % a.n_regions;
    auto issue_load = [&](uint32_t j) {      // lane 0 only; must be called for j = 0,1,2,... in order
        const uint8_t *src = a.old_state + (uint64_t)a.parents[l_row] * a.row_stride + (uint64_t)l_reg * REGION_BYTES;
        const uint32_t s = j % CS_STAGES;
        mbar_arrive_expect_tx(&bars[s], REGION_BYTES);
        bulk_g2s(stages + s * REGION_BYTES, src, REGION_BYTES, &bars[s]);
        l_row += d_row; l_reg += d_reg;
        if (l_reg >= a.n_regions) { l_reg -= a.n_regions; l_row++; }
    };

    if (lane == 0) {
        const uint32_t pre = n_my < (uint32_t)(CS_STAGES - 1) ? n_my : (uint32_t)(CS_STAGES - 1);
        for (uint32_t j = 0; j < pre; j++) issue_load(j);
    }

    uint32_t row = gw / a.n_regions, reg = gw % a.n_regions;
    for (uint32_t j = 0; j < n_my; j++) {
        const uint32_t s = j % CS_STAGES;
        uint32_t *sw = reinterpret_cast<uint32_t *>(stages + s * REGION_BYTES);

        // RNG work that does not need the data is done before waiting for the TMA load
        const uint32_t greg = a.region0 + reg;
        const uint32_t block_id = greg * 32u + lane;
        const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
        // positions >= pos_lim fall beyond the end of the alignment (ragged last region):
        // site-in-region = (pos>>4)*512 + lane*16 + (pos&15) < lim  <=>  pos < pos_lim
        uint32_t pos_lim = 256u;
        {
            const uint64_t rem = a.site_limit - reg_site0;
            if (rem < REGION_SITES) {
                const int q = (int)rem - (int)(lane * 16u);
                pos_lim = q <= 0 ? 0u : (uint32_t)(q >> 9) * 16u + min(16u, (uint32_t)(q & 511));
            }
        }
        const uint4 mctr = make_ctr(block_id, row, a.gen, STREAM_CORE_MUT);
        const uint4 hctr = make_ctr(block_id, row, a.gen, STREAM_CORE_HR);
        uint4 hg0 = make_uint4(0, 0, 0, 0);
        uint32_t k = 0, kh = 0;
        SnpGroup mg;
        if (RNG) {
            if (a.mut_nsub) {
                const uint4 mg0 = philox4x32_10(mctr, a.key);
                k = stream_count(mctr, a.key, mg0.x, tab_mut, a.mut_nsub, a.mut_kmax);
                mg = snp_group(mctr, a.key, 0, &mg0);
            }
            if (a.hr_nsub) {
                hg0 = philox4x32_10(hctr, a.key);
                kh = stream_count(hctr, a.key, hg0.x, tab_hr, a.hr_nsub, a.hr_kmax);
            }
        }

        mbar_wait(&bars[s], (j / CS_STAGES) & 1u);

        if (RNG) {
            // ---- SNP mutation (population.rs:512-539) ----
            if (a.mut_nsub) {
                const MutApply<DUMP> f{smem_u32(sw) + lane * 4u, lane, pos_lim, k, row, reg_site0, &a};
                f.group(mg, 0u);
                for (uint32_t grp = 1; grp * SNP_GROUP < k; grp++)       // rare: more than 20 events
                    f.group(snp_group(mctr, a.key, grp, nullptr), grp * SNP_GROUP);
            }

            // ---- homologous recombination (population.rs:544-751, core) ----
            if (a.hr_nsub) {
                // exclusive prefix of event counts over the warp -> global event index
                uint32_t incl = kh;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += v;
                }
                const uint32_t pre = incl - kh;
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t n_other = a.n_rows - 1u;
                for (uint32_t win = 0; win < tot; win += HRQ_CAP) {
                    // 1) owners enqueue their events that fall in [win, win + HRQ_CAP)
                    if (kh && pre + kh > win && pre < win + HRQ_CAP) {
                        uint4 g = hg0;
                        uint32_t have_call = 0;
                        for (uint32_t e = 0; e < kh; e++) {
                            const uint32_t gi = pre + e;
                            if (gi < win) continue;
                            if (gi >= win + HRQ_CAP) break;
                            const uint32_t call = e >> 1;
                            if (call != have_call) {
                                uint4 c = hctr;
                                c.w += call;
                                g = philox4x32_10(c, a.key);
                                have_call = call;
                            }
                            uint32_t pos, d;
                            hr_event(g, call == 0, e & 1u, n_other, hctr, a.key, e, pos, d);
                            d += (d >= row) ? 1u : 0u;                  // population.rs:616-619
                            const uint32_t ok = pos < pos_lim ? 1u : 0u;
                            hrq[gi - win] = make_uint2(d, pos | (lane << 8) | (ok << 13));
                        }
                    }
                    __syncwarp();
                    // 2) all lanes: snapshot value of queue entries (donor after SNPs, before HR)
                    const uint32_t n_q = min((uint32_t)HRQ_CAP, tot - win);
                    for (uint32_t q0 = 0; q0 < n_q; q0 += 32) {
                        const uint32_t q = q0 + lane;
                        if (q < n_q) {
                            const uint2 ent = hrq[q];
                            if ((ent.y >> 13) & 1u) {
                                const uint32_t pos = ent.y & 255u, owner = (ent.y >> 8) & 31u;
                                const uint32_t d = ent.x;
                                const uint32_t widx = ((pos >> 4) << 5) + owner;
                                const uint32_t sh = (pos & 15u) * 2u;
                                const uint32_t *dsrc = reinterpret_cast<const uint32_t *>(
                                    a.old_state + (uint64_t)a.parents[d] * a.row_stride + (uint64_t)reg * REGION_BYTES);
                                uint32_t val = (__ldg(dsrc + widx) >> sh) & 3u;
                                if (a.mut_nsub) {
                                    const uint32_t m = snp_probe(greg * 32u + owner, d, pos, a, tab_mut);
                                    if (m) val = m;
                                }
                                hrq[q].y = ent.y | (val << 30);        // value parked in bits 30..31
                            }
                        }
                    }
                    __syncwarp();
                    // 3) owners apply their events of this window in draw order (later wins)
                    if (kh && pre + kh > win && pre < win + HRQ_CAP) {
                        const uint32_t e_lo = pre >= win ? 0u : win - pre;
                        const uint32_t e_hi = min(kh, win + HRQ_CAP - pre);
                        for (uint32_t e = e_lo; e < e_hi; e++) {
                            const uint2 ent = hrq[pre + e - win];
                            if (!((ent.y >> 13) & 1u)) continue;
                            const uint32_t pos = ent.y & 255u, val = ent.y >> 30;
                            const uint32_t widx = ((pos >> 4) << 5) + lane;
                            const uint32_t sh = (pos & 15u) * 2u;
                            uint32_t w = sw[widx];
                            w = (w & ~(3u << sh)) | (val << sh);
                            sw[widx] = w;
                            if (DUMP) {
                                const uint32_t slot = atomicAdd(&a.dump_counters[1], 1u);
                                if (slot < a.dump_cap) {
                                    a.d_hr_rec[slot] = row;
                                    a.d_hr_locus[slot] = (uint32_t)(reg_site0 + widx * 16u + (pos & 15u));
                                    a.d_hr_donor[slot] = ent.x;
                                    a.d_hr_seq[slot] = e;
                                    a.d_hr_value[slot] = (uint8_t)(1u << val);
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            fence_proxy_async();     // generic-proxy writes -> visible to the bulk store
        }
        __syncwarp();

        if (lane == 0) {
            uint8_t *dst = a.new_state + (uint64_t)row * a.row_stride + (uint64_t)reg * REGION_BYTES;
            bulk_s2g(dst, sw, REGION_BYTES);
            bulk_commit();
            if (j + CS_STAGES - 1 < n_my) {
                bulk_wait_read<1>();      // the store that last used stage (j-1)%S has left smem
                issue_load(j + CS_STAGES - 1);
            }
        }
        __syncwarp();
        row += d_row; reg += d_reg;
        if (reg >= a.n_regions) { reg -= a.n_regions; row++; }
    }
    if (lane == 0) bulk_wait<0>();
}

}  // namespace pansim
