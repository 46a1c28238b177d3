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
//     Poisson(256*lambda/L) number of events at uniform positions. We draw
//     k' ~ Poisson(4/3 * 256*lambda/L) 10-bit SLOTS (8-bit position, 2-bit
//     code); a slot with code 0 is void. Thinning a Poisson count with
//     probability 3/4 leaves Poisson(256*lambda/L) events whose allele is
//     uniform on codes {1,2,3} = {C,G,T}: the same law, with no rejection loop.
//   * HR: every donor d != r emits Poisson(lambda_HR/((N-1)L)) events onto cell
//     (r,l); summed over donors the block receives Poisson(256*lambda_HR/L)
//     events, each with a donor uniform on the other N-1 rows, carrying the
//     donor's post-mutation (pre-recombination) allele at the same locus
//     (snapshot semantics, population.rs:693-695). The snapshot value is
//     recomputed from the donor's parent row plus the donor's own (counter-
//     based, hence reproducible) SNP slots. Later events overwrite earlier ones
//     (population.rs:745).
//
// Control flow is warp-uniform: Philox calls are issued in whole 128-bit groups
// for the whole warp (group count = warp maximum), slots are extracted with
// compile-time shifts, and the HR events of all 32 lanes are compacted into a
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
    uint32_t mut_size, mut_nsub, mut_kmax;     // table of Poisson(4/3 * SNP mean per block / nsub)
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
           (size_t)CS_WARPS * HRQ_CAP * sizeof(uint2) + (size_t)(mut_size + hr_size) * sizeof(uint32_t);
}

// slot T (10 bits) of a little-endian bit string held in w[0..NW)
template <int T, int NW>
__device__ __forceinline__ uint32_t slot10(const uint32_t (&w)[NW])
{
    constexpr int o = 10 * T, wi = o >> 5, sh = o & 31;
    static_assert(o + 10 <= 32 * NW, "slot out of range");
    if constexpr (sh <= 22) return (w[wi] >> sh) & 1023u;
    else return __funnelshift_r(w[wi], w[wi + 1], sh) & 1023u;
}

// Poisson count of a (block,row) stream. Draw 0 uses `first` (word x of Philox
// call 0); a mean above the table range adds draws from dedicated count calls
// (exact by additivity).
__device__ __forceinline__ uint32_t stream_count(uint4 ctr, uint2 key, uint32_t first, const uint32_t *thr,
                                                 uint32_t size, uint32_t nsub, uint32_t kmax)
{
    uint32_t k = poisson_from_uniform(thr, size, kmax, first);
    for (uint32_t s = 1; s < nsub; s++) {
        uint4 c = ctr;
        c.w |= 0x8000u | ((s - 1) >> 2);
        const uint4 r = philox4x32_10(c, key);
        const uint32_t sel = (s - 1) & 3u;
        const uint32_t u = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
        k += poisson_from_uniform(thr, size, kmax, u);
    }
    return k;
}

// ---- SNP slots: apply to the lane's words in shared memory ------------------
template <bool DUMP>
struct MutApply {
    uint32_t *sw;
    uint32_t lane, lim, k, row;
    uint64_t reg_site0;
    const CoreStepArgs *a;
    __device__ __forceinline__ void operator()(uint32_t x, uint32_t idx) const
    {
        const uint32_t al = x >> 8, pos = x & 255u;
        const uint32_t widx = ((pos >> 4) << 5) + lane;
        const uint32_t sir = widx * 16u + (pos & 15u);
        if (idx < k && al != 0u && sir < lim) {
            const uint32_t sh = (pos & 15u) * 2u;
            uint32_t w = sw[widx];
            w = (w & ~(3u << sh)) | (al << sh);
            sw[widx] = w;
            if (DUMP) {
                const uint32_t slot = atomicAdd(&a->dump_counters[0], 1u);
                if (slot < a->dump_cap) {
                    a->d_mut_row[slot] = row;
                    a->d_mut_site[slot] = (uint32_t)(reg_site0 + sir);
                    a->d_mut_seq[slot] = idx;
                    a->d_mut_allele[slot] = (uint8_t)(1u << al);
                }
            }
        }
    }
};

// ---- SNP slots: find the last valid slot that hits `pos` (snapshot recompute) -
struct MutProbe {
    uint32_t pos, k;
    uint32_t val;
    __device__ __forceinline__ void operator()(uint32_t x, uint32_t idx)
    {
        if (idx < k && (x & 255u) == pos && (x >> 8) != 0u) val = x >> 8;
    }
};

template <typename F, int... Is>
__device__ __forceinline__ void for_slots3(F &f, const uint32_t (&w)[3], uint32_t base, std::integer_sequence<int, Is...>)
{
    (f(slot10<Is, 3>(w), base + Is), ...);
}
template <typename F, int... Is>
__device__ __forceinline__ void for_slots4(F &f, const uint32_t (&w)[4], uint32_t base, std::integer_sequence<int, Is...>)
{
    (f(slot10<Is, 4>(w), base + Is), ...);
}

// Walk the SNP slots of stream (block,row): call 0 = {count word, 9 slots},
// calls 1.. = 12 slots each. `kwarp` (>= k) bounds the number of groups and is
// warp-uniform at the call sites that need uniform control flow.
template <typename F>
__device__ __forceinline__ void walk_snp_slots(F &f, uint4 ctr, uint2 key, const uint4 g0, uint32_t kwarp)
{
    {
        const uint32_t w[3] = {g0.y, g0.z, g0.w};
        for_slots3(f, w, 0u, std::make_integer_sequence<int, 9>{});
    }
    uint32_t call = 1;
    for (uint32_t base = 9; base < kwarp; base += 12, call++) {
        uint4 c = ctr;
        c.w += call;
        const uint4 g = philox4x32_10(c, key);
        const uint32_t w[4] = {g.x, g.y, g.z, g.w};
        for_slots4(f, w, base, std::make_integer_sequence<int, 12>{});
    }
}

__device__ __forceinline__ uint32_t warp_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }

// HR event e of a (block,row) stream: position and donor. Call c = e/2 of the
// HR stream carries events 2c and 2c+1; call 0 additionally carries the count.
//   call 0:  x = count word, y = donor rnd (e0), w = donor rnd (e1), z = positions
//   call c:  x = donor rnd (e 2c), y = donor rnd (e 2c+1), z = positions, w = retry word
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
__global__ void __launch_bounds__(CS_THREADS) core_step_kernel(const CoreStepArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *bars_all = reinterpret_cast<uint64_t *>(smem_raw + (size_t)CS_WARPS * CS_STAGES * REGION_BYTES);
    uint2 *hrq_all = reinterpret_cast<uint2 *>(bars_all + CS_WARPS * CS_STAGES);
    uint32_t *tab_mut = reinterpret_cast<uint32_t *>(hrq_all + CS_WARPS * HRQ_CAP);
    uint32_t *tab_hr = tab_mut + a.mut_size;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *stages = smem_raw + (size_t)warp * CS_STAGES * REGION_BYTES;
    uint64_t *bars = bars_all + warp * CS_STAGES;
    uint2 *hrq = hrq_all + warp * HRQ_CAP;

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

        // the RNG work that does not need the data is done before waiting for the TMA load
        const uint32_t greg = a.region0 + reg;
        const uint32_t block_id = greg * 32u + lane;
        const uint64_t reg_site0 = (uint64_t)greg * REGION_SITES;
        const uint64_t rem = a.site_limit - reg_site0;
        const uint32_t lim = rem >= REGION_SITES ? REGION_SITES : (uint32_t)rem;
        uint4 mctr = make_ctr(block_id, row, a.gen, STREAM_CORE_MUT);
        uint4 hctr = make_ctr(block_id, row, a.gen, STREAM_CORE_HR);
        uint4 mg0 = make_uint4(0, 0, 0, 0), hg0 = make_uint4(0, 0, 0, 0);
        uint32_t k = 0, kh = 0;
        if (RNG) {
            if (a.mut_nsub) {
                mg0 = philox4x32_10(mctr, a.key);
                k = stream_count(mctr, a.key, mg0.x, tab_mut, a.mut_size, a.mut_nsub, a.mut_kmax);
            }
            if (a.hr_nsub) {
                hg0 = philox4x32_10(hctr, a.key);
                kh = stream_count(hctr, a.key, hg0.x, tab_hr, a.hr_size, a.hr_nsub, a.hr_kmax);
            }
        }

        mbar_wait(&bars[s], (j / CS_STAGES) & 1u);

        if (RNG) {
            // ---- SNP mutation (population.rs:512-539) ----
            if (a.mut_nsub) {
                MutApply<DUMP> f{sw, lane, lim, k, row, reg_site0, &a};
                walk_snp_slots(f, mctr, a.key, mg0, warp_max(k));
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
                            const uint32_t widx = ((pos >> 4) << 5) + lane;
                            const uint32_t sir = widx * 16u + (pos & 15u);
                            const uint32_t ok = sir < lim ? 1u : 0u;
                            hrq[gi - win] = make_uint2(d, pos | (lane << 8) | (ok << 13));
                        }
                    }
                    __syncwarp();
                    // 2) all lanes: snapshot value of queue entries (donor after SNPs, before HR)
                    const uint32_t n_q = min((uint32_t)HRQ_CAP, tot - win);
                    for (uint32_t q0 = 0; q0 < n_q; q0 += 32) {
                        const uint32_t q = q0 + lane;
                        const bool active = q < n_q;
                        uint2 ent = active ? hrq[q] : make_uint2(0, 0);
                        const uint32_t pos = ent.y & 255u, owner = (ent.y >> 8) & 31u;
                        const bool ok = active && ((ent.y >> 13) & 1u);
                        const uint32_t d = ent.x;
                        const uint32_t widx = ((pos >> 4) << 5) + owner;
                        const uint32_t sh = (pos & 15u) * 2u;
                        uint32_t val = 0;
                        if (ok) {
                            const uint32_t *dsrc = reinterpret_cast<const uint32_t *>(
                                a.old_state + (uint64_t)a.parents[d] * a.row_stride + (uint64_t)reg * REGION_BYTES);
                            val = (__ldg(dsrc + widx) >> sh) & 3u;
                        }
                        if (a.mut_nsub) {
                            // the donor's SNP slots in the owner's site block: same counter => same slots
                            const uint4 dctr = make_ctr(greg * 32u + owner, d, a.gen, STREAM_CORE_MUT);
                            uint32_t kd = 0;
                            uint4 dg0 = make_uint4(0, 0, 0, 0);
                            if (ok) {
                                dg0 = philox4x32_10(dctr, a.key);
                                kd = stream_count(dctr, a.key, dg0.x, tab_mut, a.mut_size, a.mut_nsub, a.mut_kmax);
                            }
                            MutProbe pr{pos, kd, 0u};
                            walk_snp_slots(pr, dctr, a.key, dg0, warp_max(kd));
                            if (pr.val) val = pr.val;
                        }
                        if (active) hrq[q].y = ent.y | (val << 30);        // value parked in bits 30..31
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
