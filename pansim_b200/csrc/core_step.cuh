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
// 32 lanes are renumbered lane-major and handled one per lane, so that the
// expensive snapshot recomputation runs with all lanes busy.
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
constexpr uint32_t POISSON_TABLE_MAX = 1024;

struct CoreStepArgs {
    const uint8_t *old_state;
    uint8_t *new_state;
    const uint32_t *parents;
    uint32_t n_rows;
    uint32_t n_regions;       // regions per (local) row
    uint64_t row_stride;      // bytes
    uint32_t region0;         // global index of local region 0
    uint32_t items_per_warp;  // (row, region) items per warp; a CTA covers CS_WARPS * items_per_warp consecutive items
    uint32_t snapshot_pass;   // 1 = second pass of the two-pass mode: old_state already holds the gathered and
                              // mutated rows (identity gather, no SNPs; HR donors are read from it directly)
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
           (size_t)(2 * GUIDE_ENTRIES + mut_size + hr_size) * sizeof(uint32_t);
}

// Poisson count of a (block,row) stream. Draw 0 uses `first` (word x of Philox
// call 0); a mean above the table range adds draws from dedicated count calls
// (exact by additivity).
__device__ __noinline__ uint32_t stream_count_extra(uint4 ctr, uint2 key, const uint32_t *tab, uint32_t nsub,
                                                    uint32_t kmax)
{
    uint32_t k = 0;
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

__device__ __forceinline__ uint32_t stream_count(uint4 ctr, uint2 key, uint32_t first, const uint32_t *tab,
                                                 uint32_t nsub, uint32_t kmax)
{
    uint32_t k = poisson_from_uniform(tab, kmax, first);
    if (nsub > 1) k += stream_count_extra(ctr, key, tab, nsub, kmax);
    return k;
}

// SNP stream layout of (block,row); group g = events 28g .. 28g+27, calls 3g..3g+2:
//   call 3g   : x = Poisson count word (g == 0 only), y = digit bytes, z = reserve digit bytes,
//               w = positions 0..3
//   call 3g+1 : x,y,z,w = positions 4..19
//   call 3g+2 : x = digit bytes, y = reserve digit bytes, z,w = positions 20..27
// Part 0 = events 0..19, part 1 = events 20..27. Part 1 is only generated when
// some lane of the warp needs it, so the common case costs two Philox calls.
//
// Alleles: one digit byte v serves five consecutive events as its base-3 digits
// (event i of the chunk gets floor(v / 3^i) mod 3). v must be uniform on
// [0,243): a byte >= 243 is replaced by the reserve byte, and if that is >= 243
// too (p = 0.26 %) by a dedicated Philox word scaled to [0,243) (bias 6e-8).
constexpr uint32_t SNP_PART0 = 20, SNP_GROUP = 28;

__device__ __forceinline__ uint4 snp_call(uint4 ctr, uint2 key, uint32_t call)
{
    ctr.w += call;
    return philox4x32_10(ctr, key);
}

__device__ __noinline__ uint32_t digit_byte_fallback(uint4 ctr, uint2 key, uint32_t tag)
{
    ctr.w |= 0x2000u;
    ctr.x ^= 0x5bd1e995u * (tag + 1u);
    return __umulhi(philox4x32_10(ctr, key).x, 243u);
}

__device__ __forceinline__ uint32_t digit_byte(uint32_t tp, uint32_t tr, uint32_t c, uint4 ctr, uint2 key,
                                               uint32_t tag)
{
    uint32_t v = (tp >> (8u * c)) & 255u;
    if (v >= 243u) {
        v = (tr >> (8u * c)) & 255u;
        if (v >= 243u) v = digit_byte_fallback(ctr, key, tag);
    }
    return v;
}

template <bool DUMP>
struct MutApply {
    uint32_t lane_base;       // shared-space byte address of this lane's word 0
    uint32_t lane, pos_lim, k, row;
    uint64_t reg_site0;
    const CoreStepArgs *a;

    // `code` = new 2-bit allele (1..3). Branch-free: the read-modify-write is computed for
    // every slot and only the store is predicated on the slot being a real event, so the
    // five slots of a chunk form one basic block the scheduler can interleave.
    __device__ __forceinline__ void slot(uint32_t pos, uint32_t code, uint32_t idx, uint32_t kk) const
    {
        // positions beyond the end of a ragged last region are written too and
        // cleared again by the padding fix-up at the end of the item
        const uint32_t addr = lane_base + ((pos & 0xF0u) << 3);          // word (pos>>4)*32 + lane
        const uint32_t sh = (pos & 15u) * 2u;
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
        w = (w & ~(3u << sh)) | (code << sh);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %2, %3;\n\t@p st.shared.u32 [%0], %1;\n\t}"
                     ::"r"(addr), "r"(w), "r"(idx), "r"(kk) : "memory");
        if (DUMP) {
            if (idx < kk && pos < pos_lim) {
                const uint32_t s = atomicAdd(&a->dump_counters[0], 1u);
                if (s < a->dump_cap) {
                    const uint32_t sir = (((pos >> 4) << 5) + lane) * 16u + (pos & 15u);
                    a->d_mut_row[s] = row;
                    a->d_mut_site[s] = (uint32_t)(reg_site0 + sir);
                    a->d_mut_seq[s] = idx;
                    a->d_mut_allele[s] = (uint8_t)(1u << code);
                }
            }
        }
    }

    // apply the first n_ev events of one part (positions packed in p0..p4, 4 per word);
    // `base` = index of its first event. Loop over chunks of 5 events = one digit byte.
    __device__ __forceinline__ void part(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4,
                                         uint32_t n_ev, uint32_t kw, uint32_t tp, uint32_t tr, uint32_t base,
                                         uint4 ctr, uint2 key, uint32_t tag0) const
    {
        const uint32_t kk = min(k, base + n_ev);           // events of this part that exist for this lane
        const uint32_t n_run = min(n_ev, kw - base);       // ... and for the busiest lane of the warp
#pragma unroll 1
        for (uint32_t c = 0; 5u * c < n_run; c++) {
            uint32_t v = digit_byte(tp, tr, c, ctr, key, tag0 + c);
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const uint32_t q = (v * 171u) >> 9;        // v / 3 for v < 256
                const uint32_t code = v + 1u - 3u * q;     // base-3 digit + 1 = allele code of {C,G,T}
                v = q;
                const uint32_t pos = i < 4 ? ((p0 >> (8 * i)) & 255u) : (p1 & 255u);
                slot(pos, code, base + 5u * c + i, kk);
            }
            // advance the position string by 5 bytes
            p0 = __funnelshift_r(p1, p2, 8);
            p1 = __funnelshift_r(p2, p3, 8);
            p2 = __funnelshift_r(p3, p4, 8);
            p3 = p4 >> 8;
            p4 = 0;
        }
    }
};

// index (within the part) of the last of its first `n` position bytes equal to `pos`, or -1
template <int NW>
__device__ __forceinline__ int find_last_pos(const uint32_t (&p)[NW], uint32_t pos, int n)
{
    const uint32_t rep = pos * 0x01010101u;
    int last = -1;
#pragma unroll
    for (int j = 0; j < NW; j++) {
        const uint32_t t = p[j] ^ rep;
        // exact per-byte zero test: bit 7 of each byte of z set iff that byte of t is 0
        uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
        const int nv = n - 4 * j;                         // valid bytes in this word
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
    uint4 c0 = snp_call(ctr, a.key, 0);
    const uint32_t kd = stream_count(ctr, a.key, c0.x, tab_mut, a.mut_nsub, a.mut_kmax);
    int last = -1;
    uint32_t tp = 0, tr = 0, tag = 0;
#pragma unroll 1
    for (uint32_t base = 0, g = 0; base < kd; base += SNP_GROUP, g++) {
        if (g) c0 = snp_call(ctr, a.key, 3u * g);
        const uint4 c1 = snp_call(ctr, a.key, 3u * g + 1u);
        const uint32_t p0[5] = {c0.w, c1.x, c1.y, c1.z, c1.w};
        const int l0 = find_last_pos<5>(p0, pos, (int)(kd - base));
        if (l0 >= 0) { last = l0; tp = c0.y; tr = c0.z; tag = g * 8u; }
        if (kd > base + SNP_PART0) {
            const uint4 c2 = snp_call(ctr, a.key, 3u * g + 2u);
            const uint32_t p1[2] = {c2.z, c2.w};
            const int l1 = find_last_pos<2>(p1, pos, (int)(kd - base - SNP_PART0));
            if (l1 >= 0) { last = l1; tp = c2.x; tr = c2.y; tag = g * 8u + 4u; }
        }
    }
    if (last < 0) return 0u;
    const uint32_t c = (uint32_t)last / 5u, i = (uint32_t)last % 5u;
    uint32_t v = digit_byte(tp, tr, c, ctr, a.key, tag + c);
    for (uint32_t t = 0; t < i; t++) v = (v * 171u) >> 9;
    return v - 3u * ((v * 171u) >> 9) + 1u;
}

__device__ __noinline__ uint32_t hr_retry_word(uint4 c, uint2 key, uint32_t e)
{
    c.w |= 0x4000u;
    c.y ^= 0x80000000u;
    c.x += e * 0x9E3779B9u;
    return philox4x32_10(c, key).x;
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
        if (l < t) m = (uint64_t)hr_retry_word(hctr, key, e) * n_other;
    }
    donor_raw = (uint32_t)(m >> 32);
}

template <bool RNG, bool DUMP>
__global__ void __launch_bounds__(CS_THREADS, 4) core_step_kernel(const CoreStepArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t *bars_all = reinterpret_cast<uint64_t *>(smem_raw + (size_t)CS_WARPS * CS_STAGES * REGION_BYTES);
    uint32_t *tab_mut = reinterpret_cast<uint32_t *>(bars_all + CS_WARPS * CS_STAGES);
    uint32_t *tab_hr = tab_mut + GUIDE_ENTRIES + a.mut_size;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *stages = smem_raw + (size_t)warp * CS_STAGES * REGION_BYTES;
    uint64_t *bars = bars_all + warp * CS_STAGES;

    if (RNG) {
#pragma unroll 1
        for (uint32_t i = threadIdx.x; i < GUIDE_ENTRIES + a.mut_size; i += CS_THREADS) tab_mut[i] = a.mut_tab[i];
#pragma unroll 1
        for (uint32_t i = threadIdx.x; i < GUIDE_ENTRIES + a.hr_size; i += CS_THREADS) tab_hr[i] = a.hr_tab[i];
    }
    if (lane == 0) {
        for (int s = 0; s < CS_STAGES; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // CTA b covers items [b*C, (b+1)*C), C = CS_WARPS*items_per_warp; warp w takes b*C + w + 8j.
    // CTAs are short-lived on purpose: SM slots turn over every few tens of microseconds, so the
    // (higher-priority) accessory/selection kernels of the next generation can slip in between.
    const uint32_t total = a.n_rows * a.n_regions;
    const uint32_t cta_items = CS_WARPS * a.items_per_warp;
    const uint32_t cta_base = blockIdx.x * cta_items;
    const uint32_t cta_end = min(total, cta_base + cta_items);
    const uint32_t gw = cta_base + warp;
    if (gw >= cta_end) return;
    const uint32_t n_my = (cta_end - gw + CS_WARPS - 1) / CS_WARPS;
    // item t = gw + j*CS_WARPS  ->  (row, reg), advanced incrementally
    const uint32_t d_row = CS_WARPS / a.n_regions, d_reg = CS_WARPS % a.n_regions;

    // the load side runs CS_STAGES-1 items ahead with its own (row, reg) cursor (lane 0 only)
    uint32_t l_row = gw / a.n_regions, l_reg = gw % a.n_regions, l_j = 0;
#define PANSIM_ISSUE_LOAD()                                                                                  \
    do {                                                                                                     \
        const uint8_t *src_ = a.old_state + (uint64_t)(a.snapshot_pass ? l_row : a.parents[l_row]) * a.row_stride + \
                              (uint64_t)l_reg * REGION_BYTES;                                                \
        const uint32_t s_ = l_j % CS_STAGES;                                                                 \
        mbar_arrive_expect_tx(&bars[s_], REGION_BYTES);                                                      \
        bulk_g2s(stages + s_ * REGION_BYTES, src_, REGION_BYTES, &bars[s_]);                                 \
        l_j++; l_row += d_row; l_reg += d_reg;                                                               \
        if (l_reg >= a.n_regions) { l_reg -= a.n_regions; l_row++; }                                         \
    } while (0)

    if (lane == 0) {
        const uint32_t pre = n_my < (uint32_t)(CS_STAGES - 1) ? n_my : (uint32_t)(CS_STAGES - 1);
#pragma unroll 1
        for (uint32_t jj = 0; jj < pre; jj++) PANSIM_ISSUE_LOAD();
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
        uint32_t pos_lim = 256u, pos_lim_any = 256u;       // pos_lim_any < 256: ragged region (warp-uniform)
        {
            const uint64_t rem = a.site_limit - reg_site0;
            if (rem < REGION_SITES) {
                pos_lim_any = 0u;
                const int q = (int)rem - (int)(lane * 16u);
                pos_lim = q <= 0 ? 0u : (uint32_t)(q >> 9) * 16u + min(16u, (uint32_t)(q & 511));
            }
        }
        const uint4 mctr = make_ctr(block_id, row, a.gen, STREAM_CORE_MUT);
        const uint4 hctr = make_ctr(block_id, row, a.gen, STREAM_CORE_HR);
        uint4 hg0 = make_uint4(0, 0, 0, 0), mc0 = make_uint4(0, 0, 0, 0), mc1 = make_uint4(0, 0, 0, 0);
        uint32_t k = 0, kh = 0;
        if (RNG) {
            if (a.mut_nsub && !a.snapshot_pass) {
                mc0 = snp_call(mctr, a.key, 0);
                mc1 = snp_call(mctr, a.key, 1);
                k = stream_count(mctr, a.key, mc0.x, tab_mut, a.mut_nsub, a.mut_kmax);
            }
            if (a.hr_nsub) {
                hg0 = philox4x32_10(hctr, a.key);
                kh = stream_count(hctr, a.key, hg0.x, tab_hr, a.hr_nsub, a.hr_kmax);
            }
        }

        mbar_wait(&bars[s], (j / CS_STAGES) & 1u);

        if (RNG) {
            // ---- SNP mutation (population.rs:512-539) ----
            if (a.mut_nsub && !a.snapshot_pass) {
                const MutApply<DUMP> f{smem_u32(sw) + lane * 4u, lane, pos_lim, k, row, reg_site0, &a};
                const uint32_t kw = __reduce_max_sync(0xffffffffu, k);      // warp-uniform trip counts
#pragma unroll 1
                for (uint32_t base = 0, g = 0; base < kw; base += SNP_GROUP, g++) {
                    if (g) {
                        mc0 = snp_call(mctr, a.key, 3u * g);
                        mc1 = snp_call(mctr, a.key, 3u * g + 1u);
                    }
                    f.part(mc0.w, mc1.x, mc1.y, mc1.z, mc1.w, SNP_PART0, kw, mc0.y, mc0.z, base, mctr, a.key, g * 8u);
                    if (kw > base + SNP_PART0) {
                        const uint4 mc2 = snp_call(mctr, a.key, 3u * g + 2u);
                        f.part(mc2.z, mc2.w, 0u, 0u, 0u, SNP_GROUP - SNP_PART0, kw, mc2.x, mc2.y, base + SNP_PART0,
                               mctr, a.key, g * 8u + 4u);
                    }
                }
            }

            // ---- homologous recombination (population.rs:544-751, core) ----
            // The warp's events are numbered lane-major (owner lane, then draw order) and handled
            // 32 at a time, one event per lane, so the snapshot recomputation runs with all lanes
            // busy. Events of one owner stay in draw order: within a window they are applied in
            // rounds of increasing rank, and windows are processed in order (later wins, :745).
            if (a.hr_nsub) {
                uint32_t incl = kh;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += v;
                }
                const uint32_t pre = incl - kh;
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t n_other = a.n_rows - 1u;
#pragma unroll 1
                for (uint32_t win = 0; win < tot; win += 32) {
                    const uint32_t q = win + lane;
                    const bool active = q < tot;
                    // owner = first lane whose inclusive count exceeds q
                    uint32_t owner = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t v = __shfl_sync(0xffffffffu, incl, (owner + step - 1) & 31u);
                        if (v <= q && owner + step <= 31u) owner += step;
                    }
                    const uint32_t o_pre = __shfl_sync(0xffffffffu, pre, owner);
                    const uint32_t o_lim = __shfl_sync(0xffffffffu, pos_lim, owner);
                    uint4 g;
                    g.x = 0;
                    g.y = __shfl_sync(0xffffffffu, hg0.y, owner);
                    g.z = __shfl_sync(0xffffffffu, hg0.z, owner);
                    g.w = __shfl_sync(0xffffffffu, hg0.w, owner);
                    const uint32_t e = q - o_pre;                       // draw index within the owner's stream
                    const uint32_t rank = q - max(o_pre, win);          // order among the owner's events of this window
                    uint32_t pos = 0, d = 0, val = 0;
                    bool ok = false;
                    if (active) {
                        const uint4 octr = make_ctr(greg * 32u + owner, row, a.gen, STREAM_CORE_HR);
                        const uint32_t call = e >> 1;
                        if (call) {                                      // third and later events: rare
                            uint4 c = octr;
                            c.w += call;
                            g = philox4x32_10(c, a.key);
                        }
                        hr_event(g, call == 0, e & 1u, n_other, octr, a.key, e, pos, d);
                        d += (d >= row) ? 1u : 0u;                       // population.rs:616-619
                        ok = pos < o_lim;
                        if (ok) {
                            // donor's allele after gather and SNPs, before any HR (snapshot, :693-695)
                            const uint32_t widx = ((pos >> 4) << 5) + owner;
                            const uint32_t *dsrc = reinterpret_cast<const uint32_t *>(
                                a.old_state + (uint64_t)(a.snapshot_pass ? d : a.parents[d]) * a.row_stride +
                                (uint64_t)reg * REGION_BYTES);
                            val = (__ldg(dsrc + widx) >> ((pos & 15u) * 2u)) & 3u;
                            if (a.mut_nsub && !a.snapshot_pass) {
                                const uint32_t m = snp_probe(greg * 32u + owner, d, pos, a, tab_mut);
                                if (m) val = m;
                            }
                        }
                    }
                    const uint32_t rmax = __reduce_max_sync(0xffffffffu, ok ? rank : 0u);
#pragma unroll 1
                    for (uint32_t r = 0; r <= rmax; r++) {
                        if (ok && rank == r) {
                            const uint32_t widx = ((pos >> 4) << 5) + owner;
                            const uint32_t sh = (pos & 15u) * 2u;
                            uint32_t w = sw[widx];
                            w = (w & ~(3u << sh)) | (val << sh);
                            sw[widx] = w;
                            if (DUMP) {
                                const uint32_t slot = atomicAdd(&a.dump_counters[1], 1u);
                                if (slot < a.dump_cap) {
                                    a.d_hr_rec[slot] = row;
                                    a.d_hr_locus[slot] = (uint32_t)(reg_site0 + widx * 16u + (pos & 15u));
                                    a.d_hr_donor[slot] = d;
                                    a.d_hr_seq[slot] = e;
                                    a.d_hr_value[slot] = (uint8_t)(1u << val);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
            // ragged last region: clear whatever events wrote beyond the end of the alignment
            if (pos_lim_any < 256u) {
                const uint32_t lim = (uint32_t)(a.site_limit - reg_site0);
#pragma unroll 1
                for (uint32_t kk = 0; kk < WORDS_PER_LANE; kk++) {
                    const uint32_t site0 = (kk * 32u + lane) * 16u;
                    const uint32_t nv = site0 >= lim ? 0u : min(16u, lim - site0);
                    if (nv < 16u) sw[kk * 32u + lane] &= (1u << (2u * nv)) - 1u;
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
                PANSIM_ISSUE_LOAD();
            }
        }
        __syncwarp();
        row += d_row; reg += d_reg;
        if (reg >= a.n_regions) { reg -= a.n_regions; row++; }
    }
    if (lane == 0) bulk_wait<0>();
#undef PANSIM_ISSUE_LOAD
}

}  // namespace pansim
