// distance.cuh -- K7: pairwise core Hamming / accessory Jaccard counts
// (population.rs:787-837 with distances.rs:22-77) on the packed layouts.
//
// Core: the reference XORs one-hot bytes and halves the popcount
// (population.rs:817) = number of differing sites. On 2-bit codes a site
// differs iff either bit of the XOR is set:
//   d = a ^ b;  m = (d | d >> 1) & 0x5555...;  differing sites = popc(m)
// Two words share one POPC by parking the second mask in the odd bits.
// Row padding is zero in every row, so it never contributes.
#pragma once
#include "common.cuh"

namespace pansim {

__device__ __forceinline__ uint32_t diff_sites4(const uint4 a, const uint4 b)
{
    const uint32_t d0 = a.x ^ b.x, d1 = a.y ^ b.y, d2 = a.z ^ b.z, d3 = a.w ^ b.w;
    const uint32_t m0 = (d0 | (d0 >> 1)) & 0x55555555u;
    const uint32_t m1 = (d1 | (d1 << 1)) & 0xAAAAAAAAu;
    const uint32_t m2 = (d2 | (d2 >> 1)) & 0x55555555u;
    const uint32_t m3 = (d3 | (d3 << 1)) & 0xAAAAAAAAu;
    return __popc(m0 | m1) + __popc(m2 | m3);
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// v1: one CTA per (pair, column chunk). Chunks are the slow grid dimension so
// that concurrently running CTAs touch the same column slab of all rows
// (N * chunk bytes is sized to stay L2-resident); partial counts are combined
// with integer atomics (order-independent, exact).
constexpr int PAIR_THREADS = 256;

__global__ void __launch_bounds__(PAIR_THREADS) pair_core_kernel(const uint8_t *state, uint64_t row_stride,
                                                                 uint32_t chunk_vec4, uint32_t row_vec4,
                                                                 const uint32_t *range1, const uint32_t *range2,
                                                                 uint32_t n_pairs, uint32_t *core_diff)
{
    __shared__ uint32_t wsum[PAIR_THREADS / 32];
    const uint32_t chunk = blockIdx.y;
    const uint32_t v_lo = chunk * chunk_vec4;
    const uint32_t v_hi = min(row_vec4, v_lo + chunk_vec4);
    for (uint32_t p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        const uint4 *ra = reinterpret_cast<const uint4 *>(state + (uint64_t)range1[p] * row_stride);
        const uint4 *rb = reinterpret_cast<const uint4 *>(state + (uint64_t)range2[p] * row_stride);
        uint32_t acc = 0;
        uint32_t v = v_lo + threadIdx.x;
        for (; v + 3 * PAIR_THREADS < v_hi; v += 4 * PAIR_THREADS) {
            const uint4 a0 = ld_stream(ra + v), b0 = ld_stream(rb + v);
            const uint4 a1 = ld_stream(ra + v + PAIR_THREADS), b1 = ld_stream(rb + v + PAIR_THREADS);
            const uint4 a2 = ld_stream(ra + v + 2 * PAIR_THREADS), b2 = ld_stream(rb + v + 2 * PAIR_THREADS);
            const uint4 a3 = ld_stream(ra + v + 3 * PAIR_THREADS), b3 = ld_stream(rb + v + 3 * PAIR_THREADS);
            acc += diff_sites4(a0, b0) + diff_sites4(a1, b1) + diff_sites4(a2, b2) + diff_sites4(a3, b3);
        }
        for (; v < v_hi; v += PAIR_THREADS) acc += diff_sites4(ld_stream(ra + v), ld_stream(rb + v));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int q = 0; q < PAIR_THREADS / 32; q++) t += wsum[q];
            if (gridDim.y == 1) core_diff[p] = t; else atomicAdd(&core_diff[p], t);
        }
        __syncthreads();
    }
}

// v1.5: row-stationary groups. Pairs are sorted by their first row on the host (plan
// cached while the pair list is unchanged, as in the reference where the pairs are
// fixed for the whole run, main.rs:413-427). A CTA handles one group = one row i and up
// to PAIR_GROUP partner rows j: row i's words are loaded once per thread and compared
// against every partner, which almost halves the L2 traffic of the pair-per-CTA kernel.
constexpr int PAIR_GROUP = 8;

struct PairGroup {
    uint32_t row_i;
    uint32_t first;      // index into the sorted partner / original-index arrays
    uint32_t count;      // 1..PAIR_GROUP
};

__global__ void __launch_bounds__(PAIR_THREADS) pair_core_grouped_kernel(
    const uint8_t *state, uint64_t row_stride, uint32_t chunk_vec4, uint32_t row_vec4, const PairGroup *groups,
    uint32_t n_groups, const uint32_t *partner, const uint32_t *orig_index, uint32_t *core_diff)
{
    __shared__ uint32_t wsum[PAIR_THREADS / 32][PAIR_GROUP];
    const uint32_t chunk = blockIdx.y;
    const uint32_t v_lo = chunk * chunk_vec4;
    const uint32_t v_hi = min(row_vec4, v_lo + chunk_vec4);
    for (uint32_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const PairGroup grp = groups[g];
        const uint4 *ra = reinterpret_cast<const uint4 *>(state + (uint64_t)grp.row_i * row_stride);
        const uint4 *rb[PAIR_GROUP];
#pragma unroll
        for (int q = 0; q < PAIR_GROUP; q++) {
            const uint32_t j = partner[grp.first + (q < (int)grp.count ? q : 0)];
            rb[q] = reinterpret_cast<const uint4 *>(state + (uint64_t)j * row_stride);
        }
        uint32_t acc[PAIR_GROUP];
#pragma unroll
        for (int q = 0; q < PAIR_GROUP; q++) acc[q] = 0;
        for (uint32_t v = v_lo + threadIdx.x; v < v_hi; v += PAIR_THREADS) {
            const uint4 a = ld_stream(ra + v);
            uint4 b[PAIR_GROUP];
#pragma unroll
            for (int q = 0; q < PAIR_GROUP; q++) b[q] = ld_stream(rb[q] + v);
#pragma unroll
            for (int q = 0; q < PAIR_GROUP; q++) acc[q] += diff_sites4(a, b[q]);
        }
#pragma unroll
        for (int q = 0; q < PAIR_GROUP; q++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int q = 0; q < PAIR_GROUP; q++) wsum[threadIdx.x >> 5][q] = acc[q];
        }
        __syncthreads();
        if (threadIdx.x < grp.count) {
            uint32_t t = 0;
            for (int w = 0; w < PAIR_THREADS / 32; w++) t += wsum[w][threadIdx.x];
            const uint32_t out = orig_index[grp.first + threadIdx.x];
            atomicAdd(&core_diff[out], t);
        }
        __syncthreads();
    }
}

// Exact all-pairs mode: the pair list of a row block [row_begin, row_end) -- every (i, j) with
// i < j < N, ordered by i then j -- and its row-stationary groups are generated on the device, so
// nothing but a small offset table crosses the bus. off[r] = number of pairs of the rows before
// row_begin + r, goff[r] = number of groups before it (both with nr + 1 entries).
__device__ __forceinline__ uint32_t upper_row(const uint32_t *off, uint32_t nr, uint32_t k)
{
    uint32_t lo = 0, hi = nr;            // last r with off[r] <= k
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= k) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void rows_pairs_kernel(const uint32_t *off, uint32_t nr, uint32_t row_begin, uint32_t n_pairs,
                                  uint32_t *range1, uint32_t *range2, uint32_t *partner, uint32_t *orig_index)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const uint32_t r = upper_row(off, nr, k);
    const uint32_t i = row_begin + r, j = i + 1u + (k - off[r]);
    range1[k] = i; range2[k] = j; partner[k] = j; orig_index[k] = k;
}

// groups of the pairs (i, j) of the row block with j in [jb0, jb1): goff[r] = groups of this
// column block before row row_begin + r
__global__ void rows_groups_kernel(const uint32_t *off, const uint32_t *goff, uint32_t nr, uint32_t row_begin,
                                   uint32_t jb0, uint32_t jb1, uint32_t n_groups, PairGroup *groups)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t r = upper_row(goff, nr, g);
    const uint32_t i = row_begin + r;
    const uint32_t j_lo = max(i + 1u, jb0);                                      // first partner of row i in this block
    const uint32_t j = j_lo + (uint32_t)PAIR_GROUP * (g - goff[r]);
    groups[g] = PairGroup{i, off[r] + (j - (i + 1u)), min((uint32_t)PAIR_GROUP, jb1 - j)};
}

// v2: shared-memory row tiles. The host plan (cached) buckets the pairs by the pair of
// 32-row blocks their endpoints fall in. A CTA takes one batch (<= 256 pairs of one
// block pair), stages the <= 64 rows it needs, 512 B per row per stage, with cp.async
// (LDGSTS, 16 B per thread, 3-stage ring; 64 separate 512 B TMA bulk copies per stage were
// measured slower), and walks the whole column range. Each warp owns up to 16 consecutive pairs of the batch (sorted by
// first endpoint so the first row is re-read only when it changes) and keeps their counts
// in registers until the end: every staged byte is reused by all pairs of the batch that
// touch its row, which cuts the L2 traffic of the streaming kernels ~3-6x.
constexpr int TILE_ROWS = 64;          // 2 blocks of 32 rows
constexpr int TILE_BYTES = 512;        // per row per stage
constexpr int TILE_STAGES = 3;
constexpr int TILE_WARPS = 16;
constexpr int TILE_THREADS = TILE_WARPS * 32;
constexpr int TILE_PPW = 16;           // pairs per warp
constexpr int TILE_BATCH = TILE_WARPS * TILE_PPW;   // 256

struct TileBatch {
    uint32_t block_a, block_b;         // 32-row blocks (block_a <= block_b)
    uint32_t first, count;             // range in the tile pair arrays
    uint32_t col_begin, col_end;       // column range in units of TILE_BYTES
};

static inline size_t tile_smem_bytes()
{
    return (size_t)TILE_STAGES * TILE_ROWS * TILE_BYTES + TILE_BATCH * sizeof(uint32_t);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(TILE_THREADS, 2) pair_core_tile_kernel(
    const uint8_t *state, uint64_t row_stride, uint32_t n_rows, const TileBatch *batches,
    const uint16_t *tile_slots, const uint32_t *tile_orig, uint32_t *core_diff)
{
    extern __shared__ __align__(128) uint8_t tsm[];
    uint32_t *meta = reinterpret_cast<uint32_t *>(tsm + (size_t)TILE_STAGES * TILE_ROWS * TILE_BYTES);
    const TileBatch bt = batches[blockIdx.x];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (uint32_t i = threadIdx.x; i < TILE_BATCH; i += TILE_THREADS)
        meta[i] = i < bt.count ? tile_slots[bt.first + i] : 0xFFFFu;

    // staging: a stage is TILE_ROWS x 512 B = 2048 pieces of 16 B, 4 per thread. Piece p of a
    // chunk belongs to row slot p/32 (slots 0..31 -> block_a, 32..63 -> block_b), byte (p%32)*16.
    constexpr int PIECES = TILE_ROWS * TILE_BYTES / 16 / TILE_THREADS;      // 4
    const uint8_t *src[PIECES];
    uint32_t dst_off[PIECES];
#pragma unroll
    for (int k = 0; k < PIECES; k++) {
        const uint32_t p = threadIdx.x + k * TILE_THREADS;
        const uint32_t slot = p >> 5;
        const uint32_t row = slot < 32 ? bt.block_a * 32 + slot : bt.block_b * 32 + (slot - 32);
        const bool ok = row < n_rows && (slot < 32 || bt.block_b != bt.block_a);
        src[k] = ok ? state + (uint64_t)row * row_stride + (uint64_t)bt.col_begin * TILE_BYTES + (p & 31u) * 16u : nullptr;
        dst_off[k] = slot * TILE_BYTES + (p & 31u) * 16u;
    }
    const uint32_t n_chunks = bt.col_end - bt.col_begin;
    auto issue = [&](uint32_t c) {
        uint8_t *stage = tsm + (size_t)(c % TILE_STAGES) * TILE_ROWS * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < PIECES; k++)
            if (src[k]) cp_async16(stage + dst_off[k], src[k] + (uint64_t)c * TILE_BYTES);
        cp_async_commit();
    };
    issue(0);
    if (n_chunks > 1) issue(1); else cp_async_commit();

    uint32_t acc[TILE_PPW];
#pragma unroll
    for (int q = 0; q < TILE_PPW; q++) acc[q] = 0;
    __syncthreads();                                   // meta[] is complete
    // per-warp pair metadata hoisted out of the column loop: byte offsets of both rows inside
    // a stage (packed a | b << 16), number of real pairs, and which pairs start a new first row
    uint32_t offs[TILE_PPW];
    uint32_t nq = 0, reload = 0;
    {
        uint32_t prev = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < TILE_PPW; q++) {
            const uint32_t m = meta[warp * TILE_PPW + q];
            const uint32_t sa = m & 0xFFu, sb = (m >> 8) & 0xFFu;
            offs[q] = (sa * TILE_BYTES) | ((sb * TILE_BYTES) << 16);
            if (m != 0xFFFFu) {
                nq = q + 1;
                if (sa != prev) reload |= 1u << q;
                prev = sa;
            }
        }
    }

    for (uint32_t c = 0; c < n_chunks; c++) {
        if (c + 2 < n_chunks) issue(c + 2); else cp_async_commit();     // keep the group count uniform
        cp_async_wait<2>();                                              // chunk c has landed (this thread)
        __syncthreads();                                                 // ... and everybody else's pieces
        const uint8_t *base = tsm + (size_t)(c % TILE_STAGES) * TILE_ROWS * TILE_BYTES + lane * 16;
        uint4 a = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < TILE_PPW; q++) {
            if (q < (int)nq) {                       // warp-uniform
                if ((reload >> q) & 1u) a = *reinterpret_cast<const uint4 *>(base + (offs[q] & 0xFFFFu));
                const uint4 b = *reinterpret_cast<const uint4 *>(base + (offs[q] >> 16));
                acc[q] += diff_sites4(a, b);
            }
        }
        __syncthreads();                              // stage c%3 may be overwritten by chunk c+3 next iteration
    }
#pragma unroll
    for (int q = 0; q < TILE_PPW; q++) {
        uint32_t v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const uint32_t idx = warp * TILE_PPW + q;
        if (lane == 0 && idx < bt.count) atomicAdd(&core_diff[tile_orig[bt.first + idx]], v);
    }
}

// accessory: one warp per pair; intersection and union popcounts (distances.rs:62-68)
__global__ void __launch_bounds__(256) pair_acc_kernel(const uint32_t *acc, uint32_t stride_words,
                                                       uint32_t n_words, const uint32_t *range1,
                                                       const uint32_t *range2, uint32_t n_pairs,
                                                       uint32_t *inter, uint32_t *uni)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5;
    for (uint32_t p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n_pairs; p += gridDim.x * wpb) {
        const uint32_t *ra = acc + (uint64_t)range1[p] * stride_words;
        const uint32_t *rb = acc + (uint64_t)range2[p] * stride_words;
        uint32_t in = 0, un = 0;
        for (uint32_t w = lane; w < n_words; w += 32) {
            const uint32_t x = ra[w], y = rb[w];
            in += __popc(x & y);
            un += __popc(x | y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            in += __shfl_xor_sync(0xffffffffu, in, o);
            un += __shfl_xor_sync(0xffffffffu, un, o);
        }
        if (lane == 0) {
            if (inter) inter[p] = in;
            if (uni) uni[p] = un;
        }
    }
}

// --print_dist statistics (main.rs:502-519): standard_deviation (population.rs:87-94) of the two
// distance vectors of one pass, on the device, so 32 bytes leave the GPU per generation instead of
// the three count vectors. The reference sums with `iter().sum::<f64>()`, i.e. strictly left to
// right, and the result is printed with 17 significant digits, so the summation ORDER is part of
// the result: each of the four sums (two means, then two sums of squared deviations) is one chain
// of dependent f64 additions made by a single thread in index order. Everything that is not the
// chain -- the divisions of population.rs:822 / :828-830, x - mean, the square -- is done by a
// producer warp one 1024-element chunk ahead, through shared memory. No fused multiply-add can
// form: the products are rounded into shared memory before they are added (Rust has no contraction).
// out[4] = { avg_core, std_core, avg_acc, std_acc } (the column order of _per_gen.tsv, main.rs:546).
constexpr int STATS_CHUNK = 1024;
constexpr int STATS_THREADS = 128;      // warp 0: core producer, 1: core chain, 2: accessory producer, 3: accessory chain

struct PairStatsArgs {
    const uint32_t *core_diff, *inter, *uni;
    uint32_t n_pairs;
    double core_size, core_genes;
    double *out;
};

__global__ void __launch_bounds__(STATS_THREADS) pair_stats_kernel(const PairStatsArgs a)
{
    __shared__ double buf[2][2][STATS_CHUNK];      // [vector][stage][element]
    __shared__ double mean_s[2];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t vec = warp >> 1;                // 0 = core distances, 1 = accessory distances
    const bool producer = (warp & 1u) == 0u;
    const uint32_t n = a.n_pairs;
    const uint32_t n_chunks = (n + STATS_CHUNK - 1) / STATS_CHUNK;
    const double dn = (double)n;

    auto produce = [&](uint32_t c, int pass, double mean) {
        double *dst = buf[vec][c & 1u];
        const uint32_t k0 = c * STATS_CHUNK, k1 = min(n, k0 + STATS_CHUNK);
        for (uint32_t k = k0 + lane; k < k1; k += 32) {
            double d;
            if (vec == 0) {
                d = __ddiv_rn((double)a.core_diff[k], a.core_size);                                          // population.rs:822
            } else {
                const double num = __dadd_rn((double)a.inter[k], a.core_genes), den = __dadd_rn((double)a.uni[k], a.core_genes);
                d = __dsub_rn(1.0, __ddiv_rn(num, den));                                                     // :828-830
            }
            if (pass == 1) {
                const double diff = __dsub_rn(d, mean);                                                      // :90
                d = __dmul_rn(diff, diff);
            }
            dst[k - k0] = d;
        }
    };

    double sum = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        const double mean = pass ? mean_s[vec] : 0.0;
        if (producer && n_chunks) produce(0, pass, mean);
        __syncthreads();
        sum = 0.0;
        for (uint32_t c = 0; c < n_chunks; c++) {
            if (producer) {
                if (c + 1 < n_chunks) produce(c + 1, pass, mean);
            } else if (lane == 0) {
                const double *src = buf[vec][c & 1u];
                const uint32_t len = min((uint32_t)STATS_CHUNK, n - c * STATS_CHUNK);
                // the chain: one dependent DADD per element (8.2 cycles each); the shared-memory loads of the
                // next eight elements are issued before the current eight are added, so only the adds remain
                uint32_t i = 0;
                if (len >= 8) {
                    double v[8], w[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = src[j];
                    for (i = 8; i + 8 <= len; i += 8) {
#pragma unroll
                        for (int j = 0; j < 8; j++) w[j] = src[i + j];
#pragma unroll
                        for (int j = 0; j < 8; j++) sum = __dadd_rn(sum, v[j]);
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = w[j];
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) sum = __dadd_rn(sum, v[j]);
                }
                for (; i < len; i++) sum = __dadd_rn(sum, src[i]);
            }
            __syncthreads();
        }
        if (!producer && lane == 0) {
            if (pass == 0) {
                const double mean0 = __ddiv_rn(sum, dn);                                                     // :83-85
                mean_s[vec] = mean0;
                a.out[vec * 2] = mean0;
            } else {
                a.out[vec * 2 + 1] = __dsqrt_rn(__ddiv_rn(sum, dn));                                         // :92-93
            }
        }
        __syncthreads();
    }
}

}  // namespace pansim
