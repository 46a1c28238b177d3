// acc_step.cuh -- K5: accessory-genome generation step on the bit-packed
// presence/absence matrix (1 bit per gene, gene g in word g/32 bit g%32).
//
//   gather-by-parent       population.rs:450-465
//   gene gain/loss flips   population.rs:488-510   (multi-hit = parity)
//   HGT (gain only)        population.rs:544-751, accessory branch
//
// Exact per-cell forms of the reference's per-row event loops:
//   flips: a gene of compartment c is hit Poisson(rate_c) times per row, the
//          bit flips iff the count is odd: p_c = (1 - exp(-2 rate_c)) / 2.
//   HGT:   donor d emits Poisson(lambda_c) events, recipient uniform on the
//          other N-1 rows, gene uniform on the K_dc genes of compartment c the
//          donor carries (post-mutation snapshot), value always 1
//          (population.rs:632-680). Hence cell (r,g) with bit 0 gains with
//          probability 1 - exp(-lambda_c/(N-1) * S_g),
//          S_g = sum over donors carrying g of 1/K_dc; cells are independent
//          given the snapshot.
#pragma once
#include "common.cuh"

namespace pansim {

struct AccArgs {
    const uint32_t *old_state;
    uint32_t *new_state;
    const uint32_t *parents;
    uint32_t n_rows, n_genes, stride_words;
    uint2 key;
    PhiloxKeys rk;                // round keys of `key` (accessory streams: Philox4x32-7 like the core streams, common.cuh)
    uint32_t gen;
    const uint32_t *gen_dev;      // nullptr, or a device word added to gen (replayed CUDA graphs)
    // gene compartments (main.rs:341-367): scalars, not arrays, so that the kernels never index
    // the parameter block dynamically (that would force a per-thread local copy of it)
    uint32_t lo0, hi0, lo1, hi1;   // genes with weight 1.0 in compartment 0 / 1 (empty: lo == hi)
    uint32_t flip_thr0, flip_thr1; // round(p_c * 2^32)
    double hgt_scale0, hgt_scale1; // lambda_c / (N-1); 0 = off
    double *rowInvK;               // [2][n_rows] 1 / (genes of compartment c present, post-mutation); 0 if none
    uint32_t *gain_planes;         // [gene words][32]: plane j = binary digit j (MSB first) of the 32 genes' gain thresholds
    uint32_t *dump_flip;           // optional [n_rows * stride_words]
    uint32_t *dump_gain;
};

__device__ __forceinline__ uint32_t comp_mask_for_word(uint32_t w, uint32_t lo, uint32_t hi)
{
    // bits of word w (genes 32w..32w+31) that fall in [lo,hi)
    const uint32_t g0 = w * 32u;
    if (hi <= g0 || lo >= g0 + 32u) return 0u;
    const uint32_t a = lo > g0 ? lo - g0 : 0u;
    const uint32_t b = hi < g0 + 32u ? hi - g0 : 32u;
    const uint32_t upto_b = b >= 32u ? 0xFFFFFFFFu : ((1u << b) - 1u);
    const uint32_t upto_a = (1u << a) - 1u;
    return upto_b & ~upto_a;
}

// 32 Bernoulli bits at once, bit-sliced and lazy: gene b gets result bit 1 iff U_b < thr_b, where
// U_b is a 32-bit uniform whose binary digits are bit b of successive Philox words (most
// significant digit first) and thr_b is given as 32 bit-planes (plane j = digit j of every gene's
// threshold). A gene is decided at the first digit where U_b and thr_b differ, so on average
// log2(32) + 1.3 ~ 6-7 random words settle all 32 genes (instead of one 32-bit uniform per gene);
// the comparison is exact (all 32 digits are used if needed; equality means "not less").
struct PlanesConst {            // per-compartment constant thresholds (gene gain/loss flips)
    uint32_t thr0, thr1, m0, m1;
    __device__ __forceinline__ uint32_t operator()(uint32_t j) const
    {
        return (((thr0 >> (31u - j)) & 1u) ? m0 : 0u) | (((thr1 >> (31u - j)) & 1u) ? m1 : 0u);
    }
};
struct PlanesTable {            // per-gene thresholds (HGT), bit-planes precomputed per gene word
    const uint32_t *p;
    uint4 a, b;                 // planes 0..7 prefetched: they settle almost every gene
    __device__ __forceinline__ explicit PlanesTable(const uint32_t *ptr)
        : p(ptr), a(*reinterpret_cast<const uint4 *>(ptr)), b(*reinterpret_cast<const uint4 *>(ptr + 4)) {}
    __device__ __forceinline__ uint32_t operator()(uint32_t j) const
    {
        switch (j) {
        case 0: return a.x; case 1: return a.y; case 2: return a.z; case 3: return a.w;
        case 4: return b.x; case 5: return b.y; case 6: return b.z; case 7: return b.w;
        default: return p[j];
        }
    }
};

template <typename Planes>
__device__ __forceinline__ uint32_t bernoulli_word(const PhiloxKeys &key, uint32_t gen, uint32_t stream, uint32_t row,
                                                   uint32_t w, uint32_t active, const Planes planes)
{
    uint32_t und = active, res = 0;
#pragma unroll 1
    for (uint32_t q = 0; q < 8 && und; q++) {
        uint4 ctr = make_ctr(w, row, gen, stream);
        ctr.w |= q;
        const uint4 r4 = philox_core(ctr, key);
        const uint32_t r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            const uint32_t t = planes(4u * q + c);
            const uint32_t lt = ~r[c] & t, gt = r[c] & ~t;
            res |= und & lt;
            und &= ~(lt | gt);
        }
    }
    return res;
}

// gather + flips + per-row compartment popcounts. One warp per row.
template <bool DUMP>
__global__ void __launch_bounds__(256) acc_gather_flip_kernel(const AccArgs a)
{
    pdl_launch_dependents();       // the gain-threshold kernel may take its SM slots now (it waits for this grid)
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (row >= a.n_rows) return;
    const uint32_t *src = a.old_state + (uint64_t)a.parents[row] * a.stride_words;
    uint32_t *dst = a.new_state + (uint64_t)row * a.stride_words;
    uint32_t k0 = 0, k1 = 0;
    for (uint32_t w = lane; w < a.stride_words; w += 32) {
        const uint32_t m0 = comp_mask_for_word(w, a.lo0, a.hi0);
        const uint32_t m1 = comp_mask_for_word(w, a.lo1, a.hi1);
        uint32_t active = 0;
        if (a.flip_thr0) active |= m0;
        if (a.flip_thr1) active |= m1;
        const uint32_t flips = bernoulli_word(a.rk, a.gen + (a.gen_dev ? __ldg(a.gen_dev) : 0u), STREAM_ACC_FLIP, row, w, active,
                                              PlanesConst{a.flip_thr0, a.flip_thr1, m0, m1});
        const uint32_t v = src[w] ^ flips;
        dst[w] = v;
        if (DUMP) a.dump_flip[(uint64_t)row * a.stride_words + w] = flips;
        k0 += __popc(v & m0);
        k1 += __popc(v & m1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k0 += __shfl_xor_sync(0xffffffffu, k0, o);
        k1 += __shfl_xor_sync(0xffffffffu, k1, o);
    }
    if (lane == 0) {
        a.rowInvK[row] = k0 ? 1.0 / (double)k0 : 0.0;
        a.rowInvK[a.n_rows + row] = k1 ? 1.0 / (double)k1 : 0.0;
    }
}

// per-gene HGT gain threshold from the post-mutation snapshot: S_g = sum over donors d carrying g of
// 1/K_dc, a bit-matrix x vector product. One CTA (32 warps) per 32-gene word. A warp takes 32 donor
// rows at a time (one row's word per lane), transposes the 32x32 bit tile with five shuffle stages
// so that each lane holds ONE gene's mask over the 32 donors, and adds 1/K only for the set bits
// (ascending donor order). Warp q handles row groups q, q+32, ...; the 32 partial sums are added in
// warp order: deterministic.
constexpr int GAIN_WARPS = 32;     // 8 warps per CTA: 14 us instead of 8 us alone, 23 instead of 17 in the pipeline

// in: lane l holds the word of row (31 - l); out: lane l holds, for gene (31 - l), bit r = row r
__device__ __forceinline__ uint32_t warp_transpose32_rev(uint32_t x, uint32_t lane)
{
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (uint32_t j = 16; j != 0; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        if ((lane & j) == 0) x ^= (x ^ (y >> j)) & m;
        else x ^= ((y ^ (x >> j)) & m) << j;
        m ^= m << (j >> 1);
    }
    return x;
}

__global__ void __launch_bounds__(GAIN_WARPS * 32) acc_gain_threshold_kernel(const AccArgs a)
{
    pdl_launch_dependents();
    pdl_wait();                    // rows and 1/K values of acc_gather_flip_kernel
    __shared__ double part[GAIN_WARPS][32];
    const uint32_t w = blockIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = w * 32u + (31u - lane);            // the gene this lane accumulates
    int c = -1;
    if (g >= a.lo0 && g < a.hi0 && a.hgt_scale0 > 0.0) c = 0;
    if (g >= a.lo1 && g < a.hi1 && a.hgt_scale1 > 0.0) c = 1;
    const double *invK = a.rowInvK + (c == 1 ? a.n_rows : 0u);
    double s = 0.0;
    for (uint32_t d0 = warp * 32u; d0 < a.n_rows; d0 += GAIN_WARPS * 32u) {
        const uint32_t d = d0 + (31u - lane);
        const uint32_t word = d < a.n_rows ? a.new_state[(uint64_t)d * a.stride_words + w] : 0u;
        uint32_t mask = warp_transpose32_rev(word, lane);  // bit r = donor d0 + r carries gene g
        if (c < 0) mask = 0;
        // ascending donor order, every donor of the tile: s += 1/K_d * (bit ? 1.0 : +0.0) as one FMA per donor (the product
        // is exact, so this is the plain addition for a set bit and leaves s unchanged otherwise; s starts at +0.0 and only
        // grows). The accessory genome of the default parameters is dense, so walking the set bits alone costs more.
        if (d0 + 32u <= a.n_rows) {
#pragma unroll
            for (uint32_t r = 0; r < 32; r++)                                                    // invK: same address in every lane
                s = fma(__ldg(invK + d0 + r), __hiloint2double((mask >> r) & 1u ? 0x3FF00000 : 0, 0), s);
        } else {
            while (mask) {
                const uint32_t r = __ffs(mask) - 1;
                mask &= mask - 1;
                s += invK[d0 + r];
            }
        }
    }
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        double S = 0.0;
        for (int q = 0; q < GAIN_WARPS; q++) S += part[q][lane];
        uint32_t thr = 0;
        if (g < a.n_genes && c >= 0 && S > 0.0) {
            const double p = -expm1(-(c == 1 ? a.hgt_scale1 : a.hgt_scale0) * S);
            const double t = rint(p * 4294967296.0);
            thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        }
        // transpose the 32 thresholds into bit-planes (lane l holds gene 31 - l, hence the bit
        // reversal): lane j keeps plane j
        uint32_t plane = 0;
#pragma unroll
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t bal = __brev(__ballot_sync(0xffffffffu, (thr >> (31u - j)) & 1u));
            if (lane == j) plane = bal;
        }
        a.gain_planes[(uint64_t)w * 32u + lane] = plane;
    }
}

// HGT apply: absent genes gain with their per-gene probability. Thread per word.
template <bool DUMP>
__global__ void __launch_bounds__(256) acc_hgt_apply_kernel(const AccArgs a)
{
    pdl_wait();                    // gain thresholds of acc_gain_threshold_kernel
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)a.n_rows * a.stride_words;
    if (idx >= total) return;
    const uint32_t row = (uint32_t)(idx / a.stride_words), w = (uint32_t)(idx % a.stride_words);
    uint32_t valid = 0;
    if (a.hgt_scale0 > 0.0) valid |= comp_mask_for_word(w, a.lo0, a.hi0);
    if (a.hgt_scale1 > 0.0) valid |= comp_mask_for_word(w, a.lo1, a.hi1);
    const uint32_t cur = a.new_state[idx];
    const uint32_t active = valid & ~cur;      // a hit on a present gene writes 1 over 1
    const uint32_t gain = bernoulli_word(a.rk, a.gen + (a.gen_dev ? __ldg(a.gen_dev) : 0u), STREAM_ACC_HGT, row, w, active,
                                         PlanesTable(a.gain_planes + (uint64_t)w * 32u));
    if (gain) a.new_state[idx] = cur | gain;
    if (DUMP) a.dump_gain[idx] = gain;
}

// plain gather (replay mode / next_generation alone)
__global__ void __launch_bounds__(256) acc_gather_kernel(const uint32_t *old_state, uint32_t *new_state,
                                                         const uint32_t *parents, uint32_t n_rows,
                                                         uint32_t stride_words)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * stride_words) return;
    const uint32_t row = (uint32_t)(idx / stride_words), w = (uint32_t)(idx % stride_words);
    new_state[idx] = old_state[(uint64_t)parents[row] * stride_words + w];
}

// replay: flips (XOR, order-free) and HGT sets (OR, order-free)
__global__ void acc_apply_flips_kernel(uint32_t *state, const uint32_t *row, const uint32_t *gene, size_t n,
                                       uint32_t stride_words)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    atomicXor(&state[(uint64_t)row[k] * stride_words + (gene[k] >> 5)], 1u << (gene[k] & 31u));
}

__global__ void acc_apply_sets_kernel(uint32_t *state, const uint32_t *row, const uint32_t *gene, size_t n,
                                      uint32_t stride_words)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    atomicOr(&state[(uint64_t)row[k] * stride_words + (gene[k] >> 5)], 1u << (gene[k] & 31u));
}

// K8: gene counts (population.rs:847-855). Same CTA shape as the threshold kernel.
__global__ void __launch_bounds__(256) acc_gene_counts_kernel(const uint32_t *state, uint32_t n_rows,
                                                              uint32_t n_genes, uint32_t stride_words,
                                                              uint32_t *counts)
{
    __shared__ uint32_t part[8][32];
    const uint32_t w = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t s = 0;
    for (uint32_t d = warp; d < n_rows; d += 8) s += (state[(uint64_t)d * stride_words + w] >> lane) & 1u;
    part[warp][lane] = s;
    __syncthreads();
    const uint32_t g = w * 32u + lane;
    if (warp == 0 && g < n_genes) {
        uint32_t t = 0;
        for (int q = 0; q < 8; q++) t += part[q][lane];
        counts[g] = t;
    }
}

}  // namespace pansim
