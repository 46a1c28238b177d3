// select.cuh -- K1/K2/K3: fitness, competition term and parent resampling.
//   K1 fitness           population.rs:282-322  (f64 sum in column order -> bit-exact)
//   K2 average_distance  population.rs:753-784 + get_distance :114-151 (bit-exact)
//   K3 weights + draw    population.rs:325-447
#pragma once
#include <cfloat>
#include <cuda.h>          // CUtensorMap (acc_inter_umma_kernel)
#include "common.cuh"

namespace pansim {

// ---------------------------------------------------------------------------
// K1: log-fitness l_i = sum_j ln(1 + s_j * x_ij) and gene counts.
// x_ij in {0,1}, so a term is either ln(1) = +0.0 (adding it never changes the
// running sum) or lw_j = ln(1 + s_j), precomputed on the host with the same
// libm expression the reference evaluates. One thread per row adds the present
// genes in increasing column order, reproducing the reference's sequential f64
// sum bit for bit (population.rs:303-317). A present gene with lw_j = -inf
// (s_j = -1) sets l_i := 0.0 (population.rs:312-318).
// ---------------------------------------------------------------------------
constexpr int FIT_WARPS = 4;                      // rows per CTA of the blocked / count-only kernels
constexpr int FIT_ROWS = 2;                       // rows per CTA of the sequential kernel (two warps per row)
constexpr int FIT_CHUNK_WORDS = 32;               // 1024 genes per compaction round
constexpr int FIT_BATCH = 8;                      // chain elements loaded ahead of the additions

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// every selection coefficient is 0 (neutral run: lw_j = ln(1) = +0.0 for all j): the sum is +0.0
// whatever the genome; only the row popcounts are needed
__global__ void __launch_bounds__(FIT_WARPS * 32) fitness_count_kernel(const uint32_t *acc, uint32_t n_rows,
                                                                       uint32_t n_genes, uint32_t stride_words,
                                                                       double *logfit, int32_t *num_genes)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * FIT_WARPS + warp;
    if (row >= n_rows) return;
    const uint32_t *r = acc + (uint64_t)row * stride_words;
    const uint32_t n_words = (n_genes + 31u) / 32u;
    int32_t cnt = 0;
    for (uint32_t w = lane; w < n_words; w += 32) cnt += __popc(r[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) { logfit[row] = 0.0; num_genes[row] = cnt; }
}

// Terms of absent genes are ln(1 + s*0) = +0.0: adding them never changes the running sum, so
// only present genes are added -- in increasing column order, by a single sequential f64 chain
// per row (bit-exact vs population.rs:303-317). Two warps per row: the PRODUCER compacts the lw
// values of the present genes of each 1024-gene chunk into one of two shared-memory lists (padded
// with +0.0 to whole batches: the running sum starts at +0.0 and can never become -0.0, so
// x + (+0.0) = x); the CONSUMER runs the chain, its loads one batch ahead of the additions, so it
// advances at the latency of a dependent DADD (8.2 cycles on B200) while the producer prepares
// the next chunk. Full/empty hand-over through named barriers.
__global__ void __launch_bounds__(FIT_ROWS * 64) fitness_kernel(const uint32_t *acc, uint32_t n_rows,
                                                                uint32_t n_genes, uint32_t stride_words,
                                                                const double *lw, double *logfit,
                                                                int32_t *num_genes)
{
    constexpr int LIST = FIT_CHUNK_WORDS * 32 + 2 * FIT_BATCH;
    __shared__ double list[FIT_ROWS][2][LIST];
    __shared__ uint32_t tot_s[FIT_ROWS][2];
    __shared__ int32_t fin_s[FIT_ROWS][2];           // gene count, -inf seen
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r_in_cta = warp >> 1;
    const bool producer = (warp & 1u) != 0u;
    const uint32_t row = blockIdx.x * FIT_ROWS + r_in_cta;
    if (row >= n_rows) return;                        // both warps of the row leave together
    const uint32_t n_words = (n_genes + 31u) / 32u;
    const uint32_t n_chunks = (n_words + FIT_CHUNK_WORDS - 1) / FIT_CHUNK_WORDS;
    const uint32_t bar_full = 1u + r_in_cta * 5u, bar_empty = 3u + r_in_cta * 5u, bar_done = 5u + r_in_cta * 5u;
    if (producer) {
        const uint32_t *r = acc + (uint64_t)row * stride_words;
        bool neg_inf = false;
        int32_t cnt = 0;
        for (uint32_t ch = 0; ch < n_chunks; ch++) {
            const uint32_t b = ch & 1u;
            if (ch >= 2) named_bar_sync(bar_empty + b, 64);          // the consumer is done with this list
            double *mine = list[r_in_cta][b];
            const uint32_t w = ch * FIT_CHUNK_WORDS + lane;
            uint32_t bits = (w < n_words) ? r[w] : 0u;
            const uint32_t c = __popc(bits);
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            uint32_t off = incl - c;
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            cnt += (int32_t)tot;
            if (lane < 2 * FIT_BATCH) mine[tot + lane] = 0.0;        // padding behind the last value
            if (lane == 0) tot_s[r_in_cta][b] = tot;
            while (bits) {
                const uint32_t bb = __ffs(bits) - 1;
                bits &= bits - 1;
                const double v = lw[w * 32u + bb];
                neg_inf |= (v == -INFINITY);
                mine[off++] = v;
            }
            named_bar_arrive(bar_full + b, 64);
        }
        neg_inf = __any_sync(0xffffffffu, neg_inf);
        if (lane == 0) { fin_s[r_in_cta][0] = cnt; fin_s[r_in_cta][1] = neg_inf ? 1 : 0; }
        named_bar_arrive(bar_done, 64);
    } else {
        double sum = 0.0;
        for (uint32_t ch = 0; ch < n_chunks; ch++) {
            const uint32_t b = ch & 1u;
            named_bar_sync(bar_full + b, 64);
            const double *mine = list[r_in_cta][b];
            const uint32_t tot = tot_s[r_in_cta][b];
            // every lane runs the same chain
            double nx[FIT_BATCH];
#pragma unroll
            for (int q = 0; q < FIT_BATCH; q++) nx[q] = mine[q];
            for (uint32_t t = 0; t < tot; t += FIT_BATCH) {
                double cu[FIT_BATCH];
#pragma unroll
                for (int q = 0; q < FIT_BATCH; q++) cu[q] = nx[q];
#pragma unroll
                for (int q = 0; q < FIT_BATCH; q++) nx[q] = mine[t + FIT_BATCH + q];
#pragma unroll
                for (int q = 0; q < FIT_BATCH; q++) sum += cu[q];
            }
            if (ch + 2 < n_chunks) named_bar_arrive(bar_empty + b, 64);
        }
        named_bar_sync(bar_done, 64);
        if (lane == 0) {
            logfit[row] = (n_genes > 0) ? (fin_s[r_in_cta][1] ? 0.0 : sum) : 0.0;
            num_genes[row] = fin_s[r_in_cta][0];
        }
    }
}

// The same sequential sum with ONE LANE PER ROW: every lane walks the set bits of its own row in
// column order and runs its own dependent f64 chain, so a warp advances 32 rows per DADD instead
// of one. The kernel needs 1/32 of the warps of the row-per-warp-pair kernel above and therefore
// hardly takes SM slots away from the core kernel it runs beside (measured: the row-per-warp-pair
// kernel holds ~500 CTAs for ~30 us each and costs the concurrent core kernel 6 %). The warp steps
// through ALL gene positions (G dependent DADD issues, each predicated on the lane's own bit) with
// the lw values of a 1024-gene chunk in shared memory. Genes with lw = -inf (s = -1) are given as
// a bit mask: they are left out of the chain and any of them present sets the result to 0.0
// (population.rs:312-318).
constexpr int FITL_THREADS = 128;

// sum += x iff (bits & mask). Written as the fused multiply-add sum = x * b + sum with b = 1.0 or +0.0
// built from the bit (one select for the high word of b; the low word is zero): x * 1.0 = x and
// x * 0.0 = +-0.0 are exact, so the single rounding of the FMA is the rounding of the plain addition,
// and sum + (+-0.0) = sum because the running sum starts at +0.0 and can never become -0.0. The
// operand selection is off the dependent chain (x and bits are known ahead): one DFMA per position.
__device__ __forceinline__ void add_if_bit(double &sum, double x, uint32_t bits, uint32_t mask)
{
    int hi;
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tselp.b32 %0, 0x3FF00000, 0, p;\n\t}"
        : "=r"(hi) : "r"(bits), "r"(mask));
    sum = fma(x, __hiloint2double(hi, 0), sum);
}

__global__ void __launch_bounds__(FITL_THREADS, 4) fitness_lane_kernel(const uint32_t *acc, uint32_t n_rows,
                                                                    uint32_t n_genes, uint32_t stride_words,
                                                                    const double *lw, const uint32_t *lethal,
                                                                    double *logfit, int32_t *num_genes)
{
    __shared__ double lw_s[FIT_CHUNK_WORDS * 32];
    __shared__ uint32_t lethal_s[FIT_CHUNK_WORDS];
    // the lane's own 32 row words of the chunk (transposed: conflict-free), fetched with eight 16-byte loads before the
    // chain starts -- read one by one from global memory they cost an L2 round trip per word (2.5 x the chain itself)
    __shared__ uint32_t bits_s[FIT_CHUNK_WORDS][FITL_THREADS];
    const uint32_t row = blockIdx.x * FITL_THREADS + threadIdx.x;
    const bool live = row < n_rows;
    const uint32_t *r = acc + (uint64_t)(live ? row : 0u) * stride_words;
    const uint32_t n_words = (n_genes + 31u) / 32u;
    double sum = 0.0;
    bool neg_inf = false;
    int32_t cnt = 0;
    // the chunk's inputs travel through registers: those of chunk k + 1 are requested before the chain of chunk k starts
    // and land while it runs, so only the first chunk waits for global memory
    uint4 rw[FIT_CHUNK_WORDS / 4];
    double lwr[FIT_CHUNK_WORDS * 32 / FITL_THREADS];
    uint32_t lethal_r = 0;
    auto fetch = [&](uint32_t w0) {
#pragma unroll
        for (uint32_t q = 0; q < FIT_CHUNK_WORDS / 4; q++)      // rows are padded with zero words to a multiple of four (16-byte aligned)
            rw[q] = (live && w0 + 4u * q < stride_words) ? __ldg(reinterpret_cast<const uint4 *>(r + w0) + q) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (uint32_t q = 0; q < FIT_CHUNK_WORDS * 32 / FITL_THREADS; q++) {
            const uint32_t g = w0 * 32u + q * FITL_THREADS + threadIdx.x;
            lwr[q] = g < n_genes ? __ldg(lw + g) : 0.0;
        }
        if (threadIdx.x < FIT_CHUNK_WORDS) lethal_r = (w0 + threadIdx.x < n_words) ? __ldg(lethal + w0 + threadIdx.x) : 0u;
    };
    fetch(0u);
    for (uint32_t w0 = 0; w0 < n_words; w0 += FIT_CHUNK_WORDS) {
        __syncthreads();                                        // the previous chunk's chain has left shared memory
#pragma unroll
        for (uint32_t q = 0; q < FIT_CHUNK_WORDS * 32 / FITL_THREADS; q++)
            lw_s[q * FITL_THREADS + threadIdx.x] = (lwr[q] == -INFINITY) ? 0.0 : lwr[q];
        if (threadIdx.x < FIT_CHUNK_WORDS) lethal_s[threadIdx.x] = lethal_r;
        __syncthreads();
#pragma unroll
        for (uint32_t q = 0; q < FIT_CHUNK_WORDS / 4; q++) {
            const uint32_t wq[4] = {rw[q].x, rw[q].y, rw[q].z, rw[q].w};
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                const uint32_t dead = wq[i] & lethal_s[4u * q + i];        // present genes with s = -1 (population.rs:312-318)
                cnt += __popc(wq[i]);
                neg_inf |= dead != 0u;
                bits_s[4u * q + i][threadIdx.x] = wq[i] ^ dead;
            }
        }
        if (w0 + FIT_CHUNK_WORDS < n_words) fetch(w0 + FIT_CHUNK_WORDS);
        const uint32_t nw = min((uint32_t)FIT_CHUNK_WORDS, n_words - w0);
        // The additions of 16 gene positions (one half word) run while the lw values of the next
        // half word are already on their way from shared memory (same address for every lane:
        // broadcast; unconditional volatile loads so that they stay ahead of the chain), so the
        // chain advances at the DFMA latency (8.4 cycles, tools/experiments/fp64_latency.cu).
        const uint32_t lw_sa = smem_u32(lw_s);
        double xa[16], xb[16];
        auto load_half = [&](double (&x)[16], uint32_t half) {       // half = 2 * word + (0 | 1)
#pragma unroll
            for (int q = 0; q < 8; q++)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[2 * q]), "=d"(x[2 * q + 1])
                             : "r"(lw_sa + half * 128u + q * 16u));
        };
        uint32_t bits_next = bits_s[0][threadIdx.x];
        load_half(xa, 0u);
#pragma unroll 1
        for (uint32_t w = 0; w < nw; w++) {
            const uint32_t bits = bits_next;
            bits_next = bits_s[min(w + 1u, (uint32_t)FIT_CHUNK_WORDS - 1u)][threadIdx.x];
            load_half(xb, 2u * w + 1u);
#pragma unroll
            for (int b = 0; b < 16; b++) add_if_bit(sum, xa[b], bits, 1u << b);
            load_half(xa, min(2u * w + 2u, 2u * FIT_CHUNK_WORDS - 1u));
#pragma unroll
            for (int b = 0; b < 16; b++) add_if_bit(sum, xb[b], bits, 1u << (16 + b));
        }
    }
    if (live) {
        logfit[row] = (n_genes > 0) ? (neg_inf ? 0.0 : sum) : 0.0;
        num_genes[row] = cnt;
    }
}

// Large shapes (N x G above FITNESS_EXACT_CELLS): the strictly sequential chain above costs one
// issue slot per addition per row and dominates the accessory/selection chain (0.6 ms at
// N = 10 000, G = 18 000). The blocked variant adds, per row, the present genes of each 32-gene
// word in column order (one lane per word), then the word sums of each 1024-gene chunk in word
// order, then the chunk sums in chunk order: a fixed association, independent of grid shape and
// GPU count, within a few ulp (|rel| < 1e-13) of the reference's flat left-to-right sum.
__global__ void __launch_bounds__(FIT_WARPS * 32) fitness_blocked_kernel(const uint32_t *acc, uint32_t n_rows,
                                                                         uint32_t n_genes, uint32_t stride_words,
                                                                         const double *lw, double *logfit,
                                                                         int32_t *num_genes)
{
    __shared__ double lwc[1024];                       // lw of the current 1024-gene chunk (shared by the CTA's rows)
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * FIT_WARPS + warp;
    const bool live = row < n_rows;
    const uint32_t *r = acc + (uint64_t)(live ? row : 0) * stride_words;
    const uint32_t n_words = (n_genes + 31u) / 32u;
    double sum = 0.0;
    bool neg_inf = false;
    int32_t cnt = 0;
    for (uint32_t w0 = 0; w0 < n_words; w0 += 32) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < 1024; i += FIT_WARPS * 32) {
            const uint32_t g = w0 * 32u + i;
            lwc[i] = g < n_genes ? lw[g] : 0.0;
        }
        __syncthreads();
        const uint32_t w = w0 + lane;
        uint32_t bits = (live && w < n_words) ? r[w] : 0u;
        cnt += __popc(bits);
        double ws = 0.0;
        const double *lww = lwc + lane * 32u;
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            const double v = lww[b];
            neg_inf |= (v == -INFINITY);
            ws += v;
        }
        double cs = 0.0;
#pragma unroll
        for (int t = 0; t < 32; t++) cs += __shfl_sync(0xffffffffu, ws, t);      // word order
        sum += cs;                                                               // chunk order
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    neg_inf = __any_sync(0xffffffffu, neg_inf);
    if (lane == 0 && live) {
        logfit[row] = (n_genes > 0) ? (neg_inf ? 0.0 : sum) : 0.0;
        num_genes[row] = cnt;
    }
}

// ---------------------------------------------------------------------------
// K2a: all-vs-all intersection counts I[i][j] = popc(row_i & row_j), 32x32 tiles,
// upper triangle computed, mirrored on store.
// ---------------------------------------------------------------------------
constexpr int INTER_CHUNK = 32;   // words per shared-memory chunk

__global__ void __launch_bounds__(256) acc_inter_kernel(const uint32_t *acc, uint32_t n_rows,
                                                        uint32_t stride_words, uint32_t n_words,
                                                        uint32_t *inter, int32_t *diag)
{
    const uint32_t bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    __shared__ uint32_t Ri[32][INTER_CHUNK + 1];
    __shared__ uint32_t Rj[32][INTER_CHUNK + 1];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    uint32_t acc4[4] = {0, 0, 0, 0};
    for (uint32_t w0 = 0; w0 < n_words; w0 += INTER_CHUNK) {
        for (uint32_t q = ty; q < 32; q += 8) {
            const uint32_t w = w0 + tx;
            const uint32_t ri = bi * 32 + q, rj = bj * 32 + q;
            Ri[q][tx] = (w < n_words && ri < n_rows) ? acc[(uint64_t)ri * stride_words + w] : 0u;
            Rj[q][tx] = (w < n_words && rj < n_rows) ? acc[(uint64_t)rj * stride_words + w] : 0u;
        }
        __syncthreads();
#pragma unroll 4
        for (uint32_t w = 0; w < INTER_CHUNK; w++) {
            const uint32_t vj = Rj[tx][w];
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) acc4[q] += __popc(Ri[ty + 8 * q][w] & vj);
        }
        __syncthreads();
    }
    for (uint32_t q = 0; q < 4; q++) {
        const uint32_t i = bi * 32 + ty + 8 * q, j = bj * 32 + tx;
        if (i < n_rows && j < n_rows) {
            inter[(uint64_t)i * n_rows + j] = acc4[q];
            inter[(uint64_t)j * n_rows + i] = acc4[q];
            if (i == j) diag[i] = (int32_t)acc4[q];           // |row_i|: the gene count the distances need
        }
    }
}

// K2a (tensor cores): the intersection counts are the dense contraction X * X^T of the 0/1
// presence matrix, so they run on the tensor cores as an integer MMA with exact s32 accumulation:
// warp-level mma.sync m16n8k32 u8 x u8 (IMMA.16832 in SASS), the bits of a 32-gene word expanded
// to 0/1 bytes in registers. A dot product does not care in which k slot a gene sits as long as
// both operands agree, so the expansion is the cheapest one: the thread with t = lane % 4 takes
// bits {t, t+8, t+16, t+24} ((w >> t) & 0x01010101) for the low k half of the fragment and bits
// {t+4, ...} for the high half -- over t = 0..3 every bit of the word exactly once.
// CTA = 64 x 64 pairs, 8 warps of 16 x 32; upper-triangle tiles only, mirrored on store. The
// diagonal I[i][i] = |row_i| is also written as a compact vector: the distance kernel takes its
// gene counts from there, so it does not have to wait for the (much longer) fitness kernel.
constexpr int IM_TILE = 64;
constexpr int IM_CHUNK = 32;      // words per shared-memory chunk

__device__ __forceinline__ void imma_16832_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) acc_inter_mma_kernel(const uint32_t *acc, uint32_t n_rows,
                                                            uint32_t stride_words, uint32_t n_words,
                                                            uint32_t *inter, int32_t *diag)
{
    // blockIdx.x enumerates the upper-triangle tiles row by row: (0,0..nb-1), (1,1..nb-1), ...
    const uint32_t nb = (n_rows + IM_TILE - 1) / IM_TILE;
    uint32_t t_lin = blockIdx.x, bi = 0;
    while (t_lin >= nb - bi) { t_lin -= nb - bi; bi++; }
    const uint32_t bj = bi + t_lin;
    __shared__ uint32_t Ri[IM_TILE][IM_CHUNK + 1];
    __shared__ uint32_t Rj[IM_TILE][IM_CHUNK + 1];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t mrow = (warp & 3u) * 16u;          // this warp's 16 rows of the i tile
    const uint32_t ncol = (warp >> 2) * 32u;          // and 32 columns (4 n-tiles of 8) of the j tile
    int c[4][4];
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
        for (int q = 0; q < 4; q++) c[n][q] = 0;
    for (uint32_t w0 = 0; w0 < n_words; w0 += IM_CHUNK) {
        __syncthreads();
        const uint32_t w = w0 + lane;
#pragma unroll
        for (uint32_t q = 0; q < IM_TILE / 8; q++) {
            const uint32_t rr = warp + 8u * q;
            const uint32_t ri = bi * IM_TILE + rr, rj = bj * IM_TILE + rr;
            Ri[rr][lane] = (w < n_words && ri < n_rows) ? acc[(uint64_t)ri * stride_words + w] : 0u;
            Rj[rr][lane] = (w < n_words && rj < n_rows) ? acc[(uint64_t)rj * stride_words + w] : 0u;
        }
        __syncthreads();
#pragma unroll 4
        for (uint32_t ks = 0; ks < IM_CHUNK; ks++) {
            const uint32_t wa = Ri[mrow + g][ks] >> t, wb = Ri[mrow + g + 8][ks] >> t;
            const uint32_t a0 = wa & 0x01010101u, a1 = wb & 0x01010101u;
            const uint32_t a2 = (wa >> 4) & 0x01010101u, a3 = (wb >> 4) & 0x01010101u;
#pragma unroll
            for (int n = 0; n < 4; n++) {
                const uint32_t wj = Rj[ncol + n * 8 + g][ks] >> t;
                imma_16832_u8(c[n], a0, a1, a2, a3, wj & 0x01010101u, (wj >> 4) & 0x01010101u);
            }
        }
    }
    // accumulator layout: c[n][0..1] = (row g, cols 2t, 2t+1), c[n][2..3] = (row g + 8, same cols)
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i = bi * IM_TILE + mrow + g + ((q >> 1) ? 8u : 0u);
            const uint32_t j = bj * IM_TILE + ncol + n * 8 + 2 * t + (q & 1);
            if (i < n_rows && j < n_rows) {
                inter[(uint64_t)i * n_rows + j] = (uint32_t)c[n][q];
                inter[(uint64_t)j * n_rows + i] = (uint32_t)c[n][q];
                if (i == j) diag[i] = c[n][q];                // |row_i|: the gene count the distances need
            }
        }
}

// K2a (tcgen05): the same contraction X * X^T on the 5th-generation tensor cores. The presence matrix
// is first expanded to one byte per gene (acc_expand_bytes_kernel: 0/1 bytes, K-major rows padded to
// whole 128-byte swizzle rows; the order of the 32 genes of a word inside their 32 bytes is a fixed
// permutation, which a dot product does not see). One CTA per upper-triangle tile of 128 x 128
// pairs: a producer thread streams [128 rows x 128 genes] operand tiles through a 4-stage
// TMA/mbarrier ring (2-D tensor map, 128-byte swizzle, out-of-range rows are zero-filled by the
// TMA unit), one thread issues tcgen05.mma.kind::i8 (u8 x u8 -> s32, M = N = 128, K = 32 per
// instruction, four per stage) with the accumulator tile in tensor memory (128 lanes x 128
// columns), tcgen05.commit releases the stages and finally signals the epilogue: every warp reads
// its 32 accumulator lanes with tcgen05.ld (32x32b.x32) and stores the counts directly and
// mirrored. Exact: the products are 0/1 and the s32 accumulation cannot overflow.
constexpr int UM_TILE = 128;                 // rows per operand tile = UMMA M = UMMA N
constexpr int UM_KBYTES = 128;               // genes per stage = one swizzle row
constexpr int UM_STAGES = 4;
constexpr int UM_THREADS = 128;
constexpr uint32_t UM_TILE_BYTES = UM_TILE * UM_KBYTES;
constexpr uint32_t UM_TMEM_COLS = 128;
// instruction descriptor (kind::i8): D = s32 (bits 4-5 = 2), A = B = u8 (0), both K-major (0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t UM_IDESC = (2u << 4) | ((uint32_t)(UM_TILE >> 3) << 17) | ((uint32_t)(UM_TILE >> 4) << 24);
// shared-memory matrix descriptor, K-major, 128-byte swizzle: stride between 8-row groups 1024 B (bits 32-45, in 16 B),
// leading byte offset 1 (unused with this swizzle), descriptor version 1 (bit 46), layout SWIZZLE_128B = 2 (bits 61-63)
constexpr uint32_t UM_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

static inline size_t inter_umma_smem_bytes()
{
    return 1024 /* alignment slack */ + 2 * (size_t)UM_STAGES * UM_TILE_BYTES + (2 * UM_STAGES + 1) * sizeof(uint64_t) + 16;
}

// row r, 32-gene word w -> 32 bytes: byte 4t + q = bit t + 8q of the word
__global__ void __launch_bounds__(256) acc_expand_bytes_kernel(const uint32_t *acc, uint32_t n_rows, uint32_t stride_words,
                                                               uint32_t kpad_words, uint8_t *bytes)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)n_rows * kpad_words) return;
    const uint32_t row = (uint32_t)(idx / kpad_words), w = (uint32_t)(idx % kpad_words);
    const uint32_t v = w < stride_words ? acc[(uint64_t)row * stride_words + w] : 0u;
    uint4 lo, hi;
    lo.x = v & 0x01010101u; lo.y = (v >> 1) & 0x01010101u; lo.z = (v >> 2) & 0x01010101u; lo.w = (v >> 3) & 0x01010101u;
    hi.x = (v >> 4) & 0x01010101u; hi.y = (v >> 5) & 0x01010101u; hi.z = (v >> 6) & 0x01010101u; hi.w = (v >> 7) & 0x01010101u;
    uint4 *dst = reinterpret_cast<uint4 *>(bytes + idx * 32u);
    dst[0] = lo;
    dst[1] = hi;
}

// mbarrier wait that traps instead of spinning forever (a wrong tensor map or descriptor would otherwise hang the device)
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar_s, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

__device__ __forceinline__ void tma_load_2d_s(uint32_t smem_dst, const void *tmap, uint32_t x, uint32_t y, uint32_t bar_s)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(bar_s) : "memory");
}

__global__ void __launch_bounds__(UM_THREADS, 1) acc_inter_umma_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t n_rows,
                                                                       uint32_t k_iters, uint32_t *inter, int32_t *diag)
{
    extern __shared__ uint8_t um_smem[];
    const uint32_t base_s = (smem_u32(um_smem) + 1023u) & ~1023u;      // operand tiles on 1 KiB boundaries (swizzle atom)
    const uint32_t a_s = base_s, b_s = base_s + UM_STAGES * UM_TILE_BYTES;
    const uint32_t full_s = b_s + UM_STAGES * UM_TILE_BYTES, empty_s = full_s + UM_STAGES * 8u, done_s = empty_s + UM_STAGES * 8u;
    const uint32_t slot_s = done_s + 8u;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nb = (n_rows + UM_TILE - 1) / UM_TILE;
    uint32_t t_lin = blockIdx.x, bi = 0;
    while (t_lin >= nb - bi) { t_lin -= nb - bi; bi++; }
    const uint32_t bj = bi + t_lin;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < 2 * UM_STAGES + 1; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_s + s * 8u), "r"(1u) : "memory");
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_s), "r"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot_s) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer ----
            for (uint32_t it = 0; it < k_iters; it++) {
                const uint32_t s = it % UM_STAGES, n = it / UM_STAGES;
                if (n) mbar_wait_or_trap(empty_s + s * 8u, (n - 1u) & 1u);       // the MMAs that read the stage's previous fill have completed
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_s + s * 8u), "r"(2u * UM_TILE_BYTES) : "memory");
                tma_load_2d_s(a_s + s * UM_TILE_BYTES, &tmap, it * UM_KBYTES, bi * UM_TILE, full_s + s * 8u);
                tma_load_2d_s(b_s + s * UM_TILE_BYTES, &tmap, it * UM_KBYTES, bj * UM_TILE, full_s + s * 8u);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issue ----
            for (uint32_t it = 0; it < k_iters; it++) {
                const uint32_t s = it % UM_STAGES, n = it / UM_STAGES;
                mbar_wait_or_trap(full_s + s * 8u, n & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_lo = (((a_s + s * UM_TILE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t b_lo = (((b_s + s * UM_TILE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
                for (uint32_t k = 0; k < UM_KBYTES / 32; k++) {           // K = 32 bytes per instruction: +32 B inside the swizzle row
                    const uint64_t da = ((uint64_t)UM_DESC_HI << 32) | (a_lo + 2u * k);
                    const uint64_t db = ((uint64_t)UM_DESC_HI << 32) | (b_lo + 2u * k);
                    const uint32_t accumulate = (it | k) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(UM_IDESC), "r"(accumulate) : "memory");
                }
                // arrives on the barrier once the MMAs issued so far have read their operands (implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty_s + s * 8u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_s) : "memory");
        }
        __syncwarp();
    }

    // ---- epilogue: accumulator lane 32 * warp + lane = row of the i tile, columns = rows of the j tile ----
    mbar_wait_or_trap(done_s, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t i = bi * UM_TILE + warp * 32u + lane;
    const bool vec_ok = (n_rows & 3u) == 0u;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < (uint32_t)UM_TILE; c0 += 32) {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((warp * 32u) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t j0 = bj * UM_TILE + c0;
        if (i < n_rows) {
            uint32_t *dst = inter + (uint64_t)i * n_rows + j0;
            if (vec_ok && j0 + 32u <= n_rows) {
#pragma unroll
                for (int q = 0; q < 32; q += 4) *reinterpret_cast<uint4 *>(dst + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 32; q++)
                    if (j0 + q < n_rows) dst[q] = v[q];
            }
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const uint32_t j = j0 + q;
                if (j < n_rows) {
                    if (bi != bj) inter[(uint64_t)j * n_rows + i] = v[q];      // lanes = consecutive i: coalesced
                    else if (i == j) diag[i] = (int32_t)v[q];                  // |row_i|: the gene count the distances need
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(UM_TMEM_COLS) : "memory");
}

// K2a (tcgen05, operands expanded inside the CTA): the kernel above is bound by the L2 -> SM traffic of the
// byte matrix (1 MiB per 128 x 128 tile). Here the CTA reads the BIT rows (an eighth of that), eight warps
// expand them to 0/1 bytes straight into the swizzled K-major operand layout the tensor core expects, make the
// writes visible to the async proxy (fence.proxy.async) and hand the stage to the MMA thread through an
// mbarrier; tcgen05.commit returns it. Thread = one row of the stage (128 rows of the i tile, 128 of the j
// tile), KB / 8 bytes of bits per row per stage, fetched one stage ahead. The 32 genes of a word go to 32
// consecutive bytes, byte 4t + q = bit t + 8q: a fixed permutation of the k index, which a dot product does not see.
//   KB = 128: rows of 128 bytes, 16-byte chunk c of row r stored at chunk c ^ (r % 8) (SWIZZLE_128B, what a TMA load
//             would have produced; 8-row groups 1024 bytes apart), four stages of 32 KiB.
//   KB = 64:  rows of 64 bytes, chunk c of row r at c ^ ((r / 2) % 4) (SWIZZLE_64B; 8-row groups 512 bytes apart),
//             three stages of 16 KiB: 48 KiB of shared memory and <= 56 registers x 288 threads, i.e. the CTA fits
//             into the slot ONE retiring core_mut_kernel CTA leaves on an SM (the generation pipeline keeps every SM
//             full of those), instead of waiting for three.
constexpr int UB_EXP_WARPS = 8;
constexpr int UB_THREADS = 32 * (UB_EXP_WARPS + 1);       // + the MMA warp

template <int KB> struct UbCfg {
    static constexpr uint32_t STAGES = KB == 128 ? 4u : 3u;
    static constexpr uint32_t TILE_BYTES = UM_TILE * KB;
    // matrix descriptor, high word: stride between 8-row groups (8 * KB bytes, in 16 B), version 1, layout 2 = SWIZZLE_128B / 4 = SWIZZLE_64B
    static constexpr uint32_t DESC_HI = ((8u * KB) >> 4) | (1u << 14) | ((KB == 128 ? 2u : 4u) << 29);
    static constexpr size_t smem_bytes() { return 1024 + 2 * (size_t)STAGES * TILE_BYTES + (2 * STAGES + 1) * sizeof(uint64_t) + 16; }
};

template <int KB>
__global__ void __launch_bounds__(UB_THREADS, KB == 128 ? 1 : 4) acc_inter_umma_bits_kernel(const uint32_t *__restrict__ acc, uint32_t n_rows,
                                                                                            uint32_t stride_words, uint32_t k_iters,
                                                                                            uint32_t *inter, int32_t *diag)
{
    using Cfg = UbCfg<KB>;
    constexpr uint32_t STAGES = Cfg::STAGES, TILE_BYTES = Cfg::TILE_BYTES, WORDS = KB / 32;      // bit words per row per stage
    extern __shared__ uint8_t um_smem[];
    const uint32_t base_s = (smem_u32(um_smem) + 1023u) & ~1023u;
    const uint32_t a_s = base_s, b_s = base_s + STAGES * TILE_BYTES;
    const uint32_t full_s = b_s + STAGES * TILE_BYTES, empty_s = full_s + STAGES * 8u, done_s = empty_s + STAGES * 8u;
    const uint32_t slot_s = done_s + 8u;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nb = (n_rows + UM_TILE - 1) / UM_TILE;
    uint32_t t_lin = blockIdx.x, bi = 0;
    while (t_lin >= nb - bi) { t_lin -= nb - bi; bi++; }
    const uint32_t bj = bi + t_lin;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_s + s * 8u), "r"((uint32_t)UB_EXP_WARPS) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_s + s * 8u), "r"(1u) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(done_s), "r"(1u) : "memory");
        fence_mbar_init();
    }
    if (warp == UB_EXP_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_s), "r"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot_s) : "memory");

    if (warp < UB_EXP_WARPS) {
        // ---- expansion: threads 0..127 = rows of the i tile, 128..255 = rows of the j tile ----
        const uint32_t tr = threadIdx.x & 127u;                          // row inside the tile
        const bool is_b = threadIdx.x >= 128u;
        const uint32_t grow = (is_b ? bj : bi) * UM_TILE + tr;
        const bool row_ok = grow < n_rows;
        const uint32_t *src = acc + (uint64_t)(row_ok ? grow : 0u) * stride_words;      // rows: zero-padded to a multiple of 4 words
        const uint32_t row_s = (is_b ? b_s : a_s) + tr * KB;
        const uint32_t sw = (KB == 128 ? (tr & 7u) : ((tr >> 1) & 3u)) << 4;             // XOR term of the swizzle
        auto fetch = [&](uint32_t it, uint32_t (&w)[WORDS]) {
            const bool ok = row_ok && (it + 1u) * WORDS <= stride_words;
            if (KB == 128) {
                const uint4 v = ok ? __ldg(reinterpret_cast<const uint4 *>(src) + it) : make_uint4(0, 0, 0, 0);
                w[0] = v.x; w[1] = v.y; w[WORDS - 2] = v.z; w[WORDS - 1] = v.w;
            } else {
                const uint2 v = ok ? __ldg(reinterpret_cast<const uint2 *>(src) + it) : make_uint2(0, 0);
                w[0] = v.x; w[WORDS - 1] = v.y;
            }
        };
        uint32_t nxt[WORDS];
        fetch(0u, nxt);
        for (uint32_t it = 0; it < k_iters; it++) {
            const uint32_t s = it % STAGES, n = it / STAGES;
            uint32_t cur[WORDS];
#pragma unroll
            for (uint32_t q = 0; q < WORDS; q++) cur[q] = nxt[q];
            if (it + 1u < k_iters) fetch(it + 1u, nxt);
            if (n) mbar_wait_or_trap(empty_s + s * 8u, (n - 1u) & 1u);      // the MMAs that read the stage's previous contents have completed
            const uint32_t dst = row_s + s * TILE_BYTES;
#pragma unroll
            for (uint32_t q = 0; q < WORDS; q++) {
                const uint32_t v = cur[q];
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2u * q) << 4) ^ sw)), "r"(v & 0x01010101u),
                             "r"((v >> 1) & 0x01010101u), "r"((v >> 2) & 0x01010101u), "r"((v >> 3) & 0x01010101u) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2u * q + 1u) << 4) ^ sw)), "r"((v >> 4) & 0x01010101u),
                             "r"((v >> 5) & 0x01010101u), "r"((v >> 6) & 0x01010101u), "r"((v >> 7) & 0x01010101u) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_s + s * 8u) : "memory");
        }
    } else {
        if (lane == 0) {
            // ---- MMA issue ----
            for (uint32_t it = 0; it < k_iters; it++) {
                const uint32_t s = it % STAGES, n = it / STAGES;
                mbar_wait_or_trap(full_s + s * 8u, n & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_lo = (((a_s + s * TILE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t b_lo = (((b_s + s * TILE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
                for (uint32_t k = 0; k < KB / 32; k++) {                     // K = 32 bytes per instruction: +32 B inside the swizzled row
                    const uint64_t da = ((uint64_t)Cfg::DESC_HI << 32) | (a_lo + 2u * k);
                    const uint64_t db = ((uint64_t)Cfg::DESC_HI << 32) | (b_lo + 2u * k);
                    const uint32_t accumulate = (it | k) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(UM_IDESC), "r"(accumulate) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty_s + s * 8u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_s) : "memory");
        }
        __syncwarp();
    }

    // ---- epilogue: warps w and w + 4 share the accumulator lanes 32 * (w % 4) .. + 31 and split the columns ----
    if (warp < UB_EXP_WARPS) {
        mbar_wait_or_trap(done_s, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t quad = warp & 3u;
        const uint32_t i = bi * UM_TILE + quad * 32u + lane;
        const bool vec_ok = (n_rows & 3u) == 0u;
#pragma unroll 1
        for (uint32_t c0 = (warp >> 2) * 64u; c0 < (warp >> 2) * 64u + 64u; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem + ((quad * 32u) << 16) + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const uint32_t j0 = bj * UM_TILE + c0;
            if (i < n_rows) {
                uint32_t *dst = inter + (uint64_t)i * n_rows + j0;
                if (vec_ok && j0 + 16u <= n_rows) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4) *reinterpret_cast<uint4 *>(dst + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        if (j0 + q < n_rows) dst[q] = v[q];
                }
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const uint32_t j = j0 + q;
                    if (j < n_rows) {
                        if (bi != bj) inter[(uint64_t)j * n_rows + i] = v[q];      // lanes = consecutive i: coalesced
                        else if (i == j) diag[i] = (int32_t)v[q];                  // |row_i|: the gene count the distances need
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == UB_EXP_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(UM_TMEM_COLS) : "memory");
}

// K2b: mean Jaccard distance of individual i to all j != i, summed in j order
// exactly like get_distance + the fold of population.rs:770-771.
constexpr int AVG_WARPS = 8;
constexpr int AVG_RING = 256;       // distances evaluated per round (8 per lane)
constexpr int AVG_BATCH = 8;

// RCP = true: the quotient a / b of the two small integers a = |i & j| + core_genes <= b = |i | j| + core_genes is
// formed from the correctly rounded reciprocal r = RN(1 / b) (host table indexed by b) instead of the ~100-instruction
// IEEE division sequence:
//     q0 = RN(a * r);  e = a - q0 * b (exact, one FMA);  q = RN(q0 + e * r)
// which is the correctly rounded quotient RN(a / b) (Markstein's division step; checked exhaustively for every
// 0 <= a <= b <= 131072 by tools/check_recip_division.c), i.e. bit for bit what the division of population.rs:144-145
// returns; b = 0 (two empty genomes, no core genes) gives r = inf and a NaN, as 0.0 / 0.0 does.
constexpr uint32_t AVG_RCP_MAX = 131072;        // largest denominator the table may serve (range of the exhaustive check)

template <bool RCP>
__global__ void __launch_bounds__(AVG_WARPS * 32) avg_distance_kernel(const uint32_t *inter,
                                                                      const int32_t *num_genes, uint32_t n_rows,
                                                                      uint32_t core_genes, const double *__restrict__ rcp,
                                                                      double *avgdist)
{
    // One warp per individual. Per round the 32 lanes evaluate 256 distances in parallel (the
    // f64 divisions are independent) into shared memory, then the values are added one by one in
    // j order -- the same sequential chain as the fold of population.rs:770 -- with the loads one
    // batch ahead of the additions. Entries with j == i or j >= N are +0.0, which never changes
    // the (non-negative) sum.
    __shared__ double ring[AVG_WARPS][AVG_RING + 2 * AVG_BATCH];
    pdl_launch_dependents();       // the parent-selection CTA may take its SM slot now (it waits for this grid)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * AVG_WARPS + warp;
    if (i >= n_rows) return;
    double *mine = ring[warp];
    if (lane < 2 * AVG_BATCH) mine[AVG_RING + lane] = 0.0;
    const double cg = (double)core_genes;
    const uint32_t ki = (uint32_t)num_genes[i];
    const uint32_t *irow = inter + (uint64_t)i * n_rows;
    double sum = 0.0;
    for (uint32_t j0 = 0; j0 < n_rows; j0 += AVG_RING) {
        // all loads of the round first (one latency), then the divisions
        uint32_t in[AVG_RING / 32], kj[AVG_RING / 32];
#pragma unroll
        for (int q = 0; q < AVG_RING / 32; q++) {
            const uint32_t j = min(j0 + q * 32 + lane, n_rows - 1u);
            in[q] = irow[j];
            kj[q] = (uint32_t)num_genes[j];
        }
        double rr[AVG_RING / 32];
        if (RCP) {
#pragma unroll
            for (int q = 0; q < AVG_RING / 32; q++) rr[q] = __ldg(rcp + (ki + kj[q] - in[q] + core_genes));
        }
#pragma unroll
        for (int q = 0; q < AVG_RING / 32; q++) {
            const uint32_t j = j0 + q * 32 + lane;
            const uint32_t un = ki + kj[q] - in[q];
            double d;
            if (RCP) {
                const double da = (double)(in[q] + core_genes), db = (double)(un + core_genes);
                const double r = rr[q], q0 = da * r;
                d = 1.0 - fma(fma(-q0, db, da), r, q0);
            } else {
                d = 1.0 - (((double)in[q] + 0.0 + cg) / ((double)un + 0.0 + cg));   // :144-145
            }
            mine[q * 32 + lane] = (j < n_rows && j != i) ? d : 0.0;
        }
        __syncwarp();
        const uint32_t tot = min((uint32_t)AVG_RING, n_rows - j0);
        double nx[AVG_BATCH];
#pragma unroll
        for (int q = 0; q < AVG_BATCH; q++) nx[q] = mine[q];
        for (uint32_t t = 0; t < tot; t += AVG_BATCH) {
            double cu[AVG_BATCH];
#pragma unroll
            for (int q = 0; q < AVG_BATCH; q++) cu[q] = nx[q];
#pragma unroll
            for (int q = 0; q < AVG_BATCH; q++) nx[q] = mine[t + AVG_BATCH + q];
#pragma unroll
            for (int q = 0; q < AVG_BATCH; q++) sum += cu[q];
        }
        __syncwarp();
    }
    if (lane == 0) {
        double fd = sum / (double)(n_rows - 1u);
        if (fd == 0.0) fd = DBL_MIN;                                  // :774-776
        avgdist[i] = fd;
    }
}

// The same kernel (warp per individual, reciprocal-table quotients) software-pipelined over the rounds, for the
// generation pipeline where the warp shares its scheduler with eight core-kernel warps and every exposed global-load
// latency is paid in full: while the chain of round r runs over one half of a double-buffered ring,
//   * the reciprocals of round r + 1 are fetched (their addresses come from intersection counts loaded a round earlier),
//   * its quotients are evaluated into the other half of the ring between the two halves of the chain,
//   * the intersection counts and gene counts of round r + 2 are requested.
// The loads are volatile asm so that they are issued where they are written; two values per shared-memory load in the
// chain. Same additions in the same order: same bits.
__device__ __forceinline__ uint32_t ldg_u32_v(const void *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_f64_v(const void *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(AVG_WARPS * 32, 4) avg_distance_pipe_kernel(const uint32_t *inter, const int32_t *num_genes,
                                                                             uint32_t n_rows, uint32_t core_genes,
                                                                             const double *__restrict__ rcp, double *avgdist)
{
    constexpr int PER = AVG_RING / 32;               // distances per lane per round
    __shared__ __align__(16) double ring[AVG_WARPS][2][AVG_RING];
    pdl_launch_dependents();       // the parent-selection CTA may take its SM slot now (it waits for this grid)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * AVG_WARPS + warp;
    if (i >= n_rows) return;
    const uint32_t ki = (uint32_t)num_genes[i] + core_genes;
    const uint32_t *irow = inter + (uint64_t)i * n_rows;
    uint32_t in[PER], kj[PER];
    double rr[PER];
    auto load_counts = [&](uint32_t j0) {            // past the end: clamped (the values become +0.0 below)
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const uint32_t j = min(j0 + q * 32 + lane, n_rows - 1u);
            in[q] = ldg_u32_v(irow + j);
            kj[q] = ldg_u32_v(num_genes + j);
        }
    };
    auto load_rcp = [&]() {
#pragma unroll
        for (int q = 0; q < PER; q++) rr[q] = ldg_f64_v(rcp + (ki + kj[q] - in[q]));
    };
    auto quotients = [&](uint32_t j0, double *buf) {
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const uint32_t j = j0 + q * 32 + lane;
            const double da = (double)(in[q] + core_genes), db = (double)(ki + kj[q] - in[q]);
            const double q0 = da * rr[q];
            const double d = 1.0 - fma(fma(-q0, db, da), rr[q], q0);                  // population.rs:144-145
            buf[q * 32 + lane] = (j < n_rows && j != i) ? d : 0.0;                    // +0.0 never changes the (non-negative) sum
        }
    };
    double sum = 0.0;
    auto chain = [&](uint32_t base_s) {              // AVG_RING / 2 values, in order
#pragma unroll
        for (int c = 0; c < AVG_RING / 2; c += 2) {
            double x0, x1;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(base_s + c * 8u));
            sum += x0;
            sum += x1;
        }
    };
    load_counts(0u);
    load_rcp();
    quotients(0u, ring[warp][0]);
    load_counts(AVG_RING);
    __syncwarp();
    uint32_t b = 0;
    for (uint32_t j0 = 0; j0 < n_rows; j0 += AVG_RING, b ^= 1u) {
        const uint32_t cur = smem_u32(ring[warp][b]);
        load_rcp();                                              // round r + 1
        chain(cur);
        quotients(j0 + AVG_RING, ring[warp][b ^ 1u]);
        load_counts(j0 + 2 * AVG_RING);                          // round r + 2
        chain(cur + (AVG_RING / 2) * 8u);
        __syncwarp();
    }
    if (lane == 0) {
        double fd = sum / (double)(n_rows - 1u);
        if (fd == 0.0) fd = DBL_MIN;                                  // :774-776
        avgdist[i] = fd;
    }
}

// ---------------------------------------------------------------------------
// K3: weights (three softmaxes multiplied, population.rs:325-393), the all-zero
// rule (:403,435-437), WeightedIndex<f64> (cumulative + binary search, :440)
// and N draws from Philox(seed, gen, individual). Single CTA: N is small and
// the whole thing is a chain of reductions.
// ---------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;      // one CTA slot of an SM: the kernel has to squeeze in beside the core step
constexpr int SEL_WARPS = SEL_THREADS / 32;

// block-wide reduction of three values at once (fixed order: shuffle tree, then the warp results
// in warp order), result in every thread
template <bool IS_MAX, int WARPS = SEL_WARPS>
__device__ __forceinline__ void block_reduce3(double &x, double &y, double &z, double (*scratch)[3])
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double tx = __shfl_xor_sync(0xffffffffu, x, o), ty = __shfl_xor_sync(0xffffffffu, y, o),
                     tz = __shfl_xor_sync(0xffffffffu, z, o);
        x = IS_MAX ? fmax(x, tx) : x + tx;
        y = IS_MAX ? fmax(y, ty) : y + ty;
        z = IS_MAX ? fmax(z, tz) : z + tz;
    }
    __syncthreads();
    if (lane == 0) { scratch[warp][0] = x; scratch[warp][1] = y; scratch[warp][2] = z; }
    __syncthreads();
    x = scratch[0][0]; y = scratch[0][1]; z = scratch[0][2];
#pragma unroll
    for (int q = 1; q < WARPS; q++) {
        x = IS_MAX ? fmax(x, scratch[q][0]) : x + scratch[q][0];
        y = IS_MAX ? fmax(y, scratch[q][1]) : y + scratch[q][1];
        z = IS_MAX ? fmax(z, scratch[q][2]) : z + scratch[q][2];
    }
}

struct SelectArgs {
    const double *logfit;
    const int32_t *num_genes;
    const double *avgdist;        // nullptr = all 1.0 (main.rs:435)
    uint32_t n_rows, n_genes;
    int32_t avg_gene_num;
    int32_t no_control_genome_size;
    double log_penalty;           // ln(genome_size_penalty), host libm
    double competition_strength;
    uint2 key;
    uint32_t gen;
    const uint32_t *gen_dev;      // nullptr, or a device word added to gen (replayed CUDA graphs)
    double *tmp_a, *tmp_b;        // [N] scratch
    double *weights;              // [N] out: final weights (population.rs:389-437)
    double *cumulative;           // [N] out
    uint32_t *parents;            // [N + 1] out: the parents, then a status word (1 = WeightedIndex::new would fail)
    int *err_flag;                // set to 1 if WeightedIndex::new would fail
};

// N <= 1024: the same computation with every per-individual value in registers (4 per thread) and
// the cumulative table in shared memory -- the kernel is a chain of dependent passes, so keeping
// them out of global memory is what makes it short.
constexpr int SEL_PER = 4;
constexpr uint32_t SEL_SMALL_MAX = SEL_THREADS * SEL_PER;

__global__ void __launch_bounds__(SEL_THREADS, 4) select_parents_small_kernel(const SelectArgs a)
{
    __shared__ double scratch[SEL_WARPS][3];
    __shared__ double warp_excl[SEL_WARPS + 1];
    __shared__ double w_s[SEL_SMALL_MAX];
    __shared__ double cum_s[SEL_SMALL_MAX];
    __shared__ int bad;
    const uint32_t n = a.n_rows, tid = threadIdx.x;
    const bool use_a = a.n_genes > 0, use_b = !a.no_control_genome_size;
    if (tid == 0) bad = 0;
    pdl_wait();                    // mean distances of avg_distance_kernel
    double xa[SEL_PER], xb[SEL_PER], xc[SEL_PER];
    double ma = -INFINITY, mb = -INFINITY, mc = -INFINITY;
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {
        const uint32_t i = tid + k * SEL_THREADS;
        const bool ok = i < n;
        const uint32_t ii = ok ? i : 0u;
        const double lf = a.logfit[ii];
        const int32_t ng = a.num_genes[ii];
        const double ad = a.avgdist ? a.avgdist[ii] : 1.0;
        xa[k] = ok ? (use_a ? lf : 0.0) : -INFINITY;
        xb[k] = ok ? (use_b ? (double)(ng - a.avg_gene_num) * a.log_penalty : 0.0) : -INFINITY;     // :350,355
        xc[k] = ok ? a.competition_strength * log(ad) : -INFINITY;                                  // :375
        ma = fmax(ma, xa[k]); mb = fmax(mb, xb[k]); mc = fmax(mc, xc[k]);
    }
    block_reduce3<true>(ma, mb, mc, scratch);
    double sa = 0.0, sb = 0.0, sc = 0.0;
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {        // entries beyond n are -inf: exp(-inf - m) = 0
        sa += exp(xa[k] - ma); sb += exp(xb[k] - mb); sc += exp(xc[k] - mc);
    }
    block_reduce3<false>(sa, sb, sc, scratch);
    const double la = (ma == -INFINITY) ? -INFINITY : ma + log(sa);
    const double lb = (mb == -INFINITY) ? -INFINITY : mb + log(sb);
    const double lc = (mc == -INFINITY) ? -INFINITY : mc + log(sc);
    double ta = 0.0, tb = 0.0, tc = 0.0;
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {
        xa[k] = exp(xa[k] - la); xb[k] = exp(xb[k] - lb); xc[k] = exp(xc[k] - lc);
        ta += xa[k]; tb += xb[k]; tc += xc[k];
    }
    block_reduce3<false>(ta, tb, tc, scratch);
    double w[SEL_PER];
    double mx = -INFINITY, dummy0 = -INFINITY, dummy1 = -INFINITY;
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {
        const double wa = use_a ? xa[k] / ta : 1.0;                 // a_i = 1 without an accessory genome (:293-296)
        const double w0 = use_b ? (xb[k] / tb) * wa : wa;           // :368 / :371
        w[k] = w0 * (xc[k] / tc);                                   // :391
        if (tid + k * SEL_THREADS < n) mx = fmax(mx, w[k]);
    }
    block_reduce3<true>(mx, dummy0, dummy1, scratch);
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {
        const uint32_t i = tid + k * SEL_THREADS;
        if (mx == 0.0) w[k] = 1.0;                                  // :435-437
        if (i < n) { w_s[i] = w[k]; a.weights[i] = w[k]; }
    }
    __syncthreads();
    // WeightedIndex::new: cumulative sums over contiguous chunks of 4, then a scan of the chunk totals
    const uint32_t lo = min(n, tid * SEL_PER), hi = min(n, lo + SEL_PER);
    double s = 0.0;
    bool mybad = false;
    for (uint32_t i = lo; i < hi; i++) {
        const double wi = w_s[i];
        if (!(wi >= 0.0)) mybad = true;
        s += wi;
    }
    if (mybad) bad = 1;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    double v = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += t;
    }
    if (lane == 31) scratch[warp][0] = v;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int q = 0; q < SEL_WARPS; q++) { warp_excl[q] = run; run += scratch[q][0]; }
        warp_excl[SEL_WARPS] = run;
        if (!(run > 0.0) || isinf(run)) bad = 1;
    }
    __syncthreads();
    const double total = warp_excl[SEL_WARPS];
    double run = warp_excl[warp] + (v - s);
    for (uint32_t i = lo; i < hi; i++) {
        run += w_s[i];
        cum_s[i] = run;
        a.cumulative[i] = run;
    }
    __syncthreads();
    if (tid == 0) a.parents[n] = bad ? 1u : 0u;       // status word behind the vector: one read-back serves both
    if (bad) {
        if (tid == 0) *a.err_flag = 1;
        for (uint32_t i = tid; i < n; i += SEL_THREADS) a.parents[i] = i;
        return;
    }
    // N draws: u ~ U[0,total), index = #cumulative[0..n-1) <= u
#pragma unroll
    for (int k = 0; k < SEL_PER; k++) {
        const uint32_t i = tid + k * SEL_THREADS;
        const uint4 r = philox4x32_10(make_ctr(i, 0u, a.gen + (a.gen_dev ? __ldg(a.gen_dev) : 0u), STREAM_PARENTS), a.key);
        const uint64_t bits = (((uint64_t)r.x << 32) | r.y) >> 11;
        const double u = (double)bits * 0x1.0p-53 * total;
        uint32_t l = 0, h = n - 1;
        while (l < h) {
            const uint32_t mid = l + ((h - l) >> 1);
            if (cum_s[mid] <= u) l = mid + 1; else h = mid;
        }
        if (i < n) a.parents[i] = l;
    }
}

// The three softmaxes of population.rs:325-382 (fitness a, genome size b, competition c) are
// independent of each other, so each of their passes is made once for all three:
//   m = max v; s = sum exp(v - m); lse = m + ln s; e_i = exp(v_i - lse); out_i = e_i / sum e
// THREADS = 1024 for large populations: four times the lanes for the transcendental passes and the
// binary searches, and a CTA that owns its SM (nothing else fits beside 1024 threads), so the kernel
// is not slowed by the core step's CTAs once it has got its slot.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) select_parents_kernel(const SelectArgs a)
{
    constexpr int WARPS = THREADS / 32;
    __shared__ double scratch[WARPS][3];
    __shared__ double warp_excl[WARPS + 1];
    __shared__ int bad;
    const uint32_t n = a.n_rows, tid = threadIdx.x;
    const bool use_a = a.n_genes > 0, use_b = !a.no_control_genome_size;
    double *va = a.tmp_a, *vb = a.tmp_b, *vc = a.cumulative;      // vc: scratch until the cumulative sums are written
    if (tid == 0) bad = 0;
    pdl_wait();                    // mean distances of avg_distance_kernel

    // pass 0: the three arguments and their maxima
    double ma = -INFINITY, mb = -INFINITY, mc = -INFINITY;
    for (uint32_t i = tid; i < n; i += THREADS) {
        const double xa = use_a ? a.logfit[i] : 0.0;
        const double xb = use_b ? (double)(a.num_genes[i] - a.avg_gene_num) * a.log_penalty : 0.0;     // :350,355
        const double xc = a.competition_strength * log(a.avgdist ? a.avgdist[i] : 1.0);                // :375
        vb[i] = xb; vc[i] = xc;
        ma = fmax(ma, xa); mb = fmax(mb, xb); mc = fmax(mc, xc);
    }
    block_reduce3<true, WARPS>(ma, mb, mc, scratch);
    // pass 1: log-sum-exp
    double sa = 0.0, sb = 0.0, sc = 0.0;
    for (uint32_t i = tid; i < n; i += THREADS) {
        if (use_a) sa += exp(a.logfit[i] - ma);
        if (use_b) sb += exp(vb[i] - mb);
        sc += exp(vc[i] - mc);
    }
    block_reduce3<false, WARPS>(sa, sb, sc, scratch);
    const double la = (ma == -INFINITY) ? -INFINITY : ma + log(sa);
    const double lb = (mb == -INFINITY) ? -INFINITY : mb + log(sb);
    const double lc = (mc == -INFINITY) ? -INFINITY : mc + log(sc);
    // pass 2: exp(v - lse) and their totals
    double ta = 0.0, tb = 0.0, tc = 0.0;
    for (uint32_t i = tid; i < n; i += THREADS) {
        const double ea = use_a ? exp(a.logfit[i] - la) : 1.0;
        const double eb = use_b ? exp(vb[i] - lb) : 1.0;
        const double ec = exp(vc[i] - lc);
        va[i] = ea; vb[i] = eb; vc[i] = ec;
        ta += ea; tb += eb; tc += ec;
    }
    block_reduce3<false, WARPS>(ta, tb, tc, scratch);
    // pass 3: weights (population.rs:365-393) and their maximum
    double mx = -INFINITY, dummy0 = -INFINITY, dummy1 = -INFINITY;
    for (uint32_t i = tid; i < n; i += THREADS) {
        const double wa = use_a ? va[i] / ta : 1.0;                 // a_i = 1 without an accessory genome (:293-296)
        const double w0 = use_b ? (vb[i] / tb) * wa : wa;           // :368 / :371
        const double w = w0 * (vc[i] / tc);                         // :391
        a.weights[i] = w;
        mx = fmax(mx, w);
    }
    block_reduce3<true, WARPS>(mx, dummy0, dummy1, scratch);
    if (mx == 0.0)                                                                       // :435-437
        for (uint32_t i = tid; i < n; i += THREADS) a.weights[i] = 1.0;
    __syncthreads();

    // WeightedIndex::new: cumulative sums; contiguous chunk per thread, then a scan of the chunk totals
    const uint32_t per = (n + THREADS - 1) / THREADS;
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    double s = 0.0;
    bool mybad = false;
    for (uint32_t i = lo; i < hi; i++) {
        const double w = a.weights[i];
        if (!(w >= 0.0)) mybad = true;
        s += w;
    }
    if (mybad) bad = 1;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    double v = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v += t;
    }
    if (lane == 31) scratch[warp][0] = v;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int q = 0; q < WARPS; q++) { warp_excl[q] = run; run += scratch[q][0]; }
        warp_excl[WARPS] = run;
        if (!(run > 0.0) || isinf(run)) bad = 1;
    }
    __syncthreads();
    const double total = warp_excl[WARPS];
    double run = warp_excl[warp] + (v - s);
    for (uint32_t i = lo; i < hi; i++) {
        run += a.weights[i];
        a.cumulative[i] = run;
    }
    __syncthreads();
    if (tid == 0) a.parents[n] = bad ? 1u : 0u;       // status word behind the vector: one read-back serves both
    if (bad) {
        if (tid == 0) *a.err_flag = 1;
        for (uint32_t i = tid; i < n; i += THREADS) a.parents[i] = i;
        return;
    }
    // N draws: u ~ U[0,total), index = #cumulative[0..n-1) <= u
    for (uint32_t i = tid; i < n; i += THREADS) {
        const uint4 r = philox4x32_10(make_ctr(i, 0u, a.gen + (a.gen_dev ? __ldg(a.gen_dev) : 0u), STREAM_PARENTS), a.key);
        const uint64_t bits = (((uint64_t)r.x << 32) | r.y) >> 11;
        const double u = (double)bits * 0x1.0p-53 * total;
        uint32_t l = 0, h = n - 1;
        while (l < h) {
            const uint32_t mid = l + ((h - l) >> 1);
            if (a.cumulative[mid] <= u) l = mid + 1; else h = mid;
        }
        a.parents[i] = l;
    }
}

}  // namespace pansim
