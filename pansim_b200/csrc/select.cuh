// select.cuh -- K1/K2/K3: fitness, competition term and parent resampling.
//   K1 fitness           population.rs:282-322  (f64 sum in column order -> bit-exact)
//   K2 average_distance  population.rs:753-784 + get_distance :114-151 (bit-exact)
//   K3 weights + draw    population.rs:325-447
#pragma once
#include <cfloat>
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
constexpr int FIT_WARPS = 4;
constexpr int FIT_CHUNK_WORDS = 32;               // 1024 genes per compaction round

__global__ void __launch_bounds__(FIT_WARPS * 32) fitness_kernel(const uint32_t *acc, uint32_t n_rows,
                                                                 uint32_t n_genes, uint32_t stride_words,
                                                                 const double *lw, double *logfit,
                                                                 int32_t *num_genes)
{
    // Terms of absent genes are ln(1 + s*0) = +0.0: adding them never changes the
    // running sum, so only present genes are added -- in increasing column order,
    // by a single sequential f64 chain per row (bit-exact vs population.rs:303-317).
    // The warp first compacts the present genes' lw values into shared memory so
    // the chain is not exposed to load latency.
    __shared__ double list[FIT_WARPS][FIT_CHUNK_WORDS * 32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = blockIdx.x * FIT_WARPS + warp;
    if (row >= n_rows) return;
    const uint32_t *r = acc + (uint64_t)row * stride_words;
    const uint32_t n_words = (n_genes + 31u) / 32u;
    double *mine = list[warp];
    double sum = 0.0;
    bool neg_inf = false;
    int32_t cnt = 0;
    for (uint32_t w0 = 0; w0 < n_words; w0 += FIT_CHUNK_WORDS) {
        const uint32_t w = w0 + lane;
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
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            const double v = lw[w * 32u + b];
            neg_inf |= (v == -INFINITY);
            mine[off++] = v;
        }
        __syncwarp();
        // every lane runs the same chain; loads are batched so the chain only waits on the adds
        uint32_t t = 0;
        for (; t + 8 <= tot; t += 8) {
            const double v0 = mine[t], v1 = mine[t + 1], v2 = mine[t + 2], v3 = mine[t + 3];
            const double v4 = mine[t + 4], v5 = mine[t + 5], v6 = mine[t + 6], v7 = mine[t + 7];
            sum += v0; sum += v1; sum += v2; sum += v3; sum += v4; sum += v5; sum += v6; sum += v7;
        }
        for (; t < tot; t++) sum += mine[t];
        __syncwarp();
    }
    neg_inf = __any_sync(0xffffffffu, neg_inf);
    if (lane == 0) {
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
                                                        uint32_t *inter)
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
        }
    }
}

// K2b: mean Jaccard distance of individual i to all j != i, summed in j order
// exactly like get_distance + the fold of population.rs:770-771.
constexpr int AVG_WARPS = 4;

__global__ void __launch_bounds__(AVG_WARPS * 32) avg_distance_kernel(const uint32_t *inter,
                                                                      const int32_t *num_genes, uint32_t n_rows,
                                                                      uint32_t core_genes, double *avgdist)
{
    // One warp per individual: the 32 lanes evaluate 32 distances in parallel (the
    // f64 division is the expensive part), then the values are added one by one in
    // j order -- the same sequential chain as the fold of population.rs:770.
    // Lanes with j == i or j >= N contribute +0.0, which never changes the sum.
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * AVG_WARPS + (threadIdx.x >> 5);
    if (i >= n_rows) return;
    const double cg = (double)core_genes;
    const uint32_t ki = (uint32_t)num_genes[i];
    const uint32_t *irow = inter + (uint64_t)i * n_rows;
    double sum = 0.0;
    for (uint32_t j0 = 0; j0 < n_rows; j0 += 32) {
        const uint32_t j = j0 + lane;
        double d = 0.0;
        if (j < n_rows && j != i) {
            const uint32_t in = irow[j];
            const uint32_t un = ki + (uint32_t)num_genes[j] - in;
            d = 1.0 - (((double)in + 0.0 + cg) / ((double)un + 0.0 + cg));   // :144-145
        }
#pragma unroll
        for (int t = 0; t < 32; t++) sum += __shfl_sync(0xffffffffu, d, t);
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
constexpr int SEL_THREADS = 1024;

__device__ __forceinline__ double block_reduce(double v, bool is_max, double *scratch)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, t) : v + t;
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = scratch[0];
    for (int q = 1; q < SEL_THREADS / 32; q++) r = is_max ? fmax(r, scratch[q]) : r + scratch[q];
    return r;
}

// v[i] <- exp(v[i] - lse(v)) / sum(...)   (population.rs:325-340)
__device__ void softmax_inplace(double *v, uint32_t n, double *scratch)
{
    double m = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) m = fmax(m, v[i]);
    m = block_reduce(m, true, scratch);
    double s = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) s += exp(v[i] - m);
    s = block_reduce(s, false, scratch);
    const double lse = (m == -INFINITY) ? -INFINITY : m + log(s);
    double t = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) {
        const double e = exp(v[i] - lse);
        v[i] = e;
        t += e;
    }
    t = block_reduce(t, false, scratch);
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) v[i] = v[i] / t;
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
    double *tmp_a, *tmp_b;        // [N] scratch
    double *weights;              // [N] out: final weights (population.rs:389-437)
    double *cumulative;           // [N] out
    uint32_t *parents;            // [N] out
    int *err_flag;                // set to 1 if WeightedIndex::new would fail
};

__global__ void __launch_bounds__(SEL_THREADS) select_parents_kernel(const SelectArgs a)
{
    __shared__ double scratch[SEL_THREADS / 32];
    __shared__ double chunk_sum[SEL_THREADS];
    __shared__ int bad;
    const uint32_t n = a.n_rows;
    if (threadIdx.x == 0) bad = 0;

    // a_i: softmax of log-fitness (skipped when there is no accessory genome, :293-296)
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) a.tmp_a[i] = (a.n_genes > 0) ? a.logfit[i] : 1.0;
    __syncthreads();
    if (a.n_genes > 0) softmax_inplace(a.tmp_a, n, scratch);
    __syncthreads();

    if (!a.no_control_genome_size) {
        for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS)
            a.tmp_b[i] = (double)(a.num_genes[i] - a.avg_gene_num) * a.log_penalty;     // :350,355
        __syncthreads();
        softmax_inplace(a.tmp_b, n, scratch);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) a.weights[i] = a.tmp_b[i] * a.tmp_a[i];  // :368
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) a.weights[i] = a.tmp_a[i];               // :371
    }
    __syncthreads();

    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS)
        a.tmp_b[i] = a.competition_strength * log(a.avgdist ? a.avgdist[i] : 1.0);       // :375
    __syncthreads();
    softmax_inplace(a.tmp_b, n, scratch);
    __syncthreads();
    double mx = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) {
        const double w = a.weights[i] * a.tmp_b[i];                                      // :391
        a.weights[i] = w;
        mx = fmax(mx, w);
    }
    mx = block_reduce(mx, true, scratch);
    if (mx == 0.0)                                                                       // :435-437
        for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) a.weights[i] = 1.0;
    __syncthreads();

    // WeightedIndex::new: cumulative sums; contiguous chunk per thread, then a
    // scan of the chunk totals.
    const uint32_t per = (n + SEL_THREADS - 1) / SEL_THREADS;
    const uint32_t lo = threadIdx.x * per, hi = min(n, lo + per);
    double s = 0.0;
    bool mybad = false;
    for (uint32_t i = lo; i < hi; i++) {
        const double w = a.weights[i];
        if (!(w >= 0.0)) mybad = true;
        s += w;
    }
    if (mybad) bad = 1;
    // block-wide exclusive scan of the 1024 chunk totals (warp shuffles, fixed order)
    {
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double v = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, o);
            if ((int)lane >= o) v += t;
        }
        if (lane == 31) scratch[warp] = v;
        __syncthreads();
        if (warp == 0) {
            const double wt = scratch[lane];
            double wv = wt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, wv, o);
                if ((int)lane >= o) wv += t;
            }
            chunk_sum[lane] = wv - wt;            // exclusive offset of each warp
            if (lane == 31) {
                chunk_sum[32] = wv;               // grand total
                if (!(wv > 0.0) || isinf(wv)) bad = 1;
            }
        }
        __syncthreads();
        const double excl = chunk_sum[warp] + (v - s);
        const double tot = chunk_sum[32];
        __syncthreads();
        chunk_sum[threadIdx.x] = excl;
        if (threadIdx.x == 0) scratch[0] = tot;
    }
    __syncthreads();
    const double total = scratch[0];
    double run = chunk_sum[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) {
        run += a.weights[i];
        a.cumulative[i] = run;
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) *a.err_flag = 1;
        for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) a.parents[i] = i;
        return;
    }
    // N draws: u ~ U[0,total), index = #cumulative[0..n-1) <= u
    for (uint32_t i = threadIdx.x; i < n; i += SEL_THREADS) {
        const uint4 r = philox4x32_10(make_ctr(i, 0u, a.gen, STREAM_PARENTS), a.key);
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
