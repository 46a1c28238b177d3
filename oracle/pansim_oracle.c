/*
 * pansim_oracle.c -- CPU restatement of bacpop/Pansim (see pansim_oracle.h).
 * TEST INFRASTRUCTURE ONLY: parity checker + "port" CPU baseline.
 * PARITY UNPINNED for the samplers (third-party crates absent, see header).
 * Citations are file:line under /root/reference/pansim/src/.
 */
#define _GNU_SOURCE
#include "pansim_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================= */
/* RNG                                                                      */
/* ======================================================================= */
static inline uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void ora_rng_seed(ora_rng *r, uint64_t seed)
{
    uint64_t x = seed;
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}

void ora_rng_seed4(ora_rng *r, uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t x = seed;
    uint64_t h = splitmix64(&x);
    x = h ^ (a * 0xD6E8FEB86659FD93ULL); h = splitmix64(&x);
    x = h ^ (b * 0xCA5A826395121157ULL); h = splitmix64(&x);
    x = h ^ (c * 0x9FB21C651E98DF25ULL);
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t ora_rng_next(ora_rng *r)
{
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}

double ora_rng_f64(ora_rng *r) { return (double)(ora_rng_next(r) >> 11) * 0x1.0p-53; }

/* Lemire's multiply-shift with rejection: exactly uniform on [0,n). */
uint64_t ora_rng_below(ora_rng *r, uint64_t n)
{
    if (n == 0) return 0;
    uint64_t x = ora_rng_next(r);
    __uint128_t m = (__uint128_t)x * (__uint128_t)n;
    uint64_t l = (uint64_t)m;
    if (l < n) {
        uint64_t t = (0 - n) % n;
        while (l < t) {
            x = ora_rng_next(r);
            m = (__uint128_t)x * (__uint128_t)n;
            l = (uint64_t)m;
        }
    }
    return (uint64_t)(m >> 64);
}

/* Exact Poisson: sequential inversion below 10, Hoermann's PTRS above
 * (the reference calls statrs::Poisson::sample, population.rs:498,522,599). */
uint64_t ora_poisson(ora_rng *r, double mean)
{
    if (!(mean > 0.0)) return 0;
    if (mean < 10.0) {
        double p = exp(-mean), cdf = p, u = ora_rng_f64(r);
        uint64_t k = 0;
        while (u >= cdf && k < 1000) { k++; p *= mean / (double)k; cdf += p; }
        return k;
    }
    const double slam = sqrt(mean), loglam = log(mean);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        double U = ora_rng_f64(r) - 0.5;
        double V = ora_rng_f64(r);
        double us = 0.5 - fabs(U);
        double kf = floor((2.0 * a / us + b) * U + mean + 0.43);
        if (us >= 0.07 && V <= vr) return (uint64_t)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b)
            <= -mean + kf * loglam - lgamma(kf + 1.0))
            return (uint64_t)kf;
    }
}

double ora_exponential(ora_rng *r, double rate)
{
    double u = ora_rng_f64(r);
    return -log1p(-u) / rate;
}

int ora_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ora_set_threads(int n)
{
#ifdef _OPENMP
    if (n < 1) n = 1;            /* main.rs:249-251 */
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ======================================================================= */
/* distances.rs                                                             */
/* ======================================================================= */

/* distances.rs:22-52: XOR + popcount over u64 chunks, then the byte tail. */
uint32_t ora_hamming_bitwise_fast(const uint8_t *x, const uint8_t *y, size_t n)
{
    uint32_t distance = 0;
    size_t chunks = n / 8;
    for (size_t c = 0; c < chunks; c++) {
        uint64_t xv, yv;
        memcpy(&xv, x + 8 * c, 8);
        memcpy(&yv, y + 8 * c, 8);
        distance += (uint32_t)__builtin_popcountll(xv ^ yv);
    }
    for (size_t k = 8 * chunks; k < n; k++)            /* distances.rs:41-49 */
        distance += (uint32_t)__builtin_popcount((unsigned)(x[k] ^ y[k]));
    return distance;
}

/* distances.rs:55-77 */
void ora_jaccard_distance_fast(const uint8_t *x, const uint8_t *y, size_t n,
                               uint32_t *intersection, uint32_t *uni)
{
    uint32_t in = 0, un = 0;
    size_t chunks = n / 8;
    for (size_t c = 0; c < chunks; c++) {
        uint64_t xv, yv;
        memcpy(&xv, x + 8 * c, 8);
        memcpy(&yv, y + 8 * c, 8);
        in += (uint32_t)__builtin_popcountll(xv & yv);
        un += (uint32_t)__builtin_popcountll(xv | yv);
    }
    for (size_t k = 8 * chunks; k < n; k++) {
        in += (uint32_t)__builtin_popcount((unsigned)(x[k] & y[k]));
        un += (uint32_t)__builtin_popcount((unsigned)(x[k] | y[k]));
    }
    *intersection = in;
    *uni = un;
}

/* population.rs:32-48 (commented-out cross-check in the reference) */
void ora_jaccard_distance_naive(const uint8_t *x, const uint8_t *y, size_t n,
                                uint32_t *intersection, uint32_t *uni)
{
    uint32_t in = 0, un = 0;
    for (size_t k = 0; k < n; k++) {
        if (x[k] == 1 || y[k] == 1) {
            un++;
            if (x[k] == 1 && y[k] == 1) in++;
        }
    }
    *intersection = in;
    *uni = un;
}

/* ======================================================================= */
/* small helpers of population.rs                                           */
/* ======================================================================= */

/* population.rs:83-94: returns (std, mean); population variance (/n). */
void ora_standard_deviation(const double *v, size_t n, double *std_out, double *mean_out)
{
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += v[i];
    double mean = sum / (double)n;
    double ss = 0.0;
    for (size_t i = 0; i < n; i++) { double d = v[i] - mean; ss += d * d; }
    double variance = ss / (double)n;
    *std_out = sqrt(variance);
    *mean_out = mean;
}

/* population.rs:154-162 */
char ora_int_to_base(uint8_t n)
{
    switch (n) {
    case 1: return 'A';
    case 2: return 'C';
    case 4: return 'G';
    case 8: return 'T';
    default: return 'N';
    }
}

/* ======================================================================= */
/* Population::new  (population.rs:181-242)                                 */
/* ======================================================================= */
int ora_population_new(ora_population *p, size_t size, size_t allele_count,
                       uint8_t max_variants, int core, double avg_gene_freq,
                       ora_rng *rng, size_t core_genes)
{
    memset(p, 0, sizeof(*p));
    p->nrows = size;
    p->ncols = allele_count;
    p->core = core;
    p->core_genes = core_genes;
    p->avg_gene_freq = avg_gene_freq;
    size_t total = size * allele_count;
    p->pop = (uint8_t *)malloc(total ? total : 1);
    if (!p->pop) return -1;
    uint8_t *first = (uint8_t *)malloc(allele_count ? allele_count : 1);
    if (!first) return -1;
    if (core) {
        /* population.rs:201-204: 1 << gen_range(0..max_variants) */
        for (size_t j = 0; j < allele_count; j++)
            first[j] = (uint8_t)(1u << ora_rng_below(rng, max_variants));
    } else {
        /* population.rs:214-219: gen::<f64>() < avg_gene_freq */
        for (size_t j = 0; j < allele_count; j++)
            first[j] = (ora_rng_f64(rng) < avg_gene_freq) ? 1 : 0;
    }
    /* population.rs:206-212 / 221-229: every row identical */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; i++)
        memcpy(p->pop + i * allele_count, first, allele_count);
    free(first);
    return 0;
}

void ora_population_free(ora_population *p)
{
    free(p->pop);
    p->pop = NULL;
}

/* population.rs:244-268 */
double ora_calc_gene_freq(const ora_population *p)
{
    double sum = 0.0;
    for (size_t i = 0; i < p->nrows; i++) {
        size_t s = 0;
        const uint8_t *row = p->pop + i * p->ncols;
        for (size_t j = 0; j < p->ncols; j++) s += row[j];
        sum += (double)s / (double)p->ncols;
    }
    return sum / (double)p->nrows;
}

/* ======================================================================= */
/* sample_indices  (population.rs:270-448)                                  */
/* ======================================================================= */

/* logsumexp 0.1 `ln_sum_exp`: a fold of pairwise log-add-exp starting from
 * -inf (published algorithm; crate source not on disk -> unpinned). */
static double ln_add_exp(double a, double b)
{
    if (a == b && isinf(a)) return a;            /* (-inf,-inf) and (inf,inf) */
    double mx = a > b ? a : b, mn = a > b ? b : a;
    if (isnan(a) || isnan(b)) return NAN;
    return mx + log1p(exp(mn - mx));
}

static double ln_sum_exp(const double *v, size_t n)
{
    double acc = -INFINITY;
    for (size_t i = 0; i < n; i++) acc = ln_add_exp(acc, v[i]);
    return acc;
}

/* population.rs:325-340 / 356-361 / 377-382: exp(x - lse), then / sum. */
static void softmax_like_reference(double *v, size_t n)
{
    double lse = ln_sum_exp(v, n);
    for (size_t i = 0; i < n; i++) v[i] = exp(v[i] - lse);
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += v[i];
    for (size_t i = 0; i < n; i++) v[i] = (v[i] != -INFINITY) ? v[i] / sum : 0.0;
}

void ora_selection_weights(const ora_population *p, int32_t avg_gene_num,
                           const double *avg_pairwise_dists,
                           const double *selection_coefficients,
                           int no_control_genome_size, double genome_size_penalty,
                           double competition_strength,
                           double *weights_out, int32_t *num_genes_out,
                           double *logfit_out)
{
    size_t N = p->nrows, G = p->ncols;
    int32_t *num_genes = (int32_t *)malloc(sizeof(int32_t) * (N ? N : 1));
    double *sel = (double *)malloc(sizeof(double) * (N ? N : 1));
    double *tmp = (double *)malloc(sizeof(double) * (N ? N : 1));

    /* population.rs:282-291 */
    for (size_t i = 0; i < N; i++) {
        int32_t s = 0;
        const uint8_t *row = p->pop + i * G;
        for (size_t j = 0; j < G; j++) s += row[j];
        num_genes[i] = s;
    }
    for (size_t i = 0; i < N; i++) sel[i] = 1.0;             /* :293 */

    if (G > 0) {                                             /* :296 */
        for (size_t i = 0; i < N; i++) {                     /* :299-322, serial */
            const uint8_t *row = p->pop + i * G;
            int neg_inf = 0;
            double log_sum = 0.0;
            for (size_t j = 0; j < G; j++) {
                double lv = log(1.0 + selection_coefficients[j] * (double)row[j]);
                if (lv == -INFINITY) neg_inf = 1;            /* :312 */
                log_sum += lv;                               /* :317 (column order) */
            }
            sel[i] = neg_inf ? 0.0 : log_sum;                /* :315-318 */
        }
        if (logfit_out) memcpy(logfit_out, sel, sizeof(double) * N);
        softmax_like_reference(sel, N);                      /* :325-340 */
    } else if (logfit_out) {
        for (size_t i = 0; i < N; i++) logfit_out[i] = 0.0;
    }

    if (!no_control_genome_size) {                           /* :346-369 */
        double lp = log(genome_size_penalty);
        for (size_t i = 0; i < N; i++)
            tmp[i] = (double)(num_genes[i] - avg_gene_num) * lp;   /* :350,355 */
        softmax_like_reference(tmp, N);
        for (size_t i = 0; i < N; i++) weights_out[i] = tmp[i] * sel[i];  /* :368 */
    } else {
        memcpy(weights_out, sel, sizeof(double) * N);        /* :371 */
    }

    /* :375-393 */
    for (size_t i = 0; i < N; i++) tmp[i] = competition_strength * log(avg_pairwise_dists[i]);
    softmax_like_reference(tmp, N);
    for (size_t i = 0; i < N; i++) weights_out[i] = weights_out[i] * tmp[i];

    /* :403, 435-437 */
    double mx = -INFINITY;
    for (size_t i = 0; i < N; i++) mx = fmax(mx, weights_out[i]);   /* f64::max ignores NaN */
    if (mx == 0.0)
        for (size_t i = 0; i < N; i++) weights_out[i] = 1.0;

    if (num_genes_out) memcpy(num_genes_out, num_genes, sizeof(int32_t) * N);
    free(num_genes); free(sel); free(tmp);
}

int ora_weighted_index_sample(const double *weights, size_t n, ora_rng *rng,
                              size_t n_draws, uint32_t *out)
{
    if (n == 0) return -1;
    double *cum = (double *)malloc(sizeof(double) * n);
    double total = 0.0;
    for (size_t i = 0; i < n; i++) {
        if (!(weights[i] >= 0.0)) { free(cum); return -1; }   /* InvalidWeight */
        total += weights[i];
        cum[i] = total;
    }
    if (!(total > 0.0) || isinf(total)) { free(cum); return -1; }      /* AllWeightsZero */
    for (size_t d = 0; d < n_draws; d++) {
        double u = ora_rng_f64(rng) * total;
        /* partition_point(|w| w <= u) over cumulative[0..n-1) */
        size_t lo = 0, hi = n - 1;
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (cum[mid] <= u) lo = mid + 1; else hi = mid;
        }
        out[d] = (uint32_t)lo;
    }
    free(cum);
    return 0;
}

int ora_sample_indices(const ora_population *p, ora_rng *rng, int32_t avg_gene_num,
                       const double *avg_pairwise_dists,
                       const double *selection_coefficients,
                       int no_control_genome_size, double genome_size_penalty,
                       double competition_strength, uint32_t *parents_out)
{
    double *w = (double *)malloc(sizeof(double) * (p->nrows ? p->nrows : 1));
    ora_selection_weights(p, avg_gene_num, avg_pairwise_dists, selection_coefficients,
                          no_control_genome_size, genome_size_penalty,
                          competition_strength, w, NULL, NULL);
    int rc = ora_weighted_index_sample(w, p->nrows, rng, p->nrows, parents_out); /* :440-443 */
    free(w);
    return rc;
}

/* ======================================================================= */
/* next_generation  (population.rs:450-465) -- serial, like the reference   */
/* ======================================================================= */
int ora_next_generation(ora_population *p, const uint32_t *sample, size_t n)
{
    size_t ncols = p->ncols;
    uint8_t *next = (uint8_t *)calloc((n && ncols) ? n * ncols : 1, 1);   /* Array2::zeros, :455 */
    if (!next) return -1;
    for (size_t i = 0; i < n; i++)                                      /* :458-462 */
        memcpy(next + i * ncols, p->pop + (size_t)sample[i] * ncols, ncols);
    free(p->pop);
    p->pop = next;                                                      /* :464 */
    p->nrows = n;
    return 0;
}

/* ======================================================================= */
/* event log                                                                */
/* ======================================================================= */
void ora_events_init(ora_events *e) { memset(e, 0, sizeof(*e)); }

void ora_events_clear(ora_events *e)
{
    e->n_core_mut = e->n_acc_flip = e->n_hr = e->n_hgt = 0;
}

void ora_events_free(ora_events *e)
{
    free(e->core_mut_row); free(e->core_mut_site); free(e->core_mut_allele);
    free(e->acc_flip_row); free(e->acc_flip_gene);
    free(e->hr_recipient); free(e->hr_locus); free(e->hr_donor); free(e->hr_value);
    free(e->hgt_recipient); free(e->hgt_gene); free(e->hgt_donor);
    memset(e, 0, sizeof(*e));
}

#define GROW(ptr, type, newcap) ptr = (type *)realloc(ptr, sizeof(type) * (newcap))

static void reserve_core_mut(ora_events *e, size_t need)
{
    if (need <= e->cap_core_mut) return;
    size_t c = e->cap_core_mut ? e->cap_core_mut : 1024;
    while (c < need) c *= 2;
    GROW(e->core_mut_row, uint32_t, c); GROW(e->core_mut_site, uint32_t, c);
    GROW(e->core_mut_allele, uint8_t, c);
    e->cap_core_mut = c;
}
static void reserve_acc_flip(ora_events *e, size_t need)
{
    if (need <= e->cap_acc_flip) return;
    size_t c = e->cap_acc_flip ? e->cap_acc_flip : 1024;
    while (c < need) c *= 2;
    GROW(e->acc_flip_row, uint32_t, c); GROW(e->acc_flip_gene, uint32_t, c);
    e->cap_acc_flip = c;
}
static void reserve_hr(ora_events *e, size_t need)
{
    if (need <= e->cap_hr) return;
    size_t c = e->cap_hr ? e->cap_hr : 1024;
    while (c < need) c *= 2;
    GROW(e->hr_recipient, uint32_t, c); GROW(e->hr_locus, uint32_t, c);
    GROW(e->hr_donor, uint32_t, c); GROW(e->hr_value, uint8_t, c);
    e->cap_hr = c;
}
static void reserve_hgt(ora_events *e, size_t need)
{
    if (need <= e->cap_hgt) return;
    size_t c = e->cap_hgt ? e->cap_hgt : 1024;
    while (c < need) c *= 2;
    GROW(e->hgt_recipient, uint32_t, c); GROW(e->hgt_gene, uint32_t, c);
    GROW(e->hgt_donor, uint32_t, c);
    e->cap_hgt = c;
}

/* ======================================================================= */
/* site sampling == WeightedIndex<f32> over 0/1 weights                     */
/* ======================================================================= */
float *ora_build_cumulative(size_t ncols, uint32_t lo, uint32_t hi)
{
    /* WeightedIndex::new keeps the running sums of all but the last weight */
    float *c = (float *)malloc(sizeof(float) * (ncols ? ncols : 1));
    float acc = 0.0f;
    for (size_t j = 0; j < ncols; j++) {
        acc += (j >= lo && j < hi) ? 1.0f : 0.0f;
        c[j] = acc;
    }
    return c;
}

static inline uint32_t sample_site(const ora_site_dist *d, ora_rng *rng)
{
    uint32_t k = (uint32_t)ora_rng_below(rng, (uint64_t)(d->hi - d->lo));
    if (!d->cumulative) return d->lo + k;
    /* same control flow as WeightedIndex::sample: draw a weight in [0,total),
     * then partition_point(|w| w <= chosen) over the first n-1 running sums.
     * chosen = k + 0.5 lands in the k-th unit interval, so the result equals
     * lo + k (uniform), without rand's 23-bit f32 granularity quirk
     * (SURVEY.md 8a "do not replicate"). */
    float chosen = (float)k + 0.5f;
    size_t lo = 0, hi = d->n_table - 1;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (d->cumulative[mid] <= chosen) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}

/* ======================================================================= */
/* mutate_alleles  (population.rs:467-542)                                  */
/* ======================================================================= */
typedef struct { uint32_t *site; uint8_t *allele; size_t n; } row_events;

void ora_mutate_alleles(ora_population *p, const double *mutations_vec,
                        const ora_site_dist *dists, size_t n_compartments,
                        uint64_t rng_seed, uint64_t gen, ora_events *ev)
{
    static const uint8_t core_vec0[3] = {2, 4, 8};   /* core_vec[1 >> value] == core_vec[0], :531 */
    size_t N = p->nrows, C = p->ncols;
    for (size_t site_idx = 0; site_idx < n_compartments; site_idx++) {   /* :476 */
        double mutations = mutations_vec[site_idx];
        if (mutations == 0.0) continue;                                  /* :480-482 */
        row_events *log = NULL;
        if (ev) log = (row_events *)calloc(N ? N : 1, sizeof(row_events));
        const ora_site_dist *dist = &dists[site_idx];
        const int is_core = p->core;
#pragma omp parallel for schedule(dynamic, 4)
        for (size_t i = 0; i < N; i++) {                                 /* :488-491 / 512-515 */
            ora_rng trng;                                                /* thread_rng(), :493/517 */
            ora_rng_seed4(&trng, rng_seed, gen, (is_core ? 0x100 : 0x200) + site_idx, i);
            uint8_t *row = p->pop + i * C;
            size_t n_sites = (size_t)ora_poisson(&trng, mutations);      /* :498 / 522 */
            if (log) {
                log[i].n = n_sites;
                log[i].site = (uint32_t *)malloc(sizeof(uint32_t) * (n_sites ? n_sites : 1));
                log[i].allele = (uint8_t *)malloc(n_sites ? n_sites : 1);
            }
            for (size_t k = 0; k < n_sites; k++) {                       /* :501 / 525 */
                uint32_t site = sample_site(dist, &trng);                /* :503 / 527 */
                uint8_t new_allele;
                if (!is_core) {
                    new_allele = (row[site] == 0) ? 1 : 0;               /* :504-505 */
                } else {
                    /* :530-534: values = core_vec[1 >> value] -> always {2,4,8} */
                    new_allele = core_vec0[ora_rng_below(&trng, 3)];
                }
                row[site] = new_allele;                                  /* :508 / 537 */
                if (log) { log[i].site[k] = site; log[i].allele[k] = new_allele; }
            }
        }
        if (ev) {
            for (size_t i = 0; i < N; i++) {
                if (is_core) {
                    reserve_core_mut(ev, ev->n_core_mut + log[i].n);
                    for (size_t k = 0; k < log[i].n; k++) {
                        ev->core_mut_row[ev->n_core_mut] = (uint32_t)i;
                        ev->core_mut_site[ev->n_core_mut] = log[i].site[k];
                        ev->core_mut_allele[ev->n_core_mut] = log[i].allele[k];
                        ev->n_core_mut++;
                    }
                } else {
                    reserve_acc_flip(ev, ev->n_acc_flip + log[i].n);
                    for (size_t k = 0; k < log[i].n; k++) {
                        ev->acc_flip_row[ev->n_acc_flip] = (uint32_t)i;
                        ev->acc_flip_gene[ev->n_acc_flip] = log[i].site[k];
                        ev->n_acc_flip++;
                    }
                }
                free(log[i].site); free(log[i].allele);
            }
            free(log);
        }
    }
}

/* ======================================================================= */
/* recombine  (population.rs:544-751)                                       */
/* ======================================================================= */
typedef struct { uint32_t *recipient, *locus; uint8_t *value; size_t n; } donor_events;

int ora_recombine(ora_population *p, const double *recombinations_vec,
                  const ora_site_dist *dists, size_t n_compartments,
                  ora_rng *rng, uint64_t rng_seed, uint64_t gen, ora_events *ev)
{
    size_t N = p->nrows, C = p->ncols;
    for (size_t site_idx = 0; site_idx < n_compartments; site_idx++) {   /* :554 */
        double n_recombinations = recombinations_vec[site_idx];
        if (n_recombinations == 0.0) continue;                           /* :558-560 */
        if (N < 2) return -1;           /* Uniform::new(0, nrows-1) panics, :584 */
        donor_events *prop = (donor_events *)calloc(N, sizeof(donor_events));  /* :566-568 */
        const ora_site_dist *dist = &dists[site_idx];
        const int is_core = p->core;

        /* propose, :587-721 (rayon over donor rows) */
#pragma omp parallel for schedule(dynamic, 4)
        for (size_t row_idx = 0; row_idx < N; row_idx++) {
            ora_rng trng;                                                /* :596 */
            ora_rng_seed4(&trng, rng_seed, gen, (is_core ? 0x300 : 0x400) + site_idx, row_idx);
            const uint8_t *row = p->pop + row_idx * C;
            size_t n_sites = (size_t)ora_poisson(&trng, n_recombinations);   /* :599 */
            donor_events *d = &prop[row_idx];
            d->recipient = (uint32_t *)malloc(sizeof(uint32_t) * (n_sites ? n_sites : 1));
            d->locus = (uint32_t *)malloc(sizeof(uint32_t) * (n_sites ? n_sites : 1));
            d->value = (uint8_t *)malloc(n_sites ? n_sites : 1);
            for (size_t k = 0; k < n_sites; k++) {                       /* :616-619 */
                uint32_t v = (uint32_t)ora_rng_below(&trng, N - 1);
                d->recipient[k] = v + (v >= row_idx ? 1u : 0u);
                d->value[k] = 1;                                         /* :632 */
            }
            d->n = 0;
            if (!is_core) {
                /* :636-681: eligible loci = compartment weight 1 AND gene present */
                uint32_t *elig = (uint32_t *)malloc(sizeof(uint32_t) * (dist->hi - dist->lo + 1));
                size_t K = 0;
                for (uint32_t g = dist->lo; g < dist->hi; g++)
                    if (row[g] != 0) elig[K++] = g;
                if (K > 0) {                                             /* :672 */
                    for (size_t k = 0; k < n_sites; k++)                 /* :677-680 */
                        d->locus[k] = elig[ora_rng_below(&trng, K)];
                    d->n = n_sites;
                }
                free(elig);
            } else {
                for (size_t k = 0; k < n_sites; k++) {                   /* :687-695 */
                    uint32_t l = (uint32_t)ora_rng_below(&trng, C);
                    d->locus[k] = l;
                    d->value[k] = row[l];        /* snapshot before any apply */
                }
                d->n = n_sites;
            }
        }

        /* :725-726: seeded shuffle of the donor order (Fisher-Yates from the end) */
        uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * N);
        for (size_t i = 0; i < N; i++) order[i] = (uint32_t)i;
        for (size_t i = N - 1; i >= 1; i--) {
            size_t j = (size_t)ora_rng_below(rng, i + 1);
            uint32_t t = order[i]; order[i] = order[j]; order[j] = t;
        }

        /* :728-748 serial apply, last writer wins */
        for (size_t oi = 0; oi < N; oi++) {
            uint32_t donor = order[oi];
            donor_events *d = &prop[donor];
            if (d->n > 0) {                                              /* :740 */
                if (ev) {
                    if (is_core) reserve_hr(ev, ev->n_hr + d->n);
                    else reserve_hgt(ev, ev->n_hgt + d->n);
                }
                for (size_t k = 0; k < d->n; k++) {
                    p->pop[(size_t)d->recipient[k] * C + d->locus[k]] = d->value[k];   /* :745 */
                    if (ev) {
                        if (is_core) {
                            ev->hr_recipient[ev->n_hr] = d->recipient[k];
                            ev->hr_locus[ev->n_hr] = d->locus[k];
                            ev->hr_donor[ev->n_hr] = donor;
                            ev->hr_value[ev->n_hr] = d->value[k];
                            ev->n_hr++;
                        } else {
                            ev->hgt_recipient[ev->n_hgt] = d->recipient[k];
                            ev->hgt_gene[ev->n_hgt] = d->locus[k];
                            ev->hgt_donor[ev->n_hgt] = donor;
                            ev->n_hgt++;
                        }
                    }
                }
            }
        }
        for (size_t i = 0; i < N; i++) { free(prop[i].recipient); free(prop[i].locus); free(prop[i].value); }
        free(prop); free(order);
    }
    return 0;
}

/* ======================================================================= */
/* replay                                                                   */
/* ======================================================================= */
void ora_apply_core_writes(ora_population *p, const uint32_t *row, const uint32_t *site,
                           const uint8_t *value, size_t n)
{
    for (size_t k = 0; k < n; k++) p->pop[(size_t)row[k] * p->ncols + site[k]] = value[k];
}

void ora_apply_acc_flips(ora_population *p, const uint32_t *row, const uint32_t *gene, size_t n)
{
    for (size_t k = 0; k < n; k++) {
        uint8_t *c = &p->pop[(size_t)row[k] * p->ncols + gene[k]];
        *c = (*c == 0) ? 1 : 0;                       /* population.rs:504-508 */
    }
}

void ora_apply_acc_sets(ora_population *p, const uint32_t *row, const uint32_t *gene, size_t n)
{
    for (size_t k = 0; k < n; k++) p->pop[(size_t)row[k] * p->ncols + gene[k]] = 1;
}

int ora_step_replay(ora_population *core, ora_population *pan,
                    const uint32_t *parents, const ora_events *ev)
{
    /* main.rs:445-464 */
    if (ora_next_generation(core, parents, core->nrows)) return -1;
    if (ora_next_generation(pan, parents, pan->nrows)) return -1;
    ora_apply_core_writes(core, ev->core_mut_row, ev->core_mut_site, ev->core_mut_allele, ev->n_core_mut);
    ora_apply_acc_flips(pan, ev->acc_flip_row, ev->acc_flip_gene, ev->n_acc_flip);
    ora_apply_core_writes(core, ev->hr_recipient, ev->hr_locus, ev->hr_value, ev->n_hr);
    ora_apply_acc_sets(pan, ev->hgt_recipient, ev->hgt_gene, ev->n_hgt);
    return 0;
}

/* ======================================================================= */
/* distances                                                                */
/* ======================================================================= */

/* population.rs:822 */
double ora_core_distance_from_count(uint32_t core_diff, size_t ncols)
{
    return (double)core_diff / (double)ncols;
}

/* population.rs:828-830 (and :144-145 with matches == 0.0) */
double ora_acc_distance_from_counts(uint32_t inter, uint32_t uni, size_t core_genes)
{
    return 1.0 - (((double)inter + (double)core_genes) / ((double)uni + (double)core_genes));
}

/* population.rs:753-784 with get_distance (:114-151) */
void ora_average_distance(const ora_population *p, double *out)
{
    size_t N = p->nrows, C = p->ncols;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t i = 0; i < N; i++) {
        const uint8_t *ri = p->pop + i * C;
        double sum = 0.0;
        size_t count = 0;
        for (size_t j = 0; j < N; j++) {
            if (j == i) continue;                                        /* :128-130 */
            const uint8_t *rj = p->pop + j * C;
            double d;
            if (p->core) {
                uint32_t h = ora_hamming_bitwise_fast(ri, rj, C) / 2;    /* :136 */
                d = (double)h / (double)C;                               /* :137 */
            } else {
                uint32_t in, un;
                ora_jaccard_distance_fast(ri, rj, C, &in, &un);          /* :140 */
                d = 1.0 - (((double)in + 0.0 + (double)p->core_genes)
                           / ((double)un + 0.0 + (double)p->core_genes)); /* :144-145 */
            }
            sum += d;                                                    /* :770 */
            count++;
        }
        double fd = sum / (double)count;                                 /* :771 */
        if (fd == 0.0) fd = DBL_MIN;                                     /* :774-776 MIN_POSITIVE */
        out[i] = fd;
    }
}

void ora_pair_counts(const ora_population *p, size_t max_distances,
                     const uint32_t *range1, const uint32_t *range2,
                     uint32_t *core_diff, uint32_t *inter, uint32_t *uni)
{
    size_t C = p->ncols;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t k = 0; k < max_distances; k++) {                         /* :797-801 */
        const uint8_t *r1 = p->pop + (size_t)range1[k] * C;
        const uint8_t *r2 = p->pop + (size_t)range2[k] * C;
        if (p->core) {
            uint32_t h = ora_hamming_bitwise_fast(r1, r2, C) / 2;        /* :817 */
            if (core_diff) core_diff[k] = h;
        } else {
            uint32_t in, un;
            ora_jaccard_distance_fast(r1, r2, C, &in, &un);              /* :824 */
            if (inter) inter[k] = in;
            if (uni) uni[k] = un;
        }
    }
}

void ora_pairwise_distances(const ora_population *p, size_t max_distances,
                            const uint32_t *range1, const uint32_t *range2, double *out)
{
    size_t C = p->ncols;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t k = 0; k < max_distances; k++) {
        const uint8_t *r1 = p->pop + (size_t)range1[k] * C;
        const uint8_t *r2 = p->pop + (size_t)range2[k] * C;
        if (p->core) {
            uint32_t h = ora_hamming_bitwise_fast(r1, r2, C) / 2;
            out[k] = ora_core_distance_from_count(h, C);
        } else {
            uint32_t in, un;
            ora_jaccard_distance_fast(r1, r2, C, &in, &un);
            out[k] = ora_acc_distance_from_counts(in, un, p->core_genes);
        }
    }
}

void ora_gene_counts(const ora_population *p, uint32_t *out)
{
    size_t N = p->nrows, C = p->ncols;
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < C; j++) {
        uint32_t s = 0;
        for (size_t i = 0; i < N; i++) s += p->pop[i * C + j];           /* :850-851 */
        out[j] = s;
    }
}

/* population.rs:840-863: accessory genes first, then core_genes x 1.0 */
void ora_gene_frequencies(const ora_population *p, double *out)
{
    size_t C = p->ncols;
    uint32_t *cnt = (uint32_t *)malloc(sizeof(uint32_t) * (C ? C : 1));
    ora_gene_counts(p, cnt);
    double n_individuals = (double)p->nrows;
    for (size_t j = 0; j < C; j++) out[j] = (double)cnt[j] / n_individuals;   /* :852 */
    for (size_t j = 0; j < p->core_genes; j++) out[C + j] = 1.0;              /* :858-860 */
    free(cnt);
}

/* ======================================================================= */
/* main.rs: parameters                                                      */
/* ======================================================================= */
void ora_params_default(ora_params *p)
{
    memset(p, 0, sizeof(*p));
    p->pop_size = 1000; p->core_size = 1200000; p->pan_genes = 6000; p->core_genes = 2000;
    p->avg_gene_freq = 0.5; p->n_gen = 100; p->max_distances = 100000; p->core_mu = 0.05;
    p->HR_rate = 0.05; p->HGT_rate = 0.05; p->rate_genes1 = 1.0; p->rate_genes2 = 1000.0;
    p->prop_genes2 = 0.1; p->prop_positive = -0.1; p->pos_lambda = 10.0; p->neg_lambda = 10.0;
    p->seed = 0; p->genome_size_penalty = 0.99; p->competition_strength = 0.0; p->threads = 1;
}

/* main.rs:194-247 */
int ora_validate(const ora_params *p)
{
    if (p->core_genes > p->pan_genes) return 1;
    if (p->HR_rate < 0.0 || p->HGT_rate < 0.0) return 2;
    if (p->pos_lambda <= 0.0 || p->neg_lambda <= 0.0) return 3;
    if (p->rate_genes1 < 0.0 || p->rate_genes2 < 0.0) return 4;
    if (p->prop_genes2 < 0.0 || p->prop_genes2 > 1.0) return 5;
    if (p->pop_size < 1 || p->core_size < 1 || p->pan_genes < 1 || p->n_gen < 1 || p->max_distances < 1) return 6;
    if (p->core_mu < 0.0 || p->core_mu > 1.0) return 7;
    if (p->avg_gene_freq <= 0.0 || p->avg_gene_freq > 1.0) return 8;
    return 0;
}

/* main.rs:259-287, 333-367 */
void ora_derive(const ora_params *p, ora_derived *d)
{
    memset(d, 0, sizeof(*d));
    d->pan_size = p->pan_genes - p->core_genes;                              /* :259 */
    double core_prop = (double)p->core_genes / (double)p->pan_genes;         /* :263 */
    double acc_prop = 1.0 - core_prop;
    double agf = (p->avg_gene_freq - core_prop) / acc_prop;                  /* :265 */
    if (agf < 0.0) agf = 0.0;
    d->avg_gene_freq_adj = agf;
    d->avg_gene_num = (int32_t)round(agf * (double)d->pan_size);             /* :272 */
    d->n_core_mutations = ceil((double)p->core_size * p->core_mu);           /* :275-276 */
    d->n_recombinations_core = round(d->n_core_mutations * p->HR_rate);      /* :279 */
    d->n_recombinations_pan_total = round(d->n_core_mutations * p->HGT_rate);/* :280 */
    d->num_gene1_sites = (size_t)round((double)d->pan_size * (1.0 - p->prop_genes2));   /* :334 */
    d->num_gene2_sites = d->pan_size - d->num_gene1_sites;
    double prop1 = (double)d->num_gene1_sites / (double)d->pan_size;         /* :336 */
    double prop2 = 1.0 - prop1;
    size_t c = 0;
    if (d->num_gene1_sites > 0) {                                            /* :341-352 */
        d->comp_lo[c] = 0; d->comp_hi[c] = (uint32_t)d->num_gene1_sites;
        d->n_pan_mutations[c] = p->rate_genes1 * (double)d->num_gene1_sites;
        d->n_recombinations_pan[c] = d->n_recombinations_pan_total * prop1;
        c++;
    }
    if (d->num_gene1_sites < d->pan_size) {                                  /* :355-367 */
        d->comp_lo[c] = (uint32_t)d->num_gene1_sites; d->comp_hi[c] = (uint32_t)d->pan_size;
        d->n_pan_mutations[c] = p->rate_genes2 * (double)d->num_gene2_sites;
        d->n_recombinations_pan[c] = d->n_recombinations_pan_total * prop2;
        c++;
    }
    d->n_compartments = c;
}

/* main.rs:289-319 */
void ora_selection_coefficients(const ora_params *p, size_t pan_size, ora_rng *rng, double *out)
{
    for (size_t i = 0; i < pan_size; i++) out[i] = 0.0;                       /* :287 */
    if (p->prop_positive >= 0.0) {                                           /* :292 */
        for (size_t i = 0; i < pan_size; i++) {
            double weight = ora_rng_f64(rng);                                /* :298 */
            double s;
            if (weight <= p->prop_positive) {
                s = ora_exponential(rng, p->pos_lambda);                     /* :304 */
            } else {
                s = ora_exponential(rng, p->neg_lambda);
                while (s > 1.0) s = ora_exponential(rng, p->neg_lambda);     /* :309-311 */
                s = -1.0 * s;
            }
            out[i] = s;
        }
    }
}

/* main.rs:413-427 */
void ora_sample_pairs(size_t pop_size, size_t max_distances, ora_rng *rng,
                      uint32_t *range1, uint32_t *range2)
{
    for (size_t k = 0; k < max_distances; k++) range1[k] = (uint32_t)ora_rng_below(rng, pop_size);
    for (size_t k = 0; k < max_distances; k++) {
        uint32_t entry = (uint32_t)ora_rng_below(rng, pop_size - 1);
        if (entry >= range1[k]) entry += 1;
        range2[k] = entry;
    }
}

/* ======================================================================= */
/* Rust `{}` for f64                                                        */
/* ======================================================================= */
int ora_fmt_f64(char *buf, double x)
{
    if (isnan(x)) return sprintf(buf, "NaN");
    if (isinf(x)) return sprintf(buf, x < 0 ? "-inf" : "inf");
    if (x == 0.0) return sprintf(buf, signbit(x) ? "-0" : "0");
    char tmp[64];
    int prec;
    for (prec = 1; prec <= 17; prec++) {
        snprintf(tmp, sizeof tmp, "%.*e", prec - 1, x);
        if (strtod(tmp, NULL) == x) break;
    }
    /* tmp = [-]d.ddddde[+-]XX */
    char digits[32];
    int nd = 0, neg = 0;
    const char *s = tmp;
    if (*s == '-') { neg = 1; s++; }
    while (*s && *s != 'e') { if (*s != '.') digits[nd++] = *s; s++; }
    int exp10 = atoi(s + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--;     /* shortest form */
    char *o = buf;
    if (neg) *o++ = '-';
    if (exp10 >= 0) {
        int int_digits = exp10 + 1;
        for (int i = 0; i < int_digits; i++) *o++ = (i < nd) ? digits[i] : '0';
        if (nd > int_digits) {
            *o++ = '.';
            for (int i = int_digits; i < nd; i++) *o++ = digits[i];
        }
    } else {
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -exp10 - 1; i++) *o++ = '0';
        for (int i = 0; i < nd; i++) *o++ = digits[i];
    }
    *o = 0;
    return (int)(o - buf);
}

/* ======================================================================= */
/* main.rs driver                                                           */
/* ======================================================================= */
static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

static double median_of(const double *v, size_t n)
{
    double *c = (double *)malloc(sizeof(double) * n);
    memcpy(c, v, sizeof(double) * n);
    qsort(c, n, sizeof(double), cmp_double);
    double m = (n & 1) ? c[n / 2] : 0.5 * (c[n / 2 - 1] + c[n / 2]);
    free(c);
    return m;
}

static double now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    ora_derived d;
    ora_rng rng;
    double *selection;
    ora_population core, pan;
    ora_site_dist core_dist[1], pan_dist[2];
    float *core_table, *pan_table[2];
    uint32_t *range1, *range2;
    double *avgdist;
    uint32_t *parents;
} run_state;

static int run_setup(run_state *st, const ora_params *p, int use_tables)
{
    memset(st, 0, sizeof(*st));
    ora_derive(p, &st->d);
    ora_rng_seed(&st->rng, p->seed);                                        /* main.rs:289 */
    size_t G = st->d.pan_size, N = p->pop_size;
    st->selection = (double *)malloc(sizeof(double) * (G ? G : 1));
    ora_selection_coefficients(p, G, &st->rng, st->selection);
    /* main.rs:370 sample_beta: drawn and never used -> not replicated */
    if (ora_population_new(&st->core, N, p->core_size, 4, 1, st->d.avg_gene_freq_adj, &st->rng, p->core_genes)) return -1;
    if (ora_population_new(&st->pan, N, G, 2, 0, st->d.avg_gene_freq_adj, &st->rng, p->core_genes)) return -1;
    st->core_dist[0].lo = 0; st->core_dist[0].hi = (uint32_t)p->core_size;
    if (use_tables) {
        st->core_table = ora_build_cumulative(p->core_size, 0, (uint32_t)p->core_size);
        st->core_dist[0].cumulative = st->core_table; st->core_dist[0].n_table = p->core_size;
    }
    for (size_t c = 0; c < st->d.n_compartments; c++) {
        st->pan_dist[c].lo = st->d.comp_lo[c]; st->pan_dist[c].hi = st->d.comp_hi[c];
        if (use_tables) {
            st->pan_table[c] = ora_build_cumulative(G, st->d.comp_lo[c], st->d.comp_hi[c]);
            st->pan_dist[c].cumulative = st->pan_table[c]; st->pan_dist[c].n_table = G;
        }
    }
    st->range1 = (uint32_t *)malloc(sizeof(uint32_t) * p->max_distances);
    st->range2 = (uint32_t *)malloc(sizeof(uint32_t) * p->max_distances);
    ora_sample_pairs(N, p->max_distances, &st->rng, st->range1, st->range2);
    st->avgdist = (double *)malloc(sizeof(double) * N);
    st->parents = (uint32_t *)malloc(sizeof(uint32_t) * N);
    return 0;
}

static void run_teardown(run_state *st)
{
    free(st->selection); ora_population_free(&st->core); ora_population_free(&st->pan);
    free(st->core_table); free(st->pan_table[0]); free(st->pan_table[1]);
    free(st->range1); free(st->range2); free(st->avgdist); free(st->parents);
}

/* one generation, main.rs:435-464 */
static int run_generation(run_state *st, const ora_params *p, int j)
{
    size_t N = p->pop_size;
    for (size_t i = 0; i < N; i++) st->avgdist[i] = 1.0;                    /* :435 */
    if (p->competition_strength > 0.0) ora_average_distance(&st->pan, st->avgdist);  /* :438-440 */
    if (ora_sample_indices(&st->pan, &st->rng, st->d.avg_gene_num, st->avgdist, st->selection,
                           p->no_control_genome_size, p->genome_size_penalty,
                           p->competition_strength, st->parents)) return -1; /* :442-443 */
    if (ora_next_generation(&st->core, st->parents, N)) return -1;          /* :445 */
    if (ora_next_generation(&st->pan, st->parents, N)) return -1;           /* :447 */
    ora_mutate_alleles(&st->core, &st->d.n_core_mutations, st->core_dist, 1, p->seed, (uint64_t)j, NULL);  /* :452 */
    ora_mutate_alleles(&st->pan, st->d.n_pan_mutations, st->pan_dist, st->d.n_compartments, p->seed, (uint64_t)j, NULL); /* :455 */
    if (p->HR_rate > 0.0)                                                   /* :459-461 */
        if (ora_recombine(&st->core, &st->d.n_recombinations_core, st->core_dist, 1, &st->rng, p->seed, (uint64_t)j, NULL)) return -1;
    if (p->HGT_rate > 0.0)                                                  /* :462-464 */
        if (ora_recombine(&st->pan, st->d.n_recombinations_pan, st->pan_dist, st->d.n_compartments, &st->rng, p->seed, (uint64_t)j, NULL)) return -1;
    return 0;
}

int ora_run(const ora_params *p, const char *outpref, int use_tables, ora_summary *summary)
{
    if (ora_validate(p)) return 1;              /* reference prints and returns Ok(()) */
    ora_set_threads(p->threads);                /* main.rs:249-257 */
    run_state st;
    if (run_setup(&st, p, use_tables)) return -1;
    size_t N = p->pop_size, G = st.d.pan_size, P = p->max_distances;
    char path[4096], num[512];

    if (p->print_selection && outpref) {                                    /* main.rs:321-331 */
        snprintf(path, sizeof path, "%s_selection.tsv", outpref);
        FILE *f = fopen(path, "w");
        if (f) {
            for (size_t i = 0; i < G; i++) { ora_fmt_f64(num, st.selection[i]); fprintf(f, "%s%s", i ? "\n" : "", num); }
            fprintf(f, "\n");
            fclose(f);
        }
    }

    double *avg_core = (double *)calloc(p->n_gen, sizeof(double)), *avg_acc = (double *)calloc(p->n_gen, sizeof(double));
    double *std_core = (double *)calloc(p->n_gen, sizeof(double)), *std_acc = (double *)calloc(p->n_gen, sizeof(double));
    double *cd = (double *)malloc(sizeof(double) * P), *ad = (double *)malloc(sizeof(double) * P);

    for (int j = 0; j < p->n_gen; j++) {                                     /* main.rs:429 */
        if (run_generation(&st, p, j)) { run_teardown(&st); return -1; }
        if (j == p->n_gen - 1) {                                            /* :467-499 */
            ora_pairwise_distances(&st.core, P, st.range1, st.range2, cd);
            ora_pairwise_distances(&st.pan, P, st.range1, st.range2, ad);
            double *freqs = (double *)malloc(sizeof(double) * (G + p->core_genes + 1));
            ora_gene_frequencies(&st.pan, freqs);
            if (outpref) {
                snprintf(path, sizeof path, "%s.tsv", outpref);
                FILE *f = fopen(path, "w");
                if (f) {
                    for (size_t k = 0; k < P; k++) {
                        char a[512], b[512];
                        ora_fmt_f64(a, cd[k]); ora_fmt_f64(b, ad[k]);
                        fprintf(f, "%s\t%s\n", a, b);                        /* :481 */
                    }
                    fclose(f);
                }
                snprintf(path, sizeof path, "%s_freqs.txt", outpref);
                f = fopen(path, "w");
                if (f) {
                    for (size_t k = 0; k < G + p->core_genes; k++) { ora_fmt_f64(num, freqs[k]); fprintf(f, "%s\n", num); }
                    fclose(f);
                }
            }
            if (summary) {
                memset(summary, 0, sizeof(*summary));
                ora_standard_deviation(cd, P, &summary->std_core, &summary->mean_core);
                ora_standard_deviation(ad, P, &summary->std_acc, &summary->mean_acc);
                summary->median_core = median_of(cd, P);
                summary->median_acc = median_of(ad, P);
                double s = 0.0; size_t lt = 0, gt = 0;
                for (size_t k = 0; k < G; k++) { s += freqs[k]; if (freqs[k] < 0.1) lt++; if (freqs[k] > 0.9) gt++; }
                summary->mean_gene_freq = G ? s / (double)G : 0.0;
                summary->frac_freq_lt_01 = G ? (double)lt / (double)G : 0.0;
                summary->frac_freq_gt_09 = G ? (double)gt / (double)G : 0.0;
                summary->mean_genes_per_row = summary->mean_gene_freq * (double)G;
            }
            free(freqs);
        }
        if (p->print_dist) {                                                /* :502-519 */
            ora_pairwise_distances(&st.core, P, st.range1, st.range2, cd);
            ora_pairwise_distances(&st.pan, P, st.range1, st.range2, ad);
            ora_standard_deviation(cd, P, &std_core[j], &avg_core[j]);
            ora_standard_deviation(ad, P, &std_acc[j], &avg_acc[j]);
        }
        if (p->verbose) {                                                   /* :522-526 */
            printf("Finished gen: %d\n", j + 1);
            ora_fmt_f64(num, ora_calc_gene_freq(&st.pan));
            printf("avg_gene_freq: %s\n", num);
        }
    }

    if (p->print_dist && outpref) {                                         /* :531-548 */
        snprintf(path, sizeof path, "%s_per_gen.tsv", outpref);
        FILE *f = fopen(path, "w");
        if (f) {
            for (int j = 0; j < p->n_gen; j++) {
                char a[512], b[512], c[512], d[512];
                ora_fmt_f64(a, avg_core[j]); ora_fmt_f64(b, std_core[j]);
                ora_fmt_f64(c, avg_acc[j]); ora_fmt_f64(d, std_acc[j]);
                fprintf(f, "%s\t%s\t%s\t%s\n", a, b, c, d);                  /* :546 */
            }
            fclose(f);
        }
    }

    if (p->print_matrices && outpref) {                                     /* :550-553, population.rs:865-897 */
        snprintf(path, sizeof path, "%s_core_genome.csv", outpref);
        FILE *f = fopen(path, "w");
        if (f) {
            for (size_t i = 0; i < N; i++) {
                const uint8_t *row = st.core.pop + i * st.core.ncols;
                for (size_t k = 0; k < st.core.ncols; k++) { if (k) fputc(',', f); fputc(ora_int_to_base(row[k]), f); }
                fputc('\n', f);
            }
            fclose(f);
        }
        snprintf(path, sizeof path, "%s_pangenome.csv", outpref);
        f = fopen(path, "w");
        if (f) {
            for (size_t i = 0; i < N; i++) {
                const uint8_t *row = st.pan.pop + i * G;
                int first = 1;
                for (size_t k = 0; k < p->core_genes; k++) { if (!first) fputc(',', f); fputc('1', f); first = 0; }   /* :891 */
                for (size_t k = 0; k < G; k++) { if (!first) fputc(',', f); fputc('0' + row[k], f); first = 0; }
                fputc('\n', f);
            }
            fclose(f);
        }
    }

    free(avg_core); free(avg_acc); free(std_core); free(std_acc); free(cd); free(ad);
    run_teardown(&st);
    return 0;
}

double ora_time_generations(const ora_params *p, int n_gen, int with_distances,
                            int use_tables, double *dist_seconds)
{
    ora_set_threads(p->threads);
    run_state st;
    if (run_setup(&st, p, use_tables)) return -1.0;
    size_t P = p->max_distances;
    double *cd = (double *)malloc(sizeof(double) * P), *ad = (double *)malloc(sizeof(double) * P);
    double tg = 0.0, td = 0.0;
    for (int j = 0; j < n_gen; j++) {
        double t0 = now_seconds();
        if (run_generation(&st, p, j)) { tg = -1.0; break; }
        double t1 = now_seconds();
        tg += t1 - t0;
        if (with_distances) {
            ora_pairwise_distances(&st.core, P, st.range1, st.range2, cd);
            ora_pairwise_distances(&st.pan, P, st.range1, st.range2, ad);
            td += now_seconds() - t1;
        }
    }
    if (dist_seconds) *dist_seconds = td;
    free(cd); free(ad);
    run_teardown(&st);
    return tg;
}
