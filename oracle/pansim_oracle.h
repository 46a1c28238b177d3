/*
 * pansim_oracle.h -- CPU restatement of bacpop/Pansim's per-generation
 * Wright-Fisher step and pairwise distance pass.
 *
 * THIS IS TEST INFRASTRUCTURE. It is the parity checker for the CUDA path and
 * the "port" CPU baseline of bench.py. Nothing in pansim_b200/ (the product)
 * may include, link or call it. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference (Rust) cannot be compiled in this image (no
 * cargo/rustc) and ships no tests, golden vectors or fixtures (SURVEY.md 4,
 * 8c).  The integer operators (gather, event apply, XOR/AND/OR popcount
 * distances, gene counts) have closed-form semantics and are pinned by the
 * hand-derived known-answer tests of SURVEY.md 4 (tests/test_oracle_kat.py).
 * The samplers live in third-party crates that are absent from
 * /root/reference and unpinned (Cargo.toml:10-17 semver ranges, no
 * Cargo.lock): rand 0.8.5 (StdRng = ChaCha12, Uniform, WeightedIndex,
 * shuffle), statrs 0.16 (Poisson, Exp), logsumexp 0.1.  For those the oracle
 * restates the published algorithm (exact Poisson / uniform / categorical /
 * Fisher-Yates samplers, streaming log-add-exp) over its own generator
 * (xoshiro256**), so only distributional fidelity is claimed there.
 *
 * Data layout is the reference's: one byte per core site, one-hot {1,2,4,8}
 * (population.rs:201-204), one byte per accessory gene {0,1}
 * (population.rs:214-219), C-order [nrows x ncols] (population.rs:164-178).
 *
 * All file:line citations are relative to /root/reference/pansim/src/.
 */
#ifndef PANSIM_ORACLE_H
#define PANSIM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ RNG -- */
/* xoshiro256** seeded through splitmix64. Stands in for rand::StdRng
 * (main.rs:289) and for rand::thread_rng() (population.rs:493,517,596). */
typedef struct { uint64_t s[4]; } ora_rng;

void     ora_rng_seed(ora_rng *r, uint64_t seed);
/* independent stream for (seed, a, b, c): replaces the OS-seeded per-thread
 * generators of the reference so that the oracle is reproducible. */
void     ora_rng_seed4(ora_rng *r, uint64_t seed, uint64_t a, uint64_t b, uint64_t c);
uint64_t ora_rng_next(ora_rng *r);
double   ora_rng_f64(ora_rng *r);                 /* [0,1), 53 bits          */
uint64_t ora_rng_below(ora_rng *r, uint64_t n);   /* exact uniform on [0,n)  */
uint64_t ora_poisson(ora_rng *r, double mean);    /* exact Poisson(mean)     */
double   ora_exponential(ora_rng *r, double rate);

/* ------------------------------------------------- distances.rs:22-77 ---- */
uint32_t ora_hamming_bitwise_fast(const uint8_t *x, const uint8_t *y, size_t n);
void     ora_jaccard_distance_fast(const uint8_t *x, const uint8_t *y, size_t n,
                                   uint32_t *intersection, uint32_t *uni);
/* population.rs:32-48 (commented-out naive cross-check) */
void     ora_jaccard_distance_naive(const uint8_t *x, const uint8_t *y, size_t n,
                                    uint32_t *intersection, uint32_t *uni);

/* --------------------------------------------- population.rs:83-94,154 --- */
void ora_standard_deviation(const double *v, size_t n, double *std_out, double *mean_out);
char ora_int_to_base(uint8_t n);

/* ------------------------------------------------ struct Population ------ */
typedef struct {
    uint8_t *pop;          /* [nrows x ncols], C order                     */
    size_t   nrows, ncols;
    int      core;         /* population.rs:166                            */
    size_t   core_genes;   /* population.rs:168                            */
    double   avg_gene_freq;
} ora_population;

/* population.rs:181-242. Returns 0 on success. */
int  ora_population_new(ora_population *p, size_t size, size_t allele_count,
                        uint8_t max_variants, int core, double avg_gene_freq,
                        ora_rng *rng, size_t core_genes);
void ora_population_free(ora_population *p);

/* population.rs:244-268 */
double ora_calc_gene_freq(const ora_population *p);

/* population.rs:282-437: the weight vector handed to WeightedIndex.
 * weights_out[N]; num_genes_out[N] and logfit_out[N] optional (may be NULL).
 * avg_pairwise_dists[N] as main.rs:435-440 passes it. */
void ora_selection_weights(const ora_population *p, int32_t avg_gene_num,
                           const double *avg_pairwise_dists,
                           const double *selection_coefficients,
                           int no_control_genome_size, double genome_size_penalty,
                           double competition_strength,
                           double *weights_out, int32_t *num_genes_out,
                           double *logfit_out);

/* WeightedIndex<f64>::new + sample (rand 0.8.5; population.rs:440-443):
 * cumulative sums, u ~ U[0,total), index = #cumulative[0..n-1) <= u.
 * Returns -1 if a weight is negative/NaN or the total is not > 0 (the
 * reference panics there). */
int  ora_weighted_index_sample(const double *weights, size_t n, ora_rng *rng,
                               size_t n_draws, uint32_t *out);

/* population.rs:270-448 == ora_selection_weights + ora_weighted_index_sample */
int  ora_sample_indices(const ora_population *p, ora_rng *rng, int32_t avg_gene_num,
                        const double *avg_pairwise_dists,
                        const double *selection_coefficients,
                        int no_control_genome_size, double genome_size_penalty,
                        double competition_strength, uint32_t *parents_out);

/* population.rs:450-465 */
int  ora_next_generation(ora_population *p, const uint32_t *sample, size_t n);

/* ---------------------------------------------------- event log ---------- */
/* Everything the apply step consumes for one generation, in the order
 * main.rs:445-464 consumes it. Flat lists in APPLY ORDER (what an
 * instrumented reference run would emit at population.rs:508, :537, :745). */
typedef struct {
    /* core SNPs: population.rs:525-538, row-major, draw order within a row */
    size_t n_core_mut, cap_core_mut;
    uint32_t *core_mut_row, *core_mut_site; uint8_t *core_mut_allele;
    /* accessory flips: population.rs:501-509 (both compartments) */
    size_t n_acc_flip, cap_acc_flip;
    uint32_t *acc_flip_row, *acc_flip_gene;
    /* HR: population.rs:728-748 on the core population, serial apply order */
    size_t n_hr, cap_hr;
    uint32_t *hr_recipient, *hr_locus, *hr_donor; uint8_t *hr_value;
    /* HGT: same loop on the accessory population (value == 1 always, :632) */
    size_t n_hgt, cap_hgt;
    uint32_t *hgt_recipient, *hgt_gene, *hgt_donor;
} ora_events;

void ora_events_init(ora_events *e);
void ora_events_clear(ora_events *e);   /* keep capacity */
void ora_events_free(ora_events *e);

/* Site sampler == one WeightedIndex<f32> over 0/1 weights that are 1 on
 * [lo,hi) (main.rs:284, 341-363, 394-403). When cumulative != NULL the draw
 * goes through a binary search in that cumulative f32 table like rand's
 * WeightedIndex::sample does (cost-faithful CPU baseline); otherwise it is the
 * equivalent direct uniform integer. */
typedef struct { uint32_t lo, hi; const float *cumulative; size_t n_table; } ora_site_dist;
float *ora_build_cumulative(size_t ncols, uint32_t lo, uint32_t hi); /* malloc'd */

/* population.rs:467-542. One entry of mutations_vec/dists per compartment.
 * rng_seed/gen key the per-row generators. ev may be NULL. Uses OpenMP over
 * rows where the reference uses rayon (population.rs:488-490, 512-514). */
void ora_mutate_alleles(ora_population *p, const double *mutations_vec,
                        const ora_site_dist *dists, size_t n_compartments,
                        uint64_t rng_seed, uint64_t gen, ora_events *ev);

/* population.rs:544-751. rng is the seeded main generator (shuffle, :726);
 * proposals use per-row generators keyed by (rng_seed, gen, row). */
int  ora_recombine(ora_population *p, const double *recombinations_vec,
                   const ora_site_dist *dists, size_t n_compartments,
                   ora_rng *rng, uint64_t rng_seed, uint64_t gen, ora_events *ev);

/* Replay: apply a flat event list with the reference's store semantics. */
void ora_apply_core_writes(ora_population *p, const uint32_t *row, const uint32_t *site,
                           const uint8_t *value, size_t n);    /* pop[[r,s]] = v  */
void ora_apply_acc_flips(ora_population *p, const uint32_t *row, const uint32_t *gene,
                         size_t n);                            /* pop[[r,g]] ^= 1 */
void ora_apply_acc_sets(ora_population *p, const uint32_t *row, const uint32_t *gene,
                        size_t n);                             /* pop[[r,g]] = 1  */
/* whole generation from an event list, order of main.rs:445-464 */
int  ora_step_replay(ora_population *core, ora_population *pan,
                     const uint32_t *parents, const ora_events *ev);

/* population.rs:753-784 + 114-151 */
void ora_average_distance(const ora_population *p, double *out);

/* population.rs:787-837: integer counts (core: h/2; accessory: inter, union)
 * and the f64 distances exactly as :817-830 forms them. Any output may be NULL. */
void ora_pair_counts(const ora_population *p, size_t max_distances,
                     const uint32_t *range1, const uint32_t *range2,
                     uint32_t *core_diff, uint32_t *inter, uint32_t *uni);
void ora_pairwise_distances(const ora_population *p, size_t max_distances,
                            const uint32_t *range1, const uint32_t *range2,
                            double *out);
/* the two f64 formulas alone (population.rs:822 and :828-830) */
double ora_core_distance_from_count(uint32_t core_diff, size_t ncols);
double ora_acc_distance_from_counts(uint32_t inter, uint32_t uni, size_t core_genes);

/* population.rs:840-863: out[ncols + core_genes] */
void ora_gene_frequencies(const ora_population *p, double *out);
void ora_gene_counts(const ora_population *p, uint32_t *out /* [ncols] */);

/* -------------------------------------------------- main.rs driver ------- */
typedef struct {        /* the 27 flags of main.rs:21-151 that affect results */
    size_t pop_size, core_size, pan_genes, core_genes;
    double avg_gene_freq, HR_rate, HGT_rate;
    int    n_gen;
    size_t max_distances;
    double core_mu, rate_genes1, rate_genes2, prop_genes2, prop_positive,
           pos_lambda, neg_lambda;
    uint64_t seed;
    int    print_dist, print_matrices, print_selection, verbose,
           no_control_genome_size;
    double genome_size_penalty, competition_strength;
    int    threads;
} ora_params;

void ora_params_default(ora_params *p);            /* main.rs default_value()s */

typedef struct {        /* main.rs:259-287, 333-367 */
    size_t pan_size;
    double avg_gene_freq_adj;
    int32_t avg_gene_num;
    double n_core_mutations;
    double n_recombinations_core;
    double n_recombinations_pan_total;
    size_t num_gene1_sites, num_gene2_sites;
    size_t n_compartments;            /* 0, 1 or 2 */
    uint32_t comp_lo[2], comp_hi[2];  /* gene ranges with weight 1.0          */
    double n_pan_mutations[2];
    double n_recombinations_pan[2];
} ora_derived;

/* returns 0, or the 1-based index of the validation block of main.rs:194-247
 * that rejects the parameters (the reference prints and exits 0 there). */
int  ora_validate(const ora_params *p);
void ora_derive(const ora_params *p, ora_derived *d);

/* main.rs:289-319: selection coefficients out[pan_size] */
void ora_selection_coefficients(const ora_params *p, size_t pan_size, ora_rng *rng, double *out);
/* main.rs:413-427 */
void ora_sample_pairs(size_t pop_size, size_t max_distances, ora_rng *rng,
                      uint32_t *range1, uint32_t *range2);

/* Rust `{}` Display of f64 (shortest round trip, never scientific). Returns
 * strlen. buf must hold >= 400 bytes. */
int  ora_fmt_f64(char *buf, double x);

/* Full run == main.rs:15-564. Writes the same files. If summary != NULL it
 * receives per-run scalars for the KS tests (see ora_summary). */
typedef struct {
    double mean_core, mean_acc, median_core, median_acc, std_core, std_acc;
    double mean_gene_freq;       /* over accessory genes only                */
    double frac_freq_lt_01, frac_freq_gt_09;
    double mean_genes_per_row;
} ora_summary;

int ora_run(const ora_params *p, const char *outpref /* NULL = no files */,
            int use_tables, ora_summary *summary);

/* Timing helper for the CPU baseline: runs `n_gen` generations (no output
 * files, optional per-generation distance pass) and returns seconds for the
 * generation loop alone; *dist_seconds gets the time of the distance passes. */
double ora_time_generations(const ora_params *p, int n_gen, int with_distances,
                            int use_tables, double *dist_seconds);

int  ora_max_threads(void);
void ora_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
