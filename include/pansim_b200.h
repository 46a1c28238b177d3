/*
 * pansim_b200.h -- C ABI of libpansim_b200.so: the B200 (sm_100a) replacement
 * for Pansim's per-generation Wright-Fisher step and pairwise distance pass.
 *
 * The reference has no plugin/FFI layer: the boundary it exposes is the public
 * Rust API of `pansim::population::Population` as `main.rs` drives it
 * (pansim/src/main.rs:429-528). Each entry point below names the reference
 * interface it replaces (file:line under /root/reference/pansim/src/).
 * One opaque context holds BOTH populations (core alignment + accessory
 * presence/absence), because the reference always moves them together with
 * the same parent vector (main.rs:442-464).
 *
 * Conventions
 *  - every pointer argument is caller-owned HOST memory valid for the call,
 *    except arguments whose name starts with `d_` (CUDA device pointers on the
 *    context's device, used by the multi-GPU host plumbing);
 *  - return value 0 = ok, <0 = error (message: pansim_last_error); nothing
 *    unwinds across the boundary (the reference panics via unwrap());
 *  - a context is single-owner and not thread-safe (all Population methods
 *    are called sequentially from the main thread, main.rs:429-528);
 *  - host byte layouts are the reference's: core = one byte per site, one-hot
 *    {1,2,4,8} = A,C,G,T (population.rs:154-162, 201-204); accessory = one
 *    byte per gene in {0,1} (population.rs:214-219); matrices are C-order
 *    [pop_size x columns] (population.rs:164-178).
 *  - there is NO CPU fallback: every call fails with PANSIM_ERR_CUDA when no
 *    sm_100-class device is usable.
 */
#ifndef PANSIM_B200_H
#define PANSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PANSIM_OK              0
#define PANSIM_ERR_INVALID    (-1)   /* bad argument / unsupported parameter        */
#define PANSIM_ERR_CUDA       (-2)   /* CUDA runtime failure or no usable device    */
#define PANSIM_ERR_NOMEM      (-3)
#define PANSIM_ERR_WEIGHTS    (-4)   /* WeightedIndex::new would panic (population.rs:440) */
#define PANSIM_ERR_STATE      (-5)   /* call order (e.g. step before set_initial)   */

/* Column shards must start on a multiple of this many core sites (one 2 KiB
 * region of the 2-bit packed row; the RNG is keyed per region lane, so results
 * do not depend on where the shard boundaries are). */
#define PANSIM_SITE_ALIGN 8192u

typedef struct pansim_ctx pansim_ctx;

/* The numbers main.rs:259-287 and :333-367 derive from the command line; they
 * are the kernel parameters. Fill with pansim_config_init() then override. */
typedef struct {
    uint32_t struct_size;          /* = sizeof(pansim_config)                         */
    int32_t  device;               /* CUDA device ordinal                             */
    uint32_t pop_size;             /* N   (main.rs:155-156)                           */
    uint32_t pan_size;             /* G = pan_genes - core_genes (main.rs:259)        */
    uint64_t core_size;            /* L, whole core alignment (main.rs:157-158)       */
    uint64_t site_begin, site_end; /* column shard [begin,end) of the core held by
                                      this context; 0,0 = whole alignment. begin must
                                      be a multiple of PANSIM_SITE_ALIGN. The accessory
                                      matrix is replicated in every context.          */
    uint32_t core_genes;           /* population.rs:168; added to both Jaccard terms  */
    uint32_t n_compartments;       /* 0..2 (main.rs:341-367)                          */
    uint32_t comp_lo[2], comp_hi[2];/* genes with weight 1.0 in compartment c         */
    double   core_mut_mean;        /* n_core_mutations[0]      (main.rs:275-276)      */
    double   hr_mean;              /* n_recombinations_core[0] (main.rs:279); 0 = off */
    double   acc_mut_mean[2];      /* n_pan_mutations[c]       (main.rs:348, 361)     */
    double   hgt_mean[2];          /* n_recombinations_pan[c]  (main.rs:349-366); 0 = off */
    int32_t  avg_gene_num;         /* main.rs:272                                     */
    int32_t  no_control_genome_size;
    double   genome_size_penalty;
    double   competition_strength;
    uint64_t seed;                 /* Philox key (north_star: counter-based RNG keyed
                                      by seed, generation, individual, site block)    */
} pansim_config;

void pansim_config_init(pansim_config *cfg);

/* ---- lifecycle --------------------------------------------------------- */
/* replaces Population::new x2 (population.rs:181-242) minus the random draws,
 * which the host keeps (pansim_set_initial). */
int  pansim_create(const pansim_config *cfg, pansim_ctx **out);
void pansim_destroy(pansim_ctx *ctx);
/* message of the last failure on ctx (ctx == NULL: last pansim_create failure) */
const char *pansim_last_error(const pansim_ctx *ctx);
/* library build info, e.g. "pansim_b200 0.1 sm_100a" */
const char *pansim_version(void);

/* ---- state in / out ---------------------------------------------------- */
/* population.rs:199-230: all N rows start as one row. core_row_onehot is the
 * WHOLE alignment row [core_size]; the context keeps its [site_begin,site_end)
 * slice. acc_row is [pan_size]. */
int pansim_set_initial(pansim_ctx *ctx, const uint8_t *core_row_onehot, const uint8_t *acc_row);
/* arbitrary state (tests, restarts): core [N x local_sites], acc [N x G] */
int pansim_upload_core(pansim_ctx *ctx, const uint8_t *core_onehot);
int pansim_upload_acc(pansim_ctx *ctx, const uint8_t *acc);
/* population.rs:865-897 `write` consumes exactly these matrices */
int pansim_download_core(pansim_ctx *ctx, uint8_t *core_onehot_out);
int pansim_download_acc(pansim_ctx *ctx, uint8_t *acc_out);
/* `A,C,G,T\n` text rows of _core_genome.csv (population.rs:877-879) expanded on
 * the GPU: out holds rows [row_begin,row_end), each 2*local_sites bytes. */
int pansim_export_core_csv(pansim_ctx *ctx, uint32_t row_begin, uint32_t row_end, char *out);
/* the whole _core_genome.csv (population.rs:865-882) written by the library: text expanded on the GPU,
 * streamed through two pinned host chunks (kernel + copy of chunk k+1 overlap the write of chunk k) */
int pansim_write_core_csv(pansim_ctx *ctx, const char *path, uint64_t *bytes_out);
/* selection coefficients s[G] (main.rs:287-319), used as ln(1 + s_j) */
int pansim_set_selection(pansim_ctx *ctx, const double *s);

/* ---- per-generation operators ------------------------------------------ */
/* Population::average_distance on the accessory population
 * (population.rs:753-784 + get_distance :114-151). Result stays on the device
 * for the next pansim_sample_indices; out[N] may be NULL. */
int pansim_average_distance(pansim_ctx *ctx, double *out);

/* Population::sample_indices (population.rs:270-448). avg_pairwise_dists[N]:
 * host vector as main.rs:435-443 passes it; NULL = use the device copy left by
 * pansim_average_distance, or all 1.0 if competition_strength <= 0
 * (main.rs:435-440). parents_out[N] may be NULL. The N draws come from
 * Philox(seed, gen, individual). */
int pansim_sample_indices(pansim_ctx *ctx, uint32_t gen, const double *avg_pairwise_dists,
                          uint32_t *parents_out);
/* main.rs:435-443 in one call: average_distance when competition_strength > 0 (else all 1.0), then
 * sample_indices; both vectors come back behind a single synchronisation. Either output may be NULL. */
int pansim_select_parents(pansim_ctx *ctx, uint32_t gen, double *avg_pairwise_dists_out, uint32_t *parents_out);
/* the intermediate vectors of population.rs:282-437 for parity checks; any may
 * be NULL. Valid after pansim_sample_indices. */
int pansim_get_weights(pansim_ctx *ctx, double *weights, int32_t *num_genes, double *logfit);

/* main.rs:445-464 with parents supplied by the host: next_generation x2 +
 * mutate_alleles x2 + recombine x2 as ONE fused device pass per population
 * (generate mode: events drawn from Philox keyed by seed/gen/row/site-block).
 * The call returns once `parents` has been consumed and the step is enqueued:
 * the core pass keeps running on the device beside the selection calls of the
 * next generation. Every entry point that reads or writes the core alignment
 * (pair counts, downloads, CSV export, uploads, replay steps, timing) waits for
 * it first, and applies the recombination events the pass defers to the next
 * generation, so callers always observe the state of
 * population.rs after recombine(). */
int pansim_step_with_parents(pansim_ctx *ctx, uint32_t gen, const uint32_t *parents);
/* whole generation main.rs:435-464 on the device: competition (if
 * competition_strength > 0) -> fitness -> parents -> fused step. */
int pansim_step(pansim_ctx *ctx, uint32_t gen);
/* n consecutive generations gen0..gen0+n-1 without returning to the host */
int pansim_run_generations(pansim_ctx *ctx, uint32_t gen0, uint32_t n);
/* Population::next_generation alone (population.rs:450-465), both populations */
int pansim_next_generation(pansim_ctx *ctx, const uint32_t *parents);
/* parents of the last step (device -> host) */
int pansim_get_parents(pansim_ctx *ctx, uint32_t *parents_out);

/* Replay mode: apply an explicit event list with the reference's store
 * semantics (bit-exact). Lists are flat and in APPLY ORDER, i.e. the order the
 * reference executes `row[site] = allele` (population.rs:508, 537) and
 * `self.pop[[row_idx, col_idx]] = value` (population.rs:745): later entries
 * overwrite earlier ones. Sites/loci are global core coordinates; a sharded
 * context applies the ones that fall in its slice. */
typedef struct {
    const uint32_t *parents;          /* [N]                                          */
    size_t n_core_mut;                /* population.rs:525-538                        */
    const uint32_t *core_mut_row, *core_mut_site;
    const uint8_t  *core_mut_allele;  /* one-hot {1,2,4,8}                            */
    size_t n_acc_flip;                /* population.rs:501-509                        */
    const uint32_t *acc_flip_row, *acc_flip_gene;
    size_t n_hr;                      /* population.rs:728-748 (core)                 */
    const uint32_t *hr_recipient, *hr_locus;
    const uint8_t  *hr_value;         /* snapshotted donor allele, one-hot            */
    size_t n_hgt;                     /* population.rs:728-748 (accessory), value 1   */
    const uint32_t *hgt_recipient, *hgt_gene;
} pansim_events;
int pansim_step_replay(pansim_ctx *ctx, const pansim_events *ev);

/* ---- distance pass / reductions ---------------------------------------- */
/* Population::pairwise_distances for both populations (population.rs:787-837,
 * distances.rs:22-77). Integers out; the host does the f64 divisions of
 * population.rs:822 and :828-830 so text output is bit-identical:
 *   core_diff[k] = hamming_bitwise_fast(row_i,row_j)/2   (differing sites)
 *   inter[k], uni[k] = jaccard_distance_fast(row_i,row_j)
 * Any output may be NULL. In a sharded context core_diff is the partial count
 * over the context's sites (sum over shards = whole count). */
int pansim_pair_counts(pansim_ctx *ctx, const uint32_t *range1, const uint32_t *range2,
                       size_t n_pairs, uint32_t *core_diff, uint32_t *inter, uint32_t *uni);
/* same, outputs left in caller-provided DEVICE buffers (u32[n_pairs]) for an
 * NCCL all-reduce by the host plumbing; returns after the kernels finished. */
int pansim_pair_counts_device(pansim_ctx *ctx, const uint32_t *range1, const uint32_t *range2,
                              size_t n_pairs, void *d_core_diff, void *d_inter, void *d_uni);
/* Exact all-pairs mode (an extension: the reference only samples max_distances
 * pairs with replacement, main.rs:413-427; BASELINE config 5 asks for all
 * N(N-1)/2). Counts for every pair (i, j) with row_begin <= i < row_end and
 * i < j < N, in (i ascending, j ascending) order; the pair list and its plan are
 * generated on the device. n_pairs_out = sum over i of (N-1-i), at most 2^31-1
 * per call (walk the rows in blocks). Outputs may be NULL. */
int pansim_pair_counts_rows(pansim_ctx *ctx, uint32_t row_begin, uint32_t row_end, uint32_t *core_diff,
                            uint32_t *inter, uint32_t *uni, size_t *n_pairs_out);
/* same, outputs left in caller-provided DEVICE buffers (for the NCCL all-reduce
 * of column-sharded contexts) */
int pansim_pair_counts_rows_device(pansim_ctx *ctx, uint32_t row_begin, uint32_t row_end, void *d_core_diff,
                                   void *d_inter, void *d_uni, size_t *n_pairs_out);
/* The per-generation statistics of --print_dist (main.rs:502-519) computed on the device: the
 * distance pass above followed by standard_deviation (population.rs:87-94) of both distance vectors.
 * out[4] = { avg_core, std_core, avg_acc, std_acc }, the columns of _per_gen.tsv (main.rs:546). The
 * four f64 sums are taken strictly left to right like `iter().sum::<f64>()`, so the values are the
 * reference's bit for bit; 32 bytes leave the device instead of three count vectors. */
int pansim_pair_stats(pansim_ctx *ctx, const uint32_t *range1, const uint32_t *range2, size_t n_pairs, double *out);
/* n generations gen0..gen0+n-1, each followed by that pass and its statistics (the loop of
 * main.rs:429-519 under --print_dist), without returning to the host: stats_out[n][4]. */
int pansim_run_generations_stats(pansim_ctx *ctx, uint32_t gen0, uint32_t n, const uint32_t *range1,
                                 const uint32_t *range2, size_t n_pairs, double *stats_out);
/* the two f64 formulas (population.rs:822, :828-830), exported so every host
 * language forms the distances identically */
double pansim_core_distance(uint32_t core_diff, uint64_t core_size);
double pansim_acc_distance(uint32_t inter, uint32_t uni, uint32_t core_genes);

/* Population::gene_frequencies numerators (population.rs:840-856): counts[G] */
int pansim_gene_counts(pansim_ctx *ctx, uint32_t *counts);

/* ---- multi-GPU: column shards of one alignment --------------------------
 * (the reference is one shared-memory process, population.rs has no analogue.)
 * Every shard context holds all individuals for [site_begin, site_end) and a replica of the
 * accessory matrix; parents, flips and HGT are recomputed identically everywhere from the same
 * counters, so the generation step needs no exchange. The distance pass has one: the per-pair partial
 * core counts are summed over the shards with NCCL inside the library.
 *
 * (a) one process per GPU (torchrun, MPI): rank 0 calls pansim_comm_unique_id, the host plumbing
 *     broadcasts the 128 bytes, every rank calls pansim_comm_init_rank on its context. From then on
 *     pansim_pair_counts / _device / _rows / _rows_device / pansim_pair_stats return WHOLE-alignment
 *     core counts on every rank (ncclAllReduce on the context's stream before the read-back).
 * (b) one process, several GPUs: pansim_group_* below (ncclCommInitAll). */
#define PANSIM_COMM_ID_BYTES 128
int pansim_comm_unique_id(void *id_out);
int pansim_comm_init_rank(pansim_ctx *ctx, int n_ranks, int rank, const void *id);
int pansim_comm_info(pansim_ctx *ctx, int *n_ranks, int *rank);

/* (b) one process, several GPUs. The configuration describes the WHOLE alignment (site_begin =
 *     site_end = 0, `device` ignored); devices == NULL means 0..n_devices-1. Calls mirror the
 *     single-context ones and cite the same reference lines; each enqueues on every shard before it
 *     synchronises, so a single host thread keeps all devices busy. */
typedef struct pansim_group pansim_group;
/* column shard [begin, end) that shard i of n holds (whole 8192-site regions; needs no device) */
int  pansim_shard_bounds(uint64_t core_size, int n_shards, int shard, uint64_t *site_begin, uint64_t *site_end);
int  pansim_group_create(const pansim_config *cfg, int n_devices, const int *devices, pansim_group **out);
void pansim_group_destroy(pansim_group *g);
const char *pansim_group_last_error(const pansim_group *g);   /* g == NULL: last pansim_group_create failure */
int  pansim_group_size(const pansim_group *g);
pansim_ctx *pansim_group_ctx(pansim_group *g, int shard);     /* borrowed; shard i holds sites [begin_i, end_i) */
int pansim_group_set_initial(pansim_group *g, const uint8_t *core_row_onehot, const uint8_t *acc_row);
int pansim_group_set_selection(pansim_group *g, const double *s);
int pansim_group_run_generations(pansim_group *g, uint32_t gen0, uint32_t n);            /* main.rs:435-464 x n */
int pansim_group_pair_counts(pansim_group *g, const uint32_t *range1, const uint32_t *range2, size_t n_pairs,
                             uint32_t *core_diff, uint32_t *inter, uint32_t *uni);      /* population.rs:787-837 */
int pansim_group_run_generations_stats(pansim_group *g, uint32_t gen0, uint32_t n, const uint32_t *range1,
                                       const uint32_t *range2, size_t n_pairs, double *stats_out);
/* Exact all-pairs mode over the group (extension, BASELINE config 5): row blocks of about chunk_pairs
 * pairs; per block the partial core counts are reduce-scattered over the shards while the next block
 * is being computed. cb is called once per block, in (i, j) order; the vectors are valid during the
 * call; a non-zero return stops the walk and is returned. */
typedef int (*pansim_pairs_cb)(void *user, uint32_t row_begin, uint32_t row_end, size_t n_pairs,
                               const uint32_t *core_diff, const uint32_t *inter, const uint32_t *uni);
int pansim_group_all_pairs(pansim_group *g, size_t chunk_pairs, pansim_pairs_cb cb, void *user);
/* last walk: wall time (callbacks included), device time of shard 0 inside the reduce-scatters, pairs walked */
int pansim_group_all_pairs_timing(pansim_group *g, float *wall_ms, float *nccl_ms, uint64_t *pairs);
int pansim_group_gene_counts(pansim_group *g, uint32_t *counts);                          /* population.rs:840-856 */
int pansim_group_download_acc(pansim_group *g, uint8_t *acc_out);
int pansim_group_download_core(pansim_group *g, uint8_t *core_onehot_out);               /* [N x core_size] */
int pansim_group_export_core_csv(pansim_group *g, uint32_t row_begin, uint32_t row_end, char *out);

/* ---- introspection / instrumentation ----------------------------------- */
typedef struct {
    uint64_t core_row_stride_bytes;   /* packed row pitch on the device              */
    uint64_t acc_row_stride_bytes;
    uint64_t local_sites;             /* site_end - site_begin                       */
    uint64_t core_state_bytes;        /* one of the two packed core buffers          */
    uint64_t algorithmic_bytes_per_generation;  /* 2*N*ceil(Ll/4) + 2*N*ceil(G/8)   */
    uint64_t algorithmic_bytes_per_pair;        /* 2*ceil(Ll/4) + 2*ceil(G/8)       */
    uint32_t sm_count;
    uint32_t core_step_grid, core_step_block, core_step_smem;
} pansim_info;
int pansim_get_info(pansim_ctx *ctx, pansim_info *out);

/* Device time (ms, CUDA events on the context's stream) of the last
 * pansim_step / pansim_step_with_parents / pansim_run_generations /
 * pansim_pair_counts* call, split by kernel group. */
typedef struct {
    float total_ms;
    float core_step_ms;       /* core genome: gather+SNP kernel and the recombination pass (sum over generations) */
    float acc_step_ms;
    float select_ms;          /* competition + fitness + parent draw               */
    float pair_core_ms, pair_acc_ms;
    uint32_t launches;        /* kernels launched by that call                     */
    float core_hr_ms;         /* the recombination pass alone (part of core_step_ms) */
} pansim_timing;
int pansim_get_timing(pansim_ctx *ctx, pansim_timing *out);
/* enable/disable per-kernel event timing (off = no extra events; default on) */
int pansim_set_timing(pansim_ctx *ctx, int enabled);

/* Generate-mode event dump for parity tests: when enabled, the next
 * pansim_step* call also records every event it draws, so the oracle can
 * replay them with the reference's semantics. seq = order among events of one
 * (row, site block); the per-cell winner is the highest seq. */
typedef struct {
    size_t n_core_mut;  uint32_t *core_mut_row, *core_mut_site, *core_mut_seq; uint8_t *core_mut_allele;
    size_t n_hr;        uint32_t *hr_recipient, *hr_locus, *hr_donor, *hr_seq;  uint8_t *hr_value;
    uint8_t *acc_flip_mask;   /* [N x G] 1 = flipped (population.rs:504-508 parity) */
    uint8_t *acc_gain_mask;   /* [N x G] 1 = received by HGT                        */
} pansim_event_dump;
int  pansim_enable_event_dump(pansim_ctx *ctx, size_t max_core_events);
/* arrays are malloc'd by the library; release with pansim_free_event_dump */
int  pansim_fetch_event_dump(pansim_ctx *ctx, pansim_event_dump *out);
void pansim_free_event_dump(pansim_event_dump *d);

/* per-cell probabilities the generate-mode kernels use (documentation of the
 * thinning); out has 4 doubles:
 * [0] core site overwritten by U{C,G,T} per generation, [1] core cell
 * overwritten by a donor allele, [2],[3] accessory flip prob. compartment 0,1 */
int pansim_get_rates(pansim_ctx *ctx, double *out);

#ifdef __cplusplus
}
#endif
#endif /* PANSIM_B200_H */
