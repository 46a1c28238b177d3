// population.hpp -- C++ host-side mirror of the reference's `Population` API
// (pansim/src/population.rs) over the C ABI of include/pansim_b200.h.
//
// The reference keeps two Population objects (core_genome, pan_genome) that are
// always advanced together with the same parent vector (main.rs:442-464); here one
// `Populations` object owns both of them, on one GPU or column-sharded over several
// (pansim_group: one process, NCCL inside the library, `pansim --gpus N`). Method names, argument
// meaning and error behaviour follow population.rs; errors surface as
// pansim::Error (the reference panics via unwrap()).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pansim_b200.h"

namespace pansim {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// population.rs:87-94: returns (std, mean); population variance, sequential sums
inline std::pair<double, double> standard_deviation(const std::vector<double> &v)
{
    double sum = 0.0;
    for (double x : v) sum += x;
    const double mean = sum / (double)v.size();
    double ss = 0.0;
    for (double x : v) ss += (x - mean) * (x - mean);
    return {std::sqrt(ss / (double)v.size()), mean};
}

// population.rs:154-162
inline char int_to_base(uint8_t n)
{
    switch (n) { case 1: return 'A'; case 2: return 'C'; case 4: return 'G'; case 8: return 'T'; default: return 'N'; }
}

class Populations {
public:
    explicit Populations(const pansim_config &cfg, int n_gpus = 1) : cfg_(cfg)
    {
        std::vector<int> devs;
        for (int i = 0; i < n_gpus; i++) devs.push_back(cfg.device + i);
        const int rc = pansim_group_create(&cfg_, n_gpus, devs.data(), &grp_);
        if (rc != PANSIM_OK) throw Error(rc, pansim_group_last_error(nullptr));
        ctx_ = pansim_group_ctx(grp_, 0);
    }
    ~Populations() { pansim_group_destroy(grp_); }
    int n_gpus() const { return pansim_group_size(grp_); }
    Populations(const Populations &) = delete;
    Populations &operator=(const Populations &) = delete;

    // Population::new x2 (population.rs:181-242): every row starts as the given row
    void set_initial(const std::vector<uint8_t> &core_row_onehot, const std::vector<uint8_t> &acc_row)
    {
        gcheck(pansim_group_set_initial(grp_, core_row_onehot.data(), acc_row.data()));
    }
    void set_selection(const std::vector<double> &s) { gcheck(pansim_group_set_selection(grp_, s.data())); }

    // population.rs:753-784 (accessory population)
    std::vector<double> average_distance()
    {
        single("average_distance");
        std::vector<double> out(cfg_.pop_size);
        check(pansim_average_distance(ctx_, out.data()));
        return out;
    }
    // population.rs:270-448
    std::vector<uint32_t> sample_indices(uint32_t gen, const std::vector<double> &avg_pairwise_dists)
    {
        single("sample_indices");
        std::vector<uint32_t> out(cfg_.pop_size);
        check(pansim_sample_indices(ctx_, gen, avg_pairwise_dists.empty() ? nullptr : avg_pairwise_dists.data(), out.data()));
        return out;
    }
    // population.rs:450-465 for both populations
    void next_generation(const std::vector<uint32_t> &parents) { single("next_generation"); check(pansim_next_generation(ctx_, parents.data())); }
    // main.rs:445-464 in one fused pass (next_generation + mutate_alleles + recombine, both populations)
    void step_with_parents(uint32_t gen, const std::vector<uint32_t> &parents)
    {
        single("step_with_parents");
        check(pansim_step_with_parents(ctx_, gen, parents.data()));
    }
    // main.rs:435-464 entirely on the device
    void step(uint32_t gen) { gcheck(pansim_group_run_generations(grp_, gen, 1)); }

    // population.rs:787-837 for both populations: f64 distances formed on the host from
    // the integer counts with the reference's own expressions (:822, :828-830)
    void pairwise_distances(const std::vector<uint32_t> &range1, const std::vector<uint32_t> &range2,
                            std::vector<double> &core_out, std::vector<double> &acc_out)
    {
        const size_t P = range1.size();
        std::vector<uint32_t> cd(P), in(P), un(P);
        gcheck(pansim_group_pair_counts(grp_, range1.data(), range2.data(), P, cd.data(), in.data(), un.data()));
        core_out.resize(P);
        acc_out.resize(P);
        for (size_t k = 0; k < P; k++) {
            core_out[k] = pansim_core_distance(cd[k], cfg_.core_size);
            acc_out[k] = pansim_acc_distance(in[k], un[k], cfg_.core_genes);
        }
    }
    // (extension, BASELINE config 5) exact all-pairs mode: `sink(core_distance, acc_distance)` is called for
    // every pair i < j in (i, j) order. Row blocks of about chunk_pairs pairs; over several GPUs the partial
    // core counts of a block are reduce-scattered while the next block is computed (pansim_group_all_pairs).
    template <typename Sink>
    void all_pairs(size_t chunk_pairs, Sink &&sink)
    {
        struct Ctx { Populations *self; Sink *sink; } u{this, &sink};
        auto cb = [](void *user, uint32_t, uint32_t, size_t n, const uint32_t *cd, const uint32_t *in, const uint32_t *un) -> int {
            Ctx *x = static_cast<Ctx *>(user);
            const pansim_config &cf = x->self->cfg_;
            for (size_t k = 0; k < n; k++) (*x->sink)(pansim_core_distance(cd[k], cf.core_size), pansim_acc_distance(in[k], un[k], cf.core_genes));
            return 0;
        };
        gcheck(pansim_group_all_pairs(grp_, chunk_pairs, cb, &u));
    }
    // main.rs:429-519 under --print_dist as one device-resident batch: per generation (avg_core, std_core, avg_acc, std_acc)
    std::vector<double> run_generations_stats(uint32_t gen0, uint32_t n, const std::vector<uint32_t> &range1,
                                              const std::vector<uint32_t> &range2)
    {
        std::vector<double> out((size_t)n * 4);
        gcheck(pansim_group_run_generations_stats(grp_, gen0, n, range1.data(), range2.data(), range1.size(), out.data()));
        return out;
    }
    // main.rs:435-464 for generations gen0 .. gen0+n-1 as one device-resident batch
    void run_generations(uint32_t gen0, uint32_t n) { gcheck(pansim_group_run_generations(grp_, gen0, n)); }
    // population.rs:840-863: accessory gene frequencies, then core_genes x 1.0
    std::vector<double> gene_frequencies()
    {
        std::vector<uint32_t> counts(cfg_.pan_size);
        gcheck(pansim_group_gene_counts(grp_, counts.data()));
        std::vector<double> f;
        f.reserve(cfg_.pan_size + cfg_.core_genes);
        for (uint32_t c : counts) f.push_back((double)c / (double)cfg_.pop_size);
        for (uint32_t k = 0; k < cfg_.core_genes; k++) f.push_back(1.0);
        return f;
    }
    // population.rs:244-268
    double calc_gene_freq()
    {
        std::vector<uint8_t> acc((size_t)cfg_.pop_size * cfg_.pan_size);
        gcheck(pansim_group_download_acc(grp_, acc.data()));
        double sum = 0.0;
        for (uint32_t r = 0; r < cfg_.pop_size; r++) {
            size_t s = 0;
            for (uint32_t g = 0; g < cfg_.pan_size; g++) s += acc[(size_t)r * cfg_.pan_size + g];
            sum += (double)s / (double)cfg_.pan_size;
        }
        return sum / (double)cfg_.pop_size;
    }
    // population.rs:865-897 `write` for both populations
    void write(const std::string &outpref);

    pansim_ctx *raw() { return ctx_; }          // shard 0 (the whole run when n_gpus() == 1)
    pansim_group *group() { return grp_; }
    const pansim_config &config() const { return cfg_; }

private:
    void check(int rc)
    {
        if (rc != PANSIM_OK) throw Error(rc, pansim_last_error(ctx_));
    }
    void gcheck(int rc)
    {
        if (rc != PANSIM_OK) throw Error(rc, pansim_group_last_error(grp_));
    }
    // host-driven operators work on one context; a sharded run uses the device-resident calls
    void single(const char *what)
    {
        if (n_gpus() != 1) throw Error(PANSIM_ERR_STATE, std::string(what) + " is a single-GPU call; use step() / run_generations() with --gpus > 1");
    }
    pansim_config cfg_;
    pansim_group *grp_ = nullptr;
    pansim_ctx *ctx_ = nullptr;
};

}  // namespace pansim
