// population.hpp -- C++ host-side mirror of the reference's `Population` API
// (pansim/src/population.rs) over the C ABI of include/pansim_b200.h.
//
// The reference keeps two Population objects (core_genome, pan_genome) that are
// always advanced together with the same parent vector (main.rs:442-464); here one
// `Populations` object owns one pansim_ctx = both of them. Method names, argument
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
    explicit Populations(const pansim_config &cfg) : cfg_(cfg)
    {
        const int rc = pansim_create(&cfg_, &ctx_);
        if (rc != PANSIM_OK) throw Error(rc, pansim_last_error(nullptr));
    }
    ~Populations() { pansim_destroy(ctx_); }
    Populations(const Populations &) = delete;
    Populations &operator=(const Populations &) = delete;

    // Population::new x2 (population.rs:181-242): every row starts as the given row
    void set_initial(const std::vector<uint8_t> &core_row_onehot, const std::vector<uint8_t> &acc_row)
    {
        check(pansim_set_initial(ctx_, core_row_onehot.data(), acc_row.data()));
    }
    void set_selection(const std::vector<double> &s) { check(pansim_set_selection(ctx_, s.data())); }

    // population.rs:753-784 (accessory population)
    std::vector<double> average_distance()
    {
        std::vector<double> out(cfg_.pop_size);
        check(pansim_average_distance(ctx_, out.data()));
        return out;
    }
    // population.rs:270-448
    std::vector<uint32_t> sample_indices(uint32_t gen, const std::vector<double> &avg_pairwise_dists)
    {
        std::vector<uint32_t> out(cfg_.pop_size);
        check(pansim_sample_indices(ctx_, gen, avg_pairwise_dists.empty() ? nullptr : avg_pairwise_dists.data(), out.data()));
        return out;
    }
    // population.rs:450-465 for both populations
    void next_generation(const std::vector<uint32_t> &parents) { check(pansim_next_generation(ctx_, parents.data())); }
    // main.rs:445-464 in one fused pass (next_generation + mutate_alleles + recombine, both populations)
    void step_with_parents(uint32_t gen, const std::vector<uint32_t> &parents)
    {
        check(pansim_step_with_parents(ctx_, gen, parents.data()));
    }
    // main.rs:435-464 entirely on the device
    void step(uint32_t gen) { check(pansim_step(ctx_, gen)); }

    // population.rs:787-837 for both populations: f64 distances formed on the host from
    // the integer counts with the reference's own expressions (:822, :828-830)
    void pairwise_distances(const std::vector<uint32_t> &range1, const std::vector<uint32_t> &range2,
                            std::vector<double> &core_out, std::vector<double> &acc_out)
    {
        const size_t P = range1.size();
        std::vector<uint32_t> cd(P), in(P), un(P);
        check(pansim_pair_counts(ctx_, range1.data(), range2.data(), P, cd.data(), in.data(), un.data()));
        core_out.resize(P);
        acc_out.resize(P);
        for (size_t k = 0; k < P; k++) {
            core_out[k] = pansim_core_distance(cd[k], cfg_.core_size);
            acc_out[k] = pansim_acc_distance(in[k], un[k], cfg_.core_genes);
        }
    }
    // (extension) exact all-pairs mode: distances of every pair (i, j), row_begin <= i < row_end, i < j < N,
    // ordered by i then j; at most 2^31 - 1 pairs per call
    void pairwise_distances_rows(uint32_t row_begin, uint32_t row_end, std::vector<double> &core_out,
                                 std::vector<double> &acc_out)
    {
        size_t P = 0;
        for (uint32_t i = row_begin; i < row_end && i < cfg_.pop_size; i++) P += cfg_.pop_size - 1 - i;
        std::vector<uint32_t> cd(P ? P : 1), in(P ? P : 1), un(P ? P : 1);
        size_t n = 0;
        check(pansim_pair_counts_rows(ctx_, row_begin, row_end, cd.data(), in.data(), un.data(), &n));
        core_out.resize(n);
        acc_out.resize(n);
        for (size_t k = 0; k < n; k++) {
            core_out[k] = pansim_core_distance(cd[k], cfg_.core_size);
            acc_out[k] = pansim_acc_distance(in[k], un[k], cfg_.core_genes);
        }
    }
    // main.rs:435-464 for generations gen0 .. gen0+n-1 as one device-resident batch
    void run_generations(uint32_t gen0, uint32_t n) { check(pansim_run_generations(ctx_, gen0, n)); }
    // population.rs:840-863: accessory gene frequencies, then core_genes x 1.0
    std::vector<double> gene_frequencies()
    {
        std::vector<uint32_t> counts(cfg_.pan_size);
        check(pansim_gene_counts(ctx_, counts.data()));
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
        check(pansim_download_acc(ctx_, acc.data()));
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

    pansim_ctx *raw() { return ctx_; }
    const pansim_config &config() const { return cfg_; }

private:
    void check(int rc)
    {
        if (rc != PANSIM_OK) throw Error(rc, pansim_last_error(ctx_));
    }
    pansim_config cfg_;
    pansim_ctx *ctx_ = nullptr;
};

}  // namespace pansim
