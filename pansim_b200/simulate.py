"""Host driver: the reference's `main()` (pansim/src/main.rs:15-564) over the C ABI.

Keeps on the host exactly what north_star leaves there: flag handling,
validation, selection-coefficient setup, initial rows, pair sampling and the
six output writers with the reference's text formats. Everything between
main.rs:435 and :528 runs on the GPU through `Pansim`.

Host randomness (selection coefficients, initial rows, pairs) comes from a
numpy generator seeded with --seed; the reference uses rand's StdRng there, whose
stream cannot be reproduced without the crate, so only the distributions match.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field

import numpy as np

from .params import Derived, Params, derive, fmt_f64, validate
from .population import Pansim, standard_deviation


def selection_coefficients(p: Params, pan_size: int, rng: np.random.Generator) -> np.ndarray:
    """main.rs:287-319."""
    s = np.zeros(pan_size, np.float64)
    if p.prop_positive >= 0.0:
        for i in range(pan_size):
            weight = rng.random()
            if weight <= p.prop_positive:
                s[i] = rng.exponential(1.0 / p.pos_lambda)
            else:
                v = rng.exponential(1.0 / p.neg_lambda)
                while v > 1.0:                      # main.rs:309-311
                    v = rng.exponential(1.0 / p.neg_lambda)
                s[i] = -1.0 * v
    return s


def initial_rows(p: Params, d: Derived, rng: np.random.Generator):
    """Population::new x2 (population.rs:199-204, 214-219): one row each."""
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
    return core_row, acc_row


def sample_pairs(p: Params, rng: np.random.Generator):
    """main.rs:413-427: ordered pairs i != j, with replacement, fixed for the run."""
    r1 = rng.integers(0, p.pop_size, p.max_distances).astype(np.uint32)
    if p.pop_size < 2:
        raise ValueError("pop_size must be >= 2 to sample pairs (gen_range(0..pop_size-1) panics)")
    r2 = rng.integers(0, p.pop_size - 1, p.max_distances).astype(np.uint32)
    r2 = (r2 + (r2 >= r1)).astype(np.uint32)
    return r1, r2


@dataclass
class RunResult:
    core_distances: np.ndarray = None
    acc_distances: np.ndarray = None
    gene_freqs: np.ndarray = None            # accessory genes then core_genes ones (population.rs:857-860)
    per_gen: list = field(default_factory=list)   # (avg_core, std_core, avg_acc, std_acc)
    selection: np.ndarray = None
    stdout: list = field(default_factory=list)


def run(p: Params, outpref: str | None = None, device: int = 0, out=None, all_pairs: bool = False) -> RunResult | None:
    """main.rs:15-564. Returns None where the reference prints a validation message and exits 0.
    all_pairs (extension, BASELINE config 5): `<outpref>.tsv` holds the distances of every pair i < j
    in (i, j) order instead of the max_distances sampled pairs; res.core/acc_distances stay None."""
    out = out or sys.stdout
    msgs = validate(p)
    if msgs:
        for m in msgs:
            print(m, file=out)
        return None
    d = derive(p)
    res = RunResult()
    if p.verbose:
        line = f"avg_gene_freq adjusted to {fmt_f64(d.avg_gene_freq_adj)}"      # main.rs:269-271
        print(line, file=out)
        res.stdout.append(line)
    rng = np.random.default_rng(p.seed)                                         # main.rs:289
    sel = selection_coefficients(p, d.pan_size, rng)
    res.selection = sel
    if p.print_selection and outpref:                                            # main.rs:321-331
        with open(outpref + "_selection.tsv", "w") as f:
            f.write("\n".join(fmt_f64(float(x)) for x in sel) + "\n")
    core_row, acc_row = initial_rows(p, d, rng)
    r1, r2 = sample_pairs(p, rng)

    with Pansim.from_params(p, device=device) as sim:
        sim.set_initial(core_row, acc_row)
        sim.set_selection(sel)
        def last_generation_outputs():                                           # main.rs:467-499
            if not all_pairs:
                res.core_distances, res.acc_distances = sim.pairwise_distances(r1, r2)
            res.gene_freqs = sim.gene_frequencies()
            if outpref:
                with open(outpref + ".tsv", "w") as f:
                    if all_pairs:
                        for _, _, cd_, it_, un_ in sim.iter_all_pairs(with_indices=False):
                            c_, a_ = sim.distances_from_counts(cd_, it_, un_)
                            f.write("".join(f"{fmt_f64(float(c))}\t{fmt_f64(float(a))}\n" for c, a in zip(c_, a_)))
                    else:
                        f.write("".join(f"{fmt_f64(float(c))}\t{fmt_f64(float(a))}\n"
                                        for c, a in zip(res.core_distances, res.acc_distances)))
                with open(outpref + "_freqs.txt", "w") as f:
                    f.write("".join(fmt_f64(float(x)) + "\n" for x in res.gene_freqs))

        # Nothing has to come back between generations unless --verbose asks for it: the whole run is one
        # device-resident batch (same states, tests/test_gpu_parity.py). With --print_dist the batch also
        # makes the distance pass and its mean / standard deviation every generation on the device
        # (pansim_run_generations_stats: sums in the reference's left-to-right order, population.rs:87-94).
        device_loop = not p.verbose and not (p.print_dist and all_pairs)
        first_host_gen = 0
        if device_loop:
            if p.print_dist:
                res.per_gen = [tuple(float(x) for x in row) for row in sim.run_generations_stats(0, p.n_gen, r1, r2)]
            else:
                sim.run_generations(0, p.n_gen)
            first_host_gen = p.n_gen
            last_generation_outputs()
        for j in range(first_host_gen, p.n_gen):                                 # main.rs:429
            sim.step(j)                                                          # main.rs:435-464
            if j == p.n_gen - 1:
                last_generation_outputs()
            if p.print_dist:                                                     # main.rs:502-519
                # the reference recomputes both passes here even on the last generation;
                # the result is identical, so the last one is reused
                if j == p.n_gen - 1 and not all_pairs:
                    cd, ad = res.core_distances, res.acc_distances
                else:
                    cd, ad = sim.pairwise_distances(r1, r2)
                std_core, avg_core = standard_deviation(cd)
                std_acc, avg_acc = standard_deviation(ad)
                res.per_gen.append((avg_core, std_core, avg_acc, std_acc))
            if p.verbose:                                                        # main.rs:522-526
                l1 = f"Finished gen: {j + 1}"
                l2 = f"avg_gene_freq: {fmt_f64(sim.calc_gene_freq())}"
                print(l1, file=out)
                print(l2, file=out)
                res.stdout += [l1, l2]
        if p.print_dist and outpref:                                             # main.rs:531-548
            with open(outpref + "_per_gen.tsv", "w") as f:
                for avg_core, std_core, avg_acc, std_acc in res.per_gen:
                    f.write(f"{fmt_f64(avg_core)}\t{fmt_f64(std_core)}\t{fmt_f64(avg_acc)}\t{fmt_f64(std_acc)}\n")
        if p.print_matrices and outpref:                                         # main.rs:550-553
            sim.write(outpref)
    return res


def summarize(core_d, acc_d, gene_freqs, n_acc_genes: int) -> dict:
    """Per-run scalars used by the end-to-end KS tests (mirrors oracle ora_summary)."""
    f = np.asarray(gene_freqs[:n_acc_genes], np.float64)
    return dict(
        mean_core=float(np.mean(core_d)), mean_acc=float(np.mean(acc_d)),
        median_core=float(np.median(core_d)), median_acc=float(np.median(acc_d)),
        std_core=float(np.std(core_d)), std_acc=float(np.std(acc_d)),
        mean_gene_freq=float(f.mean()) if n_acc_genes else 0.0,
        frac_freq_lt_01=float((f < 0.1).mean()) if n_acc_genes else 0.0,
        frac_freq_gt_09=float((f > 0.9).mean()) if n_acc_genes else 0.0,
    )
