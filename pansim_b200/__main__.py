"""`python -m pansim_b200 [flags]` -- the `pansim` command line (main.rs:17-186).

Same long flags (underscores, capitalised --HR_rate / --HGT_rate), same defaults,
same parsing quirks: pop_size/core_size/pan_genes/core_genes/n_gen are parsed as
f64 and rounded (main.rs:155-167) so `--pop_size 1e3` works; max_distances,
threads and seed are parsed as integers (main.rs:169, 177, 180). --threads is
accepted and ignored (the work runs on the GPU). Extra flags: --device, --all_pairs.
"""
from __future__ import annotations

import argparse
import sys

from .params import Params, rust_round
from .simulate import run


def _f64_rounded(s: str) -> int:
    return int(rust_round(float(s)))


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(
        prog="pansim",
        description="Runs Wright-Fisher simulation, simulating neutral core genome evolution and "
                    "two-speed accessory genome evolution.", allow_abbrev=False)
    d = Params()
    ap.add_argument("--pop_size", type=_f64_rounded, default=d.pop_size, help="Number of individuals in population.")
    ap.add_argument("--core_size", type=_f64_rounded, default=d.core_size, help="Number of nucleotides in core genome.")
    ap.add_argument("--pan_genes", type=_f64_rounded, default=d.pan_genes,
                    help="Total number of genes in pangenome (core + accessory).")
    ap.add_argument("--core_genes", type=_f64_rounded, default=d.core_genes, help="Number of core genes in pangenome.")
    ap.add_argument("--avg_gene_freq", type=float, default=d.avg_gene_freq)
    ap.add_argument("--n_gen", type=_f64_rounded, default=d.n_gen, help="Number of generations to simulate.")
    ap.add_argument("--max_distances", type=int, default=d.max_distances)
    ap.add_argument("--core_mu", type=float, default=d.core_mu)
    ap.add_argument("--HR_rate", type=float, default=d.HR_rate)
    ap.add_argument("--HGT_rate", type=float, default=d.HGT_rate)
    ap.add_argument("--rate_genes1", type=float, default=d.rate_genes1)
    ap.add_argument("--rate_genes2", type=float, default=d.rate_genes2)
    ap.add_argument("--prop_genes2", type=float, default=d.prop_genes2)
    ap.add_argument("--prop_positive", type=float, default=d.prop_positive)
    ap.add_argument("--pos_lambda", type=float, default=d.pos_lambda)
    ap.add_argument("--neg_lambda", type=float, default=d.neg_lambda)
    ap.add_argument("--seed", type=int, default=d.seed)
    ap.add_argument("--outpref", default=d.outpref, help="Output prefix path.")
    ap.add_argument("--print_dist", action="store_true")
    ap.add_argument("--print_matrices", action="store_true")
    ap.add_argument("--print_selection", action="store_true")
    ap.add_argument("--threads", type=int, default=1, help="Accepted for compatibility; ignored.")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no_control_genome_size", action="store_true")
    ap.add_argument("--genome_size_penalty", type=float, default=d.genome_size_penalty)
    ap.add_argument("--competition_strength", type=float, default=d.competition_strength)
    ap.add_argument("--device", type=int, default=0, help="(extension) CUDA device ordinal")
    ap.add_argument("--all_pairs", action="store_true",
                    help="(extension) <outpref>.tsv holds every pair i < j in (i, j) order instead of max_distances sampled pairs")
    return ap


def main(argv=None) -> int:
    ns = vars(build_parser().parse_args(argv))
    device = ns.pop("device")
    all_pairs = ns.pop("all_pairs")
    p = Params(**ns)
    run(p, outpref=p.outpref, device=device, all_pairs=all_pairs)
    return 0                       # the reference exits 0 even when validation fails (main.rs:195-247)


if __name__ == "__main__":
    sys.exit(main())
