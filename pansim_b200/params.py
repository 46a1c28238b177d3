"""Host-side parameter handling: the flags of main.rs:21-151, the validation of
main.rs:194-247 and the derived rates of main.rs:259-287 / 333-367.

The derived numbers are the kernel parameters (pansim_config in
include/pansim_b200.h). Pure Python, no device work.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from decimal import Decimal


@dataclass
class Params:
    """One field per result-affecting flag, defaults = clap default_value (main.rs:21-151)."""
    pop_size: int = 1000
    core_size: int = 1200000
    pan_genes: int = 6000
    core_genes: int = 2000
    avg_gene_freq: float = 0.5
    n_gen: int = 100
    max_distances: int = 100000
    core_mu: float = 0.05
    HR_rate: float = 0.05
    HGT_rate: float = 0.05
    rate_genes1: float = 1.0
    rate_genes2: float = 1000.0
    prop_genes2: float = 0.1
    prop_positive: float = -0.1
    pos_lambda: float = 10.0
    neg_lambda: float = 10.0
    seed: int = 0
    outpref: str = "distances"
    print_dist: bool = False
    print_matrices: bool = False
    print_selection: bool = False
    threads: int = 1
    verbose: bool = False
    no_control_genome_size: bool = False
    genome_size_penalty: float = 0.99
    competition_strength: float = 0.0

    def as_dict(self):
        return asdict(self)


def rust_round(x: float) -> float:
    """f64::round: half away from zero (Python's round() is half-to-even)."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def validate(p: Params) -> list[str]:
    """main.rs:194-247. Returns the lines the reference prints before `return Ok(())`
    (empty list = parameters accepted)."""
    if p.core_genes > p.pan_genes:
        return ["core_genes must be less than or equal to pan_size"]
    if p.HR_rate < 0.0 or p.HGT_rate < 0.0:
        return ["HR_rate and HGT_rate must be above 0.0", f"HR_rate: {fmt_f64(p.HR_rate)}",
                f"HGT_rate: {fmt_f64(p.HGT_rate)}"]
    if p.pos_lambda <= 0.0 or p.neg_lambda <= 0.0:
        return ["pos_lambda and neg_lambda must be above 0.0", f"pos_lambda: {fmt_f64(p.pos_lambda)}",
                f"neg_lambda: {fmt_f64(p.neg_lambda)}"]
    if p.rate_genes1 < 0.0 or p.rate_genes2 < 0.0:
        return ["rate_genes1 and rate_genes2 must be >= 0", f"rate_genes1: {fmt_f64(p.rate_genes1)}",
                f"rate_genes2: {fmt_f64(p.rate_genes2)}"]
    if p.prop_genes2 < 0.0 or p.prop_genes2 > 1.0:
        return ["prop_genes2 must be 0.0 <= prop_genes2 <= 1.0", f"prop_genes2: {fmt_f64(p.prop_genes2)}"]
    if p.pop_size < 1 or p.core_size < 1 or p.pan_genes < 1 or p.n_gen < 1 or p.max_distances < 1:
        return ["pop_size, core_size, pan_genes, n_gen and max_distances must all be above 1",
                f"pop_size: {p.pop_size}", f"core_size: {p.core_size}", f"pan_genes: {p.pan_genes}",
                f"n_gen: {p.n_gen}", f"max_distances: {p.max_distances}"]
    if p.core_mu < 0.0 or p.core_mu > 1.0:
        return ["core_mu must be between 0.0 and 1.0", f"core_mu: {fmt_f64(p.core_mu)}"]
    if p.avg_gene_freq <= 0.0 or p.avg_gene_freq > 1.0:
        return ["avg_gene_freq must be above 0.0 and below or equal to 1.0",
                f"avg_gene_freq: {fmt_f64(p.avg_gene_freq)}"]
    return []


@dataclass
class Derived:
    pan_size: int = 0
    avg_gene_freq_adj: float = 0.0
    avg_gene_num: int = 0
    n_core_mutations: float = 0.0
    n_recombinations_core: float = 0.0
    n_recombinations_pan_total: float = 0.0
    num_gene1_sites: int = 0
    num_gene2_sites: int = 0
    comp: list = field(default_factory=list)            # [(lo, hi)] genes with weight 1.0
    n_pan_mutations: list = field(default_factory=list)
    n_recombinations_pan: list = field(default_factory=list)


def derive(p: Params) -> Derived:
    """main.rs:259-287 and 333-367, same floating-point expressions."""
    d = Derived()
    d.pan_size = p.pan_genes - p.core_genes                              # :259
    core_prop = p.core_genes / p.pan_genes                               # :263
    acc_prop = 1.0 - core_prop
    agf = (p.avg_gene_freq - core_prop) / acc_prop if acc_prop != 0.0 else float("nan")   # :265
    if agf < 0.0:
        agf = 0.0
    d.avg_gene_freq_adj = agf
    d.avg_gene_num = int(rust_round(agf * d.pan_size)) if d.pan_size else 0   # :272
    d.n_core_mutations = float(math.ceil(p.core_size * p.core_mu))       # :275-276
    d.n_recombinations_core = rust_round(d.n_core_mutations * p.HR_rate)  # :279
    d.n_recombinations_pan_total = rust_round(d.n_core_mutations * p.HGT_rate)  # :280
    d.num_gene1_sites = int(rust_round(d.pan_size * (1.0 - p.prop_genes2)))     # :334
    d.num_gene2_sites = d.pan_size - d.num_gene1_sites
    prop1 = d.num_gene1_sites / d.pan_size if d.pan_size else 0.0        # :336
    prop2 = 1.0 - prop1
    if d.num_gene1_sites > 0:                                            # :341-352
        d.comp.append((0, d.num_gene1_sites))
        d.n_pan_mutations.append(p.rate_genes1 * d.num_gene1_sites)
        d.n_recombinations_pan.append(d.n_recombinations_pan_total * prop1)
    if d.num_gene1_sites < d.pan_size:                                   # :355-367
        d.comp.append((d.num_gene1_sites, d.pan_size))
        d.n_pan_mutations.append(p.rate_genes2 * d.num_gene2_sites)
        d.n_recombinations_pan.append(d.n_recombinations_pan_total * prop2)
    return d


def fmt_f64(x: float) -> str:
    """Rust `{}` Display for f64: shortest round-trip digits, never scientific,
    integral values without '.0', NaN / inf / -inf."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == 0.0:
        return "-0" if math.copysign(1.0, x) < 0 else "0"
    s = format(Decimal(repr(x)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s
