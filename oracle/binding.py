"""ctypes binding of the CPU oracle (oracle/pansim_oracle.c).

TEST INFRASTRUCTURE ONLY. Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by pansim_b200/.
Builds oracle/libpansim_oracle.so with oracle/Makefile on first use.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpansim_oracle.so")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class PopulationStruct(C.Structure):
    _fields_ = [("pop", C.POINTER(C.c_uint8)), ("nrows", C.c_size_t), ("ncols", C.c_size_t),
                ("core", C.c_int), ("core_genes", C.c_size_t), ("avg_gene_freq", C.c_double)]


class Events(C.Structure):
    _fields_ = [
        ("n_core_mut", C.c_size_t), ("cap_core_mut", C.c_size_t),
        ("core_mut_row", C.POINTER(C.c_uint32)), ("core_mut_site", C.POINTER(C.c_uint32)),
        ("core_mut_allele", C.POINTER(C.c_uint8)),
        ("n_acc_flip", C.c_size_t), ("cap_acc_flip", C.c_size_t),
        ("acc_flip_row", C.POINTER(C.c_uint32)), ("acc_flip_gene", C.POINTER(C.c_uint32)),
        ("n_hr", C.c_size_t), ("cap_hr", C.c_size_t),
        ("hr_recipient", C.POINTER(C.c_uint32)), ("hr_locus", C.POINTER(C.c_uint32)),
        ("hr_donor", C.POINTER(C.c_uint32)), ("hr_value", C.POINTER(C.c_uint8)),
        ("n_hgt", C.c_size_t), ("cap_hgt", C.c_size_t),
        ("hgt_recipient", C.POINTER(C.c_uint32)), ("hgt_gene", C.POINTER(C.c_uint32)),
        ("hgt_donor", C.POINTER(C.c_uint32)),
    ]


class SiteDist(C.Structure):
    _fields_ = [("lo", C.c_uint32), ("hi", C.c_uint32),
                ("cumulative", C.POINTER(C.c_float)), ("n_table", C.c_size_t)]


class Params(C.Structure):
    _fields_ = [
        ("pop_size", C.c_size_t), ("core_size", C.c_size_t), ("pan_genes", C.c_size_t),
        ("core_genes", C.c_size_t), ("avg_gene_freq", C.c_double), ("HR_rate", C.c_double),
        ("HGT_rate", C.c_double), ("n_gen", C.c_int), ("max_distances", C.c_size_t),
        ("core_mu", C.c_double), ("rate_genes1", C.c_double), ("rate_genes2", C.c_double),
        ("prop_genes2", C.c_double), ("prop_positive", C.c_double), ("pos_lambda", C.c_double),
        ("neg_lambda", C.c_double), ("seed", C.c_uint64), ("print_dist", C.c_int),
        ("print_matrices", C.c_int), ("print_selection", C.c_int), ("verbose", C.c_int),
        ("no_control_genome_size", C.c_int), ("genome_size_penalty", C.c_double),
        ("competition_strength", C.c_double), ("threads", C.c_int),
    ]


class Derived(C.Structure):
    _fields_ = [
        ("pan_size", C.c_size_t), ("avg_gene_freq_adj", C.c_double), ("avg_gene_num", C.c_int32),
        ("n_core_mutations", C.c_double), ("n_recombinations_core", C.c_double),
        ("n_recombinations_pan_total", C.c_double), ("num_gene1_sites", C.c_size_t),
        ("num_gene2_sites", C.c_size_t), ("n_compartments", C.c_size_t),
        ("comp_lo", C.c_uint32 * 2), ("comp_hi", C.c_uint32 * 2),
        ("n_pan_mutations", C.c_double * 2), ("n_recombinations_pan", C.c_double * 2),
    ]


class Summary(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "mean_core", "mean_acc", "median_core", "median_acc", "std_core", "std_acc",
        "mean_gene_freq", "frac_freq_lt_01", "frac_freq_gt_09", "mean_genes_per_row")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pansim_oracle.c")
    hdr = os.path.join(_HERE, "pansim_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libpansim_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    sz, dbl, u64, i32, cint = C.c_size_t, C.c_double, C.c_uint64, C.c_int32, C.c_int
    RP, PP, EP = C.POINTER(Rng), C.POINTER(PopulationStruct), C.POINTER(Events)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ora_rng_seed", None, RP, u64)
    sig("ora_rng_seed4", None, RP, u64, u64, u64, u64)
    sig("ora_rng_next", u64, RP)
    sig("ora_rng_f64", dbl, RP)
    sig("ora_rng_below", u64, RP, u64)
    sig("ora_poisson", u64, RP, dbl)
    sig("ora_exponential", dbl, RP, dbl)
    sig("ora_hamming_bitwise_fast", C.c_uint32, u8p, u8p, sz)
    sig("ora_jaccard_distance_fast", None, u8p, u8p, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
    sig("ora_jaccard_distance_naive", None, u8p, u8p, sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
    sig("ora_standard_deviation", None, f64p, sz, C.POINTER(dbl), C.POINTER(dbl))
    sig("ora_int_to_base", C.c_char, C.c_uint8)
    sig("ora_population_new", cint, PP, sz, sz, C.c_uint8, cint, dbl, RP, sz)
    sig("ora_population_free", None, PP)
    sig("ora_calc_gene_freq", dbl, PP)
    sig("ora_selection_weights", None, PP, i32, f64p, f64p, cint, dbl, dbl, f64p,
        C.c_void_p, C.c_void_p)
    sig("ora_weighted_index_sample", cint, f64p, sz, RP, sz, u32p)
    sig("ora_sample_indices", cint, PP, RP, i32, f64p, f64p, cint, dbl, dbl, u32p)
    sig("ora_next_generation", cint, PP, u32p, sz)
    sig("ora_events_init", None, EP)
    sig("ora_events_clear", None, EP)
    sig("ora_events_free", None, EP)
    sig("ora_mutate_alleles", None, PP, C.POINTER(dbl), C.POINTER(SiteDist), sz, u64, u64, EP)
    sig("ora_recombine", cint, PP, C.POINTER(dbl), C.POINTER(SiteDist), sz, RP, u64, u64, EP)
    sig("ora_apply_core_writes", None, PP, u32p, u32p, u8p, sz)
    sig("ora_apply_acc_flips", None, PP, u32p, u32p, sz)
    sig("ora_apply_acc_sets", None, PP, u32p, u32p, sz)
    sig("ora_average_distance", None, PP, f64p)
    sig("ora_pair_counts", None, PP, sz, u32p, u32p, C.c_void_p, C.c_void_p, C.c_void_p)
    sig("ora_pairwise_distances", None, PP, sz, u32p, u32p, f64p)
    sig("ora_core_distance_from_count", dbl, C.c_uint32, sz)
    sig("ora_acc_distance_from_counts", dbl, C.c_uint32, C.c_uint32, sz)
    sig("ora_gene_frequencies", None, PP, f64p)
    sig("ora_gene_counts", None, PP, u32p)
    sig("ora_params_default", None, C.POINTER(Params))
    sig("ora_validate", cint, C.POINTER(Params))
    sig("ora_derive", None, C.POINTER(Params), C.POINTER(Derived))
    sig("ora_selection_coefficients", None, C.POINTER(Params), sz, RP, f64p)
    sig("ora_sample_pairs", None, sz, sz, RP, u32p, u32p)
    sig("ora_fmt_f64", cint, C.c_char_p, dbl)
    sig("ora_run", cint, C.POINTER(Params), C.c_char_p, cint, C.POINTER(Summary))
    sig("ora_time_generations", dbl, C.POINTER(Params), cint, cint, cint, C.POINTER(dbl))
    sig("ora_max_threads", cint)
    sig("ora_set_threads", None, cint)
    _lib = L
    return L


# --------------------------------------------------------------------------
# Pythonic wrappers
# --------------------------------------------------------------------------
def make_rng(seed: int) -> Rng:
    r = Rng()
    lib().ora_rng_seed(C.byref(r), seed)
    return r


def default_params(**kw) -> Params:
    p = Params()
    lib().ora_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def derive(p: Params) -> Derived:
    d = Derived()
    lib().ora_derive(C.byref(p), C.byref(d))
    return d


def fmt_f64(x: float) -> str:
    buf = C.create_string_buffer(512)
    lib().ora_fmt_f64(buf, x)
    return buf.value.decode()


class Population:
    """Oracle population over a numpy [nrows x ncols] uint8 matrix (reference layout)."""

    def __init__(self, matrix: np.ndarray, core: bool, core_genes: int = 0):
        self.m = np.array(matrix, dtype=np.uint8, order="C", copy=True)   # never alias the caller's buffer
        self.core = bool(core)
        self.core_genes = int(core_genes)

    # the C struct borrows self.m's buffer; rebuilt per call because
    # next_generation replaces the buffer
    def _struct(self) -> PopulationStruct:
        s = PopulationStruct()
        s.pop = self.m.ctypes.data_as(C.POINTER(C.c_uint8))
        s.nrows, s.ncols = self.m.shape
        s.core = 1 if self.core else 0
        s.core_genes = self.core_genes
        s.avg_gene_freq = 0.0
        return s

    @property
    def shape(self):
        return self.m.shape

    def copy(self) -> "Population":
        return Population(self.m.copy(), self.core, self.core_genes)

    def next_generation(self, parents: np.ndarray) -> None:
        # population.rs:450-465 (numpy gather == row-by-row assign)
        self.m = np.ascontiguousarray(self.m[np.asarray(parents, dtype=np.int64)])

    def selection_weights(self, avg_gene_num, avg_pairwise_dists, sel, no_control=False,
                          penalty=0.99, competition_strength=0.0):
        n = self.m.shape[0]
        w = np.zeros(n, np.float64)
        ng = np.zeros(n, np.int32)
        lf = np.zeros(n, np.float64)
        s = self._struct()
        lib().ora_selection_weights(C.byref(s), int(avg_gene_num),
                                    np.ascontiguousarray(avg_pairwise_dists, np.float64),
                                    np.ascontiguousarray(sel, np.float64), int(no_control),
                                    float(penalty), float(competition_strength), w,
                                    ng.ctypes.data, lf.ctypes.data)
        return w, ng, lf

    def average_distance(self) -> np.ndarray:
        out = np.zeros(self.m.shape[0], np.float64)
        s = self._struct()
        lib().ora_average_distance(C.byref(s), out)
        return out

    def pair_counts(self, r1, r2):
        r1 = np.ascontiguousarray(r1, np.uint32)
        r2 = np.ascontiguousarray(r2, np.uint32)
        P = len(r1)
        s = self._struct()
        if self.core:
            cd = np.zeros(P, np.uint32)
            lib().ora_pair_counts(C.byref(s), P, r1, r2, cd.ctypes.data, None, None)
            return cd
        inter = np.zeros(P, np.uint32)
        uni = np.zeros(P, np.uint32)
        lib().ora_pair_counts(C.byref(s), P, r1, r2, None, inter.ctypes.data, uni.ctypes.data)
        return inter, uni

    def pairwise_distances(self, r1, r2) -> np.ndarray:
        r1 = np.ascontiguousarray(r1, np.uint32)
        r2 = np.ascontiguousarray(r2, np.uint32)
        out = np.zeros(len(r1), np.float64)
        s = self._struct()
        lib().ora_pairwise_distances(C.byref(s), len(r1), r1, r2, out)
        return out

    def gene_counts(self) -> np.ndarray:
        out = np.zeros(self.m.shape[1], np.uint32)
        s = self._struct()
        lib().ora_gene_counts(C.byref(s), out)
        return out

    def gene_frequencies(self) -> np.ndarray:
        out = np.zeros(self.m.shape[1] + self.core_genes, np.float64)
        s = self._struct()
        lib().ora_gene_frequencies(C.byref(s), out)
        return out

    def calc_gene_freq(self) -> float:
        s = self._struct()
        return lib().ora_calc_gene_freq(C.byref(s))

    def mutate_alleles(self, means, ranges, seed, gen, events: "EventLog | None" = None,
                       use_tables=False):
        n = len(means)
        mv = (C.c_double * n)(*[float(x) for x in means])
        dists = (SiteDist * n)()
        keep = []
        for k, (lo, hi) in enumerate(ranges):
            dists[k].lo, dists[k].hi = int(lo), int(hi)
            if use_tables:
                tab = np.cumsum(((np.arange(self.m.shape[1]) >= lo) & (np.arange(self.m.shape[1]) < hi))
                                .astype(np.float32), dtype=np.float32)
                keep.append(tab)
                dists[k].cumulative = tab.ctypes.data_as(C.POINTER(C.c_float))
                dists[k].n_table = self.m.shape[1]
        s = self._struct()
        lib().ora_mutate_alleles(C.byref(s), mv, dists, n, int(seed), int(gen),
                                 C.byref(events.e) if events is not None else None)

    def recombine(self, means, ranges, rng: Rng, seed, gen, events: "EventLog | None" = None):
        n = len(means)
        mv = (C.c_double * n)(*[float(x) for x in means])
        dists = (SiteDist * n)()
        for k, (lo, hi) in enumerate(ranges):
            dists[k].lo, dists[k].hi = int(lo), int(hi)
        s = self._struct()
        rc = lib().ora_recombine(C.byref(s), mv, dists, n, C.byref(rng), int(seed), int(gen),
                                 C.byref(events.e) if events is not None else None)
        if rc:
            raise RuntimeError("ora_recombine failed (pop_size < 2?)")

    def apply_core_writes(self, row, site, value):
        s = self._struct()
        lib().ora_apply_core_writes(C.byref(s), np.ascontiguousarray(row, np.uint32),
                                    np.ascontiguousarray(site, np.uint32),
                                    np.ascontiguousarray(value, np.uint8), len(row))

    def apply_acc_flips(self, row, gene):
        s = self._struct()
        lib().ora_apply_acc_flips(C.byref(s), np.ascontiguousarray(row, np.uint32),
                                  np.ascontiguousarray(gene, np.uint32), len(row))

    def apply_acc_sets(self, row, gene):
        s = self._struct()
        lib().ora_apply_acc_sets(C.byref(s), np.ascontiguousarray(row, np.uint32),
                                 np.ascontiguousarray(gene, np.uint32), len(row))


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class EventLog:
    """Owns an ora_events; `.arrays()` copies it out as numpy arrays."""

    def __init__(self):
        self.e = Events()
        lib().ora_events_init(C.byref(self.e))

    def clear(self):
        lib().ora_events_clear(C.byref(self.e))

    def __del__(self):
        try:
            lib().ora_events_free(C.byref(self.e))
        except Exception:
            pass

    def arrays(self) -> dict:
        e = self.e
        return dict(
            core_mut_row=_arr(e.core_mut_row, e.n_core_mut, np.uint32),
            core_mut_site=_arr(e.core_mut_site, e.n_core_mut, np.uint32),
            core_mut_allele=_arr(e.core_mut_allele, e.n_core_mut, np.uint8),
            acc_flip_row=_arr(e.acc_flip_row, e.n_acc_flip, np.uint32),
            acc_flip_gene=_arr(e.acc_flip_gene, e.n_acc_flip, np.uint32),
            hr_recipient=_arr(e.hr_recipient, e.n_hr, np.uint32),
            hr_locus=_arr(e.hr_locus, e.n_hr, np.uint32),
            hr_donor=_arr(e.hr_donor, e.n_hr, np.uint32),
            hr_value=_arr(e.hr_value, e.n_hr, np.uint8),
            hgt_recipient=_arr(e.hgt_recipient, e.n_hgt, np.uint32),
            hgt_gene=_arr(e.hgt_gene, e.n_hgt, np.uint32),
            hgt_donor=_arr(e.hgt_donor, e.n_hgt, np.uint32),
        )


def new_population(size, allele_count, max_variants, core, avg_gene_freq, rng: Rng,
                   core_genes) -> Population:
    """Population::new (population.rs:181-242) -> numpy-backed Population."""
    s = PopulationStruct()
    rc = lib().ora_population_new(C.byref(s), size, allele_count, max_variants, int(core),
                                  float(avg_gene_freq), C.byref(rng), core_genes)
    if rc:
        raise MemoryError
    m = np.ctypeslib.as_array(s.pop, shape=(size, allele_count)).copy() if size * allele_count \
        else np.zeros((size, allele_count), np.uint8)
    lib().ora_population_free(C.byref(s))
    return Population(m, core, core_genes)


def weighted_index_sample(weights, rng: Rng, n_draws) -> np.ndarray:
    w = np.ascontiguousarray(weights, np.float64)
    out = np.zeros(n_draws, np.uint32)
    rc = lib().ora_weighted_index_sample(w, len(w), C.byref(rng), n_draws, out)
    if rc:
        raise ValueError("WeightedIndex::new would panic (negative/NaN weight or zero total)")
    return out


def standard_deviation(v):
    v = np.ascontiguousarray(v, np.float64)
    s, m = C.c_double(), C.c_double()
    lib().ora_standard_deviation(v, len(v), C.byref(s), C.byref(m))
    return s.value, m.value


def sample_pairs(pop_size, max_distances, rng: Rng):
    r1 = np.zeros(max_distances, np.uint32)
    r2 = np.zeros(max_distances, np.uint32)
    lib().ora_sample_pairs(pop_size, max_distances, C.byref(rng), r1, r2)
    return r1, r2


def selection_coefficients(p: Params, pan_size, rng: Rng) -> np.ndarray:
    out = np.zeros(pan_size, np.float64)
    lib().ora_selection_coefficients(C.byref(p), pan_size, C.byref(rng), out)
    return out


def run(p: Params, outpref: str | None = None, use_tables: bool = False) -> Summary:
    s = Summary()
    rc = lib().ora_run(C.byref(p), outpref.encode() if outpref else None, int(use_tables), C.byref(s))
    if rc < 0:
        raise RuntimeError("ora_run failed")
    return s


def time_generations(p: Params, n_gen: int, with_distances: bool, use_tables: bool = True):
    d = C.c_double(0.0)
    t = lib().ora_time_generations(C.byref(p), n_gen, int(with_distances), int(use_tables), C.byref(d))
    if t < 0:
        raise RuntimeError("ora_time_generations failed")
    return t, d.value
