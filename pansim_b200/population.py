"""Host-side mirror of the reference's `Population` API over the C ABI.

`Pansim` owns one `pansim_ctx` = the reference's `core_genome` + `pan_genome`
pair (main.rs:372-391), because the reference always moves them together with
the same parent vector (main.rs:442-464). Method names follow
pansim/src/population.rs; argument meaning and error behaviour are documented
per method. numpy arrays are host buffers; nothing here computes on the CPU
except the f64 divisions the reference also does on the host side of the
boundary (population.rs:822, 828-830, 852).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .params import Derived, Params, derive, fmt_f64

OK = 0


class PansimError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_config(p: Params, d: Derived | None = None, device: int = 0, site_begin: int = 0,
                site_end: int = 0) -> _ffi.Config:
    """pansim_config from the command-line parameters (main.rs:259-287, 333-367)."""
    d = d or derive(p)
    cfg = _ffi.Config()
    _ffi.lib().pansim_config_init(C.byref(cfg))
    cfg.device = device
    cfg.pop_size = p.pop_size
    cfg.pan_size = d.pan_size
    cfg.core_size = p.core_size
    cfg.site_begin, cfg.site_end = site_begin, site_end
    cfg.core_genes = p.core_genes
    cfg.n_compartments = len(d.comp)
    for k, (lo, hi) in enumerate(d.comp):
        cfg.comp_lo[k], cfg.comp_hi[k] = lo, hi
        cfg.acc_mut_mean[k] = d.n_pan_mutations[k]
        cfg.hgt_mean[k] = d.n_recombinations_pan[k] if p.HGT_rate > 0.0 else 0.0   # main.rs:462
    cfg.core_mut_mean = d.n_core_mutations
    cfg.hr_mean = d.n_recombinations_core if p.HR_rate > 0.0 else 0.0              # main.rs:459
    cfg.avg_gene_num = d.avg_gene_num
    cfg.no_control_genome_size = int(p.no_control_genome_size)
    cfg.genome_size_penalty = p.genome_size_penalty
    cfg.competition_strength = p.competition_strength
    cfg.seed = p.seed
    return cfg


def standard_deviation(values) -> tuple[float, float]:
    """population.rs:87-94: (std, mean) with sequential f64 sums (np.cumsum is a
    sequential scan, unlike np.sum's pairwise reduction)."""
    v = np.ascontiguousarray(values, np.float64)
    n = len(v)
    mean = float(np.cumsum(v)[-1]) / n
    d = v - mean
    ss = float(np.cumsum(d * d)[-1])
    return float(np.sqrt(ss / n)), mean


class Pansim:
    """Both populations of one run on one GPU (optionally one column shard of the core)."""

    def __init__(self, cfg: _ffi.Config):
        self._lib = _ffi.lib()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._lib.pansim_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise PansimError(rc, self._lib.pansim_last_error(None).decode())
        self._h = h
        self.N, self.G, self.L = cfg.pop_size, cfg.pan_size, cfg.core_size
        info = self.info()
        self.local_sites = info.local_sites

    # -- lifecycle ---------------------------------------------------------
    @classmethod
    def from_params(cls, p: Params, device: int = 0, site_begin: int = 0, site_end: int = 0) -> "Pansim":
        return cls(make_config(p, device=device, site_begin=site_begin, site_end=site_end))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pansim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != OK:
            raise PansimError(rc, self._lib.pansim_last_error(self._h).decode())

    def info(self) -> _ffi.Info:
        i = _ffi.Info()
        self._check(self._lib.pansim_get_info(self._h, C.byref(i)))
        return i

    def timing(self) -> _ffi.Timing:
        t = _ffi.Timing()
        self._check(self._lib.pansim_get_timing(self._h, C.byref(t)))
        return t

    def set_timing(self, enabled: bool):
        self._check(self._lib.pansim_set_timing(self._h, int(enabled)))

    def rates(self) -> np.ndarray:
        out = np.zeros(4, np.float64)
        self._check(self._lib.pansim_get_rates(self._h, _ptr(out)))
        return out

    # -- state -------------------------------------------------------------
    def set_initial(self, core_row_onehot, acc_row):
        """Population::new x2 (population.rs:199-230): every row = the given row."""
        cr = np.ascontiguousarray(core_row_onehot, np.uint8)
        ar = np.ascontiguousarray(acc_row, np.uint8)
        assert cr.shape == (self.L,) and ar.shape == (self.G,)
        self._check(self._lib.pansim_set_initial(self._h, _ptr(cr), _ptr(ar)))

    def upload(self, core_onehot=None, acc=None):
        if core_onehot is not None:
            c = np.ascontiguousarray(core_onehot, np.uint8)
            assert c.shape == (self.N, self.local_sites)
            self._check(self._lib.pansim_upload_core(self._h, _ptr(c)))
        if acc is not None:
            a = np.ascontiguousarray(acc, np.uint8)
            assert a.shape == (self.N, self.G)
            self._check(self._lib.pansim_upload_acc(self._h, _ptr(a)))

    def download_core(self) -> np.ndarray:
        out = np.empty((self.N, self.local_sites), np.uint8)
        self._check(self._lib.pansim_download_core(self._h, _ptr(out)))
        return out

    def download_acc(self) -> np.ndarray:
        out = np.empty((self.N, self.G), np.uint8)
        self._check(self._lib.pansim_download_acc(self._h, _ptr(out)))
        return out

    def export_core_csv(self, row_begin: int, row_end: int) -> bytes:
        out = np.empty((row_end - row_begin) * 2 * self.local_sites, np.uint8)
        self._check(self._lib.pansim_export_core_csv(self._h, row_begin, row_end, _ptr(out)))
        return out.tobytes()

    def set_selection(self, s):
        s = np.ascontiguousarray(s, np.float64)
        assert s.shape == (self.G,)
        self._check(self._lib.pansim_set_selection(self._h, _ptr(s)))

    # -- operators (names of population.rs) --------------------------------
    def average_distance(self) -> np.ndarray:
        """population.rs:753-784 on the accessory population."""
        out = np.empty(self.N, np.float64)
        self._check(self._lib.pansim_average_distance(self._h, _ptr(out)))
        return out

    def sample_indices(self, gen: int, avg_pairwise_dists=None) -> np.ndarray:
        """population.rs:270-448. Raises PansimError(PANSIM_ERR_WEIGHTS) where the
        reference's WeightedIndex::new(...).unwrap() panics."""
        avg = None if avg_pairwise_dists is None else np.ascontiguousarray(avg_pairwise_dists, np.float64)
        out = np.empty(self.N, np.uint32)
        self._check(self._lib.pansim_sample_indices(self._h, gen, _ptr(avg), _ptr(out)))
        return out

    def select_parents(self, gen: int, reuse: bool = False):
        """main.rs:435-443 in one call (one synchronisation): (avg_pairwise_dists, parents).
        reuse=True returns the object's own two arrays (overwritten by the next call) instead of fresh ones:
        a generation loop that only hands the parents on to step_with_parents saves the allocations."""
        if reuse:
            if not hasattr(self, "_sel_bufs"):
                a, o = np.empty(self.N, np.float64), np.empty(self.N, np.uint32)
                self._sel_bufs = (a, o, _ptr(a), _ptr(o))
            avg, out, pa, po = self._sel_bufs
        else:
            avg, out = np.empty(self.N, np.float64), np.empty(self.N, np.uint32)
            pa, po = _ptr(avg), _ptr(out)
        rc = self._lib.pansim_select_parents(self._h, gen, pa, po)
        if rc != OK:
            self._check(rc)
        return avg, out

    def weights(self):
        w = np.empty(self.N, np.float64)
        ng = np.empty(self.N, np.int32)
        lf = np.empty(self.N, np.float64)
        self._check(self._lib.pansim_get_weights(self._h, _ptr(w), _ptr(ng), _ptr(lf)))
        return w, ng, lf

    def next_generation(self, parents):
        """population.rs:450-465 for both populations (main.rs:445-447)."""
        p = np.ascontiguousarray(parents, np.uint32)
        assert p.shape == (self.N,)
        self._check(self._lib.pansim_next_generation(self._h, _ptr(p)))

    def step_with_parents(self, gen: int, parents):
        """main.rs:445-464 with host-supplied parents (generate mode)."""
        p = parents if (isinstance(parents, np.ndarray) and parents.dtype == np.uint32 and parents.flags.c_contiguous) \
            else np.ascontiguousarray(parents, np.uint32)
        assert p.shape == (self.N,)
        rc = self._lib.pansim_step_with_parents(self._h, gen, p.ctypes.data)
        if rc != OK:
            self._check(rc)

    def step(self, gen: int):
        """main.rs:435-464 entirely on the device."""
        self._check(self._lib.pansim_step(self._h, gen))

    def run_generations(self, gen0: int, n: int):
        self._check(self._lib.pansim_run_generations(self._h, gen0, n))

    def parents(self) -> np.ndarray:
        out = np.empty(self.N, np.uint32)
        self._check(self._lib.pansim_get_parents(self._h, _ptr(out)))
        return out

    def step_replay(self, parents, core_mut=None, acc_flip=None, hr=None, hgt=None):
        """Replay one generation from explicit events (apply order).
        core_mut = (row, site, allele_onehot); acc_flip = (row, gene);
        hr = (recipient, locus, value_onehot); hgt = (recipient, gene)."""
        keep = []

        def arr(x, dt):
            a = np.ascontiguousarray(x, dt)
            keep.append(a)
            return a

        e = _ffi.Events()
        e.parents = _ptr(arr(parents, np.uint32))
        if core_mut is not None and len(core_mut[0]):
            e.n_core_mut = len(core_mut[0])
            e.core_mut_row = _ptr(arr(core_mut[0], np.uint32))
            e.core_mut_site = _ptr(arr(core_mut[1], np.uint32))
            e.core_mut_allele = _ptr(arr(core_mut[2], np.uint8))
        if acc_flip is not None and len(acc_flip[0]):
            e.n_acc_flip = len(acc_flip[0])
            e.acc_flip_row = _ptr(arr(acc_flip[0], np.uint32))
            e.acc_flip_gene = _ptr(arr(acc_flip[1], np.uint32))
        if hr is not None and len(hr[0]):
            e.n_hr = len(hr[0])
            e.hr_recipient = _ptr(arr(hr[0], np.uint32))
            e.hr_locus = _ptr(arr(hr[1], np.uint32))
            e.hr_value = _ptr(arr(hr[2], np.uint8))
        if hgt is not None and len(hgt[0]):
            e.n_hgt = len(hgt[0])
            e.hgt_recipient = _ptr(arr(hgt[0], np.uint32))
            e.hgt_gene = _ptr(arr(hgt[1], np.uint32))
        self._check(self._lib.pansim_step_replay(self._h, C.byref(e)))

    # -- distances / reductions -------------------------------------------
    def pair_counts(self, range1, range2, core=True, acc=True):
        """Integer outputs of population.rs:787-837: (core_diff, inter, union)."""
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        P = len(r1)
        cd = np.empty(P, np.uint32) if core else None
        it = np.empty(P, np.uint32) if acc else None
        un = np.empty(P, np.uint32) if acc else None
        self._check(self._lib.pansim_pair_counts(self._h, _ptr(r1), _ptr(r2), P, _ptr(cd), _ptr(it), _ptr(un)))
        return cd, it, un

    def pair_counts_device(self, range1, range2, d_core_diff: int, d_inter: int, d_uni: int):
        """Same, into caller-owned device buffers (raw pointers, e.g. tensor.data_ptr())."""
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        self._check(self._lib.pansim_pair_counts_device(self._h, _ptr(r1), _ptr(r2), len(r1),
                                                        C.c_void_p(d_core_diff), C.c_void_p(d_inter),
                                                        C.c_void_p(d_uni)))

    def distances_from_counts(self, core_diff, inter, uni):
        """The f64 formulas of population.rs:822 and :828-830 (IEEE double ops,
        identical to pansim_core_distance / pansim_acc_distance)."""
        cg = float(self.cfg.core_genes)
        core_d = core_diff.astype(np.float64) / float(self.L)
        acc_d = 1.0 - ((inter.astype(np.float64) + cg) / (uni.astype(np.float64) + cg))
        return core_d, acc_d

    def pairwise_distances(self, range1, range2):
        """population.rs:787-837 for both populations -> (core_distances, acc_distances)."""
        cd, it, un = self.pair_counts(range1, range2)
        return self.distances_from_counts(cd, it, un)

    def pair_stats(self, range1, range2) -> tuple[float, float, float, float]:
        """main.rs:502-519 on the device: (avg_core, std_core, avg_acc, std_acc) of one distance pass,
        sums in the reference's left-to-right order (population.rs:87-94)."""
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        out = np.empty(4, np.float64)
        self._check(self._lib.pansim_pair_stats(self._h, _ptr(r1), _ptr(r2), len(r1), _ptr(out)))
        return tuple(float(x) for x in out)

    def run_generations_stats(self, gen0: int, n: int, range1, range2) -> np.ndarray:
        """n generations, each followed by the distance pass and its statistics, in one device-resident
        batch (the --print_dist loop, main.rs:429-519): array [n, 4] of (avg_core, std_core, avg_acc, std_acc)."""
        r1 = np.ascontiguousarray(range1, np.uint32)
        r2 = np.ascontiguousarray(range2, np.uint32)
        out = np.empty((n, 4), np.float64)
        self._check(self._lib.pansim_run_generations_stats(self._h, gen0, n, _ptr(r1), _ptr(r2), len(r1), _ptr(out)))
        return out

    # -- multi-GPU: communicator over column shards (one process per GPU) --
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * _ffi.COMM_ID_BYTES)()
        rc = _ffi.lib().pansim_comm_unique_id(C.byref(buf))
        if rc != OK:
            raise PansimError(rc, _ffi.lib().pansim_last_error(None).decode())
        return bytes(buf)

    def comm_init_rank(self, n_ranks: int, rank: int, unique_id: bytes):
        """ncclCommInitRank inside the library: from now on the pair-count calls return
        whole-alignment core counts (summed over the column shards on the device)."""
        assert len(unique_id) == _ffi.COMM_ID_BYTES
        buf = (C.c_uint8 * _ffi.COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._lib.pansim_comm_init_rank(self._h, n_ranks, rank, C.byref(buf)))

    def pairs_in_rows(self, row_begin: int, row_end: int) -> int:
        """Number of pairs (i, j) with row_begin <= i < row_end, i < j < N."""
        n = row_end - row_begin
        return n * (self.N - 1) - (row_begin + row_end - 1) * n // 2

    def pair_counts_rows(self, row_begin: int, row_end: int, core: bool = True, acc: bool = True):
        """Exact all-pairs mode for one block of rows (extension; the reference only samples pairs
        with replacement, main.rs:413-427): (core_diff, inter, union) of every pair (i, j),
        row_begin <= i < row_end, i < j < N, ordered by i then j. The pair list is generated on
        the device."""
        P = self.pairs_in_rows(row_begin, row_end)
        cd = np.empty(P, np.uint32) if core else None
        it = np.empty(P, np.uint32) if acc else None
        un = np.empty(P, np.uint32) if acc else None
        n = C.c_size_t(0)
        self._check(self._lib.pansim_pair_counts_rows(self._h, row_begin, row_end, _ptr(cd), _ptr(it), _ptr(un),
                                                      C.byref(n)))
        assert n.value == P
        return cd, it, un

    def pair_counts_rows_device(self, row_begin: int, row_end: int, d_core_diff: int, d_inter: int, d_uni: int) -> int:
        """Same, into caller-owned device buffers (raw pointers); returns the number of pairs."""
        n = C.c_size_t(0)
        self._check(self._lib.pansim_pair_counts_rows_device(self._h, row_begin, row_end, C.c_void_p(d_core_diff),
                                                             C.c_void_p(d_inter), C.c_void_p(d_uni), C.byref(n)))
        return n.value

    def row_blocks(self, chunk_pairs: int = 4_000_000):
        """Row blocks [i0, i1) of about `chunk_pairs` pairs each covering every i < N - 1."""
        N, i0 = self.N, 0
        while i0 < N - 1:
            i1, n = i0, 0
            while i1 < N - 1 and (n == 0 or n + (N - 1 - i1) <= chunk_pairs):
                n += N - 1 - i1
                i1 += 1
            yield i0, i1
            i0 = i1

    def iter_all_pairs(self, chunk_pairs: int = 4_000_000, with_indices: bool = True):
        """Exact all-pairs mode: yields (i, j, core_diff, inter, union) arrays covering every
        unordered pair i < j exactly once, in chunks of about `chunk_pairs` pairs (i, j are None
        with with_indices=False: the order is i ascending, then j ascending)."""
        N = self.N
        for i0, i1 in self.row_blocks(chunk_pairs):
            cd, it, un = self.pair_counts_rows(i0, i1)
            ii = jj = None
            if with_indices:
                ii = np.repeat(np.arange(i0, i1, dtype=np.uint32), [N - 1 - i for i in range(i0, i1)])
                jj = np.concatenate([np.arange(i + 1, N, dtype=np.uint32) for i in range(i0, i1)])
            yield ii, jj, cd, it, un

    def gene_counts(self) -> np.ndarray:
        out = np.empty(self.G, np.uint32)
        self._check(self._lib.pansim_gene_counts(self._h, _ptr(out)))
        return out

    def gene_frequencies(self) -> np.ndarray:
        """population.rs:840-863: accessory frequencies then core_genes x 1.0."""
        f = self.gene_counts().astype(np.float64) / float(self.N)
        return np.concatenate([f, np.ones(self.cfg.core_genes, np.float64)])

    def calc_gene_freq(self) -> float:
        """population.rs:244-268: mean over rows of (row sum / G), sequential sums."""
        acc = self.download_acc()
        props = acc.sum(axis=1).astype(np.float64) / float(self.G)
        return float(np.cumsum(props)[-1]) / float(self.N)

    # -- instrumentation ---------------------------------------------------
    def enable_event_dump(self, max_core_events: int):
        self._check(self._lib.pansim_enable_event_dump(self._h, max_core_events))

    def fetch_event_dump(self) -> dict:
        d = _ffi.EventDump()
        self._check(self._lib.pansim_fetch_event_dump(self._h, C.byref(d)))

        def cp(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

        out = dict(
            core_mut_row=cp(d.core_mut_row, d.n_core_mut, np.uint32),
            core_mut_site=cp(d.core_mut_site, d.n_core_mut, np.uint32),
            core_mut_seq=cp(d.core_mut_seq, d.n_core_mut, np.uint32),
            core_mut_allele=cp(d.core_mut_allele, d.n_core_mut, np.uint8),
            hr_recipient=cp(d.hr_recipient, d.n_hr, np.uint32),
            hr_locus=cp(d.hr_locus, d.n_hr, np.uint32),
            hr_donor=cp(d.hr_donor, d.n_hr, np.uint32),
            hr_seq=cp(d.hr_seq, d.n_hr, np.uint32),
            hr_value=cp(d.hr_value, d.n_hr, np.uint8),
            acc_flip_mask=cp(d.acc_flip_mask, self.N * self.G, np.uint8).reshape(self.N, self.G),
            acc_gain_mask=cp(d.acc_gain_mask, self.N * self.G, np.uint8).reshape(self.N, self.G),
        )
        self._lib.pansim_free_event_dump(C.byref(d))
        return out

    # -- writers (population.rs:865-897) ------------------------------------
    def write_core_csv(self, path: str) -> int:
        """_core_genome.csv written by the library (GPU text expansion, pinned double-buffered streaming); bytes written."""
        n = C.c_uint64(0)
        self._check(self._lib.pansim_write_core_csv(self._h, path.encode(), C.byref(n)))
        return n.value

    def write(self, outpref: str):
        self.write_core_csv(outpref + "_core_genome.csv")
        acc = self.download_acc()
        with open(outpref + "_pangenome.csv", "w") as f:
            ones = ["1"] * self.cfg.core_genes                                   # :891
            for r in range(self.N):
                f.write(",".join(ones + [str(int(x)) for x in acc[r]]) + "\n")


__all__ = ["Pansim", "PansimError", "make_config", "standard_deviation", "fmt_f64"]
