"""CPU-side checks of the drop-in boundary: the shared library loads and exports
every symbol include/pansim_b200.h declares; host logic (parameters, formatting).
No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pansim_b200 as pb
from pansim_b200 import _ffi
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pansim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pansim_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_ffi.build())
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/pansim_b200.h but not exported"
    assert sorted(_ffi.EXPORTED) == declared


def test_version_and_formula_helpers_without_gpu():
    L = _ffi.lib()
    assert b"sm_100a" in L.pansim_version()
    # population.rs:822 and :828-830 (KATs H1, J1 of SURVEY.md section 4)
    assert L.pansim_core_distance(4, 9) == 0.4444444444444444
    assert L.pansim_acc_distance(3, 7, 2) == 0.4444444444444444


def test_config_struct_matches_header_size():
    cfg = _ffi.Config()
    _ffi.lib().pansim_config_init(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(_ffi.Config)
    assert cfg.genome_size_penalty == 0.99


def test_create_fails_loudly_without_gpu_or_bad_args():
    import torch
    cfg = pb.make_config(pb.Params(pop_size=8, core_size=100, pan_genes=60, core_genes=20))
    h = C.c_void_p()
    if not torch.cuda.is_available():
        rc = _ffi.lib().pansim_create(C.byref(cfg), C.byref(h))
        assert rc == -2 and not h.value
        assert b"no CPU fallback" in _ffi.lib().pansim_last_error(None)
        with pytest.raises(pb.PansimError):
            pb.Pansim(cfg)
    cfg.site_begin = 10         # not a multiple of PANSIM_SITE_ALIGN
    cfg.site_end = 100
    rc = _ffi.lib().pansim_create(C.byref(cfg), C.byref(h))
    assert rc == -1 and b"site_begin" in _ffi.lib().pansim_last_error(None)


def test_derive_matches_oracle_over_parameter_grid():
    # main.rs:259-287, 333-367: python host == C oracle, field by field
    rng = np.random.default_rng(0)
    for _ in range(200):
        kw = dict(pop_size=int(rng.integers(1, 5000)), core_size=int(rng.integers(1, 3_000_000)),
                  pan_genes=int(rng.integers(2, 9000)), core_mu=float(rng.random()),
                  HR_rate=float(rng.random() * 2), HGT_rate=float(rng.random() * 2),
                  rate_genes1=float(rng.random() * 3), rate_genes2=float(rng.random() * 2000),
                  prop_genes2=float(rng.random()), avg_gene_freq=float(rng.random() * 0.99 + 0.01))
        kw["core_genes"] = int(rng.integers(0, kw["pan_genes"]))
        d = pb.derive(pb.Params(**kw))
        o = ob.derive(ob.default_params(**kw))
        assert d.pan_size == o.pan_size
        assert d.avg_gene_freq_adj == o.avg_gene_freq_adj
        assert d.avg_gene_num == o.avg_gene_num
        assert d.n_core_mutations == o.n_core_mutations
        assert d.n_recombinations_core == o.n_recombinations_core
        assert (d.num_gene1_sites, d.num_gene2_sites) == (o.num_gene1_sites, o.num_gene2_sites)
        assert len(d.comp) == o.n_compartments
        for k in range(len(d.comp)):
            assert d.comp[k] == (o.comp_lo[k], o.comp_hi[k])
            assert d.n_pan_mutations[k] == o.n_pan_mutations[k]
            assert d.n_recombinations_pan[k] == o.n_recombinations_pan[k]


def test_validate_messages():
    assert pb.validate(pb.Params()) == []
    assert pb.validate(pb.Params(core_genes=7000))[0] == "core_genes must be less than or equal to pan_size"
    assert pb.validate(pb.Params(core_mu=1.5)) == ["core_mu must be between 0.0 and 1.0", "core_mu: 1.5"]


def test_fmt_f64_python_equals_oracle():
    rng = np.random.default_rng(2)
    xs = list(rng.random(300)) + list(rng.standard_normal(300) * 1e7) + [0.0, 1.0, 1e-7, 1e22, 299.99999999999994]
    for x in xs:
        assert pb.fmt_f64(float(x)) == ob.fmt_f64(float(x))


def test_standard_deviation_sequential_matches_oracle():
    rng = np.random.default_rng(3)
    v = rng.random(100000)
    assert pb.standard_deviation(v) == ob.standard_deviation(v)


def test_all_pairs_row_block_helpers():
    """Pure host logic of the exact all-pairs mode: pair counts of a row block and the row-block
    partition used by Pansim.iter_all_pairs (no device needed)."""
    from pansim_b200.population import Pansim

    class Stub:
        N = 57
    s = Stub()
    for b, e in [(0, 57), (0, 1), (56, 57), (10, 10), (3, 29)]:
        assert Pansim.pairs_in_rows(s, b, e) == sum(57 - 1 - i for i in range(b, e))
    blocks = list(Pansim.row_blocks(s, 200))
    assert blocks[0][0] == 0 and blocks[-1][1] == 56
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert sum(Pansim.pairs_in_rows(s, b, e) for b, e in blocks) == 57 * 56 // 2
    assert all(Pansim.pairs_in_rows(s, b, e) <= 200 or e == b + 1 for b, e in blocks)
    s.N = 1
    assert list(Pansim.row_blocks(s, 10)) == []
