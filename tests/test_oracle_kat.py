"""Known-answer tests that pin the oracle.

The reference ships no tests; these KATs are hand-derived from its formulas
(file:line cited per test, relative to /root/reference/pansim/src/).
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import binding as ob


def test_H1_core_hamming():
    # distances.rs:22-52 (8-byte chunk + 1-byte tail), population.rs:817-822
    r1 = np.array([1, 2, 4, 8, 1, 2, 4, 8, 1], np.uint8)
    r2 = np.array([1, 4, 4, 2, 8, 2, 4, 8, 2], np.uint8)
    h = ob.lib().ora_hamming_bitwise_fast(r1, r2, 9)
    assert h == 8
    assert h // 2 == int((r1 != r2).sum()) == 4        # population.rs:819 cross-check
    pop = ob.Population(np.stack([r1, r2]), core=True)
    d = pop.pairwise_distances([0], [1])
    assert d[0] == 4 / 9 == 0.4444444444444444
    assert pop.pair_counts([0], [1])[0] == 4


def test_J1_accessory_jaccard():
    # distances.rs:55-77, population.rs:824-830
    x = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1], np.uint8)
    y = np.array([1, 1, 0, 1, 0, 0, 0, 0, 1, 0], np.uint8)
    i, u = C.c_uint32(), C.c_uint32()
    ob.lib().ora_jaccard_distance_fast(x, y, 10, C.byref(i), C.byref(u))
    assert (i.value, u.value) == (3, 7)
    i2, u2 = C.c_uint32(), C.c_uint32()
    ob.lib().ora_jaccard_distance_naive(x, y, 10, C.byref(i2), C.byref(u2))   # population.rs:32-48
    assert (i2.value, u2.value) == (3, 7)
    pop = ob.Population(np.stack([x, y]), core=False, core_genes=2)
    d = pop.pairwise_distances([0], [1])
    assert d[0] == 1.0 - (5.0 / 9.0) == 0.4444444444444444


def test_S1_standard_deviation():
    # population.rs:87-94 returns (std, mean), population variance
    s, m = ob.standard_deviation([1.0, 2.0, 3.0, 4.0])
    assert s == 1.118033988749895
    assert m == 2.5


def test_F1_parent_weights():
    # population.rs:282-437
    s = np.array([0.5, -0.2, 0.0, 0.1])
    rows = np.array([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], np.uint8)
    pop = ob.Population(rows, core=False, core_genes=0)
    avgdist = np.array([0.2, 0.5, 0.4])
    w, ng, lf = pop.selection_weights(2, avgdist, s, no_control=False, penalty=0.99,
                                      competition_strength=0.5)
    assert ng.tolist() == [2, 3, 2]
    np.testing.assert_allclose(lf, [math.log(1.5) + math.log(0.8), math.log(0.8) + math.log(1.1),
                                    math.log(1.5) + math.log(1.1)], rtol=1e-15)
    a = np.array([0.32171581769436997, 0.23592493297587133, 0.44235924932975873])
    b = np.array([0.33444816053511706, 0.33110367892976594, 0.33444816053511706])
    c = np.array([0.25029081336802045, 0.395744523829532, 0.3539646628024476])
    np.testing.assert_allclose(w, a * b * c, rtol=1e-13)
    np.testing.assert_allclose(w / w.sum(), [0.24435237883241887, 0.28049375497104545,
                                             0.47515386619653566], rtol=1e-13)
    # closed form: softmax(l + d + cs*ln D)
    logw = lf + (ng - 2) * math.log(0.99) + 0.5 * np.log(avgdist)
    sm = np.exp(logw - logw.max())
    np.testing.assert_allclose(w / w.sum(), sm / sm.sum(), rtol=1e-13)


def test_F_lethal_gene_sets_logfit_zero():
    # population.rs:312-318: s_j == -1 and present -> ln(0) = -inf -> log_sum := 0.0
    s = np.array([-1.0, 0.3])
    rows = np.array([[1, 1], [0, 1]], np.uint8)
    pop = ob.Population(rows, core=False)
    _, _, lf = pop.selection_weights(1, np.ones(2), s, no_control=True)
    assert lf[0] == 0.0
    assert lf[1] == math.log(1.3)


def test_W_all_zero_weights_become_ones():
    # population.rs:403, 435-437
    s = np.zeros(3)
    rows = np.array([[1, 1, 1], [0, 0, 0]], np.uint8)
    pop = ob.Population(rows, core=False)
    # the genome-size softmax underflows row 0 to exactly 0, the competition softmax row 1
    w, _, _ = pop.selection_weights(0, np.array([1.0, 1e-300]), s, no_control=False,
                                    penalty=1e-300, competition_strength=5.0)
    assert w.tolist() == [1.0, 1.0]


def test_P1_derived_parameters_defaults():
    # main.rs:259-287, 333-367
    d = ob.derive(ob.default_params())
    assert d.pan_size == 4000
    assert d.avg_gene_freq_adj == 0.25
    assert d.avg_gene_num == 1000
    assert d.n_core_mutations == 60000.0
    assert d.n_recombinations_core == 3000.0
    assert d.n_recombinations_pan_total == 3000.0
    assert (d.num_gene1_sites, d.num_gene2_sites) == (3600, 400)
    assert d.n_compartments == 2
    assert list(d.comp_lo) == [0, 3600] and list(d.comp_hi) == [3600, 4000]
    assert list(d.n_pan_mutations) == [3600.0, 400000.0]
    assert d.n_recombinations_pan[0] == 2700.0
    assert d.n_recombinations_pan[1] == 299.99999999999994


def test_P_validation_blocks():
    # main.rs:194-247
    L = ob.lib()
    assert L.ora_validate(C.byref(ob.default_params())) == 0
    assert L.ora_validate(C.byref(ob.default_params(core_genes=7000))) == 1
    assert L.ora_validate(C.byref(ob.default_params(HR_rate=-1.0))) == 2
    assert L.ora_validate(C.byref(ob.default_params(pos_lambda=0.0))) == 3
    assert L.ora_validate(C.byref(ob.default_params(rate_genes1=-0.5))) == 4
    assert L.ora_validate(C.byref(ob.default_params(prop_genes2=1.5))) == 5
    assert L.ora_validate(C.byref(ob.default_params(n_gen=0))) == 6
    assert L.ora_validate(C.byref(ob.default_params(core_mu=1.5))) == 7
    assert L.ora_validate(C.byref(ob.default_params(avg_gene_freq=0.0))) == 8


def test_average_distance_matches_formula_and_min_positive():
    # population.rs:753-784, 114-151
    rng = np.random.default_rng(5)
    m = (rng.random((7, 37)) < 0.3).astype(np.uint8)
    m[3] = m[2]
    pop = ob.Population(m, core=False, core_genes=3)
    got = pop.average_distance()
    for i in range(7):
        acc = 0.0
        for j in range(7):
            if j == i:
                continue
            inter = int((m[i] & m[j]).sum())
            uni = int((m[i] | m[j]).sum())
            acc += 1.0 - ((inter + 0.0 + 3.0) / (uni + 0.0 + 3.0))
        assert got[i] == acc / 6
    same = ob.Population(np.ones((4, 5), np.uint8), core=False, core_genes=1)
    assert (same.average_distance() == np.finfo(np.float64).tiny).all()   # f64::MIN_POSITIVE


def test_gene_frequencies_order():
    # population.rs:840-863: accessory first, then core_genes ones
    m = np.array([[1, 0, 1], [1, 0, 0], [1, 1, 0], [1, 0, 0]], np.uint8)
    pop = ob.Population(m, core=False, core_genes=2)
    assert pop.gene_frequencies().tolist() == [1.0, 0.25, 0.25, 1.0, 1.0]
    assert pop.gene_counts().tolist() == [4, 1, 1]
    assert pop.calc_gene_freq() == pytest.approx((2 / 3 + 1 / 3 + 2 / 3 + 1 / 3) / 4, rel=1e-15)


def test_int_to_base():
    L = ob.lib()
    assert [L.ora_int_to_base(v) for v in (1, 2, 4, 8, 3)] == [b"A", b"C", b"G", b"T", b"N"]


@pytest.mark.parametrize("x,s", [
    (1.0, "1"), (0.0, "0"), (-0.0, "-0"), (0.5, "0.5"), (0.4444444444444444, "0.4444444444444444"),
    (1e-5, "0.00001"), (1e21, "1000000000000000000000"), (299.99999999999994, "299.99999999999994"),
    (2.2250738585072014e-308, "0." + "0" * 307 + "22250738585072014"),
    (float("nan"), "NaN"), (float("inf"), "inf"), (-float("inf"), "-inf"),
    (0.1 + 0.2, "0.30000000000000004"), (123456.789, "123456.789"), (-2.5, "-2.5"),
])
def test_fmt_f64_rust_display(x, s):
    # Rust `{}` for f64: shortest round-trip digits, positional notation
    assert ob.fmt_f64(x) == s


def test_fmt_f64_round_trip_random():
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.random(200), rng.standard_normal(200) * 1e6, rng.random(50) * 1e-9]):
        assert float(ob.fmt_f64(float(x))) == float(x)
        assert "e" not in ob.fmt_f64(float(x))


def test_poisson_sampler_moments():
    r = ob.make_rng(11)
    L = ob.lib()
    for mean in (0.3, 3.2, 12.8, 300.0, 60000.0):
        xs = np.array([L.ora_poisson(C.byref(r), mean) for _ in range(20000)], np.float64)
        se = math.sqrt(mean / len(xs))
        assert abs(xs.mean() - mean) < 5 * se
        assert abs(xs.var() / mean - 1.0) < 0.06


def test_rng_below_uniform_and_bounds():
    r = ob.make_rng(3)
    L = ob.lib()
    xs = np.array([L.ora_rng_below(C.byref(r), 7) for _ in range(70000)])
    assert xs.min() == 0 and xs.max() == 6
    counts = np.bincount(xs, minlength=7)
    chi2 = ((counts - 10000) ** 2 / 10000).sum()
    assert chi2 < 30


def test_weighted_index_semantics():
    # rand 0.8.5 WeightedIndex: zero-weight entries are never drawn; bad weights error
    r = ob.make_rng(2)
    out = ob.weighted_index_sample([0.0, 2.0, 0.0, 1.0, 0.0], r, 30000)
    counts = np.bincount(out, minlength=5)
    assert counts[0] == counts[2] == counts[4] == 0
    assert abs(counts[1] / 30000 - 2 / 3) < 0.01
    with pytest.raises(ValueError):
        ob.weighted_index_sample([0.0, 0.0], r, 1)
    with pytest.raises(ValueError):
        ob.weighted_index_sample([1.0, float("nan")], r, 1)
    with pytest.raises(ValueError):
        ob.weighted_index_sample([1.0, -1.0], r, 1)


def test_sample_pairs_no_self_pairs():
    # main.rs:413-427
    r1, r2 = ob.sample_pairs(5, 5000, ob.make_rng(4))
    assert (r1 != r2).all() and r1.max() == 4 and r2.max() == 4 and r2.min() == 0


def test_population_new_clonal():
    # population.rs:199-230
    rng = ob.make_rng(9)
    core = ob.new_population(6, 500, 4, True, 0.25, rng, 10)
    assert set(np.unique(core.m)) <= {1, 2, 4, 8}
    assert (core.m == core.m[0]).all()
    pan = ob.new_population(6, 4000, 2, False, 0.25, rng, 10)
    assert set(np.unique(pan.m)) <= {0, 1} and (pan.m == pan.m[0]).all()
    assert abs(pan.m[0].mean() - 0.25) < 0.03


def test_selection_coefficients_ranges():
    # main.rs:289-319
    p = ob.default_params(prop_positive=0.1)
    s = ob.selection_coefficients(p, 20000, ob.make_rng(0))
    assert (s >= -1.0).all()
    assert abs((s > 0).mean() - 0.1) < 0.01
    assert abs(s[s > 0].mean() - 0.1) < 0.01          # Exp(10) mean
    neutral = ob.selection_coefficients(ob.default_params(), 100, ob.make_rng(0))
    assert (neutral == 0).all()
