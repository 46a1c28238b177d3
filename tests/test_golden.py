"""Golden fixture tests (tests/golden/replay_small.npz, made by tests/golden/make_golden.py).
CPU: the oracle's replay-apply reproduces the committed final state from the committed
events, and its reductions reproduce the committed outputs. GPU: the CUDA replay path and
distance pass reproduce the same committed outputs bit for bit."""
import os

import numpy as np
import pytest

import pansim_b200 as pb
from oracle import binding as ob

G_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_small.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(G_PATH))


def test_oracle_replay_reproduces_golden(gold):
    N, L, G, CG = gold["meta"]
    core, pan = ob.Population(gold["core0"], True, CG), ob.Population(gold["acc0"], False, CG)
    for g in range(2):
        core.next_generation(gold[f"parents{g}"])
        pan.next_generation(gold[f"parents{g}"])
        core.apply_core_writes(gold[f"core_mut_row{g}"], gold[f"core_mut_site{g}"], gold[f"core_mut_allele{g}"])
        pan.apply_acc_flips(gold[f"acc_flip_row{g}"], gold[f"acc_flip_gene{g}"])
        core.apply_core_writes(gold[f"hr_recipient{g}"], gold[f"hr_locus{g}"], gold[f"hr_value{g}"])
        pan.apply_acc_sets(gold[f"hgt_recipient{g}"], gold[f"hgt_gene{g}"])
    assert (core.m == gold["core_final"]).all() and (pan.m == gold["acc_final"]).all()
    assert (core.pair_counts(gold["r1"], gold["r2"]) == gold["core_diff"]).all()
    i, u = pan.pair_counts(gold["r1"], gold["r2"])
    assert (i == gold["inter"]).all() and (u == gold["uni"]).all()
    assert (core.pairwise_distances(gold["r1"], gold["r2"]) == gold["core_dist"]).all()
    assert (pan.pairwise_distances(gold["r1"], gold["r2"]) == gold["acc_dist"]).all()
    assert (pan.average_distance() == gold["avg_dist"]).all()
    assert (pan.gene_frequencies() == gold["gene_freqs"]).all()


def test_golden_events_have_reference_structure(gold):
    # HR values are the donor's allele at the locus (one-hot), recipients differ from donors,
    # SNP alleles are never A (population.rs:531), HGT only moves present genes (:636-655)
    for g in range(2):
        assert set(np.unique(gold[f"core_mut_allele{g}"])) <= {2, 4, 8}
        assert set(np.unique(gold[f"hr_value{g}"])) <= {1, 2, 4, 8}
        assert (gold[f"hr_recipient{g}"] != gold[f"hr_donor{g}"]).all()
        assert (gold[f"hgt_recipient{g}"] != gold[f"hgt_donor{g}"]).all()
        assert len(gold[f"core_mut_row{g}"]) > 5000 and len(gold[f"hr_recipient{g}"]) > 3000


@pytest.mark.gpu
def test_gpu_replay_reproduces_golden(gold):
    N, L, G, CG = (int(x) for x in gold["meta"])
    p = pb.Params(pop_size=N, core_size=L, pan_genes=G + CG, core_genes=CG, competition_strength=1.0)
    with pb.Pansim.from_params(p) as sim:
        sim.upload(gold["core0"], gold["acc0"])
        for g in range(2):
            sim.step_replay(gold[f"parents{g}"],
                            core_mut=(gold[f"core_mut_row{g}"], gold[f"core_mut_site{g}"], gold[f"core_mut_allele{g}"]),
                            acc_flip=(gold[f"acc_flip_row{g}"], gold[f"acc_flip_gene{g}"]),
                            hr=(gold[f"hr_recipient{g}"], gold[f"hr_locus{g}"], gold[f"hr_value{g}"]),
                            hgt=(gold[f"hgt_recipient{g}"], gold[f"hgt_gene{g}"]))
        assert (sim.download_core() == gold["core_final"]).all()
        assert (sim.download_acc() == gold["acc_final"]).all()
        cd, it, un = sim.pair_counts(gold["r1"], gold["r2"])
        assert (cd == gold["core_diff"]).all() and (it == gold["inter"]).all() and (un == gold["uni"]).all()
        core_d, acc_d = sim.pairwise_distances(gold["r1"], gold["r2"])
        assert (core_d == gold["core_dist"]).all() and (acc_d == gold["acc_dist"]).all()
        assert (sim.average_distance() == gold["avg_dist"]).all()
        assert (sim.gene_frequencies() == gold["gene_freqs"]).all()
