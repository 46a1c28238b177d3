"""Shared helpers for the parity tests."""
import numpy as np

import pansim_b200 as pb
from oracle import binding as ob


def random_state(rng, N, L, G, p_gene=0.3):
    core = (1 << rng.integers(0, 4, (N, L))).astype(np.uint8)
    acc = (rng.random((N, G)) < p_gene).astype(np.uint8)
    return core, acc


def small_params(**kw):
    base = dict(pop_size=48, core_size=30000, pan_genes=500, core_genes=100, n_gen=3,
                max_distances=300, seed=1)
    base.update(kw)
    return pb.Params(**base)


def sample_pairs(rng, N, P):
    r1 = rng.integers(0, N, P).astype(np.uint32)
    r2 = rng.integers(0, N - 1, P).astype(np.uint32)
    r2 = (r2 + (r2 >= r1)).astype(np.uint32)
    return r1, r2


def oracle_apply_gpu_events(ev, parents, ocore: ob.Population, opan: ob.Population):
    """Replay one generation of GPU-drawn events on the oracle with the reference's
    operator order and store semantics (main.rs:445-464, population.rs:508, 537, 745)."""
    ocore.next_generation(parents)
    opan.next_generation(parents)
    o = np.lexsort((ev["core_mut_seq"], ev["core_mut_row"]))
    ocore.apply_core_writes(ev["core_mut_row"][o], ev["core_mut_site"][o], ev["core_mut_allele"][o])
    fr, fg = np.nonzero(ev["acc_flip_mask"])
    opan.apply_acc_flips(fr, fg)
    o = np.lexsort((ev["hr_seq"], ev["hr_recipient"]))
    ocore.apply_core_writes(ev["hr_recipient"][o], ev["hr_locus"][o], ev["hr_value"][o])
    gr, gg = np.nonzero(ev["acc_gain_mask"])
    opan.apply_acc_sets(gr, gg)
