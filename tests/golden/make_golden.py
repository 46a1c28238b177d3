"""Generates tests/golden/replay_small.npz: a two-generation replay fixture.

The Rust reference cannot be built in this image (no cargo), and it ships no golden
vectors, so this fixture is produced by the ORACLE's restatement of the reference's
event loops (oracle/pansim_oracle.c: ora_mutate_alleles / ora_recombine,
population.rs:467-751). It pins the oracle's deterministic behaviour and gives the
GPU replay path a committed input/output pair. Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

N, L, G, CG = 16, 3000, 96, 7
rng = np.random.default_rng(20261018)
core0 = (1 << rng.integers(0, 4, (N, L))).astype(np.uint8)
acc0 = (rng.random((N, G)) < 0.3).astype(np.uint8)
core, pan = ob.Population(core0, True, CG), ob.Population(acc0, False, CG)
main = ob.make_rng(42)
out = dict(core0=core0, acc0=acc0, meta=np.array([N, L, G, CG]))
for gen in range(2):
    parents = rng.integers(0, N, N).astype(np.uint32)
    core.next_generation(parents)
    pan.next_generation(parents)
    ev = ob.EventLog()
    core.mutate_alleles([400.0], [(0, L)], 42, gen, ev)
    pan.mutate_alleles([20.0, 90.0], [(0, 80), (80, G)], 42, gen, ev)
    core.recombine([250.0], [(0, L)], main, 42, gen, ev)
    pan.recombine([12.0, 5.0], [(0, 80), (80, G)], main, 42, gen, ev)
    out[f"parents{gen}"] = parents
    for k, v in ev.arrays().items():
        out[f"{k}{gen}"] = v
r1 = rng.integers(0, N, 200).astype(np.uint32)
r2 = ((r1 + 1 + rng.integers(0, N - 1, 200)) % N).astype(np.uint32)
out.update(core_final=core.m, acc_final=pan.m, r1=r1, r2=r2, core_diff=core.pair_counts(r1, r2))
inter, uni = pan.pair_counts(r1, r2)
out.update(inter=inter, uni=uni, core_dist=core.pairwise_distances(r1, r2), acc_dist=pan.pairwise_distances(r1, r2),
           avg_dist=pan.average_distance(), gene_freqs=pan.gene_frequencies())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "replay_small.npz"), **out)
print("wrote replay_small.npz")
