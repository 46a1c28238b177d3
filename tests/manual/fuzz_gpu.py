#!/usr/bin/env python
"""Randomised cross-checks on a GPU (not part of the pytest suite: run by hand under gpurun).
For random shapes and rates: (1) deferred == immediate recombination after several generations,
(2) competition term / fitness / gene counts bit-exact against the oracle, (3) sampled and
all-pairs distance counts against the oracle, (4) generate-mode events replayed on the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pansim_b200 as pb  # noqa: E402
from oracle import binding as ob  # noqa: E402
from helpers import oracle_apply_gpu_events, random_state, sample_pairs  # noqa: E402


def one_case(rng, idx):
    N = int(rng.integers(2, int(os.environ.get('FUZZ_NMAX', '400'))))
    L = int(rng.integers(1, int(os.environ.get('FUZZ_LMAX', '60000'))))
    G = int(rng.integers(1, 3000))
    cg = int(rng.integers(0, 50))
    kw = dict(pop_size=N, core_size=L, pan_genes=G + cg, core_genes=cg, n_gen=3, seed=int(rng.integers(0, 1 << 30)),
              core_mu=float(rng.choice([0.0, 0.01, 0.05, 0.3, 0.9])), HR_rate=float(rng.choice([0.0, 0.05, 1.0, 4.0])),
              HGT_rate=float(rng.choice([0.0, 0.05, 1.0])), competition_strength=float(rng.choice([0.0, 0.5, 2.0])),
              prop_positive=float(rng.choice([-1.0, 0.1])))
    p = pb.Params(**kw)
    d = pb.derive(p)
    core, acc = random_state(rng, N, L, d.pan_size, float(rng.uniform(0.05, 0.95)))
    sel = rng.normal(0, 0.2, d.pan_size).clip(-0.95, None)
    if d.pan_size > 3 and rng.random() < 0.5:
        sel[int(rng.integers(0, d.pan_size))] = -1.0
    if rng.random() < 0.3:
        sel[:] = 0.0
    states = {}
    for defer in ("1", "0"):
        os.environ["PANSIM_HR_DEFER"] = defer
        with pb.Pansim.from_params(p) as sim:
            sim.upload(core, acc)
            sim.set_selection(sel)
            sim.run_generations(0, 3)
            states[defer] = (sim.download_core(), sim.download_acc(), sim.parents())
    os.environ.pop("PANSIM_HR_DEFER", None)
    for a, b in zip(states["0"], states["1"]):
        assert (a == b).all(), f"case {idx}: deferred != immediate {kw}"
    # batches through the chain graphs (and the --print_dist batch) == kernel-by-kernel launches
    rs1, rs2 = sample_pairs(rng, N, 500)
    gstates = {}
    for graph in ("1", "0"):
        os.environ["PANSIM_GRAPH"] = graph
        with pb.Pansim.from_params(p) as sim:
            sim.upload(core, acc)
            sim.set_selection(sel)
            sim.run_generations(0, 9)
            st = sim.run_generations_stats(9, 8, rs1, rs2)
            gstates[graph] = (sim.download_core(), sim.download_acc(), sim.parents(), st)
    os.environ.pop("PANSIM_GRAPH", None)
    for a, b in zip(gstates["0"], gstates["1"]):
        assert (a == b).all(), f"case {idx}: graph batch != plain launches {kw}"
    ocore, opan = ob.Population(core.copy(), True, cg), ob.Population(acc.copy(), False, cg)
    with pb.Pansim.from_params(p) as sim:
        sim.upload(core, acc)
        sim.set_selection(sel)
        if N >= 2:
            oavg = opan.average_distance()
            assert (sim.average_distance() == oavg).all(), f"case {idx}: average_distance {kw}"
            sim.sample_indices(0, oavg)
            w, ng, lf = sim.weights()
            ow, ong, olf = opan.selection_weights(d.avg_gene_num, oavg, sel, False, p.genome_size_penalty, p.competition_strength)
            assert (ng == ong).all() and (lf == olf).all(), f"case {idx}: fitness {kw}"
            if np.isfinite(ow).all() and ow.sum() > 0:
                np.testing.assert_allclose(w / w.sum(), ow / ow.sum(), rtol=1e-10, atol=1e-300)
        r1, r2 = sample_pairs(rng, N, int(rng.choice([7, 300, 3000, 20000])))      # few pairs: groups only; many: TMA tile batches
        cd, it, un = sim.pair_counts(r1, r2)
        assert (cd == ocore.pair_counts(r1, r2)).all(), f"case {idx}: pair core {kw}"
        oi, ou = opan.pair_counts(r1, r2)
        assert (it == oi).all() and (un == ou).all(), f"case {idx}: pair acc {kw}"
        # on-device statistics == the oracle's left-to-right sums over the same distances
        core_d, acc_d = sim.distances_from_counts(cd, it, un)
        sc, ac = ob.standard_deviation(core_d)
        sa, aa = ob.standard_deviation(acc_d)
        assert sim.pair_stats(r1, r2) == (ac, sc, aa, sa), f"case {idx}: pair_stats {kw}"
        i0 = int(rng.integers(0, N))
        i1 = int(rng.integers(i0, N + 1))
        cd, it, un = sim.pair_counts_rows(i0, i1)
        ii = np.repeat(np.arange(i0, i1, dtype=np.uint32), [max(0, N - 1 - i) for i in range(i0, i1)]).astype(np.uint32)
        if len(ii):
            jj = np.concatenate([np.arange(i + 1, N, dtype=np.uint32) for i in range(i0, i1)])
            assert (cd == ocore.pair_counts(ii, jj)).all(), f"case {idx}: all-pairs core {kw}"
        n_ev = N * L * kw['core_mu'] * max(1.0, kw['HR_rate'])
        if n_ev > 2.5e6:
            return kw                                  # too many events for the dump buffers of this check
        sim.enable_event_dump(4_000_000)
        for gen in range(2):
            sim.step(gen)
            ev = sim.fetch_event_dump()
            oracle_apply_gpu_events(ev, sim.parents(), ocore, opan)
            assert (sim.download_core() == ocore.m).all(), f"case {idx}: replay core gen {gen} {kw}"
            assert (sim.download_acc() == opan.m).all(), f"case {idx}: replay acc gen {gen} {kw}"
    return kw


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    for i in range(n):
        kw = one_case(rng, i)
        print(f"case {i} ok: N={kw['pop_size']} L={kw['core_size']} G={kw['pan_genes'] - kw['core_genes']} "
              f"mu={kw['core_mu']} HR={kw['HR_rate']} HGT={kw['HGT_rate']} comp={kw['competition_strength']}", flush=True)
    print("FUZZ_OK", n)


if __name__ == "__main__":
    main()
