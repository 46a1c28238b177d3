#!/usr/bin/env python
"""Selection-chain kernel variants against the oracle, bit for bit (run under gpurun):
PANSIM_INTER_UMMA (3 / 2 = tcgen05 with in-CTA expansion, 64- / 128-byte swizzle; 1 = TMA-fed; 0 = mma.sync),
PANSIM_AVG_RCP (reciprocal-table producer/consumer distance kernel vs IEEE division),
(the fitness kernel is checked along: one lane per row, row words staged in shared memory)."""
import itertools
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pansim_b200 as pb  # noqa: E402
from oracle import binding as ob  # noqa: E402
from helpers import random_state  # noqa: E402

SHAPES = [(1000, 6000, 2000, 0.25), (300, 900, 200, 0.5), (130, 37, 0, 0.9), (5, 40, 37, 0.3), (1025, 4229, 100, 0.02),
          (257, 300, 0, 1.0), (2, 64, 3, 0.0), (640, 12800, 0, 0.4)]


UMMA = sys.argv[1] if len(sys.argv) > 1 else "0123"


def main():
    rng = np.random.default_rng(7)
    bad = 0
    for (N, pan, cg, dens) in SHAPES:
        p = pb.Params(pop_size=N, core_size=8192, pan_genes=pan, core_genes=cg, n_gen=2, seed=3, prop_positive=0.1,
                      competition_strength=0.5)
        d = pb.derive(p)
        core, acc = random_state(rng, N, 8192, d.pan_size, dens)
        sel = rng.normal(0, 0.2, d.pan_size).clip(-0.95, None)
        if d.pan_size > 3:
            sel[int(rng.integers(0, d.pan_size))] = -1.0
        opan = ob.Population(acc.copy(), False, cg)
        oavg = opan.average_distance()
        ow, ong, olf = opan.selection_weights(d.avg_gene_num, oavg, sel, False, p.genome_size_penalty, p.competition_strength)
        for umma, lane, fit in itertools.product(UMMA, "012", "0"):
            os.environ["PANSIM_INTER_UMMA"] = umma
            os.environ["PANSIM_AVG_RCP"] = lane
            os.environ["PANSIM_FITNESS_MODE"] = fit
            with pb.Pansim.from_params(p) as sim:
                sim.upload(core, acc)
                sim.set_selection(sel)
                t0 = time.perf_counter()
                avg = sim.average_distance()
                t1 = time.perf_counter()
                try:
                    sim.sample_indices(0, oavg)
                except Exception as e:       # noqa: BLE001  (an all-zero weight vector is the reference's panic)
                    print("sample_indices:", e)
                w, ng, lf = sim.weights()
            ok_avg = bool((avg == oavg).all()) or bool(np.array_equal(avg, oavg, equal_nan=True))
            ok_fit = bool((ng == ong).all() and np.array_equal(lf, olf, equal_nan=True))
            tag = f"N={N} G={d.pan_size} cg={cg} dens={dens} umma={umma} lane={lane} fit={fit}"
            if not (ok_avg and ok_fit):
                bad += 1
                nbad = int((avg != oavg).sum())
                print("MISMATCH", tag, "avg ok", ok_avg, f"({nbad} rows differ)", "fitness ok", ok_fit, flush=True)
            else:
                print("ok", tag, f"{(t1 - t0) * 1e3:.2f} ms", flush=True)
    print("CHAIN_VARIANTS_OK" if not bad else f"CHAIN_VARIANTS_BAD {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
