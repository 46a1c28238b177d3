#!/usr/bin/env python
"""Distance pass alone at the cfg1 shape: device ms of the core / accessory kernels (CUDA events in the library)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pansim_b200 as pb  # noqa: E402

N, L, P = int(os.environ.get("N", 1000)), int(os.environ.get("L", 1_200_000)), int(os.environ.get("P", 100_000))
p = pb.Params(pop_size=N, core_size=L, pan_genes=6000, core_genes=2000, max_distances=P, seed=0)
d = pb.derive(p)
rng = np.random.default_rng(0)
core_row = (1 << rng.integers(0, 4, L)).astype(np.uint8)
acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
r1 = rng.integers(0, N, P).astype(np.uint32)
r2 = ((r1 + 1 + rng.integers(0, N - 1, P)) % N).astype(np.uint32)
with pb.Pansim.from_params(p) as sim:
    sim.set_initial(core_row, acc_row)
    sim.run_generations(0, 5)
    sim.pair_counts(r1, r2)
    ms = []
    for _ in range(int(os.environ.get("REPS", 8))):
        sim.pair_counts(r1, r2)
        t = sim.timing()
        ms.append((t.pair_core_ms, t.pair_acc_ms, t.total_ms))
    m = np.median(np.array(ms), axis=0)
    print("core %.3f ms  acc %.3f ms  total %.3f ms  -> %.3g pairs/s" % (m[0], m[1], m[2], P / (m[2] * 1e-3)))
