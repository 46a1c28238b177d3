#!/usr/bin/env python
"""Time the exact all-pairs distance mode (Pansim.iter_all_pairs) on a synthetic population.
    python tools/all_pairs_bench.py --pop_size 4000 --core_size 1200000 [--chunk_pairs 4000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pansim_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop_size", type=int, default=4000)
    ap.add_argument("--core_size", type=int, default=1_200_000)
    ap.add_argument("--pan_genes", type=int, default=6000)
    ap.add_argument("--chunk_pairs", type=int, default=4_000_000)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    p = pb.Params(pop_size=a.pop_size, core_size=a.core_size, pan_genes=a.pan_genes)
    d = pb.derive(p)
    rng = np.random.default_rng(0)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
    with pb.Pansim.from_params(p) as sim:
        sim.set_initial(core_row, acc_row)
        sim.set_selection(np.zeros(d.pan_size))
        sim.run_generations(0, a.warm)                       # diversify the clonal start
        t0 = time.perf_counter()
        n = 0
        kern_ms = 0.0
        s_core = 0
        for ii, jj, cd, it, un in sim.iter_all_pairs(a.chunk_pairs, with_indices=False):
            t = sim.timing()
            kern_ms += t.pair_core_ms + t.pair_acc_ms
            n += len(cd)
            s_core += int(cd.sum())
        wall = time.perf_counter() - t0
    assert n == p.pop_size * (p.pop_size - 1) // 2
    print(json.dumps(dict(N=p.pop_size, L=p.core_size, pairs=n, wall_s=wall, kernel_s=kern_ms * 1e-3,
                          pairs_per_s_wall=n / wall, pairs_per_s_kernel=n / (kern_ms * 1e-3), mean_core_diff=s_core / n)))


if __name__ == "__main__":
    main()
