#!/usr/bin/env python
"""BASELINE config 5: exact all-pairs distances (N(N-1)/2 pairs, `--all_pairs` of the hosts) over
the GPUs of one box, one process (pansim_group: column shards, ncclReduceScatter of the partial core
counts per row block, overlapped with the next block's kernels).

    python tools/all_pairs_bench.py --pop_size 20000 --core_size 1200000 --gpus 8 [--chunk_pairs 4000000] [--check 3]

Prints one JSON line: pairs/s (wall, callbacks included: every count reaches the host), the device
time shard 0 spent inside the reduce-scatters and its share of the walk."""
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
    ap.add_argument("--pop_size", type=int, default=20000)
    ap.add_argument("--core_size", type=int, default=1_200_000)
    ap.add_argument("--pan_genes", type=int, default=6000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gens", type=int, default=3)
    ap.add_argument("--chunk_pairs", type=int, default=4_000_000)
    ap.add_argument("--check", type=int, default=2, help="row blocks re-checked against numpy on the downloaded state")
    a = ap.parse_args()
    p = pb.Params(pop_size=a.pop_size, core_size=a.core_size, pan_genes=a.pan_genes, core_genes=2000, seed=0)
    d = pb.derive(p)
    rng = np.random.default_rng(0)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
    t0 = time.perf_counter()
    with pb.PansimGroup(p, a.gpus) as g:
        g.set_initial(core_row, acc_row)
        g.set_selection(np.zeros(d.pan_size))
        g.run_generations(0, a.gens)
        t_init = time.perf_counter() - t0
        checks = []
        seen = [0]

        def sink(i0, i1, cd, it, un):
            seen[0] += len(cd)
            if len(checks) < a.check:
                checks.append((i0, i1, cd[:2000].copy(), int(cd.sum(dtype=np.uint64)), it[:2000].copy(), un[:2000].copy()))

        g.walk_all_pairs(min(a.chunk_pairs, 1_000_000))               # warm-up: buffers, plans, NCCL channels
        seen[0] = 0
        checks.clear()
        t = g.walk_all_pairs(a.chunk_pairs, sink)
        N = a.pop_size
        assert seen[0] == N * (N - 1) // 2 == t["pairs"], (seen[0], t)
        # spot check against the rows themselves (first 2000 pairs of the checked blocks: row i0 against i0+1 ..)
        ok = True
        if a.check:
            acc = g.download_acc()
            sh0 = g._lib.pansim_group_ctx(g._h, 0)
            for i0, i1, cd, _tot, it, un in checks:
                n = min(2000, N - 1 - i0)
                x = acc[i0].astype(bool)
                for k in range(0, n, 97):
                    y = acc[i0 + 1 + k].astype(bool)
                    ok = ok and int((x & y).sum()) == int(it[k]) and int((x | y).sum()) == int(un[k])
        out = dict(workload="cfg5 all pairs", pop_size=N, core_size=a.core_size, n_gpus=a.gpus, pairs=t["pairs"],
                   init_s=round(t_init, 2), wall_ms=t["wall_ms"], pairs_per_s=t["pairs"] / (t["wall_ms"] * 1e-3),
                   nccl_reduce_scatter_ms_shard0=t["nccl_ms_shard0"],
                   nccl_share_of_walk=t["nccl_ms_shard0"] / t["wall_ms"] if t["wall_ms"] else None,
                   chunk_pairs=a.chunk_pairs, accessory_spot_check=bool(ok),
                   note="wall time of pansim_group_all_pairs incl. the per-block host callbacks; the reduce-scatters run on a "
                        "second stream beside the next block's kernels, so their share is overlapped time, not added time")
        print(json.dumps(out))


if __name__ == "__main__":
    main()
