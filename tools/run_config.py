#!/usr/bin/env python
"""Run the generation step / distance pass on an arbitrary shape and print timings.
    python tools/run_config.py --pop_size 10000 --core_size 5000000 --pan_genes 20000 --gens 5 --pairs 100000
Synthetic clonal start (set_initial) + warm-up generations, device-resident timing."""
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
    for k, v in pb.Params().as_dict().items():
        if isinstance(v, bool):
            ap.add_argument(f"--{k}", action="store_true")
        elif k != "outpref":
            ap.add_argument(f"--{k}", type=type(v), default=v)
    ap.add_argument("--gens", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=100000)
    ns = vars(ap.parse_args())
    gens, warm, pairs = ns.pop("gens"), ns.pop("warm"), ns.pop("pairs")
    p = pb.Params(**ns)
    d = pb.derive(p)
    rng = np.random.default_rng(p.seed)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
    r1 = rng.integers(0, p.pop_size, pairs).astype(np.uint32)
    r2 = ((r1 + 1 + rng.integers(0, p.pop_size - 1, pairs)) % p.pop_size).astype(np.uint32)
    t0 = time.perf_counter()
    with pb.Pansim.from_params(p) as sim:
        sim.set_initial(core_row, acc_row)
        sim.set_selection(np.zeros(d.pan_size))
        t_init = time.perf_counter() - t0
        sim.run_generations(0, warm)
        sim.run_generations(warm, gens)
        tm = sim.timing()
        info = sim.info()
        sim.pair_counts(r1, r2)
        cd, it, un = sim.pair_counts(r1, r2)
        tp = sim.timing()
        ms = tm.total_ms / gens
        core_bytes = 2 * p.pop_size * ((p.core_size + 3) // 4)
        out = dict(shape=dict(N=p.pop_size, L=p.core_size, G=d.pan_size), init_s=round(t_init, 2),
                   ms_per_generation=ms, generations_per_s=1e3 / ms, core_step_ms=tm.core_step_ms / gens,
                   core_step_GBps=core_bytes / (tm.core_step_ms / gens * 1e-3) / 1e9,
                   acc_ms=tm.acc_step_ms / gens, select_ms=tm.select_ms / gens,
                   pair_ms=tp.pair_core_ms + tp.pair_acc_ms, pairs_per_s=pairs / ((tp.pair_core_ms + tp.pair_acc_ms) * 1e-3),
                   mean_core_diff=float(cd.mean()), state_GB=2 * info.core_state_bytes / 1e9)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
