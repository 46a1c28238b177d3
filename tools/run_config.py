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
    ap.add_argument("--shard", action="store_true",
                    help="under torchrun: strong scaling, each rank takes a column shard of the SAME alignment")
    ns = vars(ap.parse_args())
    gens, warm, pairs, shard = ns.pop("gens"), ns.pop("warm"), ns.pop("pairs"), ns.pop("shard")
    if shard:
        return main_sharded(pb.Params(**ns), gens, warm, pairs)
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
                   ms_per_generation=ms, generations_per_s=1e3 / ms, core_step_ms=tm.core_step_ms / gens, core_hr_ms=tm.core_hr_ms / gens,
                   core_step_GBps=core_bytes / (tm.core_step_ms / gens * 1e-3) / 1e9,
                   acc_ms=tm.acc_step_ms / gens, select_ms=tm.select_ms / gens,
                   pair_ms=tp.pair_core_ms + tp.pair_acc_ms, pairs_per_s=pairs / ((tp.pair_core_ms + tp.pair_acc_ms) * 1e-3),
                   mean_core_diff=float(cd.mean()), state_GB=2 * info.core_state_bytes / 1e9)
        print(json.dumps(out))


def main_sharded(p, gens, warm, pairs):
    import torch
    import torch.distributed as dist
    from pansim_b200.sharding import ShardedPansim
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = pb.derive(p)
    rng = np.random.default_rng(p.seed)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < d.avg_gene_freq_adj).astype(np.uint8)
    r1 = rng.integers(0, p.pop_size, pairs).astype(np.uint32)
    r2 = ((r1 + 1 + rng.integers(0, p.pop_size - 1, pairs)) % p.pop_size).astype(np.uint32)
    sim = ShardedPansim(p, rank, world, device=local)
    sim.set_initial(core_row, acc_row)
    sim.set_selection(np.zeros(d.pan_size))
    sim.run_generations(0, warm)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run_generations(warm, gens)
    tm = sim.timing()
    dist.barrier(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([tm.total_ms, tm.core_step_ms, tm.acc_step_ms, tm.select_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sim.pair_counts(r1, r2)
    t1 = time.perf_counter()
    cd, it, un = sim.pair_counts(r1, r2)           # partial counts + NCCL all-reduce
    dist.barrier(); torch.cuda.synchronize()
    pair_s = time.perf_counter() - t1
    if rank == 0:
        ms = float(t[0]) / gens
        print(json.dumps(dict(world=world, shape=dict(N=p.pop_size, L=p.core_size, G=d.pan_size), ms_per_generation=ms,
                              generations_per_s=1e3 / ms, wall_ms_per_generation=1e3 * wall / gens,
                              core_step_ms=float(t[1]) / gens, acc_ms=float(t[2]) / gens, select_ms=float(t[3]) / gens,
                              pair_pass_ms_incl_host=1e3 * pair_s, mean_core_diff=float(cd.mean()))))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
