"""Run under torchrun (one rank per GPU): column-sharded run vs a single-context run.
Every rank checks that the all-reduced pair counts, the accessory state and the parents
equal those of an unsharded context (results are independent of the GPU count)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pansim_b200 as pb  # noqa: E402
from pansim_b200.sharding import ShardedPansim  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = pb.Params(pop_size=96, core_size=8192 * 7 + 123, pan_genes=700, core_genes=200, n_gen=4, max_distances=500,
                  seed=5, prop_positive=0.1, competition_strength=0.3, HR_rate=0.5)
    d = pb.derive(p)
    rng = np.random.default_rng(1)
    core_row = (1 << rng.integers(0, 4, p.core_size)).astype(np.uint8)
    acc_row = (rng.random(d.pan_size) < 0.25).astype(np.uint8)
    sel = rng.normal(0, 0.05, d.pan_size)
    r1 = rng.integers(0, p.pop_size, p.max_distances).astype(np.uint32)
    r2 = ((r1 + 1 + rng.integers(0, p.pop_size - 1, p.max_distances)) % p.pop_size).astype(np.uint32)

    sh = ShardedPansim(p, rank, world, device=local)
    whole = pb.Pansim.from_params(p, device=local)
    for s in (sh, whole):
        s.set_initial(core_row, acc_row)
        s.set_selection(sel)
        s.run_generations(0, p.n_gen)
    cd, it, un = sh.pair_counts(r1, r2)                  # ncclAllReduce of the partial core counts inside the library
    wcd, wit, wun = whole.pair_counts(r1, r2)
    b, e = sh.shards[rank]
    ok = (cd == wcd).all() and (it == wit).all() and (un == wun).all()
    ok = ok and (sh.download_acc() == whole.download_acc()).all() and (sh.parents() == whole.parents()).all()
    ok = ok and (sh.download_core() == whole.download_core()[:, b:e]).all()
    acd, ait, aun = sh.pair_counts_rows(10, 40)          # exact all-pairs block, all-reduced as well
    ok = ok and sh.pair_stats(r1, r2) == whole.pair_stats(r1, r2)
    wacd, wait_, waun = whole.pair_counts_rows(10, 40)
    ok = ok and (acd == wacd).all() and (ait == wait_).all() and (aun == waun).all()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(t.item()) == 1 else "MULTI_GPU_MISMATCH", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
