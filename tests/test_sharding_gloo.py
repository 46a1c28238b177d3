"""world_size-2 gloo test (CPU) of the N>1 host logic: the column-shard plan and the
all-reduce of per-pair partial core counts. Partials are computed with the oracle on
each rank's column slice (oracle = checker); the sum must equal the whole-alignment count."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pansim_b200.sharding import SITE_ALIGN, allreduce_pair_counts, column_shards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_column_shards_cover_and_align():
    for L in (1, 8191, 8192, 8193, 1_200_000, 5_000_000):
        for w in (1, 2, 3, 4, 8):
            sh = column_shards(L, w)
            assert len(sh) == w and sh[0][0] == 0 and sh[-1][1] == L
            for (b, e), (b2, _) in zip(sh, sh[1:] + [(L, L)]):
                assert b <= e and e == b2
                if e > b:                      # empty trailing shards are (L, L)
                    assert b % SITE_ALIGN == 0
                    assert e == L or e % SITE_ALIGN == 0
            sizes = [e - b for b, e in sh if e > b]
            assert max(sizes) - min(sizes) < 2 * SITE_ALIGN or len(sizes) < w   # one region + the ragged tail


def _worker(rank, world, port, L, q):
    sys.path.insert(0, ROOT)
    from oracle import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                       # same state on every rank
    N, P = 12, 64
    core = (1 << rng.integers(0, 4, (N, L))).astype(np.uint8)
    r1 = rng.integers(0, N, P).astype(np.uint32)
    r2 = ((r1 + 1 + rng.integers(0, N - 1, P)) % N).astype(np.uint32)
    b, e = column_shards(L, world)[rank]
    part = ob.Population(core[:, b:e], True).pair_counts(r1, r2) if e > b else np.zeros(P, np.uint32)
    total = allreduce_pair_counts(part)
    full = ob.Population(core, True).pair_counts(r1, r2)
    ok = bool((total == full).all())
    # tensors take the in-place path
    t = torch.from_numpy(part.astype(np.int32))
    allreduce_pair_counts(t)
    ok = ok and bool((t.numpy() == full).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("L", [3 * 8192 + 100, 5000])
def test_partial_counts_allreduce_world2(L):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, L, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_library_shards_equal_python_shards():
    """pansim_group_* (one process, N GPUs) and the torchrun plumbing must cut the alignment at the same
    sites: pansim_shard_bounds (C ABI, no device needed) against sharding.column_shards."""
    import ctypes as C
    from pansim_b200 import _ffi
    lib = _ffi.lib()
    for L in (1, 8191, 8192, 8193, 1_200_000, 5_000_000, 12_345_678):
        for n in (1, 2, 3, 4, 7, 8):
            want = column_shards(L, n)
            for i in range(n):
                b, e = C.c_uint64(), C.c_uint64()
                assert lib.pansim_shard_bounds(L, n, i, C.byref(b), C.byref(e)) == 0
                assert (b.value, e.value) == want[i], (L, n, i)
    b, e = C.c_uint64(), C.c_uint64()
    assert lib.pansim_shard_bounds(100, 2, 2, C.byref(b), C.byref(e)) == -1
